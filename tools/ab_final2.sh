#!/bin/bash
set -u
O=gpurun_out/final2; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/gputests.log 2>&1; echo "gputests rc=$?" >> $O/gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
C="python bench.py --steps 30 --warmup 3 --no-cpu-baseline"
$C > $O/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:muav_step_kernel -s 20 -c 2 -o $O/prof_step $C > $O/ncu_s.log 2>&1
python bench.py --steps 300 --warmup 5 --no-cpu-baseline > $O/b_300.json 2> $O/b_300.err
python bench.py --workload hard_local --steps 300 --warmup 5 --no-cpu-baseline > $O/b_hard_local.json 2> $O/b_hard_local.err
echo done > $O/done
