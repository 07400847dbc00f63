#!/bin/bash
set -u
O=gpurun_out/r2j; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/gputests.log 2>&1; echo "gputests rc=$?" >> $O/gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
python bench.py --workload hard_cbba --steps 150 --warmup 5 --no-cpu-baseline > $O/b_hard_cbba.json 2> $O/b_hard_cbba.err
python bench.py --workload escort_cbba --envs 8192 --unique-seeds 1024 --steps 150 --warmup 5 --no-cpu-baseline > $O/b_escort_cbba.json 2> $O/b_escort_cbba.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/b_default20.json 2> $O/b_default20.err
echo done > $O/done
