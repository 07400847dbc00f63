#!/bin/bash
set -u
O=gpurun_out/r2f; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/gputests.log 2>&1; echo "gputests rc=$?" >> $O/gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
B="python bench.py --steps 300 --warmup 20 --no-cpu-baseline"
$B > $O/b_default.json 2> $O/b_default.err
for w in 5 6 7; do MUAV_CTA_WARPS=$w $B > $O/b_w$w.json 2> $O/b_w$w.err; done
$B --task-cap -1 > $O/b_tc48.json 2> $O/b_tc48.err
python bench.py --steps 20 --warmup 5 --cpu-seconds 5 > $O/b_driver_like.json 2> $O/b_driver_like.err
python bench.py --impl reference --steps 20 --warmup 5 --cpu-seconds 5 > $O/b_ref.json 2> $O/b_ref.err
python bench.py --workload hard_local --steps 150 --warmup 5 --no-cpu-baseline > $O/b_hard_local.json 2> $O/b_hard_local.err
bash tools/scale_r2.sh 1
echo done > $O/done
