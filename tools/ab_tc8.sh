#!/bin/bash
set -u
O=gpurun_out/tc8; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/gputests.log 2>&1; echo "gputests rc=$?" >> $O/gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
python bench.py --steps 20 --warmup 5 --cpu-seconds 20 > $O/b_driver_like.json 2> $O/b_driver_like.err
python bench.py --steps 1500 --warmup 5 --no-cpu-baseline > $O/b_1500.json 2> $O/b_1500.err
MUAV_SCORER_TC=0 python bench.py --steps 1500 --warmup 5 --no-cpu-baseline > $O/b_1500_fp32.json 2> $O/b_1500_fp32.err
