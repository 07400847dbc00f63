#!/bin/bash
set -u
O=gpurun_out/tc2; mkdir -p $O
timeout 300 python tools/tc_scorer_check.py 4096 > $O/check.log 2>&1; rc=$?; echo "rc=$rc" >> $O/check.log
[ $rc -ne 0 ] && exit 0
timeout 900 python -m pytest tests -m gpu -q -k "scorer or pair or smoke or flips or context" > $O/gputests.log 2>&1; echo "rc=$?" >> $O/gputests.log
python bench.py --steps 300 --warmup 5 --no-cpu-baseline > $O/b_tc.json 2> $O/b_tc.err
MUAV_SCORER_TC=0 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > $O/b_fp32.json 2> $O/b_fp32.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/b_tc20.json 2> $O/b_tc20.err
