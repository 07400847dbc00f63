#!/bin/bash
set -u
O=gpurun_out/tc9; mkdir -p $O
timeout 300 python tools/tc_scorer_check.py 4096 > $O/check.log 2>&1; rc=$?; echo "rc=$rc" >> $O/check.log
[ $rc -ne 0 ] && exit 0
python bench.py --steps 300 --warmup 5 --no-cpu-baseline > $O/b_tc.json 2> $O/b_tc.err
