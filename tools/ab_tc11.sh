#!/bin/bash
set -u
O=gpurun_out/tc11; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "coalition or escort or scorer" > $O/gputests.log 2>&1; rc=$?; echo "rc=$rc" >> $O/gputests.log
[ $rc -ne 0 ] && exit 0
python bench.py --workload escort_att --envs 8192 --unique-seeds 2048 --steps 150 --warmup 5 --no-cpu-baseline > $O/b_escort_att.json 2> $O/b_escort_att.err
