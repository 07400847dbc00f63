#!/bin/bash
set -u
O=gpurun_out/tc12; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "commit or scorer" > $O/gputests.log 2>&1; rc=$?; echo "rc=$rc" >> $O/gputests.log
[ $rc -ne 0 ] && exit 0
python bench.py --workload commit_att --envs 16384 --unique-seeds 2048 --steps 150 --warmup 5 --no-cpu-baseline > $O/b_commit_att_tc.json 2> $O/b_commit_att_tc.err
MUAV_SCORER_TC=0 python bench.py --workload commit_att --envs 16384 --unique-seeds 2048 --steps 150 --warmup 5 --no-cpu-baseline > $O/b_commit_att_fp32.json 2> $O/b_commit_att_fp32.err
