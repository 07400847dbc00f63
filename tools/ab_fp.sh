#!/bin/bash
set -u
O=gpurun_out/fp; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/gputests.log 2>&1; echo "gputests rc=$?" >> $O/gputests.log
for v in fixed nofixed; do
  if [ $v = nofixed ]; then export MUAV_NO_FIXED_SHAPE=1; fi
  python bench.py --workload commit_urgency --envs 16384 --unique-seeds 2048 --steps 150 --warmup 5 --no-cpu-baseline > $O/b_commit_urgency_$v.json 2> $O/b_commit_urgency_$v.err
  python bench.py --workload commit_att --envs 16384 --unique-seeds 2048 --steps 150 --warmup 5 --no-cpu-baseline > $O/b_commit_att_$v.json 2> $O/b_commit_att_$v.err
  python bench.py --workload escort_pi --envs 8192 --unique-seeds 1024 --steps 150 --warmup 5 --no-cpu-baseline > $O/b_escort_pi_$v.json 2> $O/b_escort_pi_$v.err
done
