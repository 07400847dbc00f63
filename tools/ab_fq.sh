#!/bin/bash
set -u
O=gpurun_out/fq; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/gputests.log 2>&1; echo "gputests rc=$?" >> $O/gputests.log
for v in fixed nofixed; do
  if [ $v = nofixed ]; then export MUAV_NO_FIXED_SHAPE=1; fi
  python bench.py --workload hard_pi --steps 150 --warmup 5 --no-cpu-baseline > $O/b_hard_pi_$v.json 2> $O/b_hard_pi_$v.err
  python bench.py --workload burst_x2 --envs 65536 --unique-seeds 1024 --steps 150 --warmup 5 --no-cpu-baseline > $O/b_burst_x2_$v.json 2> $O/b_burst_x2_$v.err
  python bench.py --workload burst_x4 --envs 65536 --unique-seeds 512 --steps 150 --warmup 5 --no-cpu-baseline > $O/b_burst_x4_$v.json 2> $O/b_burst_x4_$v.err
  python bench.py --workload burst_x8 --envs 65536 --unique-seeds 256 --steps 60 --warmup 3 --no-cpu-baseline > $O/b_burst_x8_$v.json 2> $O/b_burst_x8_$v.err
done
