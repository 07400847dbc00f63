#!/bin/bash
# BASELINE configs 3-5 sharded over N GPUs of one box (strong scaling, fixed global batch) + config 2 (weak scaling).
#   gpurun --gpus N -- 'bash tools/scale_r2.sh N'
set -u
N=$1
O=gpurun_out/scale; mkdir -p $O
run() {  # name, extra args...
  local name=$1; shift
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --no-cpu-baseline "$@" > $O/N${N}_$name.json 2> $O/N${N}_$name.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --no-cpu-baseline "$@" > $O/N${N}_$name.json 2> $O/N${N}_$name.err
  fi
  echo "$name rc=$?" >> $O/N${N}.log
}
# whole episodes (150 steps): the first 60 steps of an episode replan much more often than the rest
[ "${SKIP_HARD:-0}" = "1" ] || run hard_pair --steps 150 --warmup 10
run commit_urgency --workload commit_urgency --global-envs 16384 --unique-seeds 2048 --steps 150 --warmup 5
run escort_coalition --workload escort_coalition --global-envs 8192 --unique-seeds 1024 --steps 150 --warmup 5
run burst_x2 --workload burst_x2 --global-envs 65536 --unique-seeds 1024 --steps 150 --warmup 5
run burst_x4 --workload burst_x4 --global-envs 65536 --unique-seeds 512 --steps 150 --warmup 5
run burst_x8 --workload burst_x8 --global-envs 65536 --unique-seeds 256 --steps 150 --warmup 3
echo done >> $O/N${N}.log
