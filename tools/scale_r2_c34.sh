#!/bin/bash
# BASELINE configs 3 and 4 sharded over N GPUs (strong scaling, fixed global batch): gpurun --gpus N -- 'bash tools/scale_r2_c34.sh N'
set -u
N=$1
O=gpurun_out/scale_c34; mkdir -p $O
run() {
  local name=$1; shift
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --no-cpu-baseline "$@" > $O/N${N}_$name.json 2> $O/N${N}_$name.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --no-cpu-baseline "$@" > $O/N${N}_$name.json 2> $O/N${N}_$name.err
  fi
  echo "$name rc=$?" >> $O/N${N}.log
}
run commit_urgency --workload commit_urgency --global-envs 16384 --unique-seeds 2048 --steps 150 --warmup 5
run escort_coalition --workload escort_coalition --global-envs 8192 --unique-seeds 1024 --steps 150 --warmup 5
echo done >> $O/N${N}.log
