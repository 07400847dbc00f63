#!/bin/bash
# BASELINE config 2 (default bench, weak scaling) on N GPUs of one box, the driver's short command and a whole-episode run.
#   gpurun --gpus N -- 'bash tools/scale_r2_hard.sh N'
set -u
N=$1
O=gpurun_out/scale_tc; mkdir -p $O
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
run hard_pair_20 --steps 20 --warmup 5
run hard_pair_150 --steps 150 --warmup 10
echo done >> $O/N${N}.log
