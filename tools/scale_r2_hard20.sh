#!/bin/bash
# the driver's short command on N GPUs: gpurun --gpus N -- 'bash tools/scale_r2_hard20.sh N'
set -u
N=$1
O=gpurun_out/scale_last; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --no-cpu-baseline --steps 20 --warmup 5 > $O/N${N}_hard_pair_20.json 2> $O/N${N}_hard_pair_20.err
echo "rc=$?" >> $O/N${N}.log
