#!/bin/bash
# the driver's bench command on the round's last build (CPU arm shortened to 10 s to fit the remaining GPU budget)
set -u
O=gpurun_out/final_r2; mkdir -p $O
timeout 100 python bench.py --steps 20 --warmup 5 --cpu-seconds 10 > $O/b_20_full.json 2> $O/b_20_full.err; echo "rc=$?" >> $O/b_20_full.err
tail -2 $O/b_20_full.err; cut -c1-200 $O/b_20_full.json
