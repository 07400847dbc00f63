#!/bin/bash
O=gpurun_out/ll; mkdir -p $O
C="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
$C > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv $C > $O/ncu_l.log 2>&1
echo done >> $O/plain.log
