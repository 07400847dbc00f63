#!/bin/bash
# traffic capture of the step kernel for the round's last sources (bench.py's roofline.traffic reads the summary)
set -u
O=gpurun_out/final_r2; mkdir -p $O
C="python bench.py --steps 30 --warmup 3 --no-cpu-baseline"
timeout 170 ncu --set full --clock-control none --import-source on -k regex:muav_step_kernel -s 20 -c 2 -o $O/prof_step $C > $O/ncu_s.log 2>&1
echo "ncu rc=$?" >> $O/ncu_s.log; tail -2 $O/ncu_s.log; ls -la $O
