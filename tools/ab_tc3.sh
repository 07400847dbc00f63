#!/bin/bash
set -u
O=gpurun_out/tc5; mkdir -p $O
timeout 300 python tools/fused_scorer_prof.py 4096 > $O/plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:att_pair_tc -s 3 -c 1 -o $O/prof_tc python tools/fused_scorer_prof.py 4096 > $O/ncu.log 2>&1
echo done >> $O/plain.log
