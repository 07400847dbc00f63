#!/bin/bash
# round-2 A/B on one B200: round-1 library vs fixed-shape instantiations (development aid)
set -u
O=gpurun_out/r2; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
R1=$PWD/build/libmuav_b200_r1.so
timeout 600 python -m pytest tests -m gpu -x -q > $O/gputests_a.log 2>&1; echo "gputests rc=$?" >> $O/gputests_a.log
for rep in 1 2; do
  MUAV_LIB_OVERRIDE=$R1 python tools/kbench.py WPS_hard 4096 > $O/kb_r1_$rep.json 2> $O/kb_r1_$rep.err
  python tools/kbench.py WPS_hard 4096 > $O/kb_fixed_$rep.json 2> $O/kb_fixed_$rep.err
  MUAV_NO_FIXED_SHAPE=1 python tools/kbench.py WPS_hard 4096 > $O/kb_lean_$rep.json 2> $O/kb_lean_$rep.err
  KB_TASK_CAP=32 python tools/kbench.py WPS_hard 4096 > $O/kb_fixed32_$rep.json 2> $O/kb_fixed32_$rep.err
  KB_TASK_CAP=32 MUAV_NO_FIXED_SHAPE=1 python tools/kbench.py WPS_hard 4096 > $O/kb_lean32_$rep.json 2> $O/kb_lean32_$rep.err
done
MUAV_LIB_OVERRIDE=$R1 python bench.py --steps 300 --warmup 20 --cpu-seconds 0 --no-cpu-baseline > $O/b_r1.json 2> $O/b_r1.err
python bench.py --steps 300 --warmup 20 --no-cpu-baseline > $O/b_fixed.json 2> $O/b_fixed.err
python bench.py --steps 300 --warmup 20 --no-cpu-baseline --task-cap 32 > $O/b_fixed32.json 2> $O/b_fixed32.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/b_fixed_20.json 2> $O/b_fixed_20.err
for w in 4 8 12; do MUAV_CTA_WARPS=$w KB_TASK_CAP=32 python tools/kbench.py WPS_hard 4096 > $O/kb_fixed32_w$w.json 2>/dev/null; done
echo done > $O/ab_r2a.done
