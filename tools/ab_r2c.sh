#!/bin/bash
set -u
O=gpurun_out/r2c; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests.log 2>&1; echo "gputests rc=$?" >> $O/gputests.log
python tools/kbench.py WPS_hard 4096 > $O/kb_fixed.json 2> $O/kb_fixed.err
MUAV_NO_FIXED_SHAPE=1 python tools/kbench.py WPS_hard 4096 > $O/kb_lean.json 2> $O/kb_lean.err
KB_TASK_CAP=32 python tools/kbench.py WPS_hard 4096 > $O/kb_fixed32.json 2> $O/kb_fixed32.err
python tools/kbench.py WPS_commit 16384 > $O/kb_commit.json 2> $O/kb_commit.err
python tools/kbench.py WPS_escort 8192 > $O/kb_escort.json 2> $O/kb_escort.err
B="python bench.py --steps 300 --warmup 20 --no-cpu-baseline"
$B > $O/b_fixed.json 2> $O/b_fixed.err
MUAV_NO_FIXED_SHAPE=1 $B > $O/b_lean.json 2> $O/b_lean.err
$B --task-cap 32 > $O/b_fixed32.json 2> $O/b_fixed32.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/b_fixed_20.json 2> $O/b_fixed_20.err
for m in 0 1 17 19 27; do MUAV_SYNC_MASK=$m $B > $O/b_mask$m.json 2>/dev/null; done
for w in 2 3 4 8 12; do MUAV_CTA_WARPS=$w $B > $O/b_w$w.json 2>/dev/null; done
for w in 4 5 8 10; do MUAV_CTA_WARPS=$w $B --task-cap 32 > $O/b32_w$w.json 2>/dev/null; done
python bench.py --steps 120 --warmup 20 --no-cpu-baseline > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:muav_step_kernel -s 90 -c 2 -o $O/prof_step_fixed python bench.py --steps 120 --warmup 20 --no-cpu-baseline > $O/ncu.log 2>&1
echo done > $O/done
