#!/bin/bash
set -u
O=gpurun_out/r2d; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/gputests.log 2>&1; echo "gputests rc=$?" >> $O/gputests.log
python tools/kbench.py WPS_hard 4096 > $O/kb_fixed.json 2> $O/kb_fixed.err
KB_TASK_CAP=32 python tools/kbench.py WPS_hard 4096 > $O/kb_fixed32.json 2> $O/kb_fixed32.err
MUAV_NO_FIXED_SHAPE=1 python tools/kbench.py WPS_hard 4096 > $O/kb_lean.json 2> $O/kb_lean.err
python tools/kbench.py WPS_commit 16384 > $O/kb_commit.json 2> $O/kb_commit.err
python tools/kbench.py WPS_escort 8192 > $O/kb_escort.json 2> $O/kb_escort.err
B="python bench.py --steps 300 --warmup 20 --no-cpu-baseline"
$B > $O/b_fixed.json 2> $O/b_fixed.err
$B --task-cap 32 > $O/b_fixed32.json 2> $O/b_fixed32.err
MUAV_NO_FIXED_SHAPE=1 $B > $O/b_lean.json 2> $O/b_lean.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/b_fixed_20.json 2> $O/b_fixed_20.err
for L in lb2 lb4; do
  MUAV_LIB_OVERRIDE=$PWD/build/libmuav_b200_$L.so $B > $O/b_${L}.json 2>/dev/null
  MUAV_LIB_OVERRIDE=$PWD/build/libmuav_b200_$L.so $B --task-cap 32 > $O/b32_${L}.json 2>/dev/null
done
for w in 4 5 6; do MUAV_CTA_WARPS=$w $B > $O/b_w$w.json 2>/dev/null; MUAV_CTA_WARPS=$w $B --task-cap 32 > $O/b32_w$w.json 2>/dev/null; done
for w in 4 5 6 8; do MUAV_LIB_OVERRIDE=$PWD/build/libmuav_b200_lb4.so MUAV_CTA_WARPS=$w $B --task-cap 32 > $O/b32_lb4_w$w.json 2>/dev/null; done
for wl in escort_coalition burst_x4 burst_x8 commit_urgency hard_local; do python bench.py --workload $wl --envs 8192 --unique-seeds 512 --steps 150 --warmup 5 --no-cpu-baseline > $O/b_$wl.json 2> $O/b_$wl.err; done
python bench.py --steps 120 --warmup 20 --no-cpu-baseline --task-cap 32 > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:muav_step_kernel -s 90 -c 2 -o $O/prof_step_fixed32 python bench.py --steps 120 --warmup 20 --no-cpu-baseline --task-cap 32 > $O/ncu.log 2>&1
echo done > $O/done
