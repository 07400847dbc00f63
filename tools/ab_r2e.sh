#!/bin/bash
set -u
O=gpurun_out/r2e; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/gputests.log 2>&1; echo "gputests rc=$?" >> $O/gputests.log
python tools/kbench.py WPS_hard 4096 > $O/kb_fixed.json 2> $O/kb_fixed.err
KB_TASK_CAP=32 python tools/kbench.py WPS_hard 4096 > $O/kb_fixed32.json 2> $O/kb_fixed32.err
python tools/kbench.py WPS_commit 16384 > $O/kb_commit.json 2> $O/kb_commit.err
python tools/kbench.py WPS_escort 8192 > $O/kb_escort.json 2> $O/kb_escort.err
MUAV_LIB_OVERRIDE=$PWD/build/libmuav_b200_lean128.so python tools/kbench.py WPS_commit 16384 > $O/kb_commit_lean128.json 2> /dev/null
MUAV_LIB_OVERRIDE=$PWD/build/libmuav_b200_lean128.so python tools/kbench.py WPS_escort 8192 > $O/kb_escort_lean128.json 2> /dev/null
MUAV_STAGE_COLD=1 python tools/kbench.py WPS_commit 16384 > $O/kb_commit_cold1.json 2> /dev/null
MUAV_STAGE_COLD=1 python tools/kbench.py WPS_escort 8192 > $O/kb_escort_cold1.json 2> /dev/null
B="python bench.py --steps 300 --warmup 20 --no-cpu-baseline"
$B > $O/b_fixed.json 2> $O/b_fixed.err
$B --task-cap 32 > $O/b_fixed32.json 2> $O/b_fixed32.err
MUAV_STAGE_COLD=0 $B > $O/b_fixed_cold0.json 2> /dev/null
MUAV_STAGE_COLD=0 $B --task-cap 32 > $O/b_fixed32_cold0.json 2> /dev/null
MUAV_STAGE_COLD=0 MUAV_LIB_OVERRIDE=$PWD/build/libmuav_b200_lb3.so $B --task-cap 32 > $O/b_fixed32_cold0_lb3.json 2> /dev/null
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --task-cap 32 > $O/b_fixed32_20.json 2> $O/b_fixed32_20.err
for wl in escort_coalition commit_urgency; do
  MUAV_LIB_OVERRIDE=$PWD/build/libmuav_b200_lean128.so python bench.py --workload $wl --envs 8192 --unique-seeds 512 --steps 150 --warmup 5 --no-cpu-baseline > $O/b_${wl}_lean128.json 2> /dev/null
  MUAV_SPLIT_STEP=0 python bench.py --workload $wl --envs 8192 --unique-seeds 512 --steps 150 --warmup 5 --no-cpu-baseline > $O/b_${wl}_nosplit.json 2> /dev/null
done
python bench.py --steps 120 --warmup 20 --no-cpu-baseline --task-cap 32 > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:muav_step_kernel -s 90 -c 2 -o $O/prof_step_fixed32 python bench.py --steps 120 --warmup 20 --no-cpu-baseline --task-cap 32 > $O/ncu.log 2>&1
echo done > $O/done
