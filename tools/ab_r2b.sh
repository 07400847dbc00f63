#!/bin/bash
# round-2 A/B (b): lane-parallel FSM / threat pursuit, fixed-shape instantiation; ncu capture of the step kernel in bench mode
set -u
O=gpurun_out/r2b; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests.log 2>&1; echo "gputests rc=$?" >> $O/gputests.log
python tools/kbench.py WPS_hard 4096 > $O/kb_fixed.json 2> $O/kb_fixed.err
MUAV_NO_FIXED_SHAPE=1 python tools/kbench.py WPS_hard 4096 > $O/kb_lean.json 2> $O/kb_lean.err
KB_TASK_CAP=32 python tools/kbench.py WPS_hard 4096 > $O/kb_fixed32.json 2> $O/kb_fixed32.err
MUAV_CTA_WARPS=8 KB_TASK_CAP=32 python tools/kbench.py WPS_hard 4096 > $O/kb_fixed32_w8.json 2> /dev/null
python tools/kbench.py WPS_commit 16384 > $O/kb_commit.json 2> $O/kb_commit.err
python tools/kbench.py WPS_escort 8192 > $O/kb_escort.json 2> $O/kb_escort.err
python bench.py --steps 300 --warmup 20 --no-cpu-baseline > $O/b_fixed.json 2> $O/b_fixed.err
MUAV_NO_FIXED_SHAPE=1 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > $O/b_lean.json 2> $O/b_lean.err
python bench.py --steps 300 --warmup 20 --no-cpu-baseline --task-cap 32 > $O/b_fixed32.json 2> $O/b_fixed32.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/b_fixed_20.json 2> $O/b_fixed_20.err
python bench.py --steps 120 --warmup 20 --no-cpu-baseline > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:muav_step_kernel -s 90 -c 2 -o $O/prof_step_fixed python bench.py --steps 120 --warmup 20 --no-cpu-baseline > $O/ncu.log 2>&1
echo done > $O/done
