#!/bin/bash
set -u
O=gpurun_out/r2i; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/gputests.log 2>&1; echo "gputests rc=$?" >> $O/gputests.log
B="python bench.py --steps 300 --warmup 20 --no-cpu-baseline"
for rep in 1 2; do
  $B > $O/b_heavy_$rep.json 2> $O/b_heavy_$rep.err
  MUAV_SYNC_MASK=159 $B > $O/b_noheavy_$rep.json 2> /dev/null
  MUAV_LIB_OVERRIDE=$PWD/build/libmuav_b200_tm.so $B > $O/b_tm_$rep.json 2> $O/b_tm_$rep.err
done
python tools/kbench.py WPS_hard 4096 > $O/kb_heavy.json 2>/dev/null
MUAV_SYNC_MASK=159 python tools/kbench.py WPS_hard 4096 > $O/kb_noheavy.json 2>/dev/null
KB_TASK_CAP=32 python tools/kbench.py WPS_hard 4096 > $O/kb32_heavy.json 2>/dev/null
KB_TASK_CAP=32 MUAV_SYNC_MASK=159 python tools/kbench.py WPS_hard 4096 > $O/kb32_noheavy.json 2>/dev/null
python tools/kbench.py WPS_commit 16384 > $O/kb_commit_heavy.json 2>/dev/null
MUAV_SYNC_MASK=159 python tools/kbench.py WPS_commit 16384 > $O/kb_commit_noheavy.json 2>/dev/null
python tools/fused_scorer_prof.py 4096 > $O/scorer_default.log 2>&1
MUAV_LIB_OVERRIDE=$PWD/build/libmuav_b200_tm.so python tools/fused_scorer_prof.py 4096 > $O/scorer_tm.log 2>&1
MUAV_LIB_OVERRIDE=$PWD/build/libmuav_b200_tm.so timeout 600 python -m pytest tests -m gpu -q -k "fused_scorer or context_scorer" > $O/gputests_tm.log 2>&1; echo "rc=$?" >> $O/gputests_tm.log
echo done > $O/done
