#!/bin/bash
set -u
O=gpurun_out/r2g; mkdir -p $O
timeout 60 python tools/tc_probe/run.py > $O/tc_probe.log 2>&1; echo "probe rc=$?" >> $O/tc_probe.log
timeout 900 python -m pytest tests -m gpu -q -x -k "golden or lean or split or task_cap" > $O/gputests.log 2>&1; echo "gputests rc=$?" >> $O/gputests.log
python tools/kbench.py WPS_commit 16384 > $O/kb_commit.json 2> $O/kb_commit.err
python tools/kbench.py WPS_escort 8192 > $O/kb_escort.json 2> $O/kb_escort.err
python tools/kbench.py WPS_hard 4096 > $O/kb_hard.json 2> $O/kb_hard.err
B="python bench.py --no-cpu-baseline"
$B --steps 300 --warmup 20 > $O/b_default.json 2> $O/b_default.err
$B --workload commit_urgency --envs 8192 --unique-seeds 512 --steps 150 --warmup 5 > $O/b_commit_8192_u512.json 2>/dev/null
$B --workload commit_urgency --envs 16384 --unique-seeds 512 --steps 60 --warmup 5 > $O/b_commit_16384_u512.json 2>/dev/null
$B --workload commit_urgency --envs 16384 --unique-seeds 2048 --steps 60 --warmup 5 > $O/b_commit_16384_u2048.json 2>/dev/null
$B --workload escort_coalition --envs 8192 --unique-seeds 512 --steps 150 --warmup 5 > $O/b_escort_8192_u512.json 2>/dev/null
$B --workload escort_coalition --envs 8192 --unique-seeds 1024 --steps 60 --warmup 5 > $O/b_escort_8192_u1024_60.json 2>/dev/null
$B --workload escort_coalition --envs 8192 --unique-seeds 1024 --steps 150 --warmup 5 > $O/b_escort_8192_u1024_150.json 2>/dev/null
echo done > $O/done
