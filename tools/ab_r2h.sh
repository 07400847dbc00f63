#!/bin/bash
set -u
O=gpurun_out/r2h; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/gputests.log 2>&1; echo "gputests rc=$?" >> $O/gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
python bench.py --steps 20 --warmup 5 --cpu-seconds 10 > $O/b_driver_like.json 2> $O/b_driver_like.err
python bench.py --steps 1500 --warmup 150 --no-cpu-baseline > $O/b_default_1500.json 2> $O/b_default_1500.err
python bench.py --workload hard_local --steps 150 --warmup 5 --no-cpu-baseline > $O/b_hard_local.json 2> $O/b_hard_local.err
python bench.py --workload attn_context --steps 150 --warmup 5 --no-cpu-baseline > $O/b_attn_context.json 2> $O/b_attn_context.err
python bench.py --workload hard_pi --steps 150 --warmup 5 --no-cpu-baseline > $O/b_hard_pi.json 2> $O/b_hard_pi.err
python bench.py --workload escort_pi --envs 8192 --unique-seeds 1024 --steps 150 --warmup 5 --no-cpu-baseline > $O/b_escort_pi.json 2> $O/b_escort_pi.err
python tools/kbench.py WPS_hard 4096 > $O/kb_hard.json 2> $O/kb_hard.err
KB_TASK_CAP=32 python tools/kbench.py WPS_hard 4096 > $O/kb_hard32.json 2> $O/kb_hard32.err
SKIP_HARD=1 bash tools/scale_r2.sh 1
C="python bench.py --steps 30 --warmup 3 --no-cpu-baseline"
$C > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 200 --csv --log-file $O/launches_bench.csv $C > $O/ncu_l.log 2>&1
$C > $O/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:muav_step_kernel -s 20 -c 2 -o $O/prof_step $C > $O/ncu_s.log 2>&1
python tools/fused_scorer_prof.py 4096 > $O/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:att_pair -s 3 -c 1 -o $O/prof_scorer python tools/fused_scorer_prof.py 4096 > $O/ncu_c.log 2>&1
echo done > $O/done
