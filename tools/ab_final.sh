#!/bin/bash
# final round-2 evidence: full GPU suite, smoke, the driver's bench command, launch list and ncu captures of both hot kernels
set -u
O=gpurun_out/final; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/gputests.log 2>&1; echo "gputests rc=$?" >> $O/gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
C="python bench.py --steps 30 --warmup 3 --no-cpu-baseline"
$C > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 200 --csv --log-file $O/launches_bench.csv $C > $O/ncu_l.log 2>&1
$C > $O/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:muav_step_kernel -s 20 -c 2 -o $O/prof_step $C > $O/ncu_s.log 2>&1
$C > $O/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:att_tc_kernel -s 20 -c 2 -o $O/prof_tc_bench $C > $O/ncu_t.log 2>&1
python tools/fused_scorer_prof.py 4096 > $O/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:att_tc_kernel -s 3 -c 1 -o $O/prof_tc_full python tools/fused_scorer_prof.py 4096 > $O/ncu_f.log 2>&1
python bench.py --steps 20 --warmup 5 > $O/b_driver_like.json 2> $O/b_driver_like.err
python bench.py --impl reference --steps 20 --warmup 5 --cpu-seconds 20 > $O/b_ref.json 2> $O/b_ref.err
echo done > $O/done
