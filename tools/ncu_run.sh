python bench.py --steps 30 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 200 --csv --log-file gpurun_out/launches_r1_lean.csv python bench.py --steps 30 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_list3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:muav_step_kernel -s 20 -c 2 -o gpurun_out/prof_step_r1_lean python bench.py --steps 30 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out/prof_step_r1_lean.ncu-rep gpurun_out/launches_r1_lean.csv
