set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/gputests_r1_final2.log
python bench.py > gpurun_out/bench_r1_final2.json 2> gpurun_out/bench_r1_final2.err
tail -c 300 gpurun_out/bench_r1_final2.err
for w in hard_local commit_urgency escort_coalition hard_pi escort_pi attn_context burst_x2 burst_x4; do
  E=4096; [ $w = commit_urgency ] && E=16384; [ $w = escort_coalition ] && E=8192; [ $w = escort_pi ] && E=8192; [ ${w:0:5} = burst ] && E=8192
  python bench.py --workload $w --envs $E --steps 300 --warmup 150 --cpu-seconds 0 > gpurun_out/cfg2_$w.json 2> gpurun_out/cfg2_$w.err
done
python - <<'PY'
import json,glob
for f in ["gpurun_out/bench_r1_final2.json"]+sorted(glob.glob("gpurun_out/cfg2_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]/1e6,3), "e2e", round(d["e2e"]["value"]/1e6,3), "k_ms", round(d["roofline"]["kernel_ms_per_launch"],4), "frac", round(d["roofline"]["frac"],4), "err", d["error_flags"])
    except Exception as e: print(f, "ERR", e)
PY
