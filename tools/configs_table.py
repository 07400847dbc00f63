"""Regenerate the table of profiles/r01_configs.md from the committed bench lines."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = [("hard_pair (2: WPS_hard, Local-Hungarian + Att-Pair scores) — default bench", "r01_bench_default_latest"),
        ("hard_local (WPS_hard, plain Local-Hungarian: the north_star target shape)", "r01_bench_hard_local"),
        ("commit_urgency (3: WPS_commit, UrgencyCommit planner on device)", "r01_bench_commit_urgency"),
        ("escort_coalition (4: WPS_escort, Coalition-Hungarian 12)", "r01_bench_escort_coalition"),
        ("attn_context (WPS_attn, Att-ContextPair: the paper's primary method; fused context scorer)", "r01_bench_attn_context"),
        ("burst_x2 (5)", "r01_bench_burst_x2"), ("burst_x4 (5)", "r01_bench_burst_x4"), ("burst_x8 (5)", "r01_bench_burst_x8")]
print("| workload (BASELINE config) | agents | envs | env-steps/s | agent-steps/s | ms/step | e2e env-steps/s | record B | frac (B_alg) | record I/O GB/s | error_flags |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for name, f in rows:
    r = json.load(open(os.path.join(ROOT, "profiles", f + ".json")))
    c = r["config"]
    A = round(c["agent_steps_per_s"] / r["value"])
    print(f"| {name} | {A} | {c['envs_per_gpu']} | {r['value'] / 1e6:.2f} M | {c['agent_steps_per_s'] / 1e6:.1f} M | "
          f"{r['ms_per_step']:.3f} | {r['e2e']['value'] / 1e6:.2f} M | {c['record_bytes']} | {r['roofline']['frac']:.3f} | "
          f"{r['roofline']['record_io_gbs']:.0f} | {r['error_flags']} |")
