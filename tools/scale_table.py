"""Build profiles/r02_scaling.md (and copy the JSON lines it cites) from the outputs of tools/scale_r2.sh:
    python tools/scale_table.py gpurun_out/scale profiles"""
import glob
import json
import os
import shutil
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = {}
for f in sorted(glob.glob(os.path.join(src, "N*_*.json"))):
    b = os.path.basename(f)[:-5]
    n, wl = b.split("_", 1)
    try:
        d = json.load(open(f))
    except Exception:
        continue
    rows[(wl, int(n[1:]))] = d
    shutil.copy(f, os.path.join(dst, f"r02_scale_{b}.json"))
order = ["hard_pair", "commit_urgency", "escort_coalition", "burst_x2", "burst_x4", "burst_x8"]
label = {"hard_pair": "config 2: WPS_hard, Local-Hungarian + Att-Pair scores, 4096 envs PER GPU (weak)",
         "commit_urgency": "config 3: WPS_commit, UrgencyCommit on the device, 16 384 envs in total (strong)",
         "escort_coalition": "config 4: WPS_escort, Coalition-Hungarian 12, 8192 envs in total (strong)",
         "burst_x2": "config 5: WPS_burst x2 (16 agents), 65 536 envs in total (strong)",
         "burst_x4": "config 5: WPS_burst x4 (32 agents), 65 536 envs in total (strong)",
         "burst_x8": "config 5: WPS_burst x8 (64 agents), 65 536 envs in total (strong)"}
out = ["| workload | N | envs / GPU | env-steps/s | agent-steps/s | efficiency vs N=1 | e2e env-steps/s | e2e efficiency | ms / step | collective us | frac (B_alg) |",
       "|---|---|---|---|---|---|---|---|---|---|---|"]
for wl in order:
    base = rows.get((wl, 1))
    for n in (1, 2, 4, 8):
        d = rows.get((wl, n))
        if not d:
            continue
        eff = d["value"] / (n * base["value"]) if base else float("nan")
        eeff = d["e2e"]["value"] / (n * base["e2e"]["value"]) if base else float("nan")
        out.append(f"| {label[wl] if n == 1 or not base else ''} | {n} | {d['config']['envs_per_gpu']} | {d['value'] / 1e6:.2f} M | "
                   f"{d['config']['agent_steps_per_s'] / 1e6:.0f} M | {eff:.3f} | {d['e2e']['value'] / 1e6:.2f} M | {eeff:.3f} | "
                   f"{d['ms_per_step']:.4f} | {d['collective_us']:.1f} | {d['roofline']['frac']:.3f} |")
print("\n".join(out))
