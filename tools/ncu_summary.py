"""Summarise an `ncu --set full` capture of one of our kernels (read here, on the CPU box) and record its DRAM traffic
for bench.py's roofline.traffic.

    python tools/ncu_summary.py gpurun_out/r2b/prof_step_fixed.ncu-rep --key hard_pair:4096 \
        --note "bench mode, steps 70/71 of an episode" [--no-traffic]

Prints a markdown table (paste into profiles/rNN_*.md) and updates profiles/step_kernel_traffic.json with
{key: {dram_bytes_per_launch, report, source_hash, note}}; the hash is the one bench.py computes over csrc/, so a
capture goes stale as soon as a kernel source changes."""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WANT = [
    ("gpu__time_duration.sum", "duration (us, under ncu)", "us"),
    ("dram__bytes_read.sum", "dram read (MB)", "MB"),
    ("dram__bytes_write.sum", "dram write (MB)", "MB"),
    ("smsp__inst_executed.sum", "warp instructions", 1),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instruction", 1),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC per SM", 1),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots busy (%)", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active (% of 64/SM)", 1),
    ("launch__registers_per_thread", "registers / thread", 1),
    ("launch__grid_size", "grid", 1),
    ("launch__block_size", "block", 1),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA (KB)", "KB"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe (%)", 1),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe (%)", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (%)", 1),
    ("sm__icc_requests_lookup_hit.avg.pct", "icache hit (%)", 1),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts", 1),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts", 1),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--key", default=None, help="<workload>:<envs per GPU> entry of profiles/step_kernel_traffic.json")
    ap.add_argument("--note", default="")
    ap.add_argument("--no-traffic", action="store_true")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    data = rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"kernels captured: {len(data)}: " + ", ".join(sorted(set(r[ix['Kernel Name']][:60] for r in data))))
    print("\n| metric | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |")
    print("|---|" + "---|" * len(data))
    vals = {}
    for key, label, scale in WANT:
        cands = [h for h in hdr if h == key] or [h for h in hdr if h.startswith(key)]
        if not cands:
            continue
        col = ix[cands[0]]
        unit = rows[1][col].lower().split("/")[0]
        # the raw page states each column's unit: normalise to the unit the label promises
        to_base = {"nsecond": 1e-9, "ns": 1e-9, "usecond": 1e-6, "us": 1e-6, "msecond": 1e-3, "ms": 1e-3, "second": 1.0,
                   "s": 1.0, "byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
        if isinstance(scale, str):
            scale = to_base / {"us": 1e-6, "MB": 1e6, "KB": 1e3}[scale]
        v = []
        for r in data:
            try:
                v.append(float(r[col].replace(",", "")) * scale)
            except ValueError:
                v.append(float("nan"))
        vals[key] = v
        print(f"| {label} | " + " | ".join(f"{x:.6g}" for x in v) + " |")
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
             or h.startswith("smsp__average_warp_latency_issue_stalled_")]
    st = {}
    for h in stall:
        try:
            st[h] = sum(float(r[ix[h]].replace(",", "")) for r in data) / len(data)
        except ValueError:
            pass
    tot = sum(st.values()) or 1.0
    top = sorted(st.items(), key=lambda kv: -kv[1])[:8]
    print("\nStall reasons (share of warp-cycles per issue): " + ", ".join(
        f"{k.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '').replace('.ratio','')} {100 * v / tot:.0f} %" for k, v in top))
    if args.key and not args.no_traffic:
        import bench
        rd, wr = vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"]
        per = sum(a + b for a, b in zip(rd, wr)) / len(rd) * 1e6
        path = bench.TRAFFIC_FILE
        rec = json.load(open(path)) if os.path.exists(path) else {}
        rec[args.key] = {"dram_bytes_per_launch": per, "report": os.path.basename(args.report),
                         "source_hash": bench.kernel_source_hash(), "note": args.note}
        json.dump(rec, open(path, "w"), indent=1, sort_keys=True)
        print(f"\nrecorded {per / 1e6:.2f} MB per launch under {args.key} in {os.path.relpath(path, ROOT)}")


if __name__ == "__main__":
    main()
