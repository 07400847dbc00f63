"""Generate tests/golden/wps_eval_scores.json: the return dicts of the UNMODIFIED reference's run_wps_episode
(experiments/wps_eval.py:76-290), seeds 0-3, floats as float.hex().  Authoring container only (/root/reference
through tests/golden/refshim.py).

    python tests/golden/gen_eval_scores.py            # rewrites every row
    python tests/golden/gen_eval_scores.py 'WPS_hard|Local-PI'   # adds / refreshes the named rows only
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

ROWS = ["WPS_hard|Local-Hungarian", "WPS_hard|Global-Hungarian", "WPS_hard|Urgency-Pair", "WPS_commit|Urgency-Commit",
        "WPS_attn|Local-Hungarian", "WPS_hard|Local-PI", "WPS_commit|Local-PI"]
SEEDS = range(4)
DROP = ("decision_ms_mean",)   # wall-clock, not a result


def main(only=None):
    refshim.install()
    # import-only dependencies of experiments/paper_eval.py:21,34 (legacy RL policies; not on this path)
    import types
    for name, attrs in (("tianshou", {}), ("tianshou.data", {"Batch": dict}),
                        ("TaskAllocation.RL_Policies", {}), ("TaskAllocation.RL_Policies.Tianshou_Policy", {"_get_model": None}),
                        ("RL_Policies", {}), ("RL_Policies.Tianshou_Policy", {"_get_model": None})):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            m.__path__ = []
            sys.modules[name] = m
    from experiments import wps_eval
    from TaskAllocation.Hybrid.AttentionCommit import UrgencyCommit
    from TaskAllocation.Hybrid.PairCostHybrid import UrgencyPair

    path = os.path.join(HERE, "wps_eval_scores.json")
    out = json.load(open(path)) if (only and os.path.exists(path)) else {}
    for key in ROWS:
        if only and key not in only:
            continue
        case, algo = key.split("|")
        kw = {}
        if algo == "Urgency-Pair":
            kw["urg_pair"] = UrgencyPair()       # one planner object for all seeds, as wps_eval.main does
        if algo == "Urgency-Commit":
            kw["urg_commit"] = UrgencyCommit()
        rows = []
        for seed in SEEDS:
            r = wps_eval.run_wps_episode(algo, case, seed, **kw)
            rows.append({k: float(v).hex() for k, v in r.items() if k not in DROP})
        out[key] = rows
        print(key, [float.fromhex(r["S_WPS"]) for r in rows])
    with open(path, "w") as f:
        json.dump(out, f, indent=0)


if __name__ == "__main__":
    main(set(sys.argv[1:]) or None)
