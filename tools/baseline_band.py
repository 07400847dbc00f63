"""Results parity against BASELINE.md section 2 (reference measured in the survey container, seeds 0..99)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_uav_ta_gym_env_b200 import evaluate  # noqa: E402

out = {}
for case, algo, key in (("WPS_hard", "Local-Hungarian", "S_WPS"), ("WPS_escort", "Coalition-Hungarian", "S_ESC"),
                        ("WPS_commit", "Local-Hungarian", "S_WPS"), ("WPS_easy", "Local-Hungarian", "S_WPS"),
                        ("WPS_attn", "Local-Hungarian", "S_WPS"), ("WPS_hard", "Global-Hungarian", "S_WPS")):
    sc = evaluate.run_episodes(case, algo, 100)
    v = np.array([s[key] for s in sc])
    out[f"{case} {algo}"] = {"score": key, "mean": float(v.mean()), "sd": float(v.std()),
                             "on_time_rate": float(np.mean([s["on_time_rate"] for s in sc])),
                             "missed": float(np.mean([s["n_missed_windows"] for s in sc]))}
print(json.dumps(out, indent=1))
