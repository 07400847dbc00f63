"""Evidence for DESIGN.md section 1a row f3: the reference's Local-CBBA-Replan episode (experiments/wps_eval.py:134-146)
depends on PYTHONHASHSEED, because CBBA.allocate_tasks shuffles `list(remaining)` of a set of strings
(TaskAllocation/MarketBased/CBBA.py:116,128).  Authoring container only.

    python tests/golden/cbba_hashseed_check.py         # runs itself under PYTHONHASHSEED = 1, 2, 3 and prints the scores
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, HERE)
    import refshim

    refshim.install()
    import types
    for name, attrs in (("tianshou", {}), ("tianshou.data", {"Batch": dict}), ("TaskAllocation.RL_Policies", {}),
                        ("TaskAllocation.RL_Policies.Tianshou_Policy", {"_get_model": None})):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            m.__path__ = []
            sys.modules[name] = m
    from experiments import wps_eval

    r = [wps_eval.run_wps_episode("Local-CBBA-Replan", "WPS_hard", seed) for seed in range(3)]
    print([(x["S_WPS"], x["total_distance"]) for x in r])
else:
    for hs in ("1", "2", "3"):
        env = dict(os.environ, PYTHONHASHSEED=hs)
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True)
        print("PYTHONHASHSEED", hs, out.stdout.strip() or out.stderr[-400:])
