"""How much the reference's own Local-CBBA-Replan result moves with the interpreter's string hash, next to the drop-in
classes (which reproduce PYTHONHASHSEED=0).  Authoring container only (needs /root/reference).

    python tools/cbba_hash_spread.py [n_seeds]      # prints a markdown table (profiles/r02_cbba_hash_spread.md)

Each hash seed is one child interpreter running the UNMODIFIED experiments/wps_eval.run_wps_episode on seeds 0..n-1; the
last row runs the same driver with MultiUAVEnv / HungarianAllocator / CBBAReplan swapped for the facade classes over the
CPU build of the kernel sources (tests/helpers.host_facade)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests"); sys.path.insert(0, %(root)r + "/tests/golden")
import test_dropin_facade as T
M = T._driver_module(%(module)r)
if %(facade)r:
    from helpers import host_facade
    from multi_uav_ta_gym_env_b200 import env as E
    M.MultiUAVEnv, M.HungarianAllocator, M.CBBAReplan = host_facade, E.HungarianAllocator, E.CBBAReplan
out = [getattr(M, %(fn)r)(%(algo)r, %(case)r, s) for s in range(%(n)d)]
print(json.dumps([{k: float(v) for k, v in r.items() if not isinstance(v, str)} for r in out]))
'''


def run(hashseed, facade, module, fn, algo, case, n):
    env = dict(os.environ, PYTHONHASHSEED=str(hashseed))
    code = CHILD % dict(root=ROOT, facade=facade, module=module, fn=fn, algo=algo, case=case, n=n)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def mean(xs):
    return sum(xs) / len(xs)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    for module, fn, algo, case, score in (("wps_eval", "run_wps_episode", "Local-CBBA-Replan", "WPS_hard", "S_WPS"),
                                          ("escort_eval", "run_escort_episode", "Local-CBBA-Coalition", "WPS_escort", "S_ESC")):
        print(f"\n### {case}, {algo}, seeds 0..{n - 1}\n")
        print(f"| interpreter | mean {score} | mean n_on_time | mean n_missed_windows | episodes equal to PYTHONHASHSEED=0 |")
        print("|---|---|---|---|---|")
        base = None
        for label, hs, facade in [("reference, PYTHONHASHSEED=0", 0, False), ("reference, PYTHONHASHSEED=1", 1, False),
                                  ("reference, PYTHONHASHSEED=2", 2, False), ("reference, PYTHONHASHSEED=3", 3, False),
                                  ("drop-in classes (any hash seed; run under 3)", 3, True)]:
            rows = run(hs, facade, module, fn, algo, case, n)
            if base is None:
                base = rows
            same = sum(1 for a, b in zip(rows, base) if all(a[k] == b[k] for k in a if "_ms" not in k))
            print(f"| {label} | {mean([r[score] for r in rows]):.3f} | {mean([r.get('n_on_time', 0.0) for r in rows]):.2f} | "
                  f"{mean([r.get('n_missed_windows', 0.0) for r in rows]):.2f} | {same} / {n} |")


if __name__ == "__main__":
    main()
