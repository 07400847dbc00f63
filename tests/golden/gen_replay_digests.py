"""Digests of replay documents (multi_uav_ta_gym_env_b200.replay) for the GPU box, which has no reference tree.

    python tests/golden/gen_replay_digests.py        # writes tests/golden/replay_digests.json

In the authoring container tests/test_dropin_facade.py::test_replay_documents_are_identical proves that the facade +
record_replay give the reference generator's document byte for byte.  This script stores the sha256 of the document of
a planner that needs no reference code (the facade's HungarianAllocator, Local / Coalition rule under the replay's own
cadence) computed on the CPU build of the kernel core; tests/test_gpu_facade.py recomputes it on the CUDA backend."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def replay_digest(scenario, seed, env_factory=None):
    from multi_uav_ta_gym_env_b200 import replay, wps_config
    from multi_uav_ta_gym_env_b200.env import HungarianAllocator, MultiUAVEnv

    cfg = wps_config(scenario)
    env = env_factory(cfg) if env_factory else MultiUAVEnv(cfg)
    _, info = env.reset(seed=seed)
    hung = HungarianAllocator(10**9, env.max_coord)

    def plan(e, events):
        open_tasks = [t for t in e.tasks if t.id != 0 and t.status != 2 and _residual(t) > 0]
        pairs = hung.allocate_tasks(e.get_live_agents(), open_tasks, time_step=e.time_steps, events=events, force=True,
                                    agent_known_ids=e.agent_visibility_map())
        return pairs, []

    doc = replay.record_replay(env, info, plan, cfg, scenario, seed)
    blob = json.dumps(doc, indent=2).encode()
    return {"sha256": hashlib.sha256(blob).hexdigest(), "frames": len(doc["frames"]), "events": len(doc["events"]),
            "s_wps": doc["final_metrics"]["s_wps"].hex()}


def _residual(t):
    if getattr(t, "kind", None) == "Escort" or float(getattr(t, "required_agents", 0) or 0) > 0:
        return max(float(getattr(t, "required_agents", 1) or 1) - len(t.allocationDetails), 0.0)
    return max(float(t.currentReqs[t.typeIdx] - t.allocatedReqs[t.typeIdx]), 0.0)


CASES = [("WPS_commit", 2), ("WPS_escort", 3)]

if __name__ == "__main__":
    from helpers import host_facade

    out = {f"{s}:{seed}": replay_digest(s, seed, host_facade) for s, seed in CASES}
    with open(os.path.join(HERE, "replay_digests.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out)
