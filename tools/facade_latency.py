#!/usr/bin/env python
"""Per-call latency of the drop-in facade envs (E = 1, through swarm_step_host) beside the reference's own Python
env on the same host (GPU box, via gpurun): what evaluate_protocol.py:427-431 or a one-env RLlib worker feels."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402

import swarm_b200  # noqa: E402


def bench(env, n_steps, single=False):
    rng = np.random.default_rng(0)
    env.reset()
    t_reset, t_step, resets, steps = 0.0, 0.0, 0, 0
    obs, _ = env.reset()
    for _ in range(n_steps):
        if single:
            a = rng.uniform(-1, 1, 3).astype(np.float32)
        else:
            a = {k: rng.uniform(-1, 1, 3).astype(np.float32) for k in env.agents}
        t0 = time.perf_counter()
        out = env.step(a)
        t_step += time.perf_counter() - t0
        steps += 1
        done = (out[2] or out[3]) if single else (out[2]["__all__"] or out[3]["__all__"])
        if done:
            t0 = time.perf_counter()
            env.reset()
            t_reset += time.perf_counter() - t0
            resets += 1
    return t_step / steps * 1e6, (t_reset / resets * 1e6) if resets else float("nan")


def main():
    rows = []
    try:
        import ref_runner
        Single, Swarm = ref_runner._import_envs()
    except Exception as exc:  # noqa: BLE001
        Single = Swarm = None
        print("reference envs unavailable:", exc, file=sys.stderr)
    for n in (3, 8, 32):
        cfg = {"num_drones": n, "num_obstacles": 8, "seed": 1, "world_size": 20.0 + 2 * n}
        ours = swarm_b200.DroneSwarmEnv(cfg)
        bench(ours, 200)
        s, r = bench(ours, 2000)
        row = {"env": f"DroneSwarmEnv N={n}", "facade_step_us": round(s, 1), "facade_reset_us": round(r, 1)}
        ours.close()
        if Swarm is not None:
            s2, r2 = bench(Swarm(cfg), 600 if n < 32 else 200)
            row.update(reference_step_us=round(s2, 1), reference_reset_us=round(r2, 1))
        rows.append(row)
    ours = swarm_b200.SingleDroneEnv({"seed": 1})
    bench(ours, 200, single=True)
    s, r = bench(ours, 3000, single=True)
    row = {"env": "SingleDroneEnv", "facade_step_us": round(s, 1), "facade_reset_us": round(r, 1)}
    if Single is not None:
        s2, r2 = bench(Single({"seed": 1}), 3000, single=True)
        row.update(reference_step_us=round(s2, 1), reference_reset_us=round(r2, 1))
    rows.append(row)
    print(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
