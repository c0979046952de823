#!/usr/bin/env python
"""Step time of the N <= 32 rotation-pass kernel when drones are PARKED (they reached the goal while the episode goes
on): goal-seeking actions instead of the bench's U(-1, 1).  Prints, per phase, the share of env instances with at least
one parked drone and the mean device time of `step()` (CUDA events around the step only; the action is computed before)."""
import argparse, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swarm_b200

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=65536)
ap.add_argument("--drones", type=int, default=32)
ap.add_argument("--world", type=float, default=44.0)
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--gain", type=float, default=1.0)
args = ap.parse_args()
cfg = {"num_drones": args.drones, "num_obstacles": 8, "world_size": args.world}
eng = swarm_b200.SwarmEngine(args.envs, cfg, device="cuda:0")
eng.seed(np.arange(args.envs, dtype=np.uint64))
eng.reset()
out = []
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
acc_ms, acc_parked, n = 0.0, 0.0, 0
for t in range(args.steps):
    d = eng.goal[:, None, :] - eng.positions
    act = (d * args.gain).clamp_(-1.0, 1.0).contiguous()
    alive = eng.alive
    parked_envs = ((alive == 0).any(dim=1) & (alive != 0).any(dim=1)).float().mean().item()
    torch.cuda.synchronize()
    ev0.record()
    eng.step(act, auto_reset=True)
    ev1.record()
    torch.cuda.synchronize()
    acc_ms += ev0.elapsed_time(ev1); acc_parked += parked_envs; n += 1
    if (t + 1) % 50 == 0:
        out.append({"steps": f"{t - 48}-{t + 1}", "ms_per_step": acc_ms / n, "envs_with_parked_drones": acc_parked / n})
        acc_ms, acc_parked, n = 0.0, 0.0, 0
print(json.dumps({"config": cfg, "envs": args.envs, "phases": out, "stats": eng.stats()}))
