#!/usr/bin/env python
"""Saturation sweep: device-resident agent-steps/s vs env instances per GPU, to show
where the step leaves the launch-/L2-bound regime.  Writes one JSON object per line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import swarm_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--drones", type=int, default=32)
ap.add_argument("--obstacles", type=int, default=8)
ap.add_argument("--world", type=float, default=20.0)
ap.add_argument("--min-log2", type=int, default=10)
ap.add_argument("--max-log2", type=int, default=20)
ap.add_argument("--steps", type=int, default=200)
args = ap.parse_args()
peak = 6553.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
for lg in range(args.min_log2, args.max_log2 + 1):
    E = 1 << lg
    cfg = {"num_drones": args.drones, "num_obstacles": args.obstacles, "world_size": args.world}
    eng = swarm_b200.SwarmEngine(E, cfg, device="cuda:0")
    eng.seed(np.arange(E, dtype=np.uint64))
    eng.reset()
    acts = [torch.rand((E, eng.N, 3), device="cuda:0") * 2 - 1 for _ in range(2)]
    for w in range(10):
        eng.step(acts[w % 2])
    eng.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for k in range(args.steps):
        eng.step(acts[k % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    B = eng.algorithmic_bytes_per_agent_step()
    rate = eng.stats()["agent_steps"] / (ms * 1e-3 * args.steps)
    print(json.dumps({"envs": E, "drones": eng.N, "obstacles": eng.M, "world": args.world, "us_per_step": ms * 1e3,
                      "agent_steps_per_sec": rate, "working_set_MB": B * E * eng.N / 1e6,
                      "hbm_frac": B * E * eng.N / (ms * 1e-3) / 1e9 / peak}), flush=True)
    eng.close()
    del eng, acts
    torch.cuda.empty_cache()
