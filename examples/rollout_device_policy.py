#!/usr/bin/env python
"""Device-resident rollout: a shared-policy MLP (+ centralized critic on `global_state`) acting on
E batched swarm envs without leaving the GPU -- what an RLlib rollout worker + `GlobalStateCallback`
(reference training/callbacks.py:51-57, training/models.py:104-152) do per env in Python, here as
one engine step and two matmuls per time step.

    python examples/rollout_device_policy.py --envs 16384 --drones 16 --steps 200
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import swarm_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=16384)
ap.add_argument("--drones", type=int, default=16)
ap.add_argument("--obstacles", type=int, default=8)
ap.add_argument("--steps", type=int, default=200)
args = ap.parse_args()

dev = torch.device("cuda", 0)
eng = swarm_b200.SwarmEngine(args.envs, {"num_drones": args.drones, "num_obstacles": args.obstacles}, device=dev)
eng.seed(np.arange(args.envs, dtype=np.uint64))
obs = eng.reset()                                   # [E, N, 37] float32, on the device
actor = torch.nn.Sequential(torch.nn.Linear(eng.D, 128), torch.nn.Tanh(), torch.nn.Linear(128, 3), torch.nn.Tanh()).to(dev)
critic = torch.nn.Sequential(torch.nn.Linear(eng.R, 256), torch.nn.Tanh(), torch.nn.Linear(256, 1)).to(dev)
# sample-batch columns a PPO learner would consume (reference: SampleBatch incl. the `global_state` column)
T, E, N = args.steps, args.envs, args.drones
buf = dict(obs=torch.empty((T, E, N, eng.D), device=dev), actions=torch.empty((T, E, N, 3), device=dev),
           rewards=torch.empty((T, E, N), device=dev), dones=torch.empty((T, E), dtype=torch.bool, device=dev),
           values=torch.empty((T, E), device=dev), valid=torch.empty((T, E, N), dtype=torch.bool, device=dev))
with torch.no_grad():                                # warm-up: cuBLAS handles, allocator, first launches
    for _ in range(5):
        critic(eng.global_state)
        obs, *_ = eng.step(actor(obs))
eng.reset_stats()
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.no_grad():
    for t in range(T):
        buf["obs"][t] = obs
        buf["valid"][t] = eng.obs_valid.bool()
        act = actor(obs)                             # shared policy: every drone, every env, one matmul
        buf["values"][t] = critic(eng.global_state).squeeze(-1)   # CTDE critic on [positions | velocities | goal]
        obs, rew, term, trunc = eng.step(act)        # auto-reset: obs is already the next episode's first obs
        buf["actions"][t], buf["rewards"][t] = act, rew
        buf["dones"][t] = (eng.all_terminated | eng.all_truncated).bool()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
st = eng.stats()
print(f"{T} steps x {E} envs x {N} drones in {dt:.2f} s = {st['agent_steps'] / dt:.3g} agent-steps/s incl. policy + critic; "
      f"episodes {st['episodes']} (success {st['success']}, collision {st['collision']}, timeout {st['timeout']}), "
      f"mean return {st['return_sum'] / max(st['episodes'], 1):.1f}")
