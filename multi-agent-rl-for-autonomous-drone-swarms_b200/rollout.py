"""Device-resident rollouts: policy -> env step -> sample-batch columns without leaving the GPU.

What an RLlib rollout worker does around the reference env -- per env, in Python: `compute_actions`, `env.step`,
dict marshaling, and `GlobalStateCallback.on_postprocess_trajectory` stacking `info["global_state"]` per step
(reference src/swarm_marl/training/callbacks.py:51-57, training/models.py:104-152) -- is here one engine step per
time step over E env instances and plain tensor writes into preallocated [T, E, ...] columns.  No host
synchronisation happens inside the loop; the env arithmetic is the CUDA step (SwarmEngine), this module is host-side
orchestration only.
"""
from __future__ import annotations

from typing import Callable

import torch


@torch.no_grad()
def collect_rollout(engine, policy: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], steps: int,
                    value_fn: Callable[[torch.Tensor], torch.Tensor] | None = None, out: dict | None = None) -> dict:
    """Roll `steps` time steps in every env instance of `engine` (seeded and reset by the caller), auto-reset on.

    policy(obs [E,N,D], obs_valid [E,N] bool) -> actions [E,N,3]; value_fn(global_state [E,6N+3]) -> [E] (optional:
    the CTDE critic).  Returns device tensors (the SampleBatch columns a PPO learner consumes):
      obs [T,E,N,D], actions [T,E,N,3], rewards [T,E,N], terminateds / truncateds [T,E,N] bool, dones [T,E] bool,
      agent_mask [T,E,N] bool (drone was in env.agents when the action was applied: the reference's dict keys),
      global_state [T,E,6N+3] (the critic's column), values [T,E] (if value_fn), next_obs [E,N,D], next_global_state."""
    E, N, D, R = engine.E, engine.N, engine.D, engine.R
    dev = engine.device
    T = int(steps)
    if out is None:
        out = dict(obs=torch.empty((T, E, N, D), device=dev), actions=torch.empty((T, E, N, 3), device=dev),
                   rewards=torch.empty((T, E, N), device=dev),
                   terminateds=torch.empty((T, E, N), dtype=torch.bool, device=dev),
                   truncateds=torch.empty((T, E, N), dtype=torch.bool, device=dev),
                   dones=torch.empty((T, E), dtype=torch.bool, device=dev),
                   agent_mask=torch.empty((T, E, N), dtype=torch.bool, device=dev),
                   global_state=torch.empty((T, E, R), device=dev))
        if value_fn is not None:
            out["values"] = torch.empty((T, E), device=dev)
    for t in range(T):
        obs, valid = engine.obs, engine.obs_valid.bool()
        out["obs"][t].copy_(obs)
        out["global_state"][t].copy_(engine.global_state)
        out["agent_mask"][t].copy_(engine.alive)
        if value_fn is not None:
            out["values"][t].copy_(value_fn(engine.global_state).reshape(E))
        act = policy(obs, valid)
        out["actions"][t].copy_(act)
        engine.step(act, auto_reset=True)
        out["rewards"][t].copy_(engine.reward)
        out["terminateds"][t].copy_(engine.terminated)
        out["truncateds"][t].copy_(engine.truncated)
        torch.logical_or(engine.all_terminated.bool(), engine.all_truncated.bool(), out=out["dones"][t])
    out["next_obs"] = engine.obs.clone()
    out["next_global_state"] = engine.global_state.clone()
    return out
