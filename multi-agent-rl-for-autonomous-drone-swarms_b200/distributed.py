"""Multi-GPU use of the engine: env instances shard by index, nothing in `step` crosses GPUs.

One process per GPU (torchrun); rank r of W owns the contiguous block
`[r * E_total / W, (r + 1) * E_total / W)` of GLOBAL env indices.  Seeds are a function of the
global env index, so an env's trajectory does not depend on how many GPUs the batch is spread
over.  The only collective is the reduction of the 8-word episode-statistics block
(`torch.distributed.all_reduce`, NCCL over NVLink on GPUs, gloo in the CPU tests) -- issued at
report time, never on the step path.  SURVEY.md 8(e).
"""
from __future__ import annotations

from typing import Any

import numpy as np

from ._abi import STAT_NAMES


def shard_range(total_envs: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous, balanced, exhaustive partition of [0, total_envs) over the ranks."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    lo = (total_envs * rank) // world_size
    hi = (total_envs * (rank + 1)) // world_size
    return lo, hi


def global_env_seeds(base_seed: int, lo: int, hi: int) -> np.ndarray:
    """Seed of GLOBAL env index e is base_seed + e (the reference gives every env a ctor seed;
    scripts/evaluate_protocol.py:425-431 uses seed + episode)."""
    return (np.uint64(base_seed) + np.arange(lo, hi, dtype=np.uint64)).astype(np.uint64)


def stats_to_vector(stats: dict[str, float]) -> np.ndarray:
    return np.asarray([float(stats[n]) for n in STAT_NAMES], dtype=np.float64)


def vector_to_stats(vec) -> dict[str, float]:
    out = {n: float(v) for n, v in zip(STAT_NAMES, vec)}
    for n in STAT_NAMES:
        if n != "return_sum":
            out[n] = int(round(out[n]))
    return out


def all_reduce_stats(stats: dict[str, float], device=None, group=None) -> dict[str, float]:
    """SUM-reduce the per-rank episode statistics over the process group (the path's one collective)."""
    import torch
    import torch.distributed as dist

    vec = torch.from_numpy(stats_to_vector(stats))
    if device is not None:
        vec = vec.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return vector_to_stats(vec.cpu().numpy())


def summarize(stats: dict[str, float]) -> dict[str, float]:
    """RLlib-style aggregates (reference utils/ray_metrics.py:42-74 keys)."""
    eps = max(int(stats["episodes"]), 1)
    return {
        "episodes_this_iter": int(stats["episodes"]),
        "episode_reward_mean": stats["return_sum"] / eps,
        "episode_len_mean": stats["length_sum"] / eps,
        "success_rate": stats["success"] / eps,
        "collision_rate": stats["collision"] / eps,
        "timeout_rate": stats["timeout"] / eps,
        "agent_steps": int(stats["agent_steps"]),
        "env_steps": int(stats["env_steps"]),
    }


class ShardedSwarm:
    """This rank's shard of a `total_envs`-instance batch (one process per GPU)."""

    def __init__(self, total_envs: int, config: dict[str, Any] | None = None, kind: str = "swarm",
                 base_seed: int = 0, rank: int | None = None, world_size: int | None = None, device=None, **engine_kw):
        import torch
        import torch.distributed as dist

        from .engine import SwarmEngine

        if world_size is None:
            world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        self.rank, self.world_size, self.total_envs = rank, world_size, int(total_envs)
        self.lo, self.hi = shard_range(self.total_envs, rank, world_size)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        # the domain-randomisation streams are keyed by the GLOBAL env index (same invariance as the seeds)
        engine_kw.setdefault("env_index_base", self.lo)
        self.engine = SwarmEngine(self.hi - self.lo, config, kind=kind, device=device, **engine_kw)
        self.engine.seed(global_env_seeds(base_seed, self.lo, self.hi))

    def reset(self):
        return self.engine.reset()

    def step(self, actions, auto_reset: bool = True):
        return self.engine.step(actions, auto_reset=auto_reset)

    def global_stats(self) -> dict[str, float]:
        return all_reduce_stats(self.engine.stats(), device=self.engine.device)
