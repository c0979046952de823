"""Batched protocol evaluation: one episode per env instance, metrics accumulated on the device.

Mirrors the reference's per-episode loops and aggregation
    scripts/evaluate_protocol.py:237-312  _run_single_episode_multi_agent
    scripts/evaluate_protocol.py:193-234  _run_single_episode_single_agent
    scripts/evaluate_protocol.py:315-331  _aggregate
for E episodes at once (the reference steps one env in Python per episode; its full protocol of
2000 episodes "appears stuck", docs/TROUBLESHOOTING.md:57-65).  `faithful=True` reproduces the
reference's scoring exactly, including its quirk: infos only ever exist for
drones that neither reached the goal nor collided, so `collision` is never seen and the
episode-ending step (empty obs dict) counts as "all reached".  `faithful=False` scores with the
engine's true per-drone flags instead.  The env arithmetic itself is the CUDA step; this module
is host-side bookkeeping over its output tensors.
"""
from __future__ import annotations

import math
from typing import Callable

import torch


def _formation_error(pos: torch.Tensor, mask: torch.Tensor, spacing: float, chunk_bytes: int = 64 << 20) -> torch.Tensor:
    """evaluate_protocol.py:103-116 over the drones selected by `mask`: [E,N,3], [E,N] -> [E].
    Chunked over the env axis so that the [e, N, N, 3] difference temporary stays below `chunk_bytes` (at BASELINE
    config 5 -- 8192 envs x 128 drones -- the unchunked temporary would be 1.6 GB per step)."""
    E, N = pos.shape[0], pos.shape[1]
    out = torch.zeros(E, dtype=torch.float64, device=pos.device)
    eye = torch.eye(N, dtype=torch.bool, device=pos.device)
    step = max(1, int(chunk_bytes // max(N * N * 3 * 4, 1)))
    for lo in range(0, E, step):
        p, m = pos[lo:lo + step], mask[lo:lo + step]
        d = torch.linalg.vector_norm(p[:, :, None, :] - p[:, None, :, :], dim=-1).double()
        pair = m[:, :, None] & m[:, None, :] & ~eye
        cnt = pair.sum(-1)
        per = (torch.abs(d - spacing) * pair).sum(-1) / cnt.clamp(min=1)
        n = m.sum(-1)
        fe = (per * m).sum(-1) / n.clamp(min=1)
        out[lo:lo + step] = torch.where(n > 1, fe, torch.zeros_like(fe))
    return out


@torch.no_grad()
def evaluate_batched(engine, policy: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], faithful: bool = True,
                     step_limit: int | None = None) -> dict:
    """Run one episode in every env instance of `engine` (seed it first) and score it.

    policy(obs [E,N,D], obs_valid [E,N] bool) -> actions [E,N,3] (CUDA tensors)."""
    E, N = engine.E, engine.N
    dev = engine.device
    swarm = engine.kind == "swarm"
    spacing = float(engine.cfg.desired_spacing)
    engine.reset()
    obs = engine.obs
    start = obs[..., 0:3].clone()
    goal = start + obs[..., 6:9]
    last = start.clone()
    traveled = torch.zeros((E, N), dtype=torch.float64, device=dev)
    ep_reward = torch.zeros(E, dtype=torch.float64, device=dev)
    fe_sum = torch.zeros(E, dtype=torch.float64, device=dev)
    steps = torch.zeros(E, dtype=torch.int64, device=dev)
    reached_step = torch.full((E,), -1, dtype=torch.int64, device=dev)
    any_col = torch.zeros(E, dtype=torch.bool, device=dev)
    done = torch.zeros(E, dtype=torch.bool, device=dev)
    valid = engine.obs_valid.bool().clone()
    limit = step_limit if step_limit is not None else int(engine.cfg.max_steps) + 1
    for _ in range(limit):
        if bool(done.all()):
            break
        live = ~done
        in_step = engine.alive.clone() if swarm else torch.ones((E, N), dtype=torch.bool, device=dev)
        actions = policy(engine.obs, valid)
        engine.step(actions, auto_reset=False)
        steps += live
        rew = engine.reward64 if engine.reward64 is not None else engine.reward.double()
        n_in = in_step.sum(-1)
        mean_rew = (rew * in_step).sum(-1) / n_in.clamp(min=1)          # np.mean(list(rewards.values())) :272
        ep_reward += torch.where(live & (n_in > 0), mean_rew, torch.zeros_like(mean_rew))
        valid = engine.obs_valid.bool() & live[:, None]
        pos = engine.obs[..., 0:3]
        seg = torch.linalg.vector_norm(last - pos, dim=-1).double()
        traveled += torch.where(valid, seg, torch.zeros_like(seg))        # :277-279
        last = torch.where(valid[..., None], pos, last)
        if swarm:
            fe_sum += torch.where(live, _formation_error(pos, valid, spacing), torch.zeros_like(fe_sum))  # :289
        reached = engine.reached.bool()
        collided = engine.collision.bool()
        if faithful:
            # infos exist only for emitted drones and always say reached_goal = collision = False (:281-287)
            if swarm:
                all_flag = ~valid.any(-1)
            else:
                all_flag = reached[:, 0]
                any_col |= live & collided[:, 0]
        else:
            parked_or_reached = reached | ~in_step
            all_flag = parked_or_reached.all(-1)
            any_col |= live & (collided & in_step).any(-1)
        reached_step = torch.where(live & all_flag & (reached_step < 0), steps, reached_step)   # :290-291
        done |= (engine.all_terminated.bool() | engine.all_truncated.bool())
    straight = torch.linalg.vector_norm(start - goal, dim=-1).double()
    pe = torch.where(traveled > 1e-8, straight / traveled.clamp(min=1e-30), torch.zeros_like(traveled)).mean(-1)
    success = (~any_col) & (reached_step >= 0)
    fe = fe_sum / steps.clamp(min=1) if swarm else torch.zeros_like(fe_sum)
    ttg = torch.where(reached_step >= 0, reached_step.double(), torch.full_like(ep_reward, math.nan))
    per_episode = dict(success=success, collision_free=~any_col, time_to_goal=ttg, formation_error=fe,
                       path_efficiency=pe, episode_reward=ep_reward, length=steps, finished=done)
    has_ttg = reached_step >= 0
    agg = {                                                              # _aggregate :315-331
        "success_rate": float(success.double().mean()),
        "collision_free_rate": float((~any_col).double().mean()),
        "mean_time_to_goal": float(ttg[has_ttg].mean()) if bool(has_ttg.any()) else math.nan,
        "formation_error": float(fe.mean()),
        "path_efficiency": float(pe.mean()),
        "episode_reward_mean": float(ep_reward.mean()),
        "episode_reward_std": float(ep_reward.std(unbiased=False)) if E > 1 else 0.0,
        "episodes": E,
    }
    return {"aggregate": agg, "per_episode": per_episode}
