"""Replay a tests/golden/*.npz fixture through a backend and compare bit-for-bit.

A backend is anything with the batched-contract attributes (numpy arrays, leading env axis):
positions, velocities, goal, obstacles, step_count, obs, reward, dist, terminated, truncated,
reached, collision, obs_valid, all_terminated, all_truncated, global_state and the methods
seed(seeds), reset(), step(actions, auto_reset=True).  The C oracle (oracle/swarm_oracle.py)
and the CUDA engine wrapper used by the gpu tests both satisfy it.
"""
from __future__ import annotations

import glob
import hashlib
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    g["meta"] = json.loads(str(g["meta"]))
    return g


def hash_rows(x):
    x = np.ascontiguousarray(x)
    out = np.zeros(x.shape[0], np.uint64)
    for e in range(x.shape[0]):
        out[e] = int.from_bytes(hashlib.blake2b(x[e].tobytes(), digest_size=8).digest(), "little")
    return out


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({4: np.uint32, 8: np.uint64, 1: np.uint8}[a.dtype.itemsize])


def assert_biteq(name, got, want, t):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (name, t, got.shape, want.shape)
    if got.dtype != want.dtype:
        got = got.astype(want.dtype)
    if not np.array_equal(bits(got), bits(want)):
        bad = np.argwhere(bits(got) != bits(want))
        i = tuple(bad[0])
        raise AssertionError(f"{name} differs at step {t}, index {i}: got {got[i]!r} want {want[i]!r} "
                             f"({len(bad)} mismatching elements)")


def _block_ok_up_to_ties(block, cand_rel, cand_dist, k_slots):
    """`block` = k_slots x (rx, ry, rz, d).  Valid iff its distances are the k smallest candidate
    distances in order and every slot is a distinct real candidate carrying exactly that distance
    (i.e. it differs from the reference only in how np.argsort broke an exact tie -- trap T5)."""
    n = len(cand_dist)
    k = min(k_slots, n)
    blk = block.reshape(k_slots, 4)
    want_d = np.sort(cand_dist, kind="stable")[:k]
    if not np.array_equal(bits(blk[:k, 3]), bits(want_d)):
        return False
    if np.any(bits(blk[k:]) != 0):
        return False
    used = set()
    for q in range(k):
        hit = [j for j in range(n) if j not in used and np.array_equal(bits(cand_rel[j]), bits(blk[q, :3]))
               and bits(cand_dist[j:j + 1])[0] == bits(blk[q, 3:4])[0]]
        if not hit:
            return False
        used.add(hit[0])
    return True


def obs_row_ok_up_to_ties(kind, cfg, pos, vel, goal, obst, i, row):
    """Check one obs row against the state using numpy's own norms (the reference's arithmetic on
    this platform), accepting any tie order.  drone_swarm_env.py:226-291."""
    K = int(cfg.get("neighbor_k", 3)) if kind == "swarm" else 0
    S = int(cfg.get("sensed_obstacles", 4))
    own = np.concatenate([pos[i], vel[i], goal - pos[i]]).astype(np.float32)
    if not np.array_equal(bits(row[:9]), bits(own)):
        return False
    N, M = pos.shape[0], obst.shape[0]
    if K > 0:
        rel = np.stack([pos[j] - pos[i] for j in range(N) if j != i]) if N > 1 else np.zeros((0, 3), np.float32)
        d = np.array([np.linalg.norm(v) for v in rel], np.float32)
        if not _block_ok_up_to_ties(row[9:9 + 4 * K], rel, d, K):
            return False
    rel = (obst - pos[i]).astype(np.float32)
    d = np.linalg.norm(rel, axis=1).astype(np.float32) if M else np.zeros(0, np.float32)
    return _block_ok_up_to_ties(row[9 + 4 * K:], rel, d, S)


def replay_and_compare(backend, g, reward_dtype=np.float64, steps=None, get=lambda x: np.asarray(x)):
    """Returns a dict of event counts.  `reward_dtype`: compare rewards after rounding the
    golden float64 reward to this dtype (float32 for the engine's f32 reward output)."""
    T = g["actions"].shape[0] if steps is None else min(steps, g["actions"].shape[0])
    head = g["obs"].shape[0]
    backend.seed(g["seeds"])
    backend.reset()
    assert_biteq("reset.positions", get(backend.positions), g["reset0_pos"], -1)
    assert_biteq("reset.velocities", get(backend.velocities), g["reset0_vel"], -1)
    assert_biteq("reset.goal", get(backend.goal), g["reset0_goal"], -1)
    assert_biteq("reset.obstacles", get(backend.obstacles), g["reset0_obst"], -1)
    assert_biteq("reset.obs", get(backend.obs), g["reset0_obs"], -1)
    assert_biteq("reset.dist", get(backend.dist), g["reset0_dist"], -1)
    assert_biteq("reset.global_state", get(backend.global_state), g["reset0_gs"], -1)
    assert np.all(get(backend.obs_valid) == 1)
    stats = dict(steps=T, tie_steps=0, episodes=0, agent_steps=0, reached=0, collided=0, truncated_eps=0)
    for t in range(T):
        backend.step(g["actions"][t], auto_reset=True)
        ins = g["in_step"][t].astype(bool)
        valid = g["obs_valid"][t].astype(bool)
        done = (g["all_term"][t] | g["all_trunc"][t]).astype(bool)
        assert_biteq("all_terminated", get(backend.all_terminated), g["all_term"][t], t)
        assert_biteq("all_truncated", get(backend.all_truncated), g["all_trunc"][t], t)
        assert_biteq("positions", get(backend.positions), g["pos"][t], t)
        assert_biteq("velocities", get(backend.velocities), g["vel"][t], t)
        assert_biteq("goal", get(backend.goal), g["goal"][t], t)
        assert_biteq("obstacles", get(backend.obstacles), g["obst"][t], t)
        assert_biteq("step_count", get(backend.step_count), g["step_count"][t], t)
        want_r = np.where(ins, g["reward"][t], 0.0).astype(reward_dtype)
        assert_biteq("reward", get(backend.reward).astype(reward_dtype), want_r, t)
        for name, key in (("terminated", "terminated"), ("truncated", "truncated"), ("reached", "reached"),
                          ("collision", "collision")):
            assert_biteq(name, get(getattr(backend, name)) * ins, g[key][t] * ins, t)
            assert not np.any(get(getattr(backend, name))[~ins]), (name, t)
        assert_biteq("obs_valid", get(backend.obs_valid), g["obs_valid"][t], t)
        obs = np.where(valid[..., None], get(backend.obs), np.float32(0.0))
        hmiss = np.nonzero(hash_rows(obs) != g["obs_hash"][t])[0]
        for e in hmiss:
            # only acceptable cause: np.argsort broke an exact distance tie differently (T5)
            for i in np.nonzero(valid[e])[0]:
                if t < head and np.array_equal(bits(obs[e, i]), bits(g["obs"][t][e, i])):
                    continue
                ok = obs_row_ok_up_to_ties(g["meta"]["kind"], g["meta"]["config"], g["pos"][t][e], g["vel"][t][e],
                                           g["goal"][t][e], g["obst"][t][e], i, obs[e, i])
                assert ok, f"obs row differs beyond tie order at step {t}, env {e}, drone {i}"
            stats["tie_steps"] += 1
        if t < head and len(hmiss) == 0:
            assert_biteq("obs", obs, g["obs"][t], t)
        dmask = ins | done[:, None]
        assert_biteq("dist", np.where(dmask, get(backend.dist), np.float32(0)), np.where(dmask, g["dist"][t], np.float32(0)), t)
        assert_biteq("gs_hash", hash_rows(get(backend.global_state)), g["gs_hash"][t], t)
        if t < head:
            assert_biteq("global_state", get(backend.global_state), g["gs"][t], t)
        stats["episodes"] += int(done.sum())
        stats["truncated_eps"] += int(g["all_trunc"][t].sum())
        stats["agent_steps"] += int(ins.sum())
        stats["reached"] += int((g["reached"][t] * ins).sum())
        stats["collided"] += int((g["collision"][t] * ins).sum())
    return stats
