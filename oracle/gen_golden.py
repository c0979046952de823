"""Generate tests/golden/*.npz from the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Run in the build container (needs /root/reference):   python oracle/gen_golden.py
The fixtures pin oracle/swarm_oracle.c (tests/test_oracle_golden.py, CPU) and the CUDA
path (tests/test_gpu_parity.py, B200).  Platform of record: numpy 2.3.5, OpenBLAS 0.3.30
x86-64 (sdot double-accumulate tail), Python 3.12.

Every fixture is a batch of E independent reference envs (ctor seed = seeds[e]) rolled T
steps with auto-reset: when an env's episode ends (`__all__` for the swarm env,
`terminated or truncated` for the single env) `env.reset()` is called with no seed, which
continues the env's PCG64 stream exactly as RLlib's sampler does.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference_envs  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


OBS_HEAD = 120


def hash_rows(x):
    """blake2b-64 digest of each [t, e] slab of a float32 array (bit-exact comparison key)."""
    import hashlib
    x = np.ascontiguousarray(x)
    T, E = x.shape[:2]
    out = np.zeros((T, E), np.uint64)
    for t in range(T):
        for e in range(E):
            out[t, e] = int.from_bytes(hashlib.blake2b(x[t, e].tobytes(), digest_size=8).digest(), "little")
    return out


def _uniform_actions(lo, hi):
    def fn(rng, env, kind, N):
        return rng.uniform(lo, hi, size=(N, 3)).astype(np.float32)
    return fn


def _goal_seek(noise):
    def fn(rng, env, kind, N):
        pos = env.positions if kind == "swarm" else env.position[None, :]
        d = env.goal[None, :].astype(np.float64) - pos.astype(np.float64)
        d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-6)
        return (d + rng.normal(0.0, noise, size=(N, 3))).astype(np.float32)
    return fn


def roll(kind, cfg, seeds, T, action_fn, action_seed=1000):
    Single, Swarm = load_reference_envs()
    E = len(seeds)
    N = int(cfg.get("num_drones", 3)) if kind == "swarm" else 1
    envs = []
    for s in seeds:
        c = dict(cfg)
        c["seed"] = int(s)
        envs.append(Swarm(c) if kind == "swarm" else Single(c))
    M = envs[0].cfg.num_obstacles
    D = envs[0].observation_space.shape[0]
    G = 6 * N + 3
    arngs = [np.random.default_rng(action_seed + e) for e in range(E)]
    ids = [f"drone_{i}" for i in range(N)]

    def state(env):
        if kind == "swarm":
            return env.positions.copy(), env.velocities.copy(), env.goal.copy(), env.obstacles.copy()
        return env.position[None].copy(), env.velocity[None].copy(), env.goal.copy(), env.obstacles.copy()

    def pack_reset(env, obs, info):
        o = np.zeros((N, D), np.float32)
        d = np.zeros(N, np.float32)
        g = np.zeros(G, np.float32)
        if kind == "swarm":
            for i, a in enumerate(ids):
                o[i] = obs[a]
                d[i] = info[a]["distance_to_goal"]
                assert float(d[i]) == info[a]["distance_to_goal"]
            g[:] = info[ids[0]]["global_state"]
        else:
            o[0] = obs
            d[0] = info["distance_to_goal"]
            p, v, gl, _ = state(env)
            g[:] = np.concatenate([p.ravel(), v.ravel(), gl])
        return o, d, g

    rec = {k: [] for k in ("actions", "pos_step", "vel_step", "pos", "vel", "goal", "obst", "reward",
                           "in_step", "terminated", "truncated", "reached", "collision", "obs_valid", "obs",
                           "dist", "all_term", "all_trunc", "gs", "step_count")}
    r0 = {k: [] for k in ("pos", "vel", "goal", "obst", "obs", "dist", "gs")}
    for env in envs:
        obs, info = env.reset()
        p, v, g, ob = state(env)
        o, d, gs = pack_reset(env, obs, info)
        for k, x in zip(("pos", "vel", "goal", "obst", "obs", "dist", "gs"), (p, v, g, ob, o, d, gs)):
            r0[k].append(x)

    for t in range(T):
        row = {k: [] for k in rec}
        for e, env in enumerate(envs):
            act = action_fn(arngs[e], env, kind, N)
            o = np.zeros((N, D), np.float32)
            d = np.zeros(N, np.float32)
            gs = np.zeros(G, np.float32)
            rew = np.zeros(N, np.float64)
            fl = np.zeros((6, N), np.uint8)  # in_step, term, trunc, reached, collision, obs_valid
            if kind == "swarm":
                active = list(env.agents)
                captured = {}
                orig_mask = env._collision_mask

                def _spy(active_indices, _orig=orig_mask, _cap=captured):
                    out = _orig(active_indices)
                    _cap.update(out)
                    return out
                env._collision_mask = _spy
                adict = {a: act[int(a.split("_")[1])] for a in active}
                for i, a in enumerate(ids):
                    if a not in active:
                        act[i] = 0.0
                obs, rewards, term, trunc, infos = env.step(adict)
                env._collision_mask = orig_mask
                for i, a in enumerate(ids):
                    if a in rewards:
                        fl[0, i] = 1
                        rew[i] = rewards[a]
                        fl[1, i] = term[a]
                        fl[2, i] = trunc[a]
                        # ground truth of the two reward/termination causes, from the reference's own
                        # helpers on the post-step state (infos only ever carry False for both)
                        fl[3, i] = env._distance_to_goal(i) <= env.cfg.goal_radius
                        fl[4, i] = bool(captured[i])
                        assert bool(fl[1, i]) == bool(fl[3, i] or fl[4, i])
                        d[i] = env._distance_to_goal(i)
                    if a in obs:
                        fl[5, i] = 1
                        o[i] = obs[a]
                        assert float(d[i]) == infos[a]["distance_to_goal"]
                        assert not infos[a]["reached_goal"] and not infos[a]["collision"]
                        gs[:] = infos[a]["global_state"]
                at, atr = term["__all__"], trunc["__all__"]
                done = at or atr
            else:
                obs, r, tm, tr, info = env.step(act[0])
                fl[0, 0] = 1
                rew[0] = r
                fl[1, 0], fl[2, 0] = tm, tr
                fl[3, 0], fl[4, 0] = info["reached_goal"], info["collision"]
                fl[5, 0] = 1
                o[0] = obs
                d[0] = info["distance_to_goal"]
                at, atr = tm, tr
                done = tm or tr
            p_s, v_s, _, _ = state(env)
            sc = env.step_count
            if kind != "swarm" or not done:
                p, v, gl, _ = state(env)
                gs[:] = np.concatenate([p.ravel(), v.ravel(), gl])
            if done:
                obs, info = env.reset()
                o, d, gs = pack_reset(env, obs, info)
                fl[5, :] = 1
                sc = 0
            p, v, g, ob = state(env)
            for k, x in zip(rec.keys(), (act, p_s, v_s, p, v, g, ob, rew, fl[0], fl[1], fl[2], fl[3], fl[4],
                                         fl[5], o, d, np.uint8(at), np.uint8(atr), gs, np.int32(sc))):
                row[k].append(x)
        for k in rec:
            rec[k].append(np.stack(row[k]))
    out = {k: np.stack(v) for k, v in rec.items()}
    # keep the repo light: full obs / global_state rows only for the first OBS_HEAD steps, a
    # 64-bit digest per (step, env) for every step
    out["obs_hash"] = hash_rows(out["obs"])
    out["gs_hash"] = hash_rows(out["gs"])
    out["obs"] = out["obs"][:OBS_HEAD]
    out["gs"] = out["gs"][:OBS_HEAD]
    out.update({"reset0_" + k: np.stack(v) for k, v in r0.items()})
    out["seeds"] = np.asarray(seeds, np.uint64)
    out["meta"] = np.asarray(json.dumps({"kind": kind, "config": cfg, "T": T, "N": N, "M": M, "D": D,
                                         "action_seed": action_seed, "numpy": np.__version__}))
    return out


CASES = {
    # C1: reference parity run -- single drone, 8 obstacles, random actions (exercise all clips)
    "single_c1": ("single", {"max_steps": 400}, [123], 1000, _uniform_actions(-1.5, 1.5)),
    "single_goalseek": ("single", {"max_steps": 120, "num_obstacles": 8}, [7, 8], 1500, _goal_seek(0.4)),
    "single_m0": ("single", {"max_steps": 50, "num_obstacles": 0}, [5], 120, _goal_seek(0.2)),
    # reference tests/test_env_smoke.py ctor args
    "swarm_smoke_n3": ("swarm", {"num_drones": 3, "max_steps": 10}, [123], 40, _uniform_actions(-1.0, 1.0)),
    # C2 shape
    "swarm_c2_n8_m4": ("swarm", {"num_drones": 8, "num_obstacles": 4}, [0, 1, 2, 3], 1000,
                       _uniform_actions(-1.0, 1.0)),
    # C3 shape (global_state checked)
    "swarm_c3_n16_m8": ("swarm", {"num_drones": 16, "num_obstacles": 8, "world_size": 30.0}, [10, 11], 300,
                        _uniform_actions(-1.5, 1.5)),
    # C4 shape, reference-default world (reset heavy) and density-matched world
    "swarm_c4_n32_w20": ("swarm", {"num_drones": 32, "num_obstacles": 8}, [20, 21], 120,
                         _uniform_actions(-1.0, 1.0)),
    "swarm_c4_n32_w44": ("swarm", {"num_drones": 32, "num_obstacles": 8, "world_size": 44.0}, [30], 250,
                         _uniform_actions(-1.5, 1.5)),
    # north_star's horizon ("over 1000 steps") at the DEFAULT world for the C3 / C4 shapes (one env each: repo size)
    "swarm_c3_n16_w20_long": ("swarm", {"num_drones": 16, "num_obstacles": 8}, [12], 1000, _uniform_actions(-1.0, 1.0)),
    "swarm_c4_n32_w20_long": ("swarm", {"num_drones": 32, "num_obstacles": 8}, [22], 1000, _uniform_actions(-1.0, 1.0)),
    "swarm_c4_n32_w44_long": ("swarm", {"num_drones": 32, "num_obstacles": 8, "world_size": 44.0}, [31], 1000,
                              _uniform_actions(-1.0, 1.0)),
    # C5 shape
    "swarm_c5_n128_w70": ("swarm", {"num_drones": 128, "num_obstacles": 8, "world_size": 70.0}, [40], 24,
                          _uniform_actions(-1.0, 1.0)),
    # goal seeking: reach events, parked drones, time-limit truncation
    "swarm_goalseek_n5": ("swarm", {"num_drones": 5, "num_obstacles": 8, "max_steps": 150}, [50, 51, 52], 1500,
                          _goal_seek(0.35)),
    "swarm_goalseek_n12": ("swarm", {"num_drones": 12, "num_obstacles": 2, "max_steps": 200,
                                     "world_size": 36.0, "collision_radius": 0.2}, [60, 61], 900, _goal_seek(0.5)),
    # thresholds / scalars that are not f32-representable, odd K/S/M
    "swarm_oddcfg_n6": ("swarm", {"num_drones": 6, "num_obstacles": 5, "sensed_obstacles": 3, "neighbor_k": 2,
                                  "world_size": 12.7, "dt": 0.07, "max_speed": 3.7, "max_accel": 2.3,
                                  "collision_radius": 0.33, "goal_radius": 0.83, "obstacle_radius": 0.71,
                                  "desired_spacing": 2.4, "reward_progress_scale": 1.7, "reward_goal": 21.3,
                                  "reward_collision": -17.9, "reward_formation_scale": 0.13, "max_steps": 90},
                        [70, 71, 72], 700, _goal_seek(0.6)),
    # padding edge cases: N-1 < K, M < S, N = 1, M = 0
    "swarm_n2_m2": ("swarm", {"num_drones": 2, "num_obstacles": 2, "max_steps": 60}, [80, 81], 300, _goal_seek(0.3)),
    "swarm_n1_m0": ("swarm", {"num_drones": 1, "num_obstacles": 0, "max_steps": 80}, [90], 300, _goal_seek(0.3)),
    # curriculum_v1.yaml stage env_configs (configs/curriculum_v1.yaml:12-16,51-55)
    "swarm_curr_stage1": ("swarm", {"num_drones": 3, "num_obstacles": 0, "max_steps": 300, "world_size": 20.0},
                          [100, 101], 700, _goal_seek(0.3)),
    "swarm_curr_stage4": ("swarm", {"num_drones": 8, "num_obstacles": 12, "max_steps": 450, "world_size": 28.0},
                          [110], 600, _goal_seek(0.5)),
}


def main(names=None):
    os.makedirs(OUT, exist_ok=True)
    for name, (kind, cfg, seeds, T, fn) in CASES.items():
        if names and name not in names:
            continue
        out = roll(kind, cfg, seeds, T, fn)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **out)
        ev = int(out["all_term"].sum()), int(out["all_trunc"].sum()), int(out["reached"].sum()), \
            int(out["collision"].sum()), int((out["in_step"] == 0).sum())
        print(f"{name}: T={T} E={len(seeds)} episodes term/trunc={ev[0]}/{ev[1]} reached={ev[2]} "
              f"collided={ev[3]} parked_slots={ev[4]} size={os.path.getsize(path)/1e6:.2f}MB")


if __name__ == "__main__":
    main(sys.argv[1:] or None)
