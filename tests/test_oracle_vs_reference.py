"""Live pin: the C oracle against the UNMODIFIED reference envs, imported from /root/reference.

Runs only where the reference tree exists (the build container; `/root/reference` does not travel to the
GPU box, where this file skips).  The committed fixtures under tests/golden/ are recordings of exactly this
comparison's reference side (oracle/gen_golden.py); this test re-does it live on fresh seeds and configs
that are NOT among the fixtures, so the oracle is pinned by more than the recorded cases.
"""
import os
import sys

import numpy as np
import pytest

import parity_util as pu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import ref_loader  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")

CASES = [
    ("swarm", {"num_drones": 4, "num_obstacles": 6, "max_steps": 40, "world_size": 14.0}, 6, 150),
    ("swarm", {"num_drones": 9, "num_obstacles": 3, "neighbor_k": 4, "sensed_obstacles": 2, "max_steps": 60,
               "desired_spacing": 1.75, "reward_formation_scale": 0.3}, 4, 120),
    ("swarm", {"num_drones": 16, "num_obstacles": 8}, 3, 60),
    ("single", {"num_obstacles": 5, "max_steps": 35, "world_size": 12.0, "goal_radius": 1.1}, 8, 200),
    # the stage env_configs of configs/curriculum_v1.yaml:12-55 (BASELINE config 4 trains through them)
    ("swarm", {"num_drones": 3, "num_obstacles": 0, "max_steps": 300, "world_size": 20.0}, 4, 120),
    ("swarm", {"num_drones": 3, "num_obstacles": 4, "max_steps": 350, "world_size": 20.0}, 4, 120),
    ("swarm", {"num_drones": 5, "num_obstacles": 8, "max_steps": 400, "world_size": 24.0}, 4, 120),
    ("swarm", {"num_drones": 8, "num_obstacles": 12, "max_steps": 450, "world_size": 28.0}, 3, 100),
]


@pytest.mark.parametrize("kind,cfg,E,T", CASES)
def test_oracle_matches_live_reference(kind, cfg, E, T):
    import swarm_oracle as so

    Single, Swarm = ref_loader.load_reference_envs()
    N = int(cfg.get("num_drones", 3)) if kind == "swarm" else 1
    seeds = np.arange(7000, 7000 + E, dtype=np.uint64)
    refs = [(Swarm if kind == "swarm" else Single)({**cfg, "seed": int(s)}) for s in seeds]
    o = so.OracleSwarm(E, cfg, kind=kind)
    o.seed(seeds)
    o.reset()
    ids = [f"drone_{i}" for i in range(N)]
    for e, r in enumerate(refs):
        obs, info = r.reset()
        if kind == "swarm":
            for i, a in enumerate(ids):
                assert np.array_equal(pu.bits(obs[a]), pu.bits(o.obs[e, i]))
                assert info[a]["distance_to_goal"] == float(o.dist[e, i])
            assert np.array_equal(pu.bits(info[ids[0]]["global_state"]), pu.bits(o.global_state[e]))
        else:
            assert np.array_equal(pu.bits(obs), pu.bits(o.obs[e, 0]))
    rng = np.random.default_rng(99)
    events = dict(done=0, reached=0, collided=0)
    for t in range(T):
        act = rng.uniform(-1.4, 1.4, size=(E, N, 3)).astype(np.float32)
        if t % 3 == 0:  # head for the goal now and then: parks drones, reaches goals
            pos = o.positions
            d = o.goal[:, None, :] - pos
            act = (d / np.maximum(np.linalg.norm(d, axis=2, keepdims=True), 1e-6)).astype(np.float32)
        was_active = o.active.astype(bool).copy()
        o.step(act, auto_reset=False)
        for e, r in enumerate(refs):
            if kind == "swarm":
                if not r.agents:
                    continue
                actions = {a: act[e, i] for i, a in enumerate(ids) if a in r.agents}
                obs, rew, term, trunc, infos = r.step(actions)
                for i, a in enumerate(ids):
                    if not was_active[e, i]:
                        continue
                    assert rew[a] == o.reward[e, i], (t, e, a)
                    assert term[a] == bool(o.terminated[e, i]) and trunc[a] == bool(o.truncated[e, i])
                    assert (a in obs) == bool(o.obs_valid[e, i])
                    if a in obs:
                        if not np.array_equal(pu.bits(obs[a]), pu.bits(o.obs[e, i])):   # exact tie ordered differently (T5)
                            assert pu.obs_row_ok_up_to_ties(kind, {**so.DEFAULTS, **cfg}, r.positions, r.velocities, r.goal,
                                                            r.obstacles, i, o.obs[e, i])
                        assert infos[a]["collision"] == bool(o.collision[e, i])
                        assert infos[a]["reached_goal"] == bool(o.reached[e, i])
                        assert np.array_equal(pu.bits(infos[a]["global_state"]), pu.bits(o.global_state[e]))
                    events["reached"] += int(o.reached[e, i])
                    events["collided"] += int(o.collision[e, i])
                assert term["__all__"] == bool(o.all_terminated[e]) and trunc["__all__"] == bool(o.all_truncated[e])
                assert np.array_equal(pu.bits(r.positions), pu.bits(o.positions[e]))
                assert np.array_equal(pu.bits(r.velocities), pu.bits(o.velocities[e]))
                done = term["__all__"] or trunc["__all__"]
            else:
                obs, rew, term, trunc, info = r.step(act[e, 0])
                assert rew == o.reward[e, 0] and term == bool(o.terminated[e, 0]) and trunc == bool(o.truncated[e, 0])
                assert np.array_equal(pu.bits(obs), pu.bits(o.obs[e, 0]))
                assert np.array_equal(pu.bits(r.position), pu.bits(o.positions[e, 0]))
                done = term or trunc
            if done:
                events["done"] += 1
                r.reset()     # continues the env's PCG64 stream, as RLlib's sampler does
                m = np.zeros(E, np.uint8)
                m[e] = 1
                o.reset(m)
                if kind == "swarm":
                    assert np.array_equal(pu.bits(r.positions), pu.bits(o.positions[e]))
                    assert np.array_equal(pu.bits(r.obstacles), pu.bits(o.obstacles[e]))
                else:
                    assert np.array_equal(pu.bits(r.goal), pu.bits(o.goal[e]))
    assert events["done"] > 0
