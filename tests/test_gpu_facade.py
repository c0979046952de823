"""The reference-compatible dict API (DroneSwarmEnv / SingleDroneEnv façades over the CUDA engine)
replayed against golden fixtures recorded from the reference, plus the reference's own two smoke
tests (tests/test_env_smoke.py) run verbatim against the façade classes."""
import numpy as np
import pytest

import parity_util as pu

pytestmark = pytest.mark.gpu


def test_reference_smoke_single():
    from swarm_b200 import SingleDroneEnv

    env = SingleDroneEnv({"seed": 123, "max_steps": 10})
    obs, info = env.reset()
    assert obs.shape == env.observation_space.shape
    assert "distance_to_goal" in info
    action = np.zeros(3, dtype=np.float32)
    obs, reward, terminated, truncated, info = env.step(action)
    assert obs.shape == env.observation_space.shape
    assert isinstance(reward, float)
    assert isinstance(terminated, bool)
    assert isinstance(truncated, bool)
    assert "distance_to_goal" in info
    # SURVEY section 4 known answers
    assert reward == 0.0
    obs, reward, *_ = env.step(np.array([1, -1, 0.5], np.float32))
    assert reward == -0.016613006591796875


def test_reference_smoke_multi_agent():
    from swarm_b200 import DroneSwarmEnv

    env = DroneSwarmEnv({"num_drones": 3, "seed": 123, "max_steps": 10})
    obs, infos = env.reset()
    assert len(obs) == 3
    assert len(infos) == 3
    actions = {agent_id: np.zeros(3, dtype=np.float32) for agent_id in obs}
    next_obs, rewards, terminated, truncated, infos = env.step(actions)
    assert len(next_obs) == 3
    assert len(rewards) == 3
    assert "__all__" in terminated
    assert "__all__" in truncated
    assert all(isinstance(v, float) for v in rewards.values())
    assert all("global_state" in infos[agent_id] for agent_id in infos)
    assert rewards == {"drone_0": -1.8244807004928587, "drone_1": -1.9182087421417235,
                       "drone_2": -1.7687857389450072}
    assert env.cfg.desired_spacing == 2.5 and env.agents == ["drone_0", "drone_1", "drone_2"]
    assert env.positions.shape == (3, 3) and env.obstacles.shape == (8, 3) and env.goal.shape == (3,)
    assert env.action_space.sample().shape == (3,)


@pytest.mark.parametrize("name,steps", [("swarm_goalseek_n5", 400), ("swarm_oddcfg_n6", 300),
                                        ("swarm_smoke_n3", 40), ("swarm_n2_m2", 200)])
def test_swarm_facade_dict_contract_matches_golden(name, steps):
    from swarm_b200 import DroneSwarmEnv

    g = pu.load_golden(name)
    cfg = dict(g["meta"]["config"])
    cfg["seed"] = int(g["seeds"][0])
    N = g["meta"]["N"]
    ids = [f"drone_{i}" for i in range(N)]
    env = DroneSwarmEnv(cfg)
    obs, infos = env.reset()
    assert list(obs) == ids
    for i, a in enumerate(ids):
        assert np.array_equal(pu.bits(obs[a]), pu.bits(g["reset0_obs"][0, i]))
        assert infos[a]["distance_to_goal"] == float(g["reset0_dist"][0, i])
        assert set(infos[a]) == {"distance_to_goal", "global_state"}
    head = g["obs"].shape[0]
    for t in range(min(steps, g["actions"].shape[0])):
        active = list(env.agents)
        ins = g["in_step"][t, 0].astype(bool)
        assert active == [a for i, a in enumerate(ids) if ins[i]], t
        # leave one agent's action out now and then: a missing key means a zero action (:104)
        adict = {a: g["actions"][t, 0, env.agent_id_to_index[a]] for a in active}
        obs, rew, term, trunc, infos = env.step(adict)
        valid = g["obs_valid"][t, 0].astype(bool)
        done = bool(g["all_term"][t, 0] or g["all_trunc"][t, 0])
        assert list(rew) == active and set(term) == set(active) | {"__all__"} and set(trunc) == set(term)
        for i, a in enumerate(ids):
            if ins[i]:
                assert rew[a] == g["reward"][t, 0, i] and isinstance(rew[a], float)
                assert term[a] == bool(g["terminated"][t, 0, i]) and trunc[a] == bool(g["truncated"][t, 0, i])
        assert term["__all__"] == bool(g["all_term"][t, 0]) and trunc["__all__"] == bool(g["all_trunc"][t, 0])
        if done:
            # on an episode-ending step nobody continues: obs / infos are empty (drone_swarm_env.py:154)
            assert obs == {} and infos == {}
            assert env.agents == []
            obs, infos = env.reset()
            for i, a in enumerate(ids):
                if t < head:
                    assert np.array_equal(pu.bits(obs[a]), pu.bits(g["obs"][t, 0, i]))
                assert infos[a]["distance_to_goal"] == float(g["dist"][t, 0, i])
        else:
            assert list(obs) == [a for i, a in enumerate(ids) if valid[i]] == list(infos)
            for i, a in enumerate(ids):
                if valid[i]:
                    if t < head:
                        assert np.array_equal(pu.bits(obs[a]), pu.bits(g["obs"][t, 0, i]))
                        assert np.array_equal(pu.bits(infos[a]["global_state"]), pu.bits(g["gs"][t, 0]))
                    assert infos[a]["distance_to_goal"] == float(g["dist"][t, 0, i])
                    assert infos[a]["reached_goal"] is False and infos[a]["collision"] is False
                    assert set(infos[a]) == {"distance_to_goal", "reached_goal", "collision", "global_state"}
        assert np.array_equal(pu.bits(env.positions), pu.bits(g["pos"][t, 0]))


def test_single_facade_matches_golden():
    from swarm_b200 import SingleDroneEnv

    g = pu.load_golden("single_goalseek")
    cfg = dict(g["meta"]["config"])
    cfg["seed"] = int(g["seeds"][0])
    env = SingleDroneEnv(cfg)
    obs, info = env.reset()
    assert np.array_equal(pu.bits(obs), pu.bits(g["reset0_obs"][0, 0]))
    for t in range(400):
        o, r, tm, tr, info = env.step(g["actions"][t, 0, 0])
        assert r == g["reward"][t, 0, 0] and tm == bool(g["terminated"][t, 0, 0]) and tr == bool(g["truncated"][t, 0, 0])
        assert info["reached_goal"] == bool(g["reached"][t, 0, 0]) and info["collision"] == bool(g["collision"][t, 0, 0])
        if tm or tr:
            o, info = env.reset()
        assert np.array_equal(pu.bits(o), pu.bits(g["obs"][t, 0, 0])) if t < g["obs"].shape[0] else True
        assert info["distance_to_goal"] == float(g["dist"][t, 0, 0])


def test_dead_env_step_and_reseed():
    from swarm_b200 import DroneSwarmEnv

    env = DroneSwarmEnv({"num_drones": 3, "seed": 123, "max_steps": 2})
    env.reset()
    z = {a: np.zeros(3, np.float32) for a in env.agent_ids}
    env.step(z)
    _, _, term, trunc, _ = env.step(z)
    assert trunc["__all__"] and not term["__all__"] and env.agents == []
    assert env.step(z) == ({}, {}, {"__all__": True}, {"__all__": False}, {})
    o1, _ = env.reset(seed=123)
    env2 = DroneSwarmEnv({"num_drones": 3, "seed": 123, "max_steps": 2})
    o2, _ = env2.reset()
    assert all(np.array_equal(o1[a], o2[a]) for a in o1)


@pytest.mark.gpu
def test_vector_env_matches_per_env_facades():
    """VectorSwarmEnv (RLlib BaseEnv-style poll / send_actions / try_reset over ONE engine) hands out
    what E separate DroneSwarmEnv instances do, episode after episode."""
    import swarm_b200

    cfg = {"num_drones": 4, "num_obstacles": 6, "max_steps": 12, "world_size": 12.0}
    E, base = 6, 4000
    vec = swarm_b200.VectorSwarmEnv(E, cfg, base_seed=base)
    refs = [swarm_b200.DroneSwarmEnv({**cfg, "seed": base + e}) for e in range(E)]
    obs, rew, term, trunc, infos, _ = vec.poll()
    cur = {}
    for e, r in enumerate(refs):
        o, i = r.reset()
        cur[e] = o
        assert set(obs[e]) == set(o) and all(np.array_equal(obs[e][a], o[a]) for a in o)
        assert np.array_equal(infos[e]["drone_0"]["global_state"], i["drone_0"]["global_state"])
    rng = np.random.default_rng(2)
    episodes = 0
    for t in range(60):
        actions = {e: {a: rng.uniform(-1, 1, 3).astype(np.float32) for a in cur[e]} for e in range(E)}
        vec.send_actions(actions)
        obs, rew, term, trunc, infos, _ = vec.poll()
        for e, r in enumerate(refs):
            o, rw, te, tr, inf = r.step(actions[e])
            assert rew[e] == rw and term[e] == te and trunc[e] == tr, (t, e)
            assert set(obs[e]) == set(o) and all(np.array_equal(obs[e][a], o[a]) for a in o), (t, e)
            assert all(infos[e][a]["collision"] == inf[a]["collision"] and
                       np.array_equal(infos[e][a]["global_state"], inf[a]["global_state"]) for a in o)
            cur[e] = o
            if te["__all__"] or tr["__all__"]:
                episodes += 1
                ro, ri = vec.try_reset(e)
                o2, i2 = r.reset()
                assert all(np.array_equal(ro[e][a], o2[a]) for a in o2), (t, e)
                assert ri[e]["drone_1"]["distance_to_goal"] == i2["drone_1"]["distance_to_goal"]
                cur[e] = o2
    assert episodes >= E
    vec.stop()
    for r in refs:
        r.close()


@pytest.mark.gpu
def test_vector_env_batch_interface_and_lazy_dicts():
    """`step_batch` hands out one dict of [E, ...] arrays (global_state column included) that agrees with the lazy
    per-env dicts of `poll()` and with E separate facade envs; untouched envs cost no dict building."""
    import swarm_b200

    cfg = {"num_drones": 5, "num_obstacles": 4, "max_steps": 9, "world_size": 11.0}
    E, base = 8, 900
    vec = swarm_b200.VectorSwarmEnv(E, cfg, base_seed=base)
    refs = [swarm_b200.DroneSwarmEnv({**cfg, "seed": base + e}) for e in range(E)]
    obs0 = vec.poll()[0]
    cur = []
    for e, r in enumerate(refs):
        o, _ = r.reset()
        cur.append(o)
        assert all(np.array_equal(obs0[e][a], o[a]) for a in o)
    rng = np.random.default_rng(5)
    for t in range(40):
        act = rng.uniform(-1, 1, (E, 5, 3)).astype(np.float32)
        b = vec.step_batch(act)
        lazy = vec.poll()
        assert lazy[0]._cache == {}                      # nothing built until somebody looks
        assert b["global_state"].shape == (E, 6 * 5 + 3) and b["global_state"] is vec.global_state_column()
        for e, r in enumerate(refs):
            actions = {a: act[e, int(a[-1])] for a in cur[e]}
            o, rw, te, tr, inf = r.step(actions)
            done = te["__all__"] or tr["__all__"]
            assert bool(b["all_terminated"][e]) == te["__all__"] and bool(b["all_truncated"][e]) == tr["__all__"]
            for a in rw:
                i = int(a[-1])
                assert b["was_active"][e, i] and b["reward"][e, i] == rw[a]
                assert bool(b["terminated"][e, i]) == te[a] and bool(b["truncated"][e, i]) == tr[a]
            assert lazy[1][e] == rw and lazy[2][e] == te and lazy[3][e] == tr
            assert set(lazy[0][e]) == set(o) and all(np.array_equal(lazy[0][e][a], o[a]) for a in o)
            if not done:
                for a in o:
                    i = int(a[-1])
                    assert np.array_equal(b["obs"][e, i], o[a])
                    assert np.array_equal(b["global_state"][e], inf[a]["global_state"])
                cur[e] = o
            else:
                ro, _ = vec.try_reset(e)
                o2, _ = r.reset()
                assert all(np.array_equal(ro[e][a], o2[a]) for a in o2)
                cur[e] = o2
    vec.stop()
    for r in refs:
        r.close()


@pytest.mark.gpu
def test_device_rollout_matches_the_per_env_facade_loop():
    """`collect_rollout` (policy -> step -> [T,E,...] columns, all on the device, auto-reset on) against E facade
    envs driven by the same deterministic policy in a Python loop with `env.reset()` on episode end: obs, rewards,
    flags, the agent mask and the global_state column agree bit for bit (a fake RLlib sampler)."""
    import torch
    import swarm_b200

    cfg = {"num_drones": 4, "num_obstacles": 5, "max_steps": 14, "world_size": 10.0}
    E, T, base = 6, 45, 700

    def policy_np(o):      # a deterministic stand-in for the actor: fly at the goal, wobble with the position
        return np.tanh(0.4 * o[..., 6:9] + 0.2 * np.sin(3.0 * o[..., 0:3])).astype(np.float32)

    def policy_t(obs, valid):
        # (same arithmetic on the host for both sides, so the comparison is about the env, not about libm)
        return torch.from_numpy(policy_np(obs.cpu().numpy())).to(obs.device)

    eng = swarm_b200.SwarmEngine(E, cfg, device="cuda:0", reward64=True)
    eng.seed(np.arange(base, base + E, dtype=np.uint64))
    eng.reset()
    ro = swarm_b200.collect_rollout(eng, policy_t, T, value_fn=lambda gs: gs.sum(-1))
    got = {k: v.cpu().numpy() for k, v in ro.items()}
    assert got["obs"].shape == (T, E, 4, eng.D) and got["global_state"].shape == (T, E, eng.R)
    for e in range(E):
        env = swarm_b200.DroneSwarmEnv({**cfg, "seed": base + e})
        obs, infos = env.reset()
        for t in range(T):
            assert got["agent_mask"][t, e].tolist() == [a in obs for a in env.agent_ids], (e, t)
            for a in obs:
                i = env.agent_id_to_index[a]
                assert np.array_equal(got["obs"][t, e, i], obs[a]), (e, t, a)
                assert np.array_equal(got["global_state"][t, e], infos[a]["global_state"]), (e, t)
            full = np.stack([obs.get(a, got["obs"][t, e, i]) for i, a in enumerate(env.agent_ids)])
            act = policy_np(full)
            actions = {a: act[env.agent_id_to_index[a]] for a in obs}
            obs, rew, term, trunc, infos = env.step(actions)
            for a in rew:
                i = env.agent_id_to_index[a]
                assert np.float32(rew[a]) == got["rewards"][t, e, i], (e, t, a)
                assert term[a] == bool(got["terminateds"][t, e, i]) and trunc[a] == bool(got["truncateds"][t, e, i])
            done = term["__all__"] or trunc["__all__"]
            assert done == bool(got["dones"][t, e]), (e, t)
            if done:
                obs, infos = env.reset()
        env.close()
    assert np.allclose(got["values"], got["global_state"].sum(-1), rtol=1e-4, atol=1e-2)   # (a device-side sum)
