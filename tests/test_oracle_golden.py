"""Pin the C oracle (oracle/swarm_oracle.c) bit-for-bit against the golden fixtures that
oracle/gen_golden.py recorded from the unmodified reference (CPU only)."""
import numpy as np
import pytest

import parity_util as pu
import swarm_oracle as so


@pytest.mark.parametrize("name", pu.golden_names())
def test_oracle_matches_reference_golden(name):
    g = pu.load_golden(name)
    m = g["meta"]
    o = so.OracleSwarm(len(g["seeds"]), m["config"], kind=m["kind"])
    assert o.D == m["D"]
    stats = pu.replay_and_compare(o, g)
    assert stats["steps"] == m["T"]


def test_seedsequence_pcg64_matches_numpy():
    L = so.lib()
    for seed in [0, 1, 123, 2**32 - 1, 2**32, 2**40 + 7, 2**63 + 11, 2**64 - 1]:
        st = np.zeros(4, np.uint64)
        L.oracle_seed(seed, st.ctypes.data)
        gen = np.random.default_rng(seed)
        s = gen.bit_generator.state["state"]
        want = (s["state"] >> 64, s["state"] & (2**64 - 1), s["inc"] >> 64, s["inc"] & (2**64 - 1))
        assert tuple(int(x) for x in st) == want
        got = [L.oracle_pcg64_next(st.ctypes.data) for _ in range(8)]
        assert got == [int(x) for x in gen.integers(0, 2**64, 8, dtype=np.uint64)]


def test_uniform_draws_match_numpy():
    o = so.OracleSwarm(3, {"num_drones": 5, "num_obstacles": 7, "world_size": 13.3}, kind="swarm")
    o.seed([5, 6, 7])
    o.reset()
    for e, seed in enumerate([5, 6, 7]):
        r = np.random.default_rng(seed)
        b = 13.3 / 2.0
        assert np.array_equal(o.positions[e], r.uniform(-b, b, size=(5, 3)).astype(np.float32))
        assert np.array_equal(o.goal[e], r.uniform(-b, b, size=3).astype(np.float32))
        assert np.array_equal(o.obstacles[e], r.uniform(-b, b, size=(7, 3)).astype(np.float32))


def test_known_answers_from_survey():
    """SURVEY.md section 4 known answers (generated from the reference, seed 123)."""
    o = so.OracleSwarm(1, {"num_drones": 3, "max_steps": 10}, kind="swarm")
    o.seed([123])
    o.reset()
    gs = np.array([3.6470373, -8.923579, -5.5928025, -6.312564, -6.481882, 6.24189, 8.4669, -4.468512,
                   6.395091] + [0.0] * 9 + [7.797854, 0.2594091, -5.100708], np.float32)
    assert np.array_equal(o.global_state[0], gs)
    nb = np.array([4.8198624, 4.455067, 11.987894, 13.667051, -9.959601, 2.4416971, 11.834692, 15.659358,
                   0, 0, 0, 0], np.float32)
    assert np.array_equal(o.obs[0, 0, 9:21], nb)
    o.step(np.zeros((1, 3, 3), np.float32))
    assert o.reward[0].tolist() == [-1.8244807004928587, -1.9182087421417235, -1.7687857389450072]
    assert not o.terminated.any() and not o.truncated.any() and o.obs_valid.all()

    s = so.OracleSwarm(1, {"max_steps": 10}, kind="single")
    s.seed([123])
    s.reset()
    assert np.array_equal(s.obs[0, 0, :9], np.array([3.6470373, -8.923579, -5.5928025, 0, 0, 0, -9.959601,
                                                     2.4416971, 11.834692], np.float32))
    assert float(s.dist[0, 0]) == 15.659358024597168
    s.step(np.zeros((1, 1, 3), np.float32))
    assert s.reward[0, 0] == 0.0
    s.step(np.array([[[1, -1, 0.5]]], np.float32))
    assert s.reward[0, 0] == -0.016613006591796875
    assert np.array_equal(s.obs[0, 0, :6], np.array([3.6670372, -8.94358, -5.5828023, 0.2, -0.2, 0.1], np.float32))
