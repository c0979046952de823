"""Adapter exposing a SwarmEngine (CUDA, through the C ABI) with the backend interface that
tests/parity_util.py replays goldens through (numpy views, leading env axis)."""
from __future__ import annotations

import numpy as np
import torch

import swarm_b200


class EngineBackend:
    def __init__(self, num_envs, config, kind="swarm", reward64=True, **kw):
        self.eng = swarm_b200.SwarmEngine(num_envs, config, kind=kind, device="cuda:0", reward64=reward64, **kw)
        self.D = self.eng.D

    def seed(self, seeds):
        self.eng.seed(np.asarray(seeds, np.uint64))

    def reset(self, mask=None):
        self.eng.reset(None if mask is None else torch.as_tensor(np.asarray(mask, np.uint8)))

    def step(self, actions, auto_reset=True):
        a = torch.from_numpy(np.ascontiguousarray(actions, np.float32)).to("cuda:0")
        self.eng.step(a, auto_reset=auto_reset)

    def _np(self, t):
        return t.detach().cpu().numpy()

    positions = property(lambda s: s._np(s.eng.positions))
    velocities = property(lambda s: s._np(s.eng.velocities))
    goal = property(lambda s: s._np(s.eng.goal))
    obstacles = property(lambda s: s._np(s.eng.obstacles))
    step_count = property(lambda s: s._np(s.eng.step_count))
    obs = property(lambda s: s._np(s.eng.obs))
    reward = property(lambda s: s._np(s.eng.reward64 if s.eng.reward64 is not None else s.eng.reward))
    reward32 = property(lambda s: s._np(s.eng.reward))
    dist = property(lambda s: s._np(s.eng.dist))
    terminated = property(lambda s: s._np(s.eng.terminated))
    truncated = property(lambda s: s._np(s.eng.truncated))
    reached = property(lambda s: s._np(s.eng.reached))
    collision = property(lambda s: s._np(s.eng.collision))
    obs_valid = property(lambda s: s._np(s.eng.obs_valid))
    all_terminated = property(lambda s: s._np(s.eng.all_terminated))
    all_truncated = property(lambda s: s._np(s.eng.all_truncated))
    global_state = property(lambda s: s._np(s.eng.global_state))
    active = property(lambda s: s._np(s.eng.alive).astype(np.uint8))
    dr_params = property(lambda s: s._np(s.eng.dr_params))
