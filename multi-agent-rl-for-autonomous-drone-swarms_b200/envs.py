"""Reference-compatible env classes backed by the CUDA engine (the drop-in façade).

`DroneSwarmEnv` / `SingleDroneEnv` keep the reference's constructor, `reset` / `step` contract,
dict population rules and public attributes, so an env creator `lambda cfg: DroneSwarmEnv(cfg)`
registered with `ray.tune.registry.register_env` (reference scripts/train_ctde.py:125,
scripts/evaluate_protocol.py:362-364, :427-431) or direct construction works unchanged:

    reference src/swarm_marl/envs/drone_swarm_env.py   DroneSwarmEnv   :17-302
    reference src/swarm_marl/envs/single_drone_env.py  SingleDroneEnv  :12-159

Each instance is one env (E = 1) stepped through the host-buffer C-ABI entry point; all
arithmetic runs in the sm_100a kernels (no CPU path).  For throughput use `SwarmEngine`
(thousands of env instances per launch) -- the façade exists for API compatibility.
"""
from __future__ import annotations

import os
from collections.abc import Mapping
from typing import Any, Callable

import numpy as np

from .config import DroneEnvConfig

try:  # gymnasium is optional here (it is not installed in the build image)
    import gymnasium as _gym
    from gymnasium import spaces as _spaces

    _GymEnv = _gym.Env
    Box = _spaces.Box
except ModuleNotFoundError:  # pragma: no cover - exercised in this image
    class _GymEnv:  # type: ignore[no-redef]
        metadata: dict = {"render_modes": []}

        def reset(self, *, seed=None, options=None):
            return None

    class Box:  # type: ignore[no-redef]
        """Minimal `gymnasium.spaces.Box` stand-in: shape / dtype / low / high / sample / contains."""

        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            self.dtype = np.dtype(dtype)
            self.shape = tuple(shape) if shape is not None else np.shape(low)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
            self._rng = np.random.default_rng(seed)

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return self._rng.uniform(lo, hi).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

try:
    from ray.rllib.env.multi_agent_env import MultiAgentEnv as _MultiAgentEnv
except ModuleNotFoundError:  # same fallback as the reference (drone_swarm_env.py:8-12)
    class _MultiAgentEnv:  # type: ignore[no-redef]
        pass


def _entropy_seed() -> int:
    """`np.random.default_rng(None)` draws OS entropy; so do we (63 bits)."""
    return int.from_bytes(os.urandom(8), "little") >> 1


class _EngineBacked:
    """Shared plumbing: one SwarmEngine with E = 1 and pinned host outputs."""

    def _make_engine(self, config: dict[str, Any], kind: str, device):
        from .engine import SwarmEngine

        self._engine = SwarmEngine(1, config, kind=kind, device=device or "cuda", global_state=True, reward64=True)
        seed = self.cfg.seed if self.cfg.seed is not None else _entropy_seed()
        self._engine.seed(np.asarray([seed], dtype=np.uint64))
        self._host = self._engine.host_buffers()
        # numpy views of the pinned host block, made once (a step is then: write the action row, one C-ABI call --
        # H2D, launches, ONE D2H of the whole output block -- and dict building from these views)
        self._hn = {k: v.numpy() for k, v in self._host.items()}

    def _fetch_reset_outputs(self):
        """obs / dist / global_state after `reset()`: the output block comes back in one copy."""
        self._engine.fetch_outputs()
        hn = self._hn
        return hn["obs"][0], hn["dist"][0], (hn["global_state"][0] if "global_state" in hn else None)

    # state attributes read by scripts/visualize_swarm.py:76-81,105-110
    @property
    def obstacles(self) -> np.ndarray:
        return self._engine.obstacles[0].cpu().numpy()

    @property
    def goal(self) -> np.ndarray:
        return self._engine.goal[0].cpu().numpy()

    @property
    def step_count(self) -> int:
        return int(self._engine.step_count[0].item())

    def close(self):
        self._engine.close()


class DroneSwarmEnv(_EngineBacked, _MultiAgentEnv):
    """Multi-agent 3D swarm environment (reference drone_swarm_env.py:17)."""

    def __init__(self, config: dict[str, Any] | None = None, device=None):
        super().__init__()
        cfg_dict = dict(config or {})
        self.num_drones = int(cfg_dict.get("num_drones", 3))                      # :32
        env_cfg = {k: v for k, v in cfg_dict.items() if k != "num_drones"}        # :33
        self.cfg = DroneEnvConfig.from_dict(env_cfg)                              # :34
        self.agent_ids = [f"drone_{i}" for i in range(self.num_drones)]           # :37
        self.agent_id_to_index = {a: i for i, a in enumerate(self.agent_ids)}
        self.agents = list(self.agent_ids)
        self._obs_dim = 9 + self.cfg.neighbor_k * 4 + self.cfg.sensed_obstacles * 4   # :41-45
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=(self._obs_dim,), dtype=np.float32)
        self.action_space = Box(low=-1.0, high=1.0, shape=(3,), dtype=np.float32)
        self._make_engine({**env_cfg, "num_drones": self.num_drones}, "swarm", device)

    # -- reference attributes
    @property
    def positions(self) -> np.ndarray:
        return self._engine.positions[0].cpu().numpy()

    @property
    def velocities(self) -> np.ndarray:
        return self._engine.velocities[0].cpu().numpy()

    def reset(self, *, seed: int | None = None, options: dict[str, Any] | None = None):
        """drone_swarm_env.py:65-90."""
        if seed is not None:
            self._engine.seed(np.asarray([seed], dtype=np.uint64))
        self.agents = list(self.agent_ids)
        self._engine.reset()
        obs, dist, gs = self._fetch_reset_outputs()
        observations = {a: obs[i].copy() for i, a in enumerate(self.agent_ids)}
        infos = {a: {"distance_to_goal": float(dist[i]), "global_state": gs.copy()}
                 for i, a in enumerate(self.agent_ids)}
        return observations, infos

    def step(self, action_dict: dict[str, np.ndarray]):
        """drone_swarm_env.py:92-174 (same dict population rules)."""
        active = list(self.agents)
        if not active:                                                            # :94-95
            return {}, {}, {"__all__": True}, {"__all__": False}, {}
        hn = self._hn
        act = hn["actions"]
        act[...] = 0.0
        for a in active:                                                          # :103-106
            raw = action_dict.get(a)
            if raw is not None:
                act[0, self.agent_id_to_index[a]] = np.asarray(raw, dtype=np.float32).reshape(3)
        self._engine.step_host(None, auto_reset=False)
        rew = hn["reward64"][0].tolist()
        term, trunc = hn["terminated"][0].tolist(), hn["truncated"][0].tolist()
        valid = hn["obs_valid"][0].tolist()
        obs_rows, dist = hn["obs"][0], hn["dist"][0].tolist()
        reached, collision = hn["reached"][0].tolist(), hn["collision"][0].tolist()
        gs = hn["global_state"][0]
        rewards: dict[str, float] = {}
        terminated: dict[str, bool] = {}
        truncated: dict[str, bool] = {}
        infos: dict[str, dict[str, Any]] = {}
        obs: dict[str, np.ndarray] = {}
        next_active: list[str] = []
        for a in active:                                                          # :141-162
            i = self.agent_id_to_index[a]
            rewards[a] = rew[i]
            terminated[a] = bool(term[i])
            truncated[a] = bool(trunc[i])
            if valid[i]:
                obs[a] = obs_rows[i].copy()
                infos[a] = {
                    "distance_to_goal": dist[i],
                    "reached_goal": bool(reached[i]),
                    "collision": bool(collision[i]),
                    "global_state": gs.copy(),
                }
                next_active.append(a)
        terminated["__all__"] = bool(hn["all_terminated"][0])                     # :164-167
        truncated["__all__"] = bool(hn["all_truncated"][0])
        self.agents = [] if (terminated["__all__"] or truncated["__all__"]) else next_active   # :169-172
        return obs, rewards, terminated, truncated, infos


class SingleDroneEnv(_EngineBacked, _GymEnv):
    """3D continuous-control single-drone environment (reference single_drone_env.py:12)."""

    metadata = {"render_modes": []}

    def __init__(self, config: dict[str, Any] | None = None, device=None):
        super().__init__()
        self.cfg = DroneEnvConfig.from_dict(config)                               # :30
        self._obs_dim = 9 + self.cfg.sensed_obstacles * 4                         # :33
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=(self._obs_dim,), dtype=np.float32)
        self.action_space = Box(low=-1.0, high=1.0, shape=(3,), dtype=np.float32)
        raw = {k: v for k, v in (config or {}).items() if k != "num_drones"}
        self._make_engine(raw, "single", device)

    @property
    def position(self) -> np.ndarray:
        return self._engine.positions[0, 0].cpu().numpy()

    @property
    def velocity(self) -> np.ndarray:
        return self._engine.velocities[0, 0].cpu().numpy()

    def reset(self, *, seed: int | None = None, options: dict[str, Any] | None = None):
        """single_drone_env.py:53-71."""
        super().reset(seed=seed)
        if seed is not None:
            self._engine.seed(np.asarray([seed], dtype=np.uint64))
        self._engine.reset()
        obs, dist, _ = self._fetch_reset_outputs()
        return obs[0].copy(), {"distance_to_goal": float(dist[0])}

    def step(self, action: np.ndarray):
        """single_drone_env.py:73-111."""
        hn = self._hn
        hn["actions"][0, 0] = np.asarray(action, dtype=np.float32).reshape(3)
        self._engine.step_host(None, auto_reset=False)
        info = {
            "distance_to_goal": float(hn["dist"][0, 0]),
            "reached_goal": bool(hn["reached"][0, 0]),
            "collision": bool(hn["collision"][0, 0]),
        }
        return (hn["obs"][0, 0].copy(), float(hn["reward64"][0, 0]),
                bool(hn["terminated"][0, 0]), bool(hn["truncated"][0, 0]), info)


class DronePhysicsEnv(_EngineBacked, _MultiAgentEnv):
    """`DronePhysicsEnv` (reference drone_physics_env.py:22) with the drones as point masses.

    Same constructor / reset / step contract: obs and infos for EVERY drone on every step
    (:365-366), rewards for the drones of `self.agents`, one terminated / truncated pair shared by all
    drones (:401-417), `self.agents = []` once the episode is over, `set_goal` (:265-277).  The force
    / drag / gravity / speed-clamp model around PyBullet's solver is restated (24 sub-steps of 1/240 s,
    thrust `a * max_accel + (0, 0, 9.5)`, gravity -9.81, per-episode linear damping 0.5 * U(0.8, 1.2));
    the rigid-body solver itself is not (PyBullet is a third-party dependency that is neither vendored
    nor installed): contacts are sphere / plane tests that end the episode, as any contact does in the
    reference.  Parity unpinned."""

    def __init__(self, config: dict[str, Any] | None = None, device=None):
        super().__init__()
        cfg_dict = dict(config or {})
        self.num_drones = int(cfg_dict.get("num_drones", 3))                      # :115
        env_cfg = {k: v for k, v in cfg_dict.items() if k not in ("num_drones", "gui")}
        self.cfg = DroneEnvConfig.from_dict(env_cfg)                              # :118
        self.gui = False                                                          # (no GUI: :119 is ignored)
        self.agent_ids = [f"drone_{i}" for i in range(self.num_drones)]           # :164
        self.agent_id_to_index = {a: i for i, a in enumerate(self.agent_ids)}
        self.agents = list(self.agent_ids)
        self._obs_dim = 9 + self.cfg.neighbor_k * 4 + self.cfg.sensed_obstacles * 4   # :152-156
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=(self._obs_dim,), dtype=np.float32)
        self.action_space = Box(low=-1.0, high=1.0, shape=(3,), dtype=np.float32)
        self._make_engine({**env_cfg, "num_drones": self.num_drones}, "physics", device)

    @property
    def positions(self) -> np.ndarray:
        return self._engine.positions[0].cpu().numpy()

    @property
    def velocities(self) -> np.ndarray:
        return self._engine.velocities[0].cpu().numpy()

    def set_goal(self, new_pos):
        """drone_physics_env.py:265-277 (interactive dashboard)."""
        import torch

        self._engine.goal4[0, :3] = torch.as_tensor(np.asarray(new_pos, dtype=np.float32), device=self._engine.device)

    def reset(self, *, seed: int | None = None, options: dict[str, Any] | None = None):
        """drone_physics_env.py:174-263.  With seed=None the reference re-seeds from OS entropy; so do we."""
        self._engine.seed(np.asarray([seed if seed is not None else _entropy_seed()], dtype=np.uint64))
        self.agents = list(self.agent_ids)
        self._engine.reset()
        obs, dist, _ = self._fetch_reset_outputs()
        observations = {a: obs[i].copy() for i, a in enumerate(self.agent_ids)}
        infos = {a: {"distance_to_goal": float(dist[i]), "reached_goal": False, "collision": False}   # :257-261
                 for i, a in enumerate(self.agent_ids)}
        return observations, infos

    def step(self, action_dict: dict[str, np.ndarray]):
        """drone_physics_env.py:279-419."""
        active = list(self.agents)
        hn = self._hn
        act = hn["actions"]
        act[...] = 0.0
        for a, raw in action_dict.items():                                        # :325 (no clip, :336)
            act[0, self.agent_id_to_index[a]] = np.asarray(raw, dtype=np.float32).reshape(3)
        if not active:
            # the reference keeps integrating a finished episode and answers terminated["__all__"] = True with
            # empty rewards (:374, :398-411); the batched contract parks the env instead
            flags = {a: True for a in self.agent_ids}
            return {}, {}, {**flags, "__all__": True}, {**{a: False for a in self.agent_ids}, "__all__": False}, {}
        self._engine.step_host(None, auto_reset=False)
        gs = hn["global_state"][0]
        obs_rows, dist = hn["obs"][0], hn["dist"][0].tolist()
        reached, collision = hn["reached"][0].tolist(), hn["collision"][0].tolist()
        rew, term, trunc = hn["reward64"][0].tolist(), hn["terminated"][0].tolist(), hn["truncated"][0].tolist()
        obs = {a: obs_rows[i].copy() for i, a in enumerate(self.agent_ids)}
        infos = {a: {"global_state": gs.copy(), "distance_to_goal": dist[i],
                     "reached_goal": bool(reached[i]), "collision": bool(collision[i])}
                 for i, a in enumerate(self.agent_ids)}
        rewards = {a: rew[self.agent_id_to_index[a]] for a in active}
        terminated = {a: bool(term[i]) for i, a in enumerate(self.agent_ids)}
        truncated = {a: bool(trunc[i]) for i, a in enumerate(self.agent_ids)}
        terminated["__all__"] = bool(hn["all_terminated"][0])
        truncated["__all__"] = bool(hn["all_truncated"][0])
        if terminated["__all__"] or truncated["__all__"]:
            self.agents = []                                                      # :411
        return obs, rewards, terminated, truncated, infos


class _LazyEnvDicts(Mapping):
    """{env_id: per-agent dict} whose values are built on first access: `poll()` itself does no per-agent Python
    work, so a sampler that only touches some envs (or only some of the five dicts) pays only for those."""

    def __init__(self, num_envs: int, build: Callable[[int], dict]):
        self._n, self._build, self._cache = num_envs, build, {}

    def __getitem__(self, e):
        if not (0 <= e < self._n):
            raise KeyError(e)
        v = self._cache.get(e)
        if v is None:
            v = self._cache[e] = self._build(e)
        return v

    def __iter__(self):
        return iter(range(self._n))

    def __len__(self):
        return self._n


class VectorSwarmEnv:
    """E `DroneSwarmEnv` instances behind RLlib's `BaseEnv` polling interface
    (`ray.rllib.env.base_env.BaseEnv`: `poll`, `send_actions`, `try_reset`, `get_sub_environments`),
    stepped by ONE engine call per step instead of E Python envs (the reference runs
    one env per rollout worker, `training/config_builders.py:19-23`).  ray is not required.

    Per step there is one C-ABI call (pinned actions -> H2D -> step -> D2H of the output block into pinned host
    memory) and O(E) vectorised numpy bookkeeping; the dicts `poll()` returns are LAZY (built per env on first
    access, by the reference's population rules, drone_swarm_env.py:141-172).  `step_batch` / `last_batch` hand out
    the same step as one dict of [E, ...] arrays -- including the `global_state` column [E, 6N+3] that the
    reference's `GlobalStateCallback` (`training/callbacks.py:51-57`) stacks from per-agent infos in Python.  An
    env whose episode ended is reset inside the same engine step (auto-reset); its first observation is handed
    out by `try_reset(env_id)`, as RLlib expects."""

    _FIELDS = ("obs", "dist", "global_state", "reward64", "terminated", "truncated", "reached", "collision",
               "obs_valid", "all_terminated", "all_truncated")

    def __init__(self, num_envs: int, config: dict[str, Any] | None = None, device=None, base_seed: int | None = None):
        from .engine import SwarmEngine

        cfg_dict = dict(config or {})
        self.num_envs = int(num_envs)
        self.num_drones = int(cfg_dict.get("num_drones", 3))
        self.cfg = DroneEnvConfig.from_dict({k: v for k, v in cfg_dict.items() if k != "num_drones"})
        self.agent_ids = [f"drone_{i}" for i in range(self.num_drones)]
        self._obs_dim = 9 + self.cfg.neighbor_k * 4 + self.cfg.sensed_obstacles * 4
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=(self._obs_dim,), dtype=np.float32)
        self.action_space = Box(low=-1.0, high=1.0, shape=(3,), dtype=np.float32)
        self._engine = SwarmEngine(self.num_envs, cfg_dict, kind="swarm", device=device or "cuda", global_state=True,
                                   reward64=True)
        seed0 = base_seed if base_seed is not None else (self.cfg.seed if self.cfg.seed is not None else _entropy_seed())
        self._engine.seed(np.uint64(seed0) + np.arange(self.num_envs, dtype=np.uint64))   # env e: seed0 + e
        self._engine.reset()
        self._hn = {k: v.numpy() for k, v in self._engine.host_buffers().items()}   # pinned host views, made once
        self._active = np.ones((self.num_envs, self.num_drones), bool)   # membership in each env's .agents
        self._pending = None       # (was_active) of the last step, not yet polled
        self._fresh = set(range(self.num_envs))   # envs whose (reset) observation has not been handed out
        self._have_host = False

    def get_sub_environments(self):
        return []

    # ------------------------------------------------------------------ batched (array) interface
    def step_batch(self, actions) -> dict[str, np.ndarray]:
        """One step of every env from an [E,N,3] float32 array (rows of drones outside `.agents` are ignored) ->
        dict of pinned-host numpy VIEWS, valid until the next step: obs [E,N,D], reward [E,N] (float64),
        terminated / truncated / reached / collision / obs_valid [E,N] u8, all_terminated / all_truncated [E] u8,
        dist [E,N], global_state [E,6N+3], plus `was_active` [E,N] bool (the drones that were stepped)."""
        act = self._hn["actions"]
        act[...] = np.asarray(actions, dtype=np.float32).reshape(act.shape)
        return self._step_from_host_actions()

    def _step_from_host_actions(self):
        was_active = self._active.copy()
        self._engine.step_host(None, auto_reset=True)
        self._have_host = True
        hn = self._hn
        done = (hn["all_terminated"] | hn["all_truncated"]).astype(bool)
        keep = hn["obs_valid"].astype(bool) & was_active & ~done[:, None]
        self._active = np.where(done[:, None], True, keep)     # a finished env has already been reset by the engine
        self._fresh.update(np.flatnonzero(done).tolist())
        self._pending = (was_active, done, keep)
        return self.last_batch()

    def last_batch(self) -> dict[str, np.ndarray]:
        hn = self._hn
        out = {k: hn[k] for k in self._FIELDS}
        out["reward"] = hn["reward64"]
        if self._pending is not None:
            out["was_active"] = self._pending[0]
        return out

    def global_state_column(self) -> np.ndarray:
        """[E, 6N+3] float32: the centralized critic's input for every env (what `GlobalStateCallback` assembles)."""
        return self._hn["global_state"]

    # ------------------------------------------------------------------ RLlib BaseEnv-shaped interface
    def _ensure_host(self):
        if not self._have_host:
            self._engine.fetch_outputs()
            self._have_host = True

    def _reset_dicts(self, e):
        hn = self._hn
        obs_rows, dist, gs = hn["obs"][e], hn["dist"][e].tolist(), hn["global_state"][e]
        obs = {a: obs_rows[i].copy() for i, a in enumerate(self.agent_ids)}
        infos = {a: {"distance_to_goal": dist[i], "global_state": gs.copy()} for i, a in enumerate(self.agent_ids)}
        return obs, infos

    def poll(self):
        """-> (obs, rewards, terminateds, truncateds, infos, off_policy_actions), each {env_id: {agent_id: ...}}
        (lazy mappings over all env ids)."""
        self._ensure_host()
        E, hn, ids = self.num_envs, self._hn, self.agent_ids
        if self._pending is None:           # first poll: the reset observations
            self._fresh.clear()
            cache: dict[int, tuple] = {}

            def both(e):
                if e not in cache:
                    cache[e] = self._reset_dicts(e)
                return cache[e]
            return (_LazyEnvDicts(E, lambda e: both(e)[0]), _LazyEnvDicts(E, lambda e: {}),
                    _LazyEnvDicts(E, lambda e: {"__all__": False}), _LazyEnvDicts(E, lambda e: {"__all__": False}),
                    _LazyEnvDicts(E, lambda e: both(e)[1]), {})
        was_active, done, keep = self._pending
        self._pending = None
        # (the views below alias the pinned block: they stay valid until the next send_actions / step_batch)
        rew64, term, trunc = hn["reward64"], hn["terminated"], hn["truncated"]

        def obs_of(e):
            rows = hn["obs"][e]
            return {ids[i]: rows[i].copy() for i in np.flatnonzero(keep[e])}

        def rew_of(e):
            r = rew64[e].tolist()
            return {ids[i]: r[i] for i in np.flatnonzero(was_active[e])}

        def flags_of(arr, all_arr):
            def f(e):
                v = arr[e].tolist()
                d = {ids[i]: bool(v[i]) for i in np.flatnonzero(was_active[e])}
                d["__all__"] = bool(all_arr[e])
                return d
            return f

        def infos_of(e):
            dist, reached, col, gs = hn["dist"][e].tolist(), hn["reached"][e].tolist(), hn["collision"][e].tolist(), hn["global_state"][e]
            return {ids[i]: {"distance_to_goal": dist[i], "reached_goal": bool(reached[i]), "collision": bool(col[i]),
                             "global_state": gs.copy()} for i in np.flatnonzero(keep[e])}

        return (_LazyEnvDicts(E, obs_of), _LazyEnvDicts(E, rew_of), _LazyEnvDicts(E, flags_of(term, hn["all_terminated"])),
                _LazyEnvDicts(E, flags_of(trunc, hn["all_truncated"])), _LazyEnvDicts(E, infos_of), {})

    def send_actions(self, action_dict):
        """{env_id: {agent_id: action(3,)}}; agents without an action get zeros (drone_swarm_env.py:104)."""
        act = self._hn["actions"]
        act[...] = 0.0
        for e, per_agent in action_dict.items():
            row = act[e]
            for a, v in per_agent.items():
                row[int(a.rsplit("_", 1)[1])] = v
        self._step_from_host_actions()

    def try_reset(self, env_id=None, *, seed=None, options=None):
        """First observation of env `env_id`'s new episode -> ({env_id: obs}, {env_id: infos})."""
        if env_id not in self._fresh:    # explicit reset of a running env
            import torch

            mask = torch.zeros(self.num_envs, dtype=torch.uint8)
            mask[env_id] = 1
            if seed is not None:
                seeds = np.zeros(self.num_envs, np.uint64)
                seeds[env_id] = seed
                self._engine.seed(seeds, mask)
            self._engine.reset(mask)
            self._have_host = False
            self._active[env_id, :] = True
        self._ensure_host()
        self._fresh.discard(env_id)
        obs, infos = self._reset_dicts(env_id)
        return {env_id: obs}, {env_id: infos}

    def stop(self):
        self._engine.close()

    close = stop


def make_env_creator(kind: str = "swarm", device=None):
    """Env creator for `ray.tune.registry.register_env(name, creator)`:
    the drop-in for `lambda cfg: DroneSwarmEnv(cfg)` (reference scripts/train_multi_agent.py:96)."""
    cls = {"swarm": DroneSwarmEnv, "single": SingleDroneEnv, "physics": DronePhysicsEnv}[kind]
    return lambda cfg: cls(dict(cfg) if cfg is not None else None, device=device)
