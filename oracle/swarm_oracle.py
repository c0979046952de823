"""ctypes front-end of the C parity oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package never does.  See the header of
oracle/swarm_oracle.c for what the oracle restates and how it is pinned.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libswarm_oracle.so")

# mirror of DroneEnvConfig defaults (reference envs/common.py:7-25)
DEFAULTS = dict(
    world_size=20.0, dt=0.1, max_steps=400, max_speed=4.0, max_accel=2.0, collision_radius=0.5,
    goal_radius=0.8, num_obstacles=8, sensed_obstacles=4, neighbor_k=3, obstacle_radius=0.8,
    desired_spacing=2.5, reward_progress_scale=2.0, reward_goal=25.0, reward_collision=-25.0,
    reward_formation_scale=0.15,
)


class OracleConfig(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "world_size", "dt", "max_speed", "max_accel", "collision_radius", "goal_radius",
        "obstacle_radius", "desired_spacing", "reward_progress_scale", "reward_goal",
        "reward_collision", "reward_formation_scale")] + [(n, C.c_int32) for n in (
        "max_steps", "num_obstacles", "sensed_obstacles", "neighbor_k", "num_drones", "env_kind",
        "norm_mode", "reserved", "dr_enabled", "dr_pad")] + [("dr_seed", C.c_uint64), ("env_index_base", C.c_int64),
                                                             ("dr_lo", C.c_double * 6), ("dr_span", C.c_double * 6)] + [
        (n, C.c_double) for n in ("dr_std_thrust", "dr_std_pos", "dr_std_vel", "dr_std_obst")] + [
        ("dr_delay_count", C.c_int32), ("dr_delay_hist", C.c_int32), ("dr_delay_values", C.c_int32 * 4),
        ("dr_delay_cum", C.c_double * 4)]

DR_RANGE_KEYS = ("mass_scale", "max_accel_scale", "max_speed_scale", "dt_scale", "obstacle_radius_scale",
                 "world_size_scale")
DR_STD_KEYS = ("thrust_noise_std", "position_noise_std", "velocity_noise_std", "obstacle_distance_noise_std")


class OracleBatch(C.Structure):
    _fields_ = [("num_envs", C.c_int32), ("pad", C.c_int32)] + [(n, C.c_void_p) for n in (
        "positions", "velocities", "goal", "obstacles", "step_count", "active", "rng", "obs", "reward",
        "dist", "terminated", "truncated", "reached", "collision", "obs_valid", "all_terminated",
        "all_truncated", "global_state", "dr_params", "damp", "act_hist")]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "swarm_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_seed.argtypes = [C.c_uint64, C.c_void_p]
        _lib.oracle_pcg64_next.argtypes = [C.c_void_p]
        _lib.oracle_pcg64_next.restype = C.c_uint64
        _lib.oracle_obs_dim.argtypes = [C.POINTER(OracleConfig)]
        _lib.oracle_obs_dim.restype = C.c_int
        _lib.oracle_seed_batch.argtypes = [C.POINTER(OracleBatch), C.c_void_p]
        _lib.oracle_reset_batch.argtypes = [C.POINTER(OracleConfig), C.POINTER(OracleBatch), C.c_void_p]
        _lib.oracle_observe_batch.argtypes = [C.POINTER(OracleConfig), C.POINTER(OracleBatch)]
        _lib.oracle_step_batch.argtypes = [C.POINTER(OracleConfig), C.POINTER(OracleBatch), C.c_void_p,
                                           C.c_int, C.c_int]
    return _lib


class OracleSwarm:
    """A batch of E reference-equivalent envs on the CPU.

    `kind` = "swarm" (DroneSwarmEnv) or "single" (SingleDroneEnv).  `config` takes the
    reference's constructor dict keys (unknown keys dropped, like DroneEnvConfig.from_dict).
    """

    def __init__(self, num_envs: int, config: dict | None = None, kind: str = "swarm", norm_mode: int = 0,
                 dr: dict | None = None, dr_seed: int = 0, env_index_base: int = 0):
        """dr: flat dict {<range key>: (min, max), <std key>: sigma} (engine semantics, not reference)."""
        cfg = dict(DEFAULTS)
        raw = dict(config or {})
        self.num_drones = int(raw.pop("num_drones", 3)) if kind in ("swarm", "physics") else 1
        raw.pop("seed", None)
        cfg.update({k: v for k, v in raw.items() if k in DEFAULTS})
        self.cfg = cfg
        self.kind = kind
        self.E, self.N, self.M = int(num_envs), self.num_drones, int(cfg["num_obstacles"])
        self.K, self.S = int(cfg["neighbor_k"]), int(cfg["sensed_obstacles"])
        c = OracleConfig()
        for k in DEFAULTS:
            setattr(c, k, cfg[k])
        c.num_drones, c.env_kind, c.norm_mode = self.N, {"single": 0, "swarm": 1, "physics": 2}[kind], norm_mode
        if dr:
            c.dr_enabled, c.dr_seed, c.env_index_base = 1, int(dr_seed), int(env_index_base)
            for k, name in enumerate(DR_RANGE_KEYS):
                lo, hi = dr.get(name, (1.0, 1.0))
                c.dr_lo[k], c.dr_span[k] = float(lo), float(hi) - float(lo)
            c.dr_std_thrust, c.dr_std_pos, c.dr_std_vel, c.dr_std_obst = (float(dr.get(n, 0.0)) for n in DR_STD_KEYS)
            delay = dr.get("control_delay_steps")   # ((values...), (probs...)), engine semantics
            if delay and any(int(x) for x in delay[0]):
                vals, probs = [int(x) for x in delay[0]], [float(x) for x in delay[1]]
                c.dr_delay_count, c.dr_delay_hist = len(vals), max(vals)
                cum = 0.0
                for k, (x, pr) in enumerate(zip(vals, probs)):
                    cum += pr
                    c.dr_delay_values[k], c.dr_delay_cum[k] = x, cum
        self._c = c
        self.D = lib().oracle_obs_dim(C.byref(c))
        E, N, M, D = self.E, self.N, self.M, self.D
        self.positions = np.zeros((E, N, 3), np.float32)
        self.velocities = np.zeros((E, N, 3), np.float32)
        self.goal = np.zeros((E, 3), np.float32)
        self.obstacles = np.zeros((E, M, 3), np.float32)
        self.step_count = np.zeros(E, np.int32)
        self.active = np.ones((E, N), np.uint8)
        self.rng = np.zeros((E, 4), np.uint64)
        self.obs = np.zeros((E, N, D), np.float32)
        self.reward = np.zeros((E, N), np.float64)
        self.dist = np.zeros((E, N), np.float32)
        self.terminated = np.zeros((E, N), np.uint8)
        self.truncated = np.zeros((E, N), np.uint8)
        self.reached = np.zeros((E, N), np.uint8)
        self.collision = np.zeros((E, N), np.uint8)
        self.obs_valid = np.zeros((E, N), np.uint8)
        self.all_terminated = np.zeros(E, np.uint8)
        self.all_truncated = np.zeros(E, np.uint8)
        self.global_state = np.zeros((E, 6 * N + 3), np.float32)
        self.dr_params = np.zeros((E, 8), np.float32)
        self.damp = np.ones((E, N), np.float32)   # physics env: per-drone sub-step velocity factor
        self.act_hist = np.zeros((E, max(int(c.dr_delay_hist), 1), N, 3), np.float32)   # control delay: command ring
        b = OracleBatch()
        b.num_envs = E
        for name, _ in OracleBatch._fields_[2:]:
            setattr(b, name, getattr(self, name).ctypes.data)
        self._b = b

    def seed(self, seeds):
        seeds = np.ascontiguousarray(np.broadcast_to(np.asarray(seeds, np.uint64), (self.E,)))
        lib().oracle_seed_batch(C.byref(self._b), seeds.ctypes.data)

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().oracle_reset_batch(C.byref(self._c), C.byref(self._b), None if m is None else m.ctypes.data)

    def observe(self):
        lib().oracle_observe_batch(C.byref(self._c), C.byref(self._b))

    def step(self, actions, auto_reset: bool = False, num_threads: int = 1):
        a = np.ascontiguousarray(actions, np.float32).reshape(self.E, self.N, 3)
        lib().oracle_step_batch(C.byref(self._c), C.byref(self._b), a.ctypes.data, int(auto_reset),
                                int(num_threads))
