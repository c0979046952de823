"""ctypes binding of libswarm_b200.so (the C ABI in include/swarm_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  If the shared object is
missing this module raises, loudly, with the build command.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SWARM_B200_LIB") or os.path.join(_PKG_DIR, "libswarm_b200.so")  # override: tuning builds

ABI_VERSION = 5
KIND_SINGLE, KIND_SWARM, KIND_PHYSICS = 0, 1, 2
MAX_DRONES, MAX_NEIGHBOR_K, MAX_SENSED = 128, 8, 8

STAT_NAMES = ("episodes", "success", "collision", "timeout", "length_sum", "return_sum", "agent_steps",
              "env_steps", "nan_actions")

_DOUBLES = ("world_size", "dt", "max_speed", "max_accel", "collision_radius", "goal_radius", "obstacle_radius",
            "desired_spacing", "reward_progress_scale", "reward_goal", "reward_collision",
            "reward_formation_scale")


DR_RANGES = ("dr_mass_scale", "dr_max_accel_scale", "dr_max_speed_scale", "dr_dt_scale",
             "dr_obstacle_radius_scale", "dr_world_size_scale")
DR_STDS = ("dr_thrust_noise_std", "dr_position_noise_std", "dr_velocity_noise_std",
           "dr_obstacle_distance_noise_std")


class SwarmConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("abi_version", "env_kind", "num_envs", "num_drones", "num_obstacles",
                                         "sensed_obstacles", "neighbor_k", "max_steps", "norm_mode", "device")] + \
               [(n, C.c_double) for n in _DOUBLES] + \
               [("dr_enabled", C.c_int32), ("dr_reserved", C.c_int32), ("dr_seed", C.c_uint64),
                ("env_index_base", C.c_int64)] + \
               [(n, C.c_double * 2) for n in DR_RANGES] + [(n, C.c_double) for n in DR_STDS] + \
               [("dr_delay_count", C.c_int32), ("dr_delay_reserved", C.c_int32), ("dr_delay_values", C.c_int32 * 4),
                ("dr_delay_probs", C.c_double * 4)]


class SwarmSizes(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("obs_dim", "state_dim", "pos4", "vel4", "goal4", "obst4", "step_count",
                                         "rng", "ep_return", "actions", "obs", "per_agent", "per_env",
                                         "global_state", "stats", "dr_params", "act_hist")]


BUFFER_FIELDS = ("pos4", "vel4", "goal4", "obst4", "step_count", "rng", "ep_return", "obs", "reward", "reward64",
                 "dist", "terminated", "truncated", "reached", "collision", "obs_valid", "all_terminated",
                 "all_truncated", "global_state", "episode_return", "episode_length", "stats", "dr_params", "act_hist")


class SwarmBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in BUFFER_FIELDS]


HOST_OUT_FIELDS = ("obs", "reward", "reward64", "dist", "terminated", "truncated", "reached", "collision", "obs_valid",
                   "all_terminated", "all_truncated", "global_state")


class SwarmHostOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in HOST_OUT_FIELDS] + \
               [("block_host", C.c_void_p), ("block_dev", C.c_void_p), ("block_bytes", C.c_int64),
                ("flags", C.c_void_p)]   # ABI 5: the five per-agent flag arrays packed into one byte (FLAG_* bits)


FLAG_TERMINATED, FLAG_TRUNCATED, FLAG_REACHED, FLAG_COLLISION, FLAG_OBS_VALID = 1, 2, 4, 8, 16
FLAG_FIELDS = ("terminated", "truncated", "reached", "collision", "obs_valid")   # bit k = FLAG_FIELDS[k]


EXPORTS = ("swarm_abi_version", "swarm_last_error", "swarm_create", "swarm_destroy", "swarm_query_sizes",
           "swarm_seed", "swarm_reset", "swarm_observe", "swarm_step", "swarm_step_many", "swarm_step_host",
           "swarm_launch_count",
           "swarm_dr_quantile_table")


class SwarmError(RuntimeError):
    pass


_lib = None


def load():
    """dlopen the in-tree library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension is the product and there is no fallback. "
            f"Build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            f"`make -C {os.path.join(_PKG_DIR, 'csrc')}`.")
    lib = C.CDLL(LIB_PATH)
    vp, i32 = C.c_void_p, C.c_int
    lib.swarm_abi_version.restype = i32
    lib.swarm_last_error.restype = C.c_char_p
    lib.swarm_create.argtypes = [C.POINTER(SwarmConfig), C.POINTER(vp)]
    lib.swarm_destroy.argtypes = [vp]
    lib.swarm_query_sizes.argtypes = [C.POINTER(SwarmConfig), C.POINTER(SwarmSizes)]
    lib.swarm_seed.argtypes = [vp, C.POINTER(SwarmBuffers), vp, vp, vp]
    lib.swarm_reset.argtypes = [vp, C.POINTER(SwarmBuffers), vp, vp]
    lib.swarm_observe.argtypes = [vp, C.POINTER(SwarmBuffers), vp]
    lib.swarm_step.argtypes = [vp, C.POINTER(SwarmBuffers), vp, i32, vp]
    lib.swarm_step_many.argtypes = [vp, C.POINTER(SwarmBuffers), vp, i32, i32, vp]
    lib.swarm_step_host.argtypes = [vp, C.POINTER(SwarmBuffers), vp, C.POINTER(SwarmHostOut), i32, vp]
    lib.swarm_launch_count.argtypes = [vp]
    lib.swarm_launch_count.restype = C.c_int64
    lib.swarm_dr_quantile_table.argtypes = [vp]
    lib.swarm_dr_quantile_table.restype = i32
    for name in ("swarm_create", "swarm_destroy", "swarm_query_sizes", "swarm_seed", "swarm_reset",
                 "swarm_observe", "swarm_step", "swarm_step_many", "swarm_step_host"):
        getattr(lib, name).restype = i32
    if lib.swarm_abi_version() != ABI_VERSION:
        raise ImportError(f"libswarm_b200.so ABI {lib.swarm_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().swarm_last_error().decode("utf-8", "replace")
        raise SwarmError(f"{what} failed with code {rc}: {msg}")
