"""SwarmEngine -- batched, device-resident drone-swarm env step (host side of the C ABI).

One engine = E independent instances of the reference's `DroneSwarmEnv` (kind="swarm",
reference src/swarm_marl/envs/drone_swarm_env.py:17), `SingleDroneEnv` (kind="single",
src/swarm_marl/envs/single_drone_env.py:12) or `DronePhysicsEnv` as a point mass (kind="physics",
src/swarm_marl/envs/drone_physics_env.py:22; PyBullet is not vendored: parity unpinned) living on one GPU.  PyTorch is used only to own
device memory and streams; every computation is a hand-written sm_100a kernel behind
include/swarm_b200.h.  There is no fallback path: without the built library this module
raises on import of the binding.
"""
from __future__ import annotations

import ctypes as C
from typing import Any

import numpy as np
import torch

from . import _abi
from .config import DroneEnvConfig, flatten_domain_randomization


class SwarmEngine:
    """E env instances stepped by one fused kernel launch.

    Tensors (all on `device`, leading env axis):
      pos4 [E,N,4] f32 (x,y,z,alive)   vel4 [E,N,4]   goal4 [E,4]   obst4 [E,M,4]
      step_count [E] i32   rng [E,4] i64 (numpy PCG64 words)   ep_return [E] f32
      obs [E,N,D] f32   reward [E,N] f32   dist [E,N] f32
      terminated / truncated / reached / collision / obs_valid [E,N] u8
      all_terminated / all_truncated [E] u8   global_state [E,6N+3] f32
      episode_return [E] f32 / episode_length [E] i32 (episode that ended on this step)
    """

    def __init__(self, num_envs: int, config: dict[str, Any] | None = None, kind: str = "swarm",
                 device: str | torch.device = "cuda", global_state: bool = True, reward64: bool = False,
                 norm_mode: int = 0, domain_randomization: dict[str, Any] | None = None, dr_seed: int = 0,
                 env_index_base: int = 0):
        """domain_randomization: None / {} = off (bit-identical to the reference).  Otherwise either the
        reference's `configs/domain_randomization_v1.yaml` document (as loaded: `randomization.dynamics.
        mass_scale.{min,max}` ...; honoured even though its `enabled` flag is false there, pass None to
        disable) or the flat form {"mass_scale": (lo, hi), ..., "thrust_noise_std": sigma, ...} -- see
        `config.flatten_domain_randomization`.  `env_index_base`: global index of env 0 (sharded runs),
        so the randomisation streams of an env do not depend on how the batch is split."""
        if kind not in ("swarm", "single", "physics"):
            raise ValueError(f"kind must be 'swarm', 'single' or 'physics', got {kind!r}")
        self._lib = _abi.load()
        raw = dict(config or {})
        self.kind = kind
        # drone_swarm_env.py:32-34: num_drones is popped, the rest goes through DroneEnvConfig.from_dict
        self.num_drones = int(raw.pop("num_drones", 3)) if kind in ("swarm", "physics") else 1
        raw.pop("num_drones", None)
        self.cfg = DroneEnvConfig.from_dict(raw)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("SwarmEngine runs on CUDA devices only (there is no CPU path)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.E, self.N, self.M = int(num_envs), self.num_drones, int(self.cfg.num_obstacles)
        self.K = int(self.cfg.neighbor_k) if kind != "single" else 0
        self.S = int(self.cfg.sensed_obstacles)

        c = _abi.SwarmConfig()
        c.abi_version = _abi.ABI_VERSION
        c.env_kind = {"swarm": _abi.KIND_SWARM, "single": _abi.KIND_SINGLE, "physics": _abi.KIND_PHYSICS}[kind]
        c.num_envs, c.num_drones, c.num_obstacles = self.E, self.N, self.M
        c.sensed_obstacles, c.neighbor_k = self.S, int(self.cfg.neighbor_k)
        c.max_steps, c.norm_mode, c.device = int(self.cfg.max_steps), int(norm_mode), self.device.index
        for name in _abi._DOUBLES:
            setattr(c, name, float(getattr(self.cfg, name)))
        self.dr = flatten_domain_randomization(domain_randomization)
        if self.dr:
            c.dr_enabled, c.dr_seed, c.env_index_base = 1, int(dr_seed) & (2**64 - 1), int(env_index_base)
            for name in _abi.DR_RANGES:
                lo, hi = self.dr.get(name[3:], (1.0, 1.0))
                getattr(c, name)[0], getattr(c, name)[1] = float(lo), float(hi)
            for name in _abi.DR_STDS:
                setattr(c, name, float(self.dr.get(name[3:], 0.0)))
            delay = self.dr.get("control_delay_steps")
            if delay:
                c.dr_delay_count = len(delay[0])
                for k, (val, pr) in enumerate(zip(*delay)):
                    c.dr_delay_values[k], c.dr_delay_probs[k] = int(val), float(pr)
        self._c = c
        sz = _abi.SwarmSizes()
        _abi.check(self._lib.swarm_query_sizes(C.byref(c), C.byref(sz)), "swarm_query_sizes")
        self.D, self.R = int(sz.obs_dim), int(sz.state_dim)
        self._handle = C.c_void_p()
        _abi.check(self._lib.swarm_create(C.byref(c), C.byref(self._handle)), "swarm_create")

        E, N, M, D, R = self.E, self.N, self.M, self.D, self.R
        dev = self.device
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)  # noqa: E731
        self.pos4 = z((E, N, 4), torch.float32)
        self.vel4 = z((E, N, 4), torch.float32)
        self.goal4 = z((E, 4), torch.float32)
        self._obst4_storage = z((E, max(M, 1), 4), torch.float32)  # never empty, even when M == 0
        self.obst4 = self._obst4_storage[:, :M]
        self.step_count = z((E,), torch.int32)
        self.rng = z((E, 4), torch.int64)
        self.ep_return = z((E,), torch.float32)
        # every host-visible output is a view into ONE device allocation, so that a small batch (the E = 1 facade
        # envs) comes back with a single device->host copy (SwarmHostOut block mode)
        spec = [("obs", (E, N, D), torch.float32), ("reward", (E, N), torch.float32)]
        if reward64:
            spec.append(("reward64", (E, N), torch.float64))
        spec += [("dist", (E, N), torch.float32)]
        spec += [(n, (E, N), torch.uint8) for n in ("terminated", "truncated", "reached", "collision", "obs_valid")]
        spec += [("all_terminated", (E,), torch.uint8), ("all_truncated", (E,), torch.uint8)]
        if global_state:
            spec.append(("global_state", (E, R), torch.float32))
        self._out_layout, off = [], 0
        for name, shape, dt in spec:
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
            self._out_layout.append((name, shape, dt, off, nbytes))
            off += (nbytes + 255) // 256 * 256
        self._out_block = z((max(off, 256),), torch.uint8)
        self.reward64 = None
        self.global_state = None
        for name, shape, dt, o, nbytes in self._out_layout:
            setattr(self, name, self._out_block[o:o + nbytes].view(dt).view(shape))
        self.episode_return = z((E,), torch.float32)
        self.episode_length = z((E,), torch.int32)
        self.stats_words = z((len(_abi.STAT_NAMES),), torch.int64)
        # per-env per-episode dynamics constants of the domain randomisation (written by reset)
        self.dr_params = z((E, 8), torch.float32) if self.dr else None
        # control delay: ring of the last H submitted commands per env
        self.act_hist = z((E, int(sz.act_hist) // max(E * N * 3, 1), N, 3), torch.float32) if int(sz.act_hist) else None
        self.pos4[..., 3] = 1.0
        self._actions_dev = None
        self._bufs = self._make_buffers()
        self._bufs_ref = C.byref(self._bufs)
        self._n_actions = E * N * 3
        self._step_result = (self.obs, self.reward, self.terminated, self.truncated)
        self._host = None

    # ------------------------------------------------------------------ plumbing
    def _make_buffers(self) -> _abi.SwarmBuffers:
        b = _abi.SwarmBuffers()
        for name in _abi.BUFFER_FIELDS:
            t = self._obst4_storage if name == "obst4" else (
                self.stats_words if name == "stats" else getattr(self, name))
            setattr(b, name, None if t is None else t.data_ptr())
        return b

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self._lib.swarm_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ reference-shaped views
    @property
    def positions(self) -> torch.Tensor:   # DroneSwarmEnv.positions (drone_swarm_env.py:59)
        return self.pos4[..., :3]

    @property
    def velocities(self) -> torch.Tensor:  # .velocities (:60)
        return self.vel4[..., :3]

    @property
    def goal(self) -> torch.Tensor:        # .goal (:61)
        return self.goal4[:, :3]

    @property
    def obstacles(self) -> torch.Tensor:   # .obstacles (:62)
        return self.obst4[..., :3]

    @property
    def damping(self) -> torch.Tensor:     # kind="physics": per-drone velocity factor of one 1/240 s sub-step
        return self.vel4[..., 3]

    @property
    def alive(self) -> torch.Tensor:       # membership in .agents (:39, :169-172)
        return self.pos4[..., 3] != 0

    @property
    def launch_count(self) -> int:
        return int(self._lib.swarm_launch_count(self._handle))

    # ------------------------------------------------------------------ API
    def seed(self, seeds, env_mask: torch.Tensor | None = None):
        """`np.random.default_rng(seed)` per env (drone_swarm_env.py:35 / reset(seed=...) :66-67)."""
        if torch.is_tensor(seeds):
            s = seeds.to(device=self.device, dtype=torch.int64).contiguous()
        else:
            arr = np.array(np.broadcast_to(np.asarray(seeds, dtype=np.uint64), (self.E,)))  # writable copy
            s = torch.from_numpy(arr.view(np.int64)).to(self.device)
        if s.numel() != self.E:
            raise ValueError(f"need {self.E} seeds, got {s.numel()}")
        m = self._mask(env_mask)
        _abi.check(self._lib.swarm_seed(self._handle, C.byref(self._bufs), s.data_ptr(),
                                        None if m is None else m.data_ptr(), self._stream()), "swarm_seed")
        self._keep = (s, m)

    def _mask(self, env_mask):
        if env_mask is None:
            return None
        m = torch.as_tensor(env_mask, device=self.device).to(torch.uint8).contiguous()
        if m.numel() != self.E:
            raise ValueError(f"env_mask needs {self.E} entries")
        return m

    def reset(self, env_mask: torch.Tensor | None = None):
        """`env.reset()` for the masked envs (all when None): drone_swarm_env.py:65-90."""
        m = self._mask(env_mask)
        _abi.check(self._lib.swarm_reset(self._handle, C.byref(self._bufs), None if m is None else m.data_ptr(),
                                         self._stream()), "swarm_reset")
        self._keep = m
        return self.obs

    def observe(self):
        """Recompute obs / dist / global_state from the (externally written) state."""
        _abi.check(self._lib.swarm_observe(self._handle, C.byref(self._bufs), self._stream()), "swarm_observe")
        return self.obs

    def step(self, actions: torch.Tensor, auto_reset: bool = True):
        """`env.step(action_dict)` for every env: drone_swarm_env.py:92-174.

        actions: [E,N,3] float32 CUDA tensor (a drone without an action gets a zero row)."""
        if not (torch.is_tensor(actions) and actions.is_cuda):
            raise TypeError("actions must be a CUDA tensor (use step_host for host buffers)")
        a = actions
        if a.dtype is not torch.float32 or not a.is_contiguous():   # (small batches are host-bound: keep the fast path lean)
            a = a.to(dtype=torch.float32).contiguous()
        if a.numel() != self._n_actions:
            raise ValueError(f"actions must have shape [{self.E},{self.N},3]")
        rc = self._lib.swarm_step(self._handle, self._bufs_ref, a.data_ptr(), 1 if auto_reset else 0, self._stream())
        if rc != 0:
            _abi.check(rc, "swarm_step")
        self._keep = a
        return self._step_result

    def step_many(self, actions: torch.Tensor, auto_reset: bool = True):
        """T consecutive `step` calls with ONE host call: `actions` is a [T,E,N,3] float32 CUDA tensor, step t
        applies actions[t].  Outputs describe the last step; statistics accumulate.  For small batches (BASELINE's
        4096 x 8 drones) the per-call host cost, not the device, bounds a Python loop over `step`."""
        if not (torch.is_tensor(actions) and actions.is_cuda):
            raise TypeError("actions must be a CUDA tensor")
        a = actions
        if a.dtype is not torch.float32 or not a.is_contiguous():
            a = a.to(dtype=torch.float32).contiguous()
        if a.dim() != 4 or a.numel() != a.shape[0] * self._n_actions:
            raise ValueError(f"actions must have shape [T,{self.E},{self.N},3]")
        rc = self._lib.swarm_step_many(self._handle, self._bufs_ref, a.data_ptr(), int(a.shape[0]),
                                       1 if auto_reset else 0, self._stream())
        if rc != 0:
            _abi.check(rc, "swarm_step_many")
        self._keep = a
        return self._step_result

    def capture_steps(self, actions: torch.Tensor, auto_reset: bool = True) -> "StepGraph":
        """Capture `step_many(actions)` into a CUDA graph (the library keeps no per-launch host state, so the
        graph can be replayed any number of times).  `actions` ([T,E,N,3] float32 CUDA) is the graph's INPUT
        buffer: write the next T actions into it (e.g. `actions.copy_(...)`, or let a policy kernel fill
        actions[t] between replays of single-step graphs) and call `.replay()`."""
        return StepGraph(self, actions, auto_reset)

    def set_state(self, positions=None, velocities=None, goal=None, obstacles=None, alive=None, step_count=None,
                  observe: bool = True):
        """State injection (parity runs): arrays shaped like the reference attributes + env axis."""
        dev = self.device
        if positions is not None:
            self.pos4[..., :3] = torch.as_tensor(positions, dtype=torch.float32, device=dev).reshape(self.E, self.N, 3)
        if velocities is not None:
            self.vel4[..., :3] = torch.as_tensor(velocities, dtype=torch.float32, device=dev).reshape(self.E, self.N, 3)
        if goal is not None:
            self.goal4[:, :3] = torch.as_tensor(goal, dtype=torch.float32, device=dev).reshape(self.E, 3)
        if obstacles is not None and self.M:
            self.obst4[..., :3] = torch.as_tensor(obstacles, dtype=torch.float32, device=dev).reshape(self.E, self.M, 3)
        if alive is not None:
            self.pos4[..., 3] = torch.as_tensor(alive, device=dev).reshape(self.E, self.N).to(torch.float32)
        if step_count is not None:
            self.step_count[:] = torch.as_tensor(step_count, dtype=torch.int32, device=dev)
        if observe:
            self.observe()

    # ------------------------------------------------------------------ (de)serialisation
    _STATE_KEYS = ("pos4", "vel4", "goal4", "obst4", "step_count", "rng", "ep_return", "stats_words")

    def _meta(self) -> dict[str, Any]:
        """Everything a trajectory depends on besides the state tensors: shapes, the whole env config, the
        randomisation spec and its stream keys."""
        import dataclasses
        return dict(kind=self.kind, E=self.E, N=self.N, M=self.M, K=self.K, S=self.S,
                    config=dataclasses.asdict(self.cfg), norm_mode=int(self._c.norm_mode),
                    domain_randomization={k: (list(v) if isinstance(v, (tuple, list)) else v) for k, v in (self.dr or {}).items()},
                    dr_seed=int(self._c.dr_seed), env_index_base=int(self._c.env_index_base),
                    global_state=self.global_state is not None, reward64=self.reward64 is not None)

    def state_dict(self) -> dict[str, Any]:
        """Everything needed to continue every env bit-for-bit (the reference never checkpoints env
        state; RLlib resumes only the policy): the state tensors, the output block of the last call
        (obs, obs_valid, rewards, flags -- so that what a caller reads before the next step is restored too) and a
        `meta` record that `load_state_dict` checks.  CPU tensors, safe to `torch.save`."""
        torch.cuda.synchronize(self.device)
        out = {k: getattr(self, k).detach().cpu().clone() for k in self._STATE_KEYS}
        out["outputs"] = self._out_block.detach().cpu().clone()
        for k in ("episode_return", "episode_length"):
            out[k] = getattr(self, k).detach().cpu().clone()
        if self.dr_params is not None:
            out["dr_params"] = self.dr_params.detach().cpu().clone()
        if self.act_hist is not None:
            out["act_hist"] = self.act_hist.detach().cpu().clone()
        out["meta"] = self._meta()
        return out

    def load_state_dict(self, state: dict[str, Any], observe: bool = False):
        """Restore `state_dict()` output.  The engine must have been built with the same shapes, env config,
        randomisation spec, `dr_seed` and `env_index_base` -- anything else would silently continue a different
        trajectory, so a mismatch raises with the differing keys."""
        meta, mine = dict(state.get("meta", {})), self._meta()

        def canon(x):
            if isinstance(x, dict):
                return {k: canon(v) for k, v in x.items()}
            return list(x) if isinstance(x, (tuple, list)) else x
        diff = sorted(k for k in set(meta) | set(mine) if canon(meta.get(k)) != canon(mine.get(k)))
        if diff:
            raise ValueError("state_dict does not belong to an engine like this one; differing: " +
                             ", ".join(f"{k}: saved {meta.get(k)!r} vs this {mine.get(k)!r}" for k in diff))
        for k in self._STATE_KEYS:
            getattr(self, k).copy_(state[k].to(self.device))
        for k in ("episode_return", "episode_length"):
            if k in state:
                getattr(self, k).copy_(state[k].to(self.device))
        if self.dr_params is not None:
            self.dr_params.copy_(state["dr_params"].to(self.device))
        if self.act_hist is not None:
            self.act_hist.copy_(state["act_hist"].to(self.device))
        if "outputs" in state and state["outputs"].numel() == self._out_block.numel():
            self._out_block.copy_(state["outputs"].to(self.device))
        elif not observe:
            observe = True
        if observe:
            self.observe()

    # ------------------------------------------------------------------ host-buffer (end-to-end) path
    def host_buffers(self, with_global_state: bool = True) -> dict[str, np.ndarray]:
        """Pinned host arrays for `step_host` (allocated once)."""
        if self._host is None:
            E, N, D, R = self.E, self.N, self.D, self.R
            pin = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory()  # noqa: E731
            h = dict(actions=pin((E, N, 3), torch.float32))
            # one pinned block mirroring the device output block (same offsets): views per field
            self._host_block = pin((self._out_block.numel(),), torch.uint8)
            for name, shape, dt, o, nbytes in self._out_layout:
                if name == "global_state" and not with_global_state:
                    continue
                h[name] = self._host_block[o:o + nbytes].view(dt).view(shape)
            h["flags"] = pin((E, N), torch.uint8)   # ABI 5: terminated | truncated | reached | collision | obs_valid bits
            self._host = h
        return self._host

    # named output sets of `step_host` (what crosses the link per agent-step, N = 32, M = 8):
    #   None      every field, flags as five byte arrays                                   185 B
    #   "packed"  the same information, the five flag arrays as one byte (`unpack_flags`)  181 B
    #   "lean"    what env.step() returns without the info dicts: obs, reward, flags       153 B
    OUTPUT_SETS = {
        "packed": ("obs", "reward", "dist", "flags", "all_terminated", "all_truncated", "global_state"),
        "lean": ("obs", "reward", "flags", "all_terminated", "all_truncated"),
    }

    def _output_names(self, outputs) -> tuple[str, ...]:
        h = self.host_buffers()
        if outputs is None:
            return tuple(n for n in _abi.HOST_OUT_FIELDS if n in h)
        if isinstance(outputs, str):
            outputs = self.OUTPUT_SETS[outputs]
        return tuple(n for n in outputs if n in h)

    @staticmethod
    def unpack_flags(flags) -> dict[str, Any]:
        """The five per-agent flag arrays (bool, same shape) of a packed `flags` byte array."""
        return {name: (flags & (1 << k)) != 0 for k, name in enumerate(_abi.FLAG_FIELDS)}

    def step_host(self, actions_host=None, auto_reset: bool = True, outputs: tuple[str, ...] | str | None = None):
        """One step through HOST buffers: H2D actions, kernel, D2H outputs (chunk-pipelined in the
        library).  `actions_host`: [E,N,3] float32 numpy / CPU tensor, or None to use the pinned
        `host_buffers()['actions']` in place.  `outputs`: field names, or one of `OUTPUT_SETS` ("packed", "lean");
        only those fields of the returned dict are fresh.  Returns the dict of pinned host tensors."""
        h = self.host_buffers()
        act = h["actions"]
        if actions_host is not None:
            src = torch.as_tensor(actions_host, dtype=torch.float32).reshape(self.E, self.N, 3)
            if src.is_pinned() and src.is_contiguous():
                act = src                      # already page-locked: DMA straight from the caller's buffer
            else:
                h["actions"].copy_(src)
        out = _abi.SwarmHostOut()
        if outputs is None and self._out_block.numel() <= self.HOST_BLOCK_MAX_BYTES:
            # small batch: the whole output block in one copy
            out.block_host, out.block_dev = self._host_block.data_ptr(), self._out_block.data_ptr()
            out.block_bytes = self._out_block.numel()
        else:
            for name in self._output_names(outputs):
                setattr(out, name, h[name].data_ptr())
        rc = self._lib.swarm_step_host(self._handle, self._bufs_ref, act.data_ptr(), C.byref(out), int(auto_reset),
                                       self._stream())
        if rc != 0:
            _abi.check(rc, "swarm_step_host")
        return h

    HOST_BLOCK_MAX_BYTES = 1 << 20   # above this the env-axis chunk pipeline (copies under the kernel) wins

    def fetch_outputs(self) -> dict[str, Any]:
        """Copy the whole output block (obs, reward, flags, global_state ...) to the pinned host mirror in one
        device->host copy and wait for it: what the facade envs call after `reset()`."""
        h = self.host_buffers()
        self._host_block.copy_(self._out_block, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return h

    def host_bytes_per_step(self, outputs: tuple[str, ...] | str | None = None) -> tuple[int, int]:
        h = self.host_buffers()
        return h["actions"].numel() * 4, sum(h[n].numel() * h[n].element_size() for n in self._output_names(outputs))

    # ------------------------------------------------------------------ statistics
    def stats(self) -> dict[str, float]:
        """Running episode statistics accumulated on the device (one D2H copy of 64 bytes)."""
        w = self.stats_words.cpu().numpy()
        out = {n: int(w[i]) for i, n in enumerate(_abi.STAT_NAMES)}
        out["return_sum"] = float(w[5:6].view(np.float64)[0])
        return out

    def reset_stats(self):
        self.stats_words.zero_()

    def algorithmic_bytes_per_agent_step(self) -> float:
        """unpadded logical bytes one agent-step must move."""
        N, M, D = self.N, self.M, self.D
        per_agent = 36 + 24 + 4 * D + 4 + 2 + 6
        per_env = 12 + 12 * M + 4 + 4 + 2
        if self.kind != "single" and self.global_state is not None:
            per_agent += 24
            per_env += 12
        if self.dr_params is not None:
            per_env += 32   # this episode's randomised constants
        return per_agent + per_env / N


class StepGraph:
    """A captured sequence of T env steps (see `SwarmEngine.capture_steps`): one `cudaGraphLaunch` per replay instead
    of 2 T kernel launches from Python.  Launch-bound batch sizes (a few thousand envs) gain the most."""

    def __init__(self, engine: SwarmEngine, actions: torch.Tensor, auto_reset: bool = True):
        if not (torch.is_tensor(actions) and actions.is_cuda and actions.dtype is torch.float32 and
                actions.is_contiguous() and actions.dim() == 4):
            raise TypeError("actions must be a contiguous [T,E,N,3] float32 CUDA tensor (it becomes the graph's input)")
        self.engine, self.actions, self.steps = engine, actions, int(actions.shape[0])
        engine.step_many(actions[:1], auto_reset=auto_reset)       # warm-up outside the capture (lazy attribute sets)
        torch.cuda.synchronize(engine.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            engine.step_many(actions, auto_reset=auto_reset)

    def replay(self):
        self.graph.replay()
        return self.engine._step_result
