"""Time the UNMODIFIED Python reference env on the host cores (TEST INFRASTRUCTURE ONLY: bench.py's CPU legs).

Looks for the staged copy under oracle/_ref/src (see oracle/make_ref.py; that is what exists on the GPU box) and
falls back to /root/reference/src (the build container).  One worker PROCESS per host thread, each stepping its own
env instances with U(-1, 1) actions and auto-reset, the way the reference is used (one env per RLlib worker:
src/swarm_marl/training/config_builders.py:19-23).
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_src() -> str | None:
    for root in (os.path.join(HERE, "_ref", "src"), os.path.join(os.environ.get("SWARM_REFERENCE_ROOT", "/root/reference"), "src")):
        if os.path.isdir(os.path.join(root, "swarm_marl", "envs")):
            return root
    return None


def _import_envs():
    src = reference_src()
    if src is None:
        raise RuntimeError("no reference env sources (oracle/_ref/src or /root/reference/src)")
    try:
        import gymnasium  # noqa: F401
    except ModuleNotFoundError:
        shim = os.path.join(HERE, "gymnasium_shim")
        if shim not in sys.path:
            sys.path.insert(0, shim)
    if src not in sys.path:
        sys.path.insert(0, src)
    from swarm_marl.envs import DroneSwarmEnv, SingleDroneEnv  # type: ignore
    return SingleDroneEnv, DroneSwarmEnv


def _worker(rank, kind, cfg, envs_per_proc, warm_steps, min_seconds, min_steps, barrier, out):
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import numpy as np
    SingleDroneEnv, DroneSwarmEnv = _import_envs()
    envs, rngs = [], []
    for k in range(envs_per_proc):
        seed = rank * envs_per_proc + k
        env = (SingleDroneEnv if kind == "single" else DroneSwarmEnv)({**cfg, "seed": seed})
        env.reset()
        envs.append(env)
        rngs.append(np.random.default_rng(1000 + seed))

    def step_all():
        n = 0
        for env, rng in zip(envs, rngs):
            if kind == "single":
                _, _, tm, tr, _ = env.step(rng.uniform(-1, 1, 3).astype(np.float32))
                n += 1
                done = tm or tr
            else:
                agents = list(env.agents)
                acts = {a: rng.uniform(-1, 1, 3).astype(np.float32) for a in agents}
                _, _, tm, tr, _ = env.step(acts)
                n += len(agents)
                done = tm["__all__"] or tr["__all__"]
            if done:
                env.reset()
        return n

    for _ in range(warm_steps):
        step_all()
    barrier.wait()
    t0 = time.perf_counter()
    steps, agent_steps = 0, 0
    while steps < min_steps or time.perf_counter() - t0 < min_seconds:
        agent_steps += step_all()
        steps += 1
    out.put((rank, agent_steps, steps, time.perf_counter() - t0))


def time_reference(kind, cfg, procs=None, envs_per_proc=1, warm_steps=3, min_seconds=2.0, min_steps=1):
    """Returns dict(rate agent-steps/s over all processes, procs, steps per process, seconds)."""
    procs = procs or len(os.sched_getaffinity(0))
    ctx = mp.get_context("fork")
    barrier, out = ctx.Barrier(procs), ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, kind, dict(cfg), envs_per_proc, warm_steps, min_seconds, min_steps, barrier, out))
          for r in range(procs)]
    for p in ps:
        p.start()
    res = [out.get(timeout=600) for _ in ps]
    for p in ps:
        p.join()
    # every process ran at least min_seconds from the common barrier: rate = sum of per-process rates
    rate = sum(a / dt for _, a, _, dt in res)
    return {"rate": rate, "procs": procs, "envs_per_proc": envs_per_proc, "steps_per_proc": min(s for _, _, s, _ in res),
            "seconds": max(dt for _, _, _, dt in res), "agent_steps": sum(a for _, a, _, _ in res)}


if __name__ == "__main__":
    import json
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    print(json.dumps(time_reference("swarm", {"num_drones": n, "num_obstacles": 8}, min_seconds=3.0)))
