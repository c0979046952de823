"""B200-native batched simulator for the drone-swarm environment step.

Drop-in for ONE hot path of nusRying/Multi-Agent-RL-for-Autonomous-Drone-Swarms: the
reset/step of `DroneSwarmEnv` / `SingleDroneEnv` (reference src/swarm_marl/envs/), computed by
hand-written sm_100a kernels behind the C ABI in include/swarm_b200.h.

    SwarmEngine      batched tensor API (E env instances per GPU, one step launch + a tiny auto-reset launch)
    evaluate_batched reference evaluation protocol (SR / CFR / TTG / FE / PE) for E episodes at once
    collect_rollout  device-resident policy loop: [T,E,...] sample-batch columns incl. global_state, no host sync
    VectorSwarmEnv   RLlib BaseEnv-shaped vector env (lazy dicts) + a batched array interface over one engine
    DroneSwarmEnv    reference-compatible multi-agent env (dict API) backed by the engine
    SingleDroneEnv   reference-compatible single-agent env backed by the engine
    DronePhysicsEnv  the PyBullet env's contract + force/drag/gravity model as point masses (parity unpinned)
    DroneEnvConfig   mirror of the reference's config dataclass

The directory name contains hyphens, so import it as `swarm_b200` (alias module at the repo
root) or via importlib.import_module("multi-agent-rl-for-autonomous-drone-swarms_b200").
"""
from .config import DroneEnvConfig  # noqa: F401
from . import _abi  # noqa: F401


def __getattr__(name):
    # torch-dependent pieces are imported lazily so `import swarm_b200` stays cheap
    if name in ("SwarmEngine", "StepGraph"):
        from . import engine
        return getattr(engine, name)
    if name in ("DroneSwarmEnv", "SingleDroneEnv", "DronePhysicsEnv", "make_env_creator", "VectorSwarmEnv"):
        from . import envs
        return getattr(envs, name)
    if name == "collect_rollout":
        from .rollout import collect_rollout
        return collect_rollout
    if name == "evaluate_batched":
        from .evaluation import evaluate_batched
        return evaluate_batched
    if name in ("ShardedSwarm", "shard_range"):
        from . import distributed
        return getattr(distributed, name)
    raise AttributeError(name)
