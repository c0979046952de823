"""Minimal stand-in for the `gymnasium` package (TEST INFRASTRUCTURE ONLY).

`gymnasium` is not installed in this image.  The reference kinematic envs
(`/root/reference/src/swarm_marl/envs/{single_drone_env,drone_swarm_env}.py`)
only need `gym.Env` (with a `reset(seed=...)` that can be super()-called) and
`spaces.Box`; this stand-in provides exactly that so the UNMODIFIED reference
can be imported in the build container to pin the C oracle and to generate the
golden fixtures under tests/golden/.  It is never imported by the product.
"""
from . import spaces  # noqa: F401


class Env:
    metadata: dict = {}
    observation_space = None
    action_space = None

    def reset(self, *, seed=None, options=None):
        return None

    def step(self, action):
        raise NotImplementedError

    def close(self):
        pass
