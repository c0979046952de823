"""Import the UNMODIFIED reference envs (TEST INFRASTRUCTURE ONLY).

Works only where `/root/reference` exists (the build container).  Used by
`oracle/gen_golden.py` to generate tests/golden/*.npz and by the `not gpu`
tests that pin the C oracle against the live reference when it is present.
Nothing in the product package, `bench.py` or the `-m gpu` tests imports this.
"""
from __future__ import annotations

import os
import sys

REFERENCE_ROOT = os.environ.get("SWARM_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gymnasium_shim")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "swarm_marl", "envs"))


def load_reference_envs():
    """Return (SingleDroneEnv, DroneSwarmEnv) classes from the reference source tree."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    try:
        import gymnasium  # noqa: F401  (a real install wins if there is one)
    except ModuleNotFoundError:
        if _SHIM not in sys.path:
            sys.path.insert(0, _SHIM)
    src = os.path.join(REFERENCE_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    from swarm_marl.envs import DroneSwarmEnv, SingleDroneEnv  # type: ignore

    return SingleDroneEnv, DroneSwarmEnv
