"""Host-side mirror of the reference's `DroneEnvConfig` (envs/common.py:7-32).

Same field names, defaults and `from_dict` behaviour (unknown keys are silently dropped) so
the reference's constructor dicts, YAML stage `env_config`s (configs/curriculum_v1.yaml) and
`env.cfg.<field>` reads (scripts/evaluate_protocol.py:239) keep working.
"""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Any


@dataclass
class DroneEnvConfig:
    world_size: float = 20.0
    dt: float = 0.1
    max_steps: int = 400
    max_speed: float = 4.0
    max_accel: float = 2.0
    collision_radius: float = 0.5
    goal_radius: float = 0.8
    num_obstacles: int = 8
    sensed_obstacles: int = 4
    neighbor_k: int = 3
    obstacle_radius: float = 0.8
    desired_spacing: float = 2.5
    reward_progress_scale: float = 2.0
    reward_goal: float = 25.0
    reward_collision: float = -25.0
    reward_formation_scale: float = 0.15
    seed: int | None = None

    @classmethod
    def from_dict(cls, raw: dict[str, Any] | None) -> "DroneEnvConfig":
        if not raw:
            return cls()
        names = {f.name for f in fields(cls)}
        return cls(**{k: v for k, v in raw.items() if k in names})


DR_RANGE_KEYS = ("mass_scale", "max_accel_scale", "max_speed_scale", "dt_scale", "obstacle_radius_scale",
                 "world_size_scale")
DR_STD_KEYS = ("thrust_noise_std", "position_noise_std", "velocity_noise_std", "obstacle_distance_noise_std")


def flatten_domain_randomization(spec: dict[str, Any] | None) -> dict[str, Any]:
    """Normalise a domain-randomisation spec to {range key: (min, max), std key: sigma}.

    Accepts the flat form itself, or the document shape of the reference's
    `configs/domain_randomization_v1.yaml:9-60` (`randomization: {dynamics|actuation|sensing|
    environment: {<key>: {distribution, min, max | std}}}`).  No reference code reads that file, so
    what each key DOES is defined by this engine.  `control_delay_steps`
    ({values, probs} or ((values...), (probs...))) is kept only when some value is non-zero."""
    if not spec:
        return {}
    flat: dict[str, Any] = {}
    groups = spec.get("randomization") if isinstance(spec.get("randomization"), dict) else None
    items = {}
    if groups is not None:
        for g in groups.values():
            if isinstance(g, dict):
                items.update(g)
    else:
        items = dict(spec)
    for k, v in items.items():
        if k in DR_RANGE_KEYS:
            lo, hi = (v["min"], v["max"]) if isinstance(v, dict) else v
            flat[k] = (float(lo), float(hi))
        elif k in DR_STD_KEYS:
            flat[k] = float(v["std"]) if isinstance(v, dict) else float(v)
        elif k == "control_delay_steps":
            vals, probs = (v.get("values", [0]), v.get("probs")) if isinstance(v, dict) else v
            vals = [int(x) for x in vals]
            probs = [1.0 / len(vals)] * len(vals) if probs is None else [float(x) for x in probs]
            if len(vals) != len(probs) or not 1 <= len(vals) <= 4 or min(vals) < 0 or max(vals) > 8 or min(probs) < 0:
                raise ValueError("control_delay_steps: 1-4 values in [0, 8] with matching non-negative probs")
            if any(x for x, pr in zip(vals, probs) if pr > 0):
                flat[k] = (tuple(vals), tuple(probs))
        elif groups is None:
            raise KeyError(f"unknown domain-randomisation key {k!r}")
    return flat
