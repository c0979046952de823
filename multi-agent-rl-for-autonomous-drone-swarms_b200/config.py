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
