"""Config -> component factories with the reference's semantics (ballbot_gym/core/factories.py:9-213)."""
from typing import Any, Callable, Dict

import numpy as np

from .registry import ComponentRegistry

_COMPONENT_LISTS = {"reward": ComponentRegistry.list_rewards, "terrain": ComponentRegistry.list_terrains,
                    "policy": ComponentRegistry.list_policies}


def _split(config: Any, what: str):
    if not isinstance(config, dict):
        raise ValueError(f"{what} config must be a dictionary, got {type(config)}")
    kind = config.get("type")
    if kind is None:
        raise ValueError(f"{what} config must have 'type' key")
    return kind, config.get("config", {})


def _as_f32(v):
    return np.array(v, dtype=np.float32) if isinstance(v, list) else v


def create_reward(config: Dict[str, Any]):
    """factories.py:9-78. Env-level knobs (scale, action_reg_coef, survival_bonus) are NOT forwarded to the built-in
    rewards; unknown reward types receive the whole config dict (so ``scale`` is used twice for plugins, SURVEY C#9)."""
    kind, params = _split(config, "Reward")
    if kind == "directional":
        if "target_direction" not in params:
            raise ValueError("DirectionalReward requires 'target_direction' in config")
        kwargs = {"target_direction": _as_f32(params["target_direction"])}
    elif kind == "distance":
        if "goal_position" not in params:
            raise ValueError("DistanceReward requires 'goal_position' in config")
        kwargs = {"goal_position": _as_f32(params["goal_position"]), "scale": params.get("scale", 1.0)}
    else:
        kwargs = params
    try:
        return ComponentRegistry.get_reward(kind, **kwargs)
    except ValueError as e:
        raise ValueError(f"Failed to create reward '{kind}': {e}")
    except TypeError as e:
        raise TypeError(f"Failed to create reward '{kind}' with parameters {list(kwargs.keys())}: {e}")


def create_terrain(config: Dict[str, Any]) -> Callable:
    """factories.py:81-126: returns ``f(n, **overrides)`` that merges the runtime overrides (seed) over the config."""
    kind, params = _split(config, "Terrain")
    try:
        fn = ComponentRegistry.get_terrain(kind)
    except ValueError as e:
        raise ValueError(f"Failed to get terrain '{kind}': {e}")

    def configured_terrain(n: int, **override_kwargs) -> np.ndarray:
        return fn(n, **{**params, **override_kwargs})

    configured_terrain.terrain_type = kind            # engine hint: lets the GPU VecEnv map built-ins to kernels
    configured_terrain.terrain_params = dict(params)
    return configured_terrain


def create_policy(config: Dict[str, Any]) -> type:
    """factories.py:129-162: the class, not an instance."""
    kind, _ = _split(config, "Policy")
    try:
        return ComponentRegistry.get_policy(kind)
    except ValueError as e:
        raise ValueError(f"Failed to get policy '{kind}': {e}")


def validate_config(config: Dict[str, Any], component_type: str) -> bool:
    """factories.py:165-213."""
    if not isinstance(config, dict):
        raise ValueError(f"Config must be a dictionary, got {type(config)}")
    if "type" not in config:
        raise ValueError(f"{component_type} config must have 'type' key")
    if component_type not in _COMPONENT_LISTS:
        raise ValueError(f"Unknown component_type '{component_type}'. Must be one of: 'reward', 'terrain', 'policy'")
    available = _COMPONENT_LISTS[component_type]()
    if config["type"] not in available:
        raise ValueError(f"Unknown {component_type} type '{config['type']}'. Available: {available}")
    return True
