"""Plugin registry with the reference's API surface (ballbot_gym/core/registry.py:8-231).

Same class-level entry points (``register_/get_/list_`` for rewards, terrains, policies and sensors, ``clear``) and the
same error types / message substrings that the reference's unit tests assert (tests/unit/test_registry.py), built on
one generic table instead of four hand-written copies.
"""
from typing import Any, Callable, Dict, List


class _Table:
    """One named component table (kind = 'reward' | 'terrain' | 'policy' | 'sensor')."""

    def __init__(self, kind: str, plural: str, check: Callable[[Any], None] = None):
        self.kind, self.plural, self.check, self.items = kind, plural, check, {}  # type: ignore[var-annotated]

    def add(self, name: str, obj: Any) -> None:
        if name in self.items:
            raise ValueError(f"{self.kind.capitalize()} '{name}' is already registered. "
                             f"Available {self.plural}: {list(self.items)}")
        if self.check is not None:
            self.check(obj)
        self.items[name] = obj

    def lookup(self, name: str) -> Any:
        try:
            return self.items[name]
        except KeyError:
            raise ValueError(f"Unknown {self.kind}: '{name}'. Available {self.plural}: {list(self.items)}") from None


def _check_reward(cls: Any) -> None:
    from ..rewards.base import BaseReward
    if not (isinstance(cls, type) and issubclass(cls, BaseReward)):
        raise ValueError(f"Reward class must inherit from BaseReward, got {cls}")


def _check_terrain(fn: Any) -> None:
    if not callable(fn):
        raise ValueError(f"Terrain must be callable, got {type(fn)}")


class ComponentRegistry:
    """Central registry for rewards, terrains, policies and sensors (process-global, like the reference's)."""

    _tables: Dict[str, _Table] = {
        "reward": _Table("reward", "rewards", _check_reward),
        "terrain": _Table("terrain", "terrains", _check_terrain),
        "policy": _Table("policy", "policies"),
        "sensor": _Table("sensor", "sensors"),
    }
    # the reference exposes these dicts as class attributes; keep them as live views
    _rewards = _tables["reward"].items
    _terrains = _tables["terrain"].items
    _policies = _tables["policy"].items
    _sensors = _tables["sensor"].items

    # ---- rewards: stored as classes, get_reward instantiates (registry.py:36-85)
    @classmethod
    def register_reward(cls, name: str, reward_class: type) -> None:
        cls._tables["reward"].add(name, reward_class)

    @classmethod
    def get_reward(cls, name: str, **kwargs):
        return cls._tables["reward"].lookup(name)(**kwargs)

    @classmethod
    def list_rewards(cls) -> List[str]:
        return list(cls._tables["reward"].items)

    # ---- terrains: plain callables (n, **cfg) -> flat array (registry.py:87-131)
    @classmethod
    def register_terrain(cls, name: str, terrain_fn: Callable) -> None:
        cls._tables["terrain"].add(name, terrain_fn)

    @classmethod
    def get_terrain(cls, name: str) -> Callable:
        return cls._tables["terrain"].lookup(name)

    @classmethod
    def list_terrains(cls) -> List[str]:
        return list(cls._tables["terrain"].items)

    # ---- policies / sensors: classes returned un-instantiated (registry.py:133-223)
    @classmethod
    def register_policy(cls, name: str, policy_class: type) -> None:
        cls._tables["policy"].add(name, policy_class)

    @classmethod
    def get_policy(cls, name: str) -> type:
        return cls._tables["policy"].lookup(name)

    @classmethod
    def list_policies(cls) -> List[str]:
        return list(cls._tables["policy"].items)

    @classmethod
    def register_sensor(cls, name: str, sensor_class: type) -> None:
        cls._tables["sensor"].add(name, sensor_class)

    @classmethod
    def get_sensor(cls, name: str) -> type:
        return cls._tables["sensor"].lookup(name)

    @classmethod
    def list_sensors(cls) -> List[str]:
        return list(cls._tables["sensor"].items)

    @classmethod
    def clear(cls) -> None:
        """Drop every registration (the reference's tests call this in setup_method)."""
        for t in cls._tables.values():
            t.items.clear()
