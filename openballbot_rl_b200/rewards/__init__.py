"""Reward plugins; importing this module registers the built-ins (ballbot_gym/rewards/__init__.py:8-9)."""
from ..core.registry import ComponentRegistry
from .base import BaseReward
from .directional import DirectionalReward
from .distance import DistanceReward


def register_builtin_rewards():
    for name, cls in (("directional", DirectionalReward), ("distance", DistanceReward)):
        if name not in ComponentRegistry.list_rewards():
            ComponentRegistry.register_reward(name, cls)


register_builtin_rewards()

__all__ = ["BaseReward", "DirectionalReward", "DistanceReward", "register_builtin_rewards"]
