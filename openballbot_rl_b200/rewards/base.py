"""Reward plugin interface (ballbot_gym/rewards/base.py:7-21).

A reward is called with the observation/state dict.  In the reference the dict holds one env's numpy vectors; in the
batched engine the same object may also be called with ``[N, 3]`` torch CUDA tensors and must then return ``[N]``.
"""
from abc import ABC, abstractmethod
from typing import Dict


class BaseReward(ABC):
    @abstractmethod
    def __call__(self, state: Dict):
        """Reward for ``state`` (float for one env, tensor ``[N]`` for a batch)."""
