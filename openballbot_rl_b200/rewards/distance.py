"""Distance-to-goal reward (ballbot_gym/rewards/distance.py:33-50): ``-scale * ||goal - pos2d||``."""
import numpy as np

from .base import BaseReward


class DistanceReward(BaseReward):
    def __init__(self, goal_position, scale: float = 1.0):
        self.goal_position = np.array(goal_position, dtype=np.float32)
        if self.goal_position.shape != (2,):
            raise ValueError(f"goal_position must be shape (2,), got {self.goal_position.shape}")
        self.scale = float(scale)

    def __call__(self, state: dict):
        if "pos2d" not in state:
            raise ValueError("DistanceReward requires 'pos2d' in state dictionary")
        pos = state["pos2d"]
        if hasattr(pos, "is_cuda"):               # torch batch [N, 2]
            import torch
            goal = torch.as_tensor(self.goal_position, device=pos.device, dtype=pos.dtype)
            return -self.scale * torch.linalg.vector_norm(goal - pos, dim=-1)
        cur = np.array(pos, dtype=np.float32)
        return -self.scale * np.linalg.norm(self.goal_position - cur)
