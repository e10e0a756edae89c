"""Directional reward (ballbot_gym/rewards/directional.py:33-54): ``state["vel"][-3:-1] . target_direction``.

Note the reference's quirk: obs["vel"] is MuJoCo cvel[0:3], i.e. the base's world angular velocity.
The engine evaluates this reward inside the step kernel (BB_REWARD_DIRECTIONAL); this class is the plugin-API object
and the reference-compatible host implementation for single states or batches.
"""
import numpy as np

from .base import BaseReward


class DirectionalReward(BaseReward):
    def __init__(self, target_direction):
        self.target_direction = target_direction

    def __call__(self, state: dict):
        vel = state["vel"]
        xy = vel[..., -3:-1]                      # last axis: works for (3,) numpy and [N, 3] tensors alike
        td = self.target_direction
        if hasattr(xy, "is_cuda"):                # torch batch
            import torch
            tdt = torch.as_tensor(np.asarray(td, dtype=np.float32), device=xy.device, dtype=xy.dtype)
            return (xy * tdt).sum(-1)
        return xy.dot(td)
