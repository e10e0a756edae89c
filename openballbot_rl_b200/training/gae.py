"""GAE(gamma, lambda) on device-resident rollout tensors through the C ABI (``bb_gae``, csrc/bb_rollout.cu).

Replaces ``RolloutBuffer.compute_returns_and_advantage`` of the reference's learner (Stable-Baselines3 PPO driven from
ballbot_rl/training/train.py:126-141, 284).  CUDA only: CPU tensors are rejected, there is no fallback.
"""
import ctypes as C

import torch

from .. import _lib


def compute_gae(rewards: torch.Tensor, values: torch.Tensor, dones: torch.Tensor, gamma: float = 0.99, gae_lambda: float = 0.95):
    """rewards [T,N] f32, values [T+1,N] f32 (row T = bootstrap value), dones [T,N] bool/uint8 -> (advantages, returns) [T,N]."""
    if not (rewards.is_cuda and values.is_cuda and dones.is_cuda):
        raise _lib.EngineError("compute_gae needs CUDA tensors (bb_gae has no CPU fallback)")
    T, N = rewards.shape
    if values.shape != (T + 1, N) or dones.shape != (T, N):
        raise ValueError(f"shapes: rewards {tuple(rewards.shape)}, values {tuple(values.shape)} (need [T+1,N]), dones {tuple(dones.shape)}")
    rewards = rewards.float().contiguous(); values = values.float().contiguous()
    dones = dones.to(torch.uint8).contiguous()
    try:
        return _lib.torch_ops().gae(rewards, values, dones, float(gamma), float(gae_lambda))      # torch.ops.ballbot.gae
    except _lib.EngineError:
        pass                                                                                        # ctypes on the same C entry point
    adv = torch.empty_like(rewards); ret = torch.empty_like(rewards)
    stream = C.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    with torch.cuda.device(rewards.device):
        rc = _lib.lib().bb_gae(p(rewards), p(values), p(dones), T, N, float(gamma), float(gae_lambda), p(adv), p(ret), stream)
    if rc != 0:
        raise _lib.EngineError(f"bb_gae failed ({rc})")
    return adv, ret
