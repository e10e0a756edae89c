"""Device-resident rollout collection for the GPU VecEnv (replaces SB3's numpy DictRolloutBuffer round trip,
reference call stack ballbot_rl/training/train.py:126-141 -> PPO.collect_rollouts) and the cross-rank statistics reduce.

Envs are sharded across ranks with no collective on the step path; the only communication is this once-per-iteration
all-reduce of a few scalars (and the PPO gradient all-reduce done by the learner).
"""
from typing import Callable, Dict, Optional

import torch
import torch.distributed as dist


def reduce_rollout_stats(ep_return: torch.Tensor, ep_length: torch.Tensor, done: torch.Tensor, steps: int) -> Dict[str, float]:
    """Sum/count of finished-episode returns and lengths over all ranks (Monitor's rollout/ep_rew_mean, ep_len_mean)."""
    done = done.bool()
    vec = torch.stack([ep_return[done].double().sum(), ep_length[done].double().sum(), done.double().sum(),
                       torch.tensor(float(steps), dtype=torch.float64, device=ep_return.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    r, l, n, s = (float(x) for x in vec.tolist())
    return {"ep_rew_mean": r / n if n else float("nan"), "ep_len_mean": l / n if n else float("nan"), "episodes": int(n), "env_steps": int(s)}


@torch.no_grad()
def collect_rollout(venv, policy: Callable, n_steps: int, gamma: float = 0.99, gae_lambda: float = 0.95,
                    value_fn: Optional[Callable] = None):
    """Collect ``n_steps`` transitions from every env of a torch-output BallbotVecEnv without leaving the device.

    ``policy(obs_dict) -> actions [N,3]`` (already clipped to [-1,1] like SB3 does before env.step);
    ``value_fn(obs_dict) -> values [N]`` enables GAE(gamma, lambda) (the reference's values: 0.99 / 0.95).
    Returns a dict of ``[T, N, ...]`` tensors plus the reduced episode statistics.
    """
    eng = venv.engine
    N, dev = venv.num_envs, eng.device
    obs = venv._obs_view()
    keys = [k for k in obs if not k.startswith("rgbd_")]
    buf = {k: torch.empty((n_steps, N) + tuple(obs[k].shape[1:]), device=dev) for k in keys}
    act = torch.empty(n_steps, N, 3, device=dev); rew = torch.empty(n_steps, N, device=dev)
    done = torch.empty(n_steps, N, dtype=torch.bool, device=dev); val = torch.zeros(n_steps + 1, N, device=dev)
    ep_r = torch.zeros(N, device=dev); ep_l = torch.zeros(N, dtype=torch.int32, device=dev); ep_d = torch.zeros(N, dtype=torch.bool, device=dev)
    for t in range(n_steps):
        for k in keys:
            buf[k][t].copy_(obs[k])
        if value_fn is not None:
            val[t] = value_fn(obs)
        a = policy(obs).clamp_(-1.0, 1.0)
        obs, r, d, info = venv.step(a)
        act[t], rew[t], done[t] = a, r, d
        ep_r = torch.where(d, info["episode_r"], ep_r); ep_l = torch.where(d, info["episode_l"], ep_l); ep_d |= d
    out = {"obs": buf, "actions": act, "rewards": rew, "dones": done}
    if value_fn is not None:
        val[n_steps] = value_fn(obs)
        adv = torch.zeros(n_steps, N, device=dev); last = torch.zeros(N, device=dev)
        for t in reversed(range(n_steps)):
            nonterminal = (~done[t]).float()
            delta = rew[t] + gamma * val[t + 1] * nonterminal - val[t]
            last = delta + gamma * gae_lambda * nonterminal * last
            adv[t] = last
        out["advantages"], out["returns"], out["values"] = adv, adv + val[:-1], val[:-1]
    out["stats"] = reduce_rollout_stats(ep_r, ep_l, ep_d, steps=n_steps * N)
    return out
