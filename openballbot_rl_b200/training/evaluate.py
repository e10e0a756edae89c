"""Fixed-policy evaluation on the GPU VecEnv.

Stands in for the episode loop of ``ballbot_rl/evaluation/evaluate.py:130-165`` (``model.predict(obs, deterministic=True)`` ->
``env.step`` until ``terminated``) and for SB3's ``EvalCallback`` (``train.py:225-229``): every env of the batch plays its first
episode to the end; observations, policy forward and env step all stay on the device.
"""
from typing import Callable, Dict

import torch


@torch.no_grad()
def evaluate_policy(venv, policy: Callable, max_steps: int = 4000, deterministic: bool = True, terrain_seeds=None,
                    reset: bool = True) -> Dict[str, torch.Tensor]:
    """Returns per-env ``returns`` / ``lengths`` / ``failure`` of the first episode of every env (torch-output ``BallbotVecEnv``).
    ``policy(obs_dict, deterministic=...) -> actions [N,3]`` (e.g. ``BallbotPolicy``).  ``terrain_seeds`` replays recorded
    ``r_seed`` values (the terrains of an archived evaluation) instead of drawing new ones; ``reset=False`` starts from the
    venv's current (freshly reset) state."""
    N, dev = venv.num_envs, venv.engine.device
    if not reset:
        obs = venv._obs_view()
    else:
        obs = venv.reset(terrain_seeds=terrain_seeds) if terrain_seeds is not None else venv.reset()
    ret = torch.zeros(N, device=dev); length = torch.zeros(N, dtype=torch.int32, device=dev)
    alive = torch.ones(N, dtype=torch.bool, device=dev); failed = torch.zeros(N, dtype=torch.bool, device=dev)
    for t in range(max_steps):
        obs, rew, dones, info = venv.step(policy(obs, deterministic=deterministic))
        ret += rew * alive; length += alive.int()
        failed |= alive & dones & info["failure"].bool()
        alive &= ~dones
        if t % 16 == 15 and not bool(alive.any()):     # one host sync every 16 steps
            break
    return {"returns": ret, "lengths": length, "failure": failed, "mean_return": float(ret.mean()), "mean_length": float(length.float().mean())}
