"""Env factories with the reference's signature (ballbot_rl/training/utils.py:11-85) plus the batched constructor."""
import numpy as np


def make_ballbot_env(terrain_type=None, reward_config=None, terrain_config=None, env_config=None, gui=False, disable_cams=False, seed=0,
                     log_options={"cams": False, "reward_terms": False}, eval_env=False, viewer_title=None, device=0, precision=64):
    """Returns ``_init`` building ONE env, like the reference (used there once per SubprocVecEnv worker).  New code should
    call :func:`make_ballbot_vec_env` instead, which builds all N envs inside a single GPU engine."""
    if terrain_config is None:
        terrain_config = {"type": terrain_type if terrain_type is not None else "perlin", "config": {}}
    if reward_config is None:
        reward_config = {"type": "directional", "config": {"target_direction": [0.0, 1.0]}}

    def _init():
        from ..envs.ballbot_env import BBotSimulation
        return BBotSimulation(xml_path=None, GUI=gui, log_options=log_options, terrain_type=terrain_config.get("type", "perlin"),
                              reward_config=reward_config, terrain_config=terrain_config, env_config=env_config,
                              eval_env=[eval_env, seed], render_mode="rgb_array" if eval_env else None, viewer_title=viewer_title,
                              disable_cameras=disable_cams, device=device, precision=precision)

    return _init


def make_ballbot_vec_env(num_envs, terrain_config=None, reward_config=None, env_config=None, seed=0, disable_cams=False, device=0,
                         precision=64, output="torch", solver="fast", rank=0, world_size=1):
    """GPU replacement of ``SubprocVecEnv([make_ballbot_env(...) for _ in range(N)])`` (train.py:82-97).  With
    ``world_size > 1`` the N envs are sharded by index across ranks (one process per GPU, no collective on the step path)."""
    from ..envs.vec_env import BallbotVecEnv
    per_rank, offset = shard_envs(num_envs, rank, world_size)
    return BallbotVecEnv(per_rank, terrain_config=terrain_config, reward_config=reward_config, env_config=env_config, seed=seed,
                         disable_cams=disable_cams, device=device, precision=precision, output=output, solver=solver, env_offset=offset)


def shard_envs(num_envs: int, rank: int, world_size: int):
    """Contiguous index sharding: rank r owns envs [offset, offset + count)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, rem = divmod(int(num_envs), int(world_size))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return count, offset


def deg2rad(d):
    return d * np.pi / 180


def rad2deg(r):
    return r * 180 / np.pi
