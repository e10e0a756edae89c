from .utils import make_ballbot_env, make_ballbot_vec_env, shard_envs

__all__ = ["make_ballbot_env", "make_ballbot_vec_env", "shard_envs"]
