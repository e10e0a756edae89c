from .ballbot_env import BBotSimulation
from .vec_env import BallbotVecEnv
from .spaces import create_observation_space, create_action_space

__all__ = ["BBotSimulation", "BallbotVecEnv", "create_observation_space", "create_action_space"]
