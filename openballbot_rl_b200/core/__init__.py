"""Plugin core: registry, factories, config (API of ballbot_gym/core)."""
from .registry import ComponentRegistry
from .factories import create_reward, create_terrain, create_policy, validate_config
from .config import load_config, merge_configs, load_training_config, get_component_config

__all__ = ["ComponentRegistry", "create_reward", "create_terrain", "create_policy", "validate_config",
           "load_config", "merge_configs", "load_training_config", "get_component_config"]
