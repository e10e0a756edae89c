"""Drop-in alias of the reference's ``ballbot_gym`` package backed by the B200 engine (openballbot_rl_b200).

``import ballbot_gym`` keeps working for reference users: the same sub-module paths resolve
(``ballbot_gym.core.registry.ComponentRegistry``, ``ballbot_gym.core.factories.create_reward``, ``ballbot_gym.rewards``,
``ballbot_gym.terrain``, ``ballbot_gym.envs.ballbot_env.BBotSimulation`` ...), the built-in rewards / terrains are
registered on import and, when gymnasium is installed, the env id ``ballbot-v0.1`` is registered
(reference: ballbot_gym/__init__.py:34-53).
"""
import importlib
import sys

_MAP = {
    "ballbot_gym.core": "openballbot_rl_b200.core",
    "ballbot_gym.core.registry": "openballbot_rl_b200.core.registry",
    "ballbot_gym.core.factories": "openballbot_rl_b200.core.factories",
    "ballbot_gym.core.config": "openballbot_rl_b200.core.config",
    "ballbot_gym.rewards": "openballbot_rl_b200.rewards",
    "ballbot_gym.rewards.base": "openballbot_rl_b200.rewards.base",
    "ballbot_gym.rewards.directional": "openballbot_rl_b200.rewards.directional",
    "ballbot_gym.rewards.distance": "openballbot_rl_b200.rewards.distance",
    "ballbot_gym.terrain": "openballbot_rl_b200.terrain",
    "ballbot_gym.envs": "openballbot_rl_b200.envs",
    "ballbot_gym.envs.ballbot_env": "openballbot_rl_b200.envs.ballbot_env",
    "ballbot_gym.envs.observation_spaces": "openballbot_rl_b200.envs.spaces",
}
for _alias, _real in _MAP.items():
    _mod = importlib.import_module(_real)
    sys.modules[_alias] = _mod
    _parent, _, _leaf = _alias.rpartition(".")
    if _parent == "ballbot_gym":
        globals()[_leaf] = _mod

try:  # pragma: no cover - gymnasium is optional
    from gymnasium.envs.registration import register, registry as _registry
    if "ballbot-v0.1" not in _registry:
        register(id="ballbot-v0.1", entry_point="openballbot_rl_b200.envs.ballbot_env:BBotSimulation", kwargs={"xml_path": None})
except Exception:
    pass
