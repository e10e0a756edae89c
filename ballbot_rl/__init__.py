"""Drop-in alias of the hot-path part of the reference's ``ballbot_rl`` package: ``ballbot_rl.training.utils``
(make_ballbot_env, reference ballbot_rl/training/utils.py:11-85) backed by the B200 engine."""
import importlib
import sys

for _alias, _real in {"ballbot_rl.training": "openballbot_rl_b200.training",
                      "ballbot_rl.training.utils": "openballbot_rl_b200.training.utils",
                      "ballbot_rl.policies": "openballbot_rl_b200.policies",                 # registers policy plugin "mlp"
                      "ballbot_rl.policies.mlp_policy": "openballbot_rl_b200.policies.mlp_policy"}.items():
    _mod = importlib.import_module(_real)
    sys.modules[_alias] = _mod
training = sys.modules["ballbot_rl.training"]
policies = sys.modules["ballbot_rl.policies"]
