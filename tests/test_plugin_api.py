"""Host-side plugin API (registry / factories / config / rewards / terrains) against golden fixtures generated from the
reference's own modules (tests/golden/make_golden.py).  Mirrors the assertions of the reference's tests/unit/*."""
import copy
import json
import os

import numpy as np
import pytest

from openballbot_rl_b200.core import (ComponentRegistry, create_policy, create_reward, create_terrain, get_component_config,
                                      load_config, load_training_config, merge_configs, validate_config)
from openballbot_rl_b200.rewards import BaseReward, DirectionalReward, DistanceReward, register_builtin_rewards
from openballbot_rl_b200.terrain import register_builtin_terrains, shapes

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "plugin_api.json")))
TERR = np.load(os.path.join(HERE, "golden", "terrains.npz"))


@pytest.fixture(autouse=True)
def _fresh_registry():
    ComponentRegistry.clear()
    register_builtin_rewards()
    register_builtin_terrains()
    yield
    ComponentRegistry.clear()
    register_builtin_rewards()
    register_builtin_terrains()


def _raises_like(label, fn):
    exp = GOLD["errors"][label]
    with pytest.raises(Exception) as ei:
        fn()
    assert type(ei.value).__name__ == exp["type"], (label, ei.value)
    return str(ei.value), exp["message"]


def test_registry_lists_all_reference_components():
    assert set(ComponentRegistry.list_rewards()) == {"directional", "distance"}
    assert set(ComponentRegistry.list_terrains()) == {"perlin", "flat", "stepped", "ramp", "sinusoidal", "ridge_valley", "hills", "bowl",
                                                      "gradient", "terraced", "wavy", "spiral", "mixed"}


def test_registry_error_messages_match_reference():
    got, exp = _raises_like("unknown_reward", lambda: ComponentRegistry.get_reward("nope"))
    assert got == exp
    got, exp = _raises_like("dup_reward", lambda: ComponentRegistry.register_reward("directional", DirectionalReward))
    assert "already registered" in got and "already registered" in exp
    got, exp = _raises_like("bad_reward_class", lambda: ComponentRegistry.register_reward("x", dict))
    assert "must inherit from BaseReward" in got
    got, exp = _raises_like("unknown_terrain", lambda: ComponentRegistry.get_terrain("nope"))
    assert got.startswith("Unknown terrain: 'nope'")
    got, exp = _raises_like("dup_terrain", lambda: ComponentRegistry.register_terrain("flat", lambda n: None))
    assert "already registered" in got
    got, exp = _raises_like("terrain_not_callable", lambda: ComponentRegistry.register_terrain("y", 3))
    assert got == exp
    got, exp = _raises_like("unknown_policy", lambda: ComponentRegistry.get_policy("nope"))
    assert got == exp
    got, exp = _raises_like("unknown_sensor", lambda: ComponentRegistry.get_sensor("nope"))
    assert got == exp


def test_registry_register_get_clear():
    class MyReward(BaseReward):
        def __init__(self, k=1.0):
            self.k = k

        def __call__(self, state):
            return self.k

    ComponentRegistry.register_reward("mine", MyReward)
    assert ComponentRegistry.get_reward("mine", k=3.0)(dict()) == 3.0
    ComponentRegistry.register_policy("p", dict); ComponentRegistry.register_sensor("s", list)
    assert ComponentRegistry.get_policy("p") is dict and ComponentRegistry.get_sensor("s") is list
    assert create_policy({"type": "p"}) is dict
    ComponentRegistry.clear()
    assert ComponentRegistry.list_rewards() == [] and ComponentRegistry.list_terrains() == [] and ComponentRegistry.list_policies() == []


def test_factory_errors_match_reference():
    for label, fn in {
        "reward_cfg_not_dict": lambda: create_reward("directional"),
        "reward_cfg_no_type": lambda: create_reward({"config": {}}),
        "directional_missing": lambda: create_reward({"type": "directional", "config": {}}),
        "distance_missing": lambda: create_reward({"type": "distance", "config": {}}),
        "validate_bad_component": lambda: validate_config({"type": "flat"}, "widget"),
        "validate_no_type": lambda: validate_config({}, "terrain"),
    }.items():
        got, exp = _raises_like(label, fn)
        assert got == exp, label
    for label, fn in {"reward_unknown_type": lambda: create_reward({"type": "zzz", "config": {}}),
                      "terrain_unknown_type": lambda: create_terrain({"type": "zzz"}),
                      "validate_unknown_reward": lambda: validate_config({"type": "zzz"}, "reward")}.items():
        got, exp = _raises_like(label, fn)
        assert got.split("Available")[0] == exp.split("Available")[0], label     # same text up to the list of names
    assert validate_config({"type": "directional"}, "reward") is True and validate_config({"type": "hills"}, "terrain") is True


def test_rewards_match_reference_values():
    rd = create_reward({"type": "directional", "config": {"target_direction": [0.6, -0.8], "scale": 0.5}})
    assert isinstance(rd, DirectionalReward) and rd.target_direction.dtype == np.float32
    for case in GOLD["directional"]:
        assert float(rd({"vel": np.array(case["vel"], dtype=np.float32)})) == case["value"]
    dd = create_reward({"type": "distance", "config": {"goal_position": [1.0, -2.0], "scale": 0.5}})
    assert isinstance(dd, DistanceReward)
    for case in GOLD["distance"]:
        assert float(dd({"pos2d": np.array(case["pos2d"], dtype=np.float32)})) == case["value"]
    got, exp = _raises_like("distance_no_pos2d", lambda: dd({"vel": np.zeros(3)}))
    assert got == exp
    with pytest.raises(ValueError):
        DistanceReward(goal_position=[1.0, 2.0, 3.0])
    # sign convention of the reference's test_rewards.py
    r = DirectionalReward(target_direction=np.array([0.0, 1.0]))
    assert r({"vel": np.array([0.0, 0.5, 0.0])}) > 0 and r({"vel": np.array([0.0, -0.5, 0.0])}) < 0


def test_config_helpers_match_reference(tmp_path):
    m = GOLD["merge"]
    assert merge_configs({"a": 1, "b": {"c": 2, "d": {"e": 3}}, "f": [1]}, {"b": {"d": {"e": 4, "g": 5}, "h": 6}, "f": [2], "i": 7}) == m
    for case in GOLD["get_component_config"]:
        assert get_component_config(copy.deepcopy(case["config"]), case["component"], case["default"]) == case["result"]
    got, exp = _raises_like("component_cfg_missing", lambda: get_component_config({"problem": {"reward": {"config": {}}}}, "reward"))
    assert got == exp
    # load_training_config: env YAML is the base, training YAML overrides, terrain/reward mirrored under problem
    (tmp_path / "configs" / "env").mkdir(parents=True); (tmp_path / "configs" / "train").mkdir(parents=True)
    (tmp_path / "configs" / "env" / "e.yaml").write_text("terrain:\n  type: perlin\n  config: {scale: 25.0}\nreward:\n  type: directional\n  config: {target_direction: [0.0, 1.0]}\nenv: {max_ep_steps: 4000}\n")
    (tmp_path / "configs" / "train" / "t.yaml").write_text("env_config: env/e.yaml\nenv: {max_ep_steps: 100}\nnum_envs: 10\n")
    cfg = load_training_config(str(tmp_path / "configs" / "train" / "t.yaml"))
    assert cfg["env"]["max_ep_steps"] == 100 and cfg["num_envs"] == 10 and "env_config" not in cfg
    assert cfg["problem"]["terrain"]["type"] == "perlin" and cfg["problem"]["reward"]["type"] == "directional"
    (tmp_path / "bad.yaml").write_text("num_envs: 1\n")
    with pytest.raises(ValueError, match="env_config"):
        load_training_config(str(tmp_path / "bad.yaml"))
    with pytest.raises(FileNotFoundError):
        load_config(str(tmp_path / "missing.yaml"))


@pytest.mark.parametrize("case", GOLD["terrain_cases"], ids=lambda c: f"{c['name']}-{c['key']}")
def test_host_terrains_match_reference(case):
    ref = TERR[case["key"]]
    out = shapes.GENERATORS[case["name"]](case["n"], **copy.deepcopy(case["kwargs"]))
    assert out.shape == (case["n"] * case["n"],)
    assert out.min() >= 0.0 and out.max() <= 1.0
    tol = 1e-6 if case.get("dtype") == "float32" else 1e-12
    np.testing.assert_allclose(out, ref, rtol=0, atol=tol)


def test_create_terrain_runtime_seed_overrides_config():
    gen = create_terrain({"type": "hills", "config": {"num_hills": 3, "seed": 1}})
    np.testing.assert_allclose(gen(33, seed=3), TERR["factory_hills_seed3"], atol=1e-12)
    assert gen.terrain_type == "hills"
    flat = create_terrain({"type": "flat"})(129)
    assert flat.shape == (129 * 129,) and not flat.any()


def test_gradient_perlin_variant_runs_the_device_noise():
    """gradient_type="perlin" evaluates the untiled 2-D snoise2 with the engine's device code (bb_snoise2_grid): like the perlin
    terrain it needs a CUDA device and has no CPU implementation in the product (GPU parity: tests/test_gpu_envs.py)."""
    import torch
    from openballbot_rl_b200._lib import EngineError
    if torch.cuda.is_available():
        pytest.skip("CUDA present: covered by the GPU parity test")
    with pytest.raises(EngineError):
        shapes.generate_gradient_terrain(33, gradient_type="perlin")


def test_drop_in_alias_packages():
    import ballbot_gym
    import ballbot_rl
    from ballbot_gym.core.registry import ComponentRegistry as R2
    from ballbot_gym.core.factories import create_reward as cr2
    from ballbot_gym.envs.ballbot_env import BBotSimulation
    from ballbot_rl.training.utils import make_ballbot_env
    assert R2 is ComponentRegistry and cr2 is create_reward
    assert callable(make_ballbot_env(terrain_type="flat")) and BBotSimulation.metadata["render_modes"] == ["rgb_array"]
    assert ballbot_gym.core is not None and ballbot_rl.training is not None


def test_policy_plugin_mlp_is_registered_and_builds_the_reference_extractor():
    """ballbot_rl/policies/__init__.py:8 registers the feature extractor as policy plugin "mlp"; `create_policy` returns the class
    (factories.py:129-162) and the instance has the reference's layout: per-key extractors in the observation space's
    (alphabetical) order, 56 features with cameras, state-dict keys `extractors.rgbd_k.*` as in an SB3 policy.pth."""
    import torch
    import importlib
    import ballbot_rl  # noqa: F401  (alias package: importing it registers the plugin, like the reference)
    import openballbot_rl_b200.policies as _pol
    importlib.reload(_pol)             # other tests clear the registry: registration happens at import time, as in the reference
    from ballbot_gym.core.factories import create_policy
    from ballbot_gym.core.registry import ComponentRegistry
    from openballbot_rl_b200.envs.spaces import create_observation_space
    assert "mlp" in ComponentRegistry.list_policies()
    cls = create_policy({"type": "mlp", "config": {"hidden_sizes": [128] * 4}})
    assert cls.__name__ == "Extractor"
    sp = create_observation_space({"h": 64, "w": 64}, 1, False)
    ex = cls(sp).eval()
    assert list(ex.extractors.keys()) == ["actions", "angular_vel", "motor_state", "orientation", "relative_image_timestamp", "rgbd_0", "rgbd_1", "vel"]
    assert ex.features_dim == 56 and "extractors.rgbd_0.7.weight" in ex.state_dict()
    obs = {k: torch.zeros((3,) + tuple(v.shape)) for k, v in sp.spaces.items()}
    assert ex(obs).shape == (3, 56)
    sp2 = create_observation_space({"h": 64, "w": 64}, 1, True)
    assert cls(sp2).features_dim == 16                       # camera-less space still declares the timestamp (App. C #6)
    with pytest.raises(ValueError, match="Failed to get policy"):
        create_policy({"type": "transformer"})
