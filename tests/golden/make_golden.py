"""Generates tests/golden/*.npz|json from the REFERENCE's own Python modules (run in the build container only).

The reference's pure-numpy modules (registry, factories, config, rewards, 11 of the 13 terrain generators) are imported
by file path from /root/reference with stub parent packages (its package __init__ needs gymnasium/mujoco/noise, which
are not installed).  The outputs are committed as small fixtures; nothing under tests/ reads /root/reference at run time.

    python tests/golden/make_golden.py
"""
import copy
import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def main():
    for pk in ["ballbot_gym", "ballbot_gym.core", "ballbot_gym.rewards", "ballbot_gym.terrain"]:
        sys.modules[pk] = types.ModuleType(pk)
        sys.modules[pk].__path__ = []
    base = load(f"{REF}/ballbot_gym/rewards/base.py", "ballbot_gym.rewards.base")
    reg = load(f"{REF}/ballbot_gym/core/registry.py", "ballbot_gym.core.registry")
    fac = load(f"{REF}/ballbot_gym/core/factories.py", "ballbot_gym.core.factories")
    cfgm = load(f"{REF}/ballbot_gym/core/config.py", "ballbot_gym.core.config")
    dirr = load(f"{REF}/ballbot_gym/rewards/directional.py", "ballbot_gym.rewards.directional")
    dist = load(f"{REF}/ballbot_gym/rewards/distance.py", "ballbot_gym.rewards.distance")
    R = reg.ComponentRegistry
    R.register_reward("directional", dirr.DirectionalReward)
    R.register_reward("distance", dist.DistanceReward)
    names = ["stepped", "ramp", "sinusoidal", "ridge_valley", "hills", "bowl", "gradient", "terraced", "wavy", "spiral", "mixed"]
    gens = {}
    for n in names:
        m = load(f"{REF}/ballbot_gym/terrain/{n}.py", "ballbot_gym.terrain." + n)
        gens[n] = getattr(m, f"generate_{n}_terrain")
        R.register_terrain(n, gens[n])
    R.register_terrain("flat", lambda n, **kw: np.zeros(n * n))

    # ---- terrains (n = 33 keeps the fixtures small; one full-size 293 case for the spawn-window logic)
    cases = [
        ("stepped", {}), ("stepped", {"num_steps": 3, "step_height": 0.2}),
        ("ramp", {}), ("ramp", {"ramp_direction": "y", "num_ramps": 2, "flat_ratio": 0.2}), ("ramp", {"ramp_direction": "radial", "ramp_angle": 30.0}),
        ("sinusoidal", {}), ("sinusoidal", {"direction": "y", "frequency": 0.05, "phase": 0.7}),
        ("ridge_valley", {}), ("ridge_valley", {"orientation": "diagonal", "smoothness": 0.6, "spacing": 1.5}),
        ("hills", {}), ("hills", {"seed": 42, "num_hills": 8, "hill_radius": 0.1}),
        ("bowl", {}), ("bowl", {"depth": 0.9, "radius": 0.25, "center_x": 0.3, "center_y": 0.6}),
        ("gradient", {}), ("gradient", {"gradient_type": "radial", "max_slope": 35.0}), ("gradient", {"direction": "y"}),
        ("terraced", {}), ("terraced", {"direction": "y", "num_terraces": 4, "transition_width": 0.3}),
        ("wavy", {}), ("wavy", {"wave_amplitudes": [0.25, 0.1], "wave_frequencies": [0.3, 1.1], "wave_directions": [10.0, 80.0], "phase_offsets": [0.1, 0.2]}),
        ("spiral", {}), ("spiral", {"direction": "ccw", "spiral_tightness": 1.5, "height_variation": 0.8}),
        ("mixed", {"components": [{"type": "hills", "weight": 0.6, "config": {"num_hills": 4}}, {"type": "wavy", "weight": 0.4}], "seed": 5}),
        ("mixed", {"components": [{"type": "bowl", "weight": 1.0}, {"type": "terraced", "weight": 0.5}], "blend_mode": "max"}),
        ("mixed", {"components": [{"type": "ramp", "weight": 2.0}, {"type": "sinusoidal", "weight": 1.0}], "blend_mode": "weighted"}),
    ]
    arrays, meta = {}, []
    for k, (name, kw) in enumerate(cases):
        arrays[f"t{k}"] = gens[name](33, **copy.deepcopy(kw)).astype(np.float64)
        meta.append({"key": f"t{k}", "name": name, "n": 33, "kwargs": kw})
    arrays["hills293"] = gens["hills"](293, seed=7).astype(np.float32)
    meta.append({"key": "hills293", "name": "hills", "n": 293, "kwargs": {"seed": 7}, "dtype": "float32"})
    # create_terrain closure with a runtime seed override (factories.py:121-124)
    arrays["factory_hills_seed3"] = fac.create_terrain({"type": "hills", "config": {"num_hills": 3, "seed": 1}})(33, seed=3)
    np.savez_compressed(os.path.join(OUT, "terrains.npz"), **arrays)

    # ---- rewards / registry / factories / config behaviour
    g = {"terrain_cases": meta}
    rd = fac.create_reward({"type": "directional", "config": {"target_direction": [0.6, -0.8], "scale": 0.5}})
    states = [[0.5, 0.3, 0.0], [-1.5, 2.0, 0.25], [0.0, 0.0, 1.0]]
    g["directional"] = [{"vel": s, "value": float(rd({"vel": np.array(s, dtype=np.float32)}))} for s in states]
    dd = fac.create_reward({"type": "distance", "config": {"goal_position": [1.0, -2.0], "scale": 0.5}})
    g["distance"] = [{"pos2d": p, "value": float(dd({"pos2d": np.array(p, dtype=np.float32)}))} for p in ([0.0, 0.0], [1.0, -2.0], [3.5, 0.25])]
    msgs = {}
    for label, fn in {
        "unknown_reward": lambda: R.get_reward("nope"),
        "dup_reward": lambda: R.register_reward("directional", dirr.DirectionalReward),
        "bad_reward_class": lambda: R.register_reward("x", dict),
        "unknown_terrain": lambda: R.get_terrain("nope"),
        "dup_terrain": lambda: R.register_terrain("flat", lambda n: None),
        "terrain_not_callable": lambda: R.register_terrain("y", 3),
        "unknown_policy": lambda: R.get_policy("nope"),
        "unknown_sensor": lambda: R.get_sensor("nope"),
        "reward_cfg_not_dict": lambda: fac.create_reward("directional"),
        "reward_cfg_no_type": lambda: fac.create_reward({"config": {}}),
        "directional_missing": lambda: fac.create_reward({"type": "directional", "config": {}}),
        "distance_missing": lambda: fac.create_reward({"type": "distance", "config": {}}),
        "reward_unknown_type": lambda: fac.create_reward({"type": "zzz", "config": {}}),
        "terrain_unknown_type": lambda: fac.create_terrain({"type": "zzz"}),
        "validate_unknown_reward": lambda: fac.validate_config({"type": "zzz"}, "reward"),
        "validate_bad_component": lambda: fac.validate_config({"type": "flat"}, "widget"),
        "validate_no_type": lambda: fac.validate_config({}, "terrain"),
        "distance_no_pos2d": lambda: dd({"vel": np.zeros(3)}),
        "component_cfg_missing": lambda: cfgm.get_component_config({"problem": {"reward": {"config": {}}}}, "reward"),
    }.items():
        try:
            fn()
            msgs[label] = None
        except Exception as e:  # noqa: BLE001
            msgs[label] = {"type": type(e).__name__, "message": str(e)}
    g["errors"] = msgs
    g["merge"] = cfgm.merge_configs({"a": 1, "b": {"c": 2, "d": {"e": 3}}, "f": [1]}, {"b": {"d": {"e": 4, "g": 5}, "h": 6}, "f": [2], "i": 7})
    gcc_inputs = [({"problem": {"terrain": {"type": "perlin", "config": {"scale": 25.0}}}}, "terrain", None),
                  ({"terrain": {"type": "flat"}}, "terrain", None), ({"problem": {"reward": "directional"}}, "reward", None),
                  ({}, "terrain", "perlin"), ({"problem": {"terrain": {"scale": 10}}}, "terrain", "perlin")]
    g["get_component_config"] = [{"config": c, "component": t, "default": d, "result": cfgm.get_component_config(copy.deepcopy(c), t, d)} for c, t, d in gcc_inputs]
    g["validate_ok"] = [fac.validate_config({"type": "directional"}, "reward"), fac.validate_config({"type": "hills"}, "terrain")]
    with open(os.path.join(OUT, "plugin_api.json"), "w") as fh:
        json.dump(g, fh, indent=1, sort_keys=True)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
