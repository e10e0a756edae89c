"""YAML config loading / merging with the reference's rules (ballbot_gym/core/config.py:7-197)."""
from pathlib import Path
from typing import Any, Dict, Optional

import yaml


def load_config(config_path: str) -> Dict[str, Any]:
    path = Path(config_path)
    if not path.exists():
        raise FileNotFoundError(f"Configuration file not found: {config_path}")
    with path.open("r") as fh:
        return yaml.safe_load(fh) or {}


def merge_configs(base: Dict[str, Any], override: Dict[str, Any]) -> Dict[str, Any]:
    """Deep merge, ``override`` wins; dict values are merged recursively (config.py:34-53)."""
    out = dict(base)
    for key, val in override.items():
        both_dicts = isinstance(out.get(key), dict) and isinstance(val, dict)
        out[key] = merge_configs(out[key], val) if both_dicts else val
    return out


def load_training_config(config_path: str) -> Dict[str, Any]:
    """Training YAML + the env YAML it references through ``env_config`` (config.py:56-135): the env config is the base,
    the training config overrides it, terrain/reward are mirrored under ``problem`` and ``env_config`` is dropped."""
    cfg = load_config(config_path)
    env_ref = cfg.get("env_config")
    if not env_ref:
        raise ValueError("Training config must specify 'env_config' key pointing to an environment config.\n"
                         "Example: env_config: 'configs/env/perlin_directional.yaml'\n"
                         f"Config file: {config_path}")
    env_path = Path(env_ref)
    if not env_path.is_absolute():
        # "configs/..." resolves from the current working directory, anything else from the training config's parent dir
        env_path = (Path.cwd() / env_ref) if env_ref.startswith("configs/") else (Path(config_path).parent.parent / env_ref)
    env_cfg = load_config(str(env_path))
    merged = merge_configs(env_cfg, cfg)
    problem = merged.setdefault("problem", {})
    for key in ("terrain", "reward"):
        if key in env_cfg and key not in problem:
            problem[key] = env_cfg[key]
    merged.pop("env_config", None)
    return merged


def get_component_config(config: Dict[str, Any], component_type: str, default_type: Optional[str] = None) -> Dict[str, Any]:
    """``problem.<type>`` first, then top-level ``<type>``; a bare string is a type name; ``default_type`` fills gaps
    (config.py:138-197)."""
    comp = config.get("problem", {}).get(component_type, {}) or config.get(component_type, {})
    if isinstance(comp, str):
        return {"type": comp, "config": {}}
    if not comp and default_type:
        return {"type": default_type, "config": {}}
    if not isinstance(comp, dict) or "type" not in comp:
        if default_type:
            return {"type": default_type, "config": comp if isinstance(comp, dict) else {}}
        raise ValueError(f"Component config for '{component_type}' must have 'type' key or be a string, got: {comp}")
    comp.setdefault("config", {})
    return comp
