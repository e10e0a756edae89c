"""The C-ABI library builds, loads without a GPU and exports every symbol include/ballbot_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ballbot_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bb_[a-z0-9_]+)\s*\(", text)))


def test_header_and_library_agree():
    from openballbot_rl_b200 import _lib
    _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_lib.EXPORTED) == declared              # the Python binding covers exactly the declared API


def test_config_struct_layout_matches_header():
    from openballbot_rl_b200 import _lib
    cfg = _lib.default_config()                           # bb_default_config runs on the host, no GPU needed
    assert cfg.abi_version == 2 and cfg.perlin_table == -1 and cfg.seed_stream == 0 and cfg.precision == 64 and cfg.perlin_octaves == 4 and cfg.max_ep_steps == 4000
    assert abs(cfg.perlin_scale - 25.0) < 1e-6 and abs(cfg.reward_scale - 0.01) < 1e-9 and abs(cfg.survival_bonus - 0.02) < 1e-9
    assert abs(cfg.action_reg_coef + 1e-4) < 1e-9 and cfg.target_direction[1] == 1.0 and cfg.auto_reset == 1
    assert cfg.step_kernel == 0 and cfg.solver_mode == 0 and cfg.im_h == 64 and abs(cfg.hfield_zscale - 2.0) < 1e-6


def test_build_info_names_the_current_sources():
    from openballbot_rl_b200 import _lib
    _lib.build()
    info = _lib.lib().bb_build_info().decode()
    assert info.startswith("libballbot_b200 abi 2 built ") and info.endswith("src " + _lib.source_hash()), info


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present: the refusal path is exercised on CPU-only hosts")
    from openballbot_rl_b200.engine import BallbotEngine, EngineError
    with pytest.raises(EngineError, match="no CPU fallback"):
        BallbotEngine(num_envs=4)
    from openballbot_rl_b200 import _lib
    cfg = _lib.default_config(); h = ctypes.c_void_p()
    assert _lib.lib().bb_create(ctypes.byref(cfg), ctypes.byref(h)) == -3            # BB_ERR_NO_DEVICE
    assert b"no CPU fallback" in _lib.lib().bb_last_error(None)
    from openballbot_rl_b200.terrain import generate_perlin_terrain
    with pytest.raises(EngineError):
        generate_perlin_terrain(33, seed=1)


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing in the product package may import, include, link or call it."""
    pkg = os.path.join(ROOT, "openballbot_rl_b200")
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b)|(\boracle\.(oracle|OracleEnv|lib)\b)|(#include\s+\"[^\"]*oracle)|(\bbbo_[a-z_]+\s*\()|(libballbot_oracle)", re.M)
    for top in (pkg, os.path.join(ROOT, "ballbot_gym"), os.path.join(ROOT, "ballbot_rl")):
        for dirpath, _, files in os.walk(top):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not bad.search(src), os.path.join(dirpath, f)


def test_torch_ops_library_builds_and_registers_the_ballbot_namespace():
    """the PyTorch extension over the C ABI (csrc/bb_torch_ops.cpp) builds in-tree and registers
    torch.ops.ballbot.{step, reset, add_reward, gae} with mutable-output schemas (no compute without a GPU)."""
    import torch
    from openballbot_rl_b200 import _lib
    ops = _lib.torch_ops()
    for name in ("step", "reset", "add_reward", "gae"):
        assert hasattr(ops, name)
    schema = str(torch.ops.ballbot.step.default._schema)
    assert "Tensor(a!)[] outs" in schema and "int engine" in schema, schema
    with pytest.raises(Exception):
        ops.step(0, torch.zeros(4, 3), [torch.zeros(1)] * 16)        # CPU tensors: no kernel registered for the CPU backend (no fallback)
