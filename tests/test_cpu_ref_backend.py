"""The CPU oracle behind the engine's own C ABI (oracle/bb_cpu_ref.cpp, "backend = cpu_ref"): exports the same symbols as
libballbot_b200.so and runs the same batch semantics.  No GPU needed."""
import ctypes
import os

import numpy as np

from tests.backend_harness import Backend, pcg64_states
from tests.test_abi import _declared_symbols


def test_cpu_ref_library_exports_every_abi_symbol(oracle_mod):
    lib = ctypes.CDLL(oracle_mod.CPU_REF_SO)
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_cpu_ref_batch_semantics(oracle_mod):
    """N envs behind bb_step: same trajectories as N stand-alone oracle envs, auto-reset with terminal observation and Monitor
    statistics, numpy-compatible PCG64 terrain seeds, explicit seeds in bb_reset."""
    N = 3
    b = Backend("cpu_ref", num_envs=N, terrain_type=1, cameras=0, max_ep_steps=12, seed_stream=1, auto_reset=1)
    env_seeds = [20, 21, 22]
    b.set_rng_state(pcg64_states(env_seeds))
    gens = [np.random.default_rng(s) for s in env_seeds]
    o = b.reset()
    first = [int(g.integers(0, 10000)) for g in gens]
    assert b.terrain_seeds().tolist() == first and not o["orientation"].any()
    ref = []
    for i in range(N):
        e = oracle_mod.OracleEnv(max_ep_steps=12); e.reset(oracle_mod.perlin_terrain(seed=first[i])); ref.append(e)
        np.testing.assert_array_equal(b.get_hfield(i), oracle_mod.perlin_terrain(seed=first[i]))
    rng = np.random.default_rng(0)
    G = np.zeros(N)
    for t in range(12):
        a = rng.uniform(-1, 1, (N, 3)).astype(np.float32)
        o = b.step(a)
        for i, e in enumerate(ref):
            ob, r, term, fail, info = e.step(a[i]); G[i] += r
            assert o["reward"][i] == np.float32(r) and bool(o["terminated"][i]) == term
            if not term:
                np.testing.assert_array_equal(o["orientation"][i], ob[0:3])
    assert o["terminated"].all() and (o["episode_length"] == 12).all()            # timeout => terminated (ballbot_env.py:982)
    np.testing.assert_allclose(o["episode_return"], G, rtol=1e-5)
    assert b.terrain_seeds().tolist() == [int(g.integers(0, 10000)) for g in gens]   # auto-reset drew the next seeds
    assert not o["orientation"].any() and o["terminal_obs"].any()                  # reset obs returned, terminal obs kept
    o = b.reset(mask=np.array([0, 1, 0], np.uint8), seeds=np.array([1, 4242, 3], np.int32))
    assert b.terrain_seeds()[1] == 4242
    np.testing.assert_array_equal(b.get_hfield(1), oracle_mod.perlin_terrain(seed=4242))
    rows = b.contacts(0)
    assert rows.shape[1] == 14
    b.close()
