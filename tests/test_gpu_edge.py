"""Edge cases of the C-ABI engine on the GPU: invalid configs, tiny / odd batches, non-finite states and actions, reset
masks, and deep impacts that overflow the shared-memory contact records (parity with the oracle on the overflow path)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

QPOS0 = np.array([0, 0, 0.24, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.26, 1, 0, 0, 0], np.float64)


def _rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(1.0, float(np.abs(b).max())))


def test_invalid_configs_are_rejected_with_a_message():
    from openballbot_rl_b200 import _lib
    L = _lib.lib()
    for field, value in (("num_envs", 0), ("precision", 16), ("im_h", 0), ("abi_version", 99)):
        cfg = _lib.default_config(); setattr(cfg, field, value)
        h = ctypes.c_void_p()
        assert L.bb_create(ctypes.byref(cfg), ctypes.byref(h)) == -1 and not h.value          # BB_ERR_INVALID
        assert b"invalid config" in L.bb_last_error(None)
    from openballbot_rl_b200.engine import BallbotEngine
    eng = BallbotEngine(num_envs=4, terrain="flat", cameras=False)
    with pytest.raises(ValueError):
        eng.step(torch.zeros(5, 3, device="cuda"))                                            # wrong batch size
    io = _lib.IO()                                                                            # all output pointers NULL
    assert L.bb_step(eng._h, ctypes.c_void_p(torch.zeros(4, 3, device="cuda").data_ptr()), ctypes.byref(io), None) == -1
    assert b"NULL" in L.bb_last_error(eng._h)
    eng.close()


@pytest.mark.parametrize("n", [1, 3, 33])
def test_tiny_and_odd_batches_match_the_oracle(oracle_mod, n):
    """Fewer envs than one CTA of the stage kernel / an odd count (the last warp carries one env)."""
    from openballbot_rl_b200.engine import BallbotEngine
    eng = BallbotEngine(num_envs=n, precision=64, terrain="flat", cameras=False, auto_reset=False)
    eng.reset()
    ora = oracle_mod.OracleEnv(); ora.reset()
    rng = np.random.default_rng(n)
    for t in range(110):                                    # long enough for the robot to land and roll on its contacts
        a = rng.uniform(-1, 1, 3).astype(np.float32)
        eng.step(torch.from_numpy(np.tile(a, (n, 1))).cuda())
        ora.step(a)
    q, v, _ = [x.cpu().numpy() for x in eng.get_state()]
    qo, vo, _, _ = ora.get_state()
    for i in range(n):
        assert np.abs(q[i] - qo).max() < 1e-8 and np.abs(v[i] - vo).max() < 1e-7, i
    eng.close()


def test_non_finite_state_or_action_fails_one_env_only():
    """MuJoCo's mj_checkPos / mj_checkVel behaviour per env: flag, report terminated + failure, reset; neighbours untouched."""
    from openballbot_rl_b200.engine import BallbotEngine
    N = 40
    kw = dict(num_envs=N, precision=64, terrain="perlin", cameras=True, seed=2)
    e1, e2 = BallbotEngine(**kw), BallbotEngine(**kw)
    e1.reset(); e2.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for t in range(30):
        a = torch.rand(N, 3, device="cuda", generator=g) * 2 - 1
        e1.step(a); e2.step(a)
    q, v, w = e1.get_state()
    q[7, 2] = float("nan"); v[21, 5] = float("inf")
    e1.set_state(q, v, w)
    a = torch.rand(N, 3, device="cuda", generator=g) * 2 - 1
    a2 = a.clone(); a2[30, 1] = float("nan")                # a NaN action poisons the control of env 30 only
    e1.step(a2); e2.step(a)
    st = e1.status.cpu().numpy()
    bad = [7, 21, 30]
    assert all(st[i] & 1 for i in bad) and not any(st[i] & 1 for i in range(N) if i not in bad)
    assert all(bool(e1.terminated[i]) and bool(e1.failure[i]) for i in bad)
    ok = [i for i in range(N) if i not in bad]
    (q1, v1, _), (q2, v2, _) = e1.get_state(), e2.get_state()
    assert torch.equal(q1[ok], q2[ok]) and torch.equal(v1[ok], v2[ok]) and torch.equal(e1.reward[ok], e2.reward[ok])
    assert torch.isfinite(q1).all() and torch.isfinite(v1).all()        # the failed envs were reset to a valid state
    for t in range(5):                                                  # and keep stepping normally afterwards
        a = torch.rand(N, 3, device="cuda", generator=g) * 2 - 1
        e1.step(a)
    assert not bool((e1.status & 1).any()) and torch.isfinite(e1.obs["rgbd_0"]).all()
    e1.close(); e2.close()


def test_reset_masks():
    from openballbot_rl_b200.engine import BallbotEngine
    N = 12
    eng = BallbotEngine(num_envs=N, precision=64, terrain="perlin", cameras=False, auto_reset=False, seed=8)
    eng.reset()
    for t in range(20):
        eng.step(torch.full((N, 3), 0.5, device="cuda"))
    q_before = eng.get_state()[0].clone(); seeds_before = eng.terrain_seeds().clone()
    eng.reset(torch.zeros(N, dtype=torch.uint8, device="cuda"))          # empty mask: nothing changes
    assert torch.equal(eng.get_state()[0], q_before) and torch.equal(eng.terrain_seeds(), seeds_before)
    mask = torch.zeros(N, dtype=torch.bool, device="cuda"); mask[[0, 5, 11]] = True
    eng.reset(mask)
    q_after = eng.get_state()[0]
    keep = ~mask
    assert torch.equal(q_after[keep], q_before[keep])
    assert float((q_after[mask][:, 7:10]).abs().max()) == 0.0 and float((q_after[mask][:, 3] - 1).abs().max()) == 0.0   # qpos0 pose
    assert float((q_after[mask][:, 12] - q_after[mask][:, 2] - 0.02).abs().max()) < 1e-12    # ball 2 cm above the base origin
    eng.close()


def test_landing_impacts_overflow_contact_records_and_match_oracle(oracle_mod):
    """More than 3 + 8 simultaneous contacts: the solver's records spill to the global scratch.  That happens when the robot
    lands on rough terrain after a reset (the ball touches a dozen prisms for a few steps).  Trajectories from reset through
    the impact against the oracle in lock-step; the drift bound after 110 steps is 1e-6 (fp64)."""
    from openballbot_rl_b200.engine import BallbotEngine
    N, T = 32, 110
    eng = BallbotEngine(num_envs=N, precision=64, terrain="perlin", cameras=False, auto_reset=False, seed=5)
    eng.reset()
    oras = []
    for i in range(N):
        o = oracle_mod.OracleEnv(); o.reset(eng.get_hfield(i).cpu().numpy()); oras.append(o)
    rng = np.random.default_rng(9)
    ncmax = np.zeros(N, int); alive = np.ones(N, bool)
    for t in range(T):
        a = rng.uniform(-1, 1, (N, 3)).astype(np.float32)
        eng.step(torch.from_numpy(a).cuda())
        ncmax = np.maximum(ncmax, (eng.status.cpu().numpy() >> 8) & 255)
        term = eng.terminated.cpu().numpy().astype(bool)
        for i in range(N):
            if alive[i]:
                _, _, to, _, _ = oras[i].step(a[i])
                assert bool(to) == bool(term[i]), (i, t)
        alive &= ~term                                        # auto_reset is off: a terminated env is no longer compared
    assert ncmax.max() > 11, ncmax                            # the overflow path was exercised
    q1, v1, _ = [x.cpu().numpy() for x in eng.get_state()]
    worst = 0.0
    for i in range(N):
        if alive[i]:
            qo, vo, _, _ = oras[i].get_state()
            worst = max(worst, np.abs(q1[i] - qo).max(), 0.1 * np.abs(v1[i] - vo).max())
    assert alive.sum() >= N // 2 and worst < 1e-6, (worst, ncmax)
    eng.close()
