"""The engine's own physics core (openballbot_rl_b200/csrc/bb_core.cuh), compiled for the CPU by tests/hostcore, against
the oracle: two independent formulations (base-local classical Newton-Euler + matrix-free contacts vs MuJoCo-shaped
c-frame spatial algebra + dense efc_J) must agree to rounding.  This is the no-GPU half of the parity story; the GPU
half (tests/test_gpu_parity.py) runs the same core inside the CUDA kernels."""
import numpy as np
import pytest

from tests.hostcore import hostcore as H

QPOS0 = np.array([0, 0, 0.24, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.26, 1, 0, 0, 0], float)


def _rot(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _contact_state(rng, pen):
    qpos = QPOS0.copy()
    q = np.array([1, 0, 0, 0.]) + rng.normal(size=4) * 0.05; q /= np.linalg.norm(q); qpos[3:7] = q
    qpos[7:10] = rng.normal(size=3)
    ql = rng.normal(size=4); ql /= np.linalg.norm(ql); qpos[13:17] = ql
    xy = rng.uniform(-0.5, 0.5, 2)
    bc = np.array([xy[0], xy[1], 0.09 - pen])
    qpos[10:13] = bc - _rot(ql) @ np.array([0, 0, -0.14])
    qpos[0:3] = bc - _rot(q) @ (np.array([0, 0, -0.12]) + rng.normal(size=3) * 0.002)
    return qpos, rng.normal(size=15) * 0.3


def test_model_constants_agree(oracle_mod):
    mo, mh = oracle_mod.model_constants(), H.model()
    iw = mo["invweight0"][:, 0]
    # diagApprox per contact type: ball x wheel_i, hfield x ball, hfield x stick_i (cam bodies 2, 3), hfield x wheel_i, ball x tower, ball x stick_i
    np.testing.assert_allclose(mh["dA"], [iw[7] + iw[4], iw[7] + iw[5], iw[7] + iw[6], iw[7], iw[2], iw[3], iw[4], iw[5], iw[6],
                                          iw[7] + iw[1], iw[7] + iw[2], iw[7] + iw[3]], rtol=1e-11)
    assert abs(mh["meaninertia"] - mo["meaninertia"]) < 1e-13
    np.testing.assert_allclose(mh["masses"], [mo["mass"][1:4].sum(), mo["mass"][4], mo["mass"][7]], rtol=1e-13)


def test_smooth_dynamics_agree_without_contacts(oracle_mod):
    rng = np.random.default_rng(0)
    e = oracle_mod.OracleEnv()
    for _ in range(10):
        qpos = QPOS0.copy(); qpos[2] += 1.0; qpos[12] += 2.0
        q = rng.normal(size=4); qpos[3:7] = q / np.linalg.norm(q)
        q = rng.normal(size=4); qpos[13:17] = q / np.linalg.norm(q)
        qpos[7:10] = rng.normal(size=3) * 3
        qvel = rng.normal(size=15); ctrl = rng.uniform(-12, 12, 3)          # beyond ctrlrange: both must clamp to +-10
        e.set_state(qpos, qvel); fo = e.forward(ctrl)
        fh = H.forward(qpos, qvel, ctrl, np.zeros(15))
        assert fo["ncon"] == 0 and fh["ncon"] == 0
        np.testing.assert_allclose(fh["qM"], fo["qM"], atol=1e-13)
        np.testing.assert_allclose(fh["qfs"], fo["qM"] @ fo["qacc_smooth"], atol=1e-11)
        np.testing.assert_allclose(fh["qacc"], fo["qacc"], rtol=1e-10, atol=1e-10)


def test_contacts_and_newton_agree(oracle_mod):
    rng = np.random.default_rng(1)
    e = oracle_mod.OracleEnv()
    seen = set()
    for t in range(60):
        qpos, qvel = _contact_state(rng, rng.uniform(0, 0.004))
        ctrl = rng.uniform(-10, 10, 3); warm = rng.normal(size=15) * (t % 2)
        e.set_state(qpos, qvel, warm); fo = e.forward(ctrl); co = e.contacts()
        fh = H.forward(qpos, qvel, ctrl, warm)
        pen = co["dist"] < 0
        assert fh["ncon"] == pen.sum()                                       # contact sets match
        np.testing.assert_allclose(fh["dist"], co["dist"][pen], atol=1e-14)
        np.testing.assert_allclose(fh["pos"], co["pos"][pen], atol=1e-14)
        np.testing.assert_allclose(fh["frame"], co["frame"][pen], atol=1e-12)
        assert fh["niter"] == fo["niter"]
        assert np.abs(fh["qacc"] - fo["qacc"]).max() / np.abs(fo["qacc"]).max() < 1e-10
        seen.add(int(fh["ncon"]))
    assert max(seen) >= 9                                                    # 3 wheel pairs + several terrain prisms


@pytest.mark.parametrize("terrain", ["flat", "perlin"])
def test_rk4_trajectory_agrees(oracle_mod, terrain):
    """100-step drift bound: 1e-10 absolute in fp64; fp32 arithmetic stays within 2e-3 (BASELINE: 1e-3 single step)."""
    rng = np.random.default_rng(2)
    hf = None if terrain == "flat" else oracle_mod.perlin_terrain(seed=123)
    e = oracle_mod.OracleEnv(); e.reset(hf)
    s64 = e.get_state()[:3]; s32 = s64
    first32 = None
    for k in range(100):
        ctrl = -10 * np.clip(rng.normal(size=3), -1, 1)
        e.mj_step(ctrl)
        s64 = H.step(*s64, ctrl, hf, prec=64)[:3]
        s32 = H.step(*s32, ctrl, hf, prec=32)[:3]
        qo, vo, wo, _ = e.get_state()
        assert max(np.abs(qo - s64[0]).max(), np.abs(vo - s64[1]).max()) < 1e-10, k
        if k == 0:
            first32 = max(np.abs(qo - s32[0]).max() / max(1, np.abs(qo).max()), np.abs(vo - s32[1]).max() / max(1, np.abs(vo).max()))
    assert first32 < 1e-3
    assert max(np.abs(qo - s32[0]).max(), np.abs(vo - s32[1]).max()) < 2e-3 * max(1.0, np.abs(vo).max())


def test_stale_observation_kinematics(oracle_mod):
    """obs kinematics come from the LAST RK4 stage evaluation, not from the returned state."""
    e = oracle_mod.OracleEnv(); e.reset()
    rng = np.random.default_rng(3)
    s = e.get_state()[:3]
    for _ in range(30):
        ctrl = rng.uniform(-10, 10, 3)
        e.mj_step(ctrl)
        q, v, w, kin, _, _ = H.step(*s, ctrl, None, prec=64); s = (q, v, w)
    xpos, xquat, cvel = e.kin()
    np.testing.assert_allclose(kin[0:4], xquat, atol=1e-12)
    np.testing.assert_allclose(kin[4:7], cvel[0:3], atol=1e-11)             # "vel" <- cvel[0:3] (angular)
    np.testing.assert_allclose(kin[7:10], cvel[3:6], atol=1e-11)            # "angular_vel" <- cvel[3:6] (linear, at subtree COM)
    np.testing.assert_allclose(kin[10:13], xpos, atol=1e-12)
    assert np.abs(xquat - q[3:7]).max() > 1e-9                               # and it really is stale w.r.t. the new qpos


def _rough_state(rng, hf2d, tilt, ball_noise, lift=0.0):
    """Robot dropped onto rough terrain with a large tilt: wheel / stick capsules reach the heightfield, the ball may touch the tower."""
    qpos = QPOS0.copy()
    ax = rng.normal(size=3); ax[2] *= 0.2; ax /= np.linalg.norm(ax); ang = np.radians(rng.uniform(0, tilt))
    q = np.r_[np.cos(ang / 2), np.sin(ang / 2) * ax]; qpos[3:7] = q
    qpos[7:10] = rng.normal(size=3)
    ql = rng.normal(size=4); ql /= np.linalg.norm(ql); qpos[13:17] = ql
    xy = rng.uniform(-1.5, 1.5, 2)
    c, r = (xy + 5) / 10 * 292
    h = 2.0 * hf2d[int(round(r)), int(round(c))]
    bc = np.array([xy[0], xy[1], h + 0.09 - rng.uniform(0, 0.02)])
    qpos[10:13] = bc - _rot(ql) @ np.array([0, 0, -0.14])
    qpos[0:3] = bc - _rot(q) @ (np.array([0, 0, -0.12 + lift]) + rng.normal(size=3) * ball_noise)
    return qpos, rng.normal(size=15) * 0.3


def test_extra_collision_pairs_agree(oracle_mod):
    """Round-2 pairs (ORACLE_ASSUMPTIONS 7b): wheel and camera-stick capsules against the heightfield, ball against the sticks
    (patched sphere-capsule) and the tower cylinder.  Contact sets (count, type, dist, pos, frame), Newton iterations and qacc of
    the engine core against the oracle on tilted robots over a rough (fast_sin) Perlin field."""
    rng = np.random.default_rng(5)
    hf = oracle_mod.perlin_terrain(seed=77); hf2d = hf.reshape(293, 293)
    e = oracle_mod.OracleEnv(); e.reset(hf)
    pair2type = {0: 0, 1: 1, 2: 2, 3: 3, 5: 4, 6: 5, 7: 6, 8: 7, 9: 8, 10: 9, 11: 10, 12: 11}
    seen = set()
    for t in range(400):
        qpos, qvel = _rough_state(rng, hf2d, 70.0 if t % 2 else 35.0, 0.03 if t % 4 == 0 else 0.004, rng.uniform(0.085, 0.1) if t % 10 == 0 else 0.0)
        ctrl = rng.uniform(-10, 10, 3); warm = rng.normal(size=15) * (t % 2)
        e.set_state(qpos, qvel, warm); fo = e.forward(ctrl); co = e.contacts()
        if co["n"] >= 60:
            continue                                                         # capacity edge (64 contacts): not compared
        fh = H.forward(qpos, qvel, ctrl, warm, hf)
        assert fh["ncon"] == co["n"], (t, fh["ncon"], co["n"])
        np.testing.assert_array_equal(fh["type"], [pair2type[int(p)] for p in co["pair"]])
        np.testing.assert_allclose(fh["dist"], co["dist"], atol=1e-13)
        np.testing.assert_allclose(fh["pos"], co["pos"], atol=1e-13)
        np.testing.assert_allclose(fh["frame"], co["frame"], atol=1e-11)
        assert fh["niter"] == fo["niter"], t
        assert np.abs(fh["qacc"] - fo["qacc"]).max() / max(1.0, np.abs(fo["qacc"]).max()) < 1e-8, t
        seen.update(int(x) for x in fh["type"])
    assert {4, 5, 6, 7, 8, 9, 10, 11} <= seen, seen                          # every new pair type occurred


@pytest.mark.parametrize("terrain,prec,tol", [("perlin", 64, 1e-6), ("flat", 64, 1e-6), ("perlin", 32, 1e-3), ("flat", 32, 1e-3)])
def test_solver_mode_1_reaches_the_reference_minimiser(oracle_mod, terrain, prec, tol):
    """solver_mode 1 (strong-Wolfe line search with cone-apex candidates, analytic p0, warm start chained through the RK
    stages) against the ORACLE's mj_step on every state of random-action rollouts that include drops, landing impacts and
    bounces: one step agrees to 1e-6 relative (BASELINE tolerance 1e-5) although the iteration path differs."""
    rng = np.random.default_rng(11)
    worst = 0.0
    o = oracle_mod.OracleEnv()
    try:
        for ep in range(3 if terrain == "perlin" else 1):
            hf = oracle_mod.perlin_terrain(seed=int(rng.integers(0, 10000))) if terrain == "perlin" else np.zeros(293 * 293, np.float32)
            o.reset(hf)
            q, v, w, _ = o.get_state()
            for t in range(400):
                ctrl = -10.0 * rng.uniform(-1, 1, 3)
                H.set_solver(1)
                q1, v1, _, _, nc, _ = H.step(q, v, w, ctrl, hf, prec=prec)
                o.set_state(q, v, w); o.mj_step(ctrl)
                qo, vo, wo, _ = o.get_state()
                worst = max(worst, np.abs(q1 - qo).max() / max(1.0, np.abs(qo).max()), np.abs(v1 - vo).max() / max(1.0, np.abs(vo).max()))
                q, v, w = qo, vo, wo
                w_, x_, y_, z_ = q[3:7]
                if np.degrees(np.arccos(np.clip(1 - 2 * (x_ * x_ + y_ * y_), -1, 1))) > 20:
                    break
    finally:
        H.set_solver(0); o.close()
    assert worst < tol, worst   # fp64: 1e-6 (BASELINE 1e-5); fp32 build of the same core: BASELINE's 1e-3


def test_solver_mode_1_on_rough_random_states(oracle_mod):
    """solver_mode 1 against the oracle's exact-line-search solve on randomly posed robots over rough Perlin terrain (tilted up to
    70 degrees, deep and shallow penetrations, random warm starts, every collision pair type): same contact count, qacc within
    1e-6 relative although the iteration paths differ, and never more Newton iterations than the iteration cap."""
    rng = np.random.default_rng(17)
    hf = oracle_mod.perlin_terrain(seed=4242); hf2d = hf.reshape(293, 293)
    e = oracle_mod.OracleEnv(); e.reset(hf)
    worst, solved = 0.0, 0
    try:
        H.set_solver(1)
        for t in range(300):
            qpos, qvel = _rough_state(rng, hf2d, 70.0 if t % 2 else 35.0, 0.03 if t % 4 == 0 else 0.004, rng.uniform(0.085, 0.1) if t % 10 == 0 else 0.0)
            ctrl = rng.uniform(-10, 10, 3); warm = rng.normal(size=15) * (t % 2)
            e.set_state(qpos, qvel, warm); fo = e.forward(ctrl)
            if fo["ncon"] >= 60 or fo["ncon"] == 0:
                continue
            fh = H.forward(qpos, qvel, ctrl, warm, hf)
            assert fh["ncon"] == fo["ncon"] and fh["niter"] < 100, (t, fh["ncon"], fo["ncon"], fh["niter"])
            worst = max(worst, np.abs(fh["qacc"] - fo["qacc"]).max() / max(1.0, np.abs(fo["qacc"]).max()))
            solved += 1
    finally:
        H.set_solver(0); e.close()
    assert solved > 150 and worst < 1e-6, (solved, worst)
