"""Round-2 GPU parity tests: contact sets through the production lane-group code, the wheel / stick / tower collision pairs,
the distance reward, fp32 with terrain contacts, the Perlin table, the numpy-compatible PCG64 seed stream, explicit terrain seeds,
and the device guard.  Everything goes through the C ABI (libballbot_b200.so); the oracle is only the checker."""
import numpy as np
import pytest
import torch

from tests.test_engine_core_cpu import QPOS0, _rough_state

pytestmark = pytest.mark.gpu

PAIR2TYPE = {0: 0, 1: 1, 2: 2, 3: 3, 5: 4, 6: 5, 7: 6, 8: 7, 9: 8, 10: 9, 11: 10, 12: 11}     # oracle pair id -> engine contact type


def _engine(**kw):
    from openballbot_rl_b200.engine import BallbotEngine
    return BallbotEngine(**kw)


def _rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


def _sorted_contacts(types, dist, pos, frame):
    key = np.lexsort((np.round(pos[:, 2], 9), np.round(pos[:, 1], 9), np.round(pos[:, 0], 9), types))
    return types[key], dist[key], pos[key], frame[key]


def _compare_contact_sets(eng, i, ora, tag):
    """mjData.contact parity: same count, same geom pairs, dist / pos / frame to rounding (the engine lists the robot pairs first)."""
    ce = eng.get_contacts(i); co = ora.contacts(80)
    assert len(ce["dist"]) == co["n"], (tag, i, len(ce["dist"]), co["n"])
    to = np.array([PAIR2TYPE[int(p)] for p in co["pair"]], np.int32)
    te, de, pe, fe = _sorted_contacts(ce["type"], ce["dist"], ce["pos"], ce["frame"])
    to, do, po, fo = _sorted_contacts(to, co["dist"], co["pos"], co["frame"])
    np.testing.assert_array_equal(te, to)
    np.testing.assert_allclose(de, do, atol=1e-12)
    np.testing.assert_allclose(pe, po, atol=1e-12)
    np.testing.assert_allclose(fe, fo, atol=1e-10)
    return te


def test_contact_sets_match_the_oracle_on_flat_random_states(oracle_mod):
    """The 64 random contact states of test_random_state_single_step_parity: per-env ncon, contact pairs, dist, pos and frame of
    the production lane-group collision code (bb_get_contacts) equal the oracle's mjData.contact restatement."""
    N = 64
    rng = np.random.default_rng(3)
    eng = _engine(num_envs=N, precision=64, terrain="flat", cameras=False, auto_reset=False)
    eng.reset()
    qpos = np.tile(QPOS0, (N, 1)); qvel = rng.normal(size=(N, 15)) * 0.3; warm = rng.normal(size=(N, 15))
    for i in range(N):
        q = np.array([1, 0, 0, 0.]) + rng.normal(size=4) * 0.05; q /= np.linalg.norm(q)
        qpos[i, 3:7] = q
        qpos[i, 7:10] = rng.normal(size=3)
        qpos[i, 2] = 0.24 - 0.031 + rng.uniform(-0.002, 0.002); qpos[i, 12] = 0.26 - 0.031 + rng.uniform(-0.003, 0.0)
        qpos[i, 0:2] = rng.uniform(-0.3, 0.3, 2); qpos[i, 10:12] = qpos[i, 0:2] + rng.normal(size=2) * 0.002
    eng.set_state(qpos, qvel, warm)
    e = oracle_mod.OracleEnv()
    ncons = []
    for i in range(N):
        e.set_state(qpos[i], qvel[i], warm[i]); e.forward(np.zeros(3))
        ncons.append(len(_compare_contact_sets(eng, i, e, "flat")))
    assert max(ncons) >= 6 and min(ncons) >= 3
    # the step kernels report the same count (status bits 8..15 = max contacts over the RK stages >= the first stage's)
    eng.step(torch.zeros(N, 3, device="cuda"))
    assert (((eng.status.cpu().numpy() >> 8) & 255) >= np.array(ncons)).all()
    eng.close()


def test_extra_collision_pairs_contact_sets_and_single_step(oracle_mod):
    """Tilted robots on a rough Perlin field: wheel / stick capsules x heightfield, ball x sticks, ball x tower.  Contact sets of
    the lane-group code against the oracle, then ONE mj_step at BASELINE's 1e-5 relative tolerance."""
    N = 96
    rng = np.random.default_rng(5)
    hf = oracle_mod.perlin_terrain(seed=77); hf2d = hf.reshape(293, 293)
    eng = _engine(num_envs=N, precision=64, terrain="external", cameras=False, auto_reset=False)
    eng.set_hfield(np.arange(N, dtype=np.int32), np.tile(hf, (N, 1)))
    eng.reset()
    qpos = np.zeros((N, 17)); qvel = np.zeros((N, 15)); warm = rng.normal(size=(N, 15))
    for i in range(N):
        qpos[i], qvel[i] = _rough_state(rng, hf2d, 70.0 if i % 2 else 35.0, 0.03 if i % 4 == 0 else 0.004, rng.uniform(0.085, 0.1) if i % 10 == 0 else 0.0)
    eng.set_state(qpos, qvel, warm)
    e = oracle_mod.OracleEnv(); e.reset(hf)
    seen, keep = set(), []
    for i in range(N):
        e.set_state(qpos[i], qvel[i], warm[i]); e.forward(np.zeros(3))
        if e.contacts(80)["n"] >= 60:
            continue                                                    # oracle capacity edge (64 contacts)
        seen.update(int(t) for t in _compare_contact_sets(eng, i, e, "rough")); keep.append(i)
    assert {4, 5, 6, 7, 8, 9, 10, 11} <= seen, seen
    a = rng.uniform(-1, 1, (N, 3)).astype(np.float32)
    eng.step(torch.from_numpy(a).cuda())
    q1, v1, w1 = [x.cpu().numpy() for x in eng.get_state()]
    worst = 0.0
    for i in keep:
        e.set_state(qpos[i], qvel[i], warm[i]); e.mj_step(-10.0 * a[i].astype(np.float64))
        qo, vo, wo, _ = e.get_state()
        worst = max(worst, _rel(q1[i], qo), _rel(v1[i], vo))
    assert worst < 1e-5, worst
    eng.close()


def test_perlin_landing_contact_sets_and_trajectory(oracle_mod):
    """A perlin landing from reset (ball x heightfield prisms, then whatever else touches): contact sets every 10 steps and the
    trajectory against the oracle on the same heightfield bits."""
    N = 6
    eng = _engine(num_envs=N, precision=64, terrain="perlin", cameras=False, auto_reset=False, seed=5)
    eng.reset()
    envs = []
    for i in range(N):
        e = oracle_mod.OracleEnv(); e.reset(eng.get_hfield(i).cpu().numpy()); envs.append(e)
    rng = np.random.default_rng(2)
    worst, most = 0.0, 0
    for t in range(140):
        a = np.clip(rng.normal(size=(N, 3)), -1, 1).astype(np.float32)
        eng.step(torch.from_numpy(a).cuda())
        qpos, qvel, _ = [x.cpu().numpy() for x in eng.get_state()]
        for i, e in enumerate(envs):
            e.step(a[i])
            qo, vo, _, _ = e.get_state()
            worst = max(worst, _rel(qpos[i], qo), _rel(qvel[i], vo))
            if t % 10 == 9:
                e.forward(np.zeros(3))
                most = max(most, len(_compare_contact_sets(eng, i, e, f"landing t={t}")))
    assert worst < 1e-7 and most >= 6, (worst, most)
    eng.close()


def test_distance_reward_parity(oracle_mod):
    """BB_REWARD_DISTANCE on the device against the oracle's restatement of DistanceReward on info['pos2d'] (rewards/distance.py:33-50)."""
    N = 4
    goal, dscale = (0.7, -0.4), 1.5
    eng = _engine(num_envs=N, precision=64, terrain="flat", cameras=False, auto_reset=False, reward="distance", goal_position=goal, distance_scale=dscale)
    eng.reset()
    envs = [oracle_mod.OracleEnv(reward_type=1, goal=goal, distance_scale=dscale) for _ in range(N)]
    for e in envs:
        e.reset()
    rng = np.random.default_rng(0)
    for t in range(80):
        a = np.clip(rng.normal(size=(N, 3)), -1, 1).astype(np.float32)
        eng.step(torch.from_numpy(a).cuda())
        rew = eng.reward.cpu().numpy(); pos = eng.pos2d.cpu().numpy()
        for i, e in enumerate(envs):
            o, r, tm, fl, info = e.step(a[i])
            assert abs(rew[i] - r) < 1e-6, (t, i, rew[i], r)
            np.testing.assert_allclose(pos[i], info[:2], atol=1e-6)
    assert rew.max() < 0.02 - 0.008                                      # survival bonus minus 0.01 * 1.5 * ~0.8 m of distance
    eng.close()


def test_fp32_perlin_single_step_with_terrain_contacts(oracle_mod):
    """fp32 mode, BASELINE tolerance 1e-3 relative on ONE step, taken where it matters: states of a perlin rollout after the
    landing, i.e. with ball x heightfield (and wheel) contacts active and the fp32 pivot floor in play."""
    N = 32
    e64 = _engine(num_envs=N, precision=64, terrain="perlin", cameras=False, auto_reset=False, seed=9)
    e64.reset()
    hfs = torch.stack([e64.get_hfield(i) for i in range(N)])
    rng = np.random.default_rng(4)
    qpos = np.zeros((N, 17)); qvel = np.zeros((N, 15)); warm = np.zeros((N, 15)); have = np.zeros(N, bool)
    for t in range(260):                                                 # drop (up to ~0.25 s on steep fields) + landing
        e64.step(torch.from_numpy(np.clip(rng.normal(size=(N, 3)), -1, 1).astype(np.float32)).cuda())
        nc = ((e64.status.cpu().numpy() >> 8) & 255); alive = ~e64.terminated.cpu().numpy().astype(bool)
        new = ((nc >= 5) & alive & ~have) | (t == 0)                     # first step with terrain contacts: keep that state
        if new.any():
            q, v, w = [x.cpu().numpy() for x in e64.get_state()]
            qpos[new], qvel[new], warm[new] = q[new], v[new], w[new]
            if t > 0:
                have |= new
    e32 = _engine(num_envs=N, precision=32, terrain="external", cameras=False, auto_reset=False)
    e32.set_hfield(np.arange(N, dtype=np.int32), hfs); e32.reset(); e32.set_state(qpos, qvel, warm)
    a = rng.uniform(-1, 1, (N, 3)).astype(np.float32)
    e32.step(torch.from_numpy(a).cuda())
    q1, v1, _ = [x.cpu().numpy() for x in e32.get_state()]
    ncon = (e32.status.cpu().numpy() >> 8) & 255
    o = oracle_mod.OracleEnv()
    worst, with_terrain = 0.0, 0
    for i in range(N):
        o.reset(hfs[i].cpu().numpy()); o.set_state(qpos[i], qvel[i], warm[i]); o.mj_step(-10.0 * a[i].astype(np.float64))
        qo, vo, _, _ = o.get_state()
        if ncon[i] >= 4 and have[i]:
            with_terrain += 1
            worst = max(worst, _rel(q1[i], qo), _rel(v1[i], vo))
    assert with_terrain >= N // 2, (ncon, have)
    assert worst < 1e-3, worst
    e64.close(); e32.close()


def test_perlin_table_equals_per_env_generation_and_explicit_seeds():
    """The table of all 10,000 possible Perlin fields (bb_create) holds bit-identical heightfields to per-env regeneration, and
    bb_reset(seeds_dev) selects / regenerates exactly the requested seeds."""
    N = 8
    seeds = np.array([0, 1, 359, 7712, 9999, 25, 4242, 5042], np.int32)
    tab = _engine(num_envs=N, precision=64, terrain="perlin", cameras=False, auto_reset=True, perlin_table=True, max_ep_steps=20)
    per = _engine(num_envs=N, precision=64, terrain="perlin", cameras=False, auto_reset=True, perlin_table=False, max_ep_steps=20)
    tab.reset(seeds=seeds); per.reset(seeds=seeds)
    np.testing.assert_array_equal(tab.terrain_seeds().cpu().numpy(), seeds)
    np.testing.assert_array_equal(per.terrain_seeds().cpu().numpy(), seeds)
    ref = per.perlin_terrain(seeds)
    for i in range(N):
        assert torch.equal(tab.get_hfield(i), per.get_hfield(i)) and torch.equal(tab.get_hfield(i), ref[i])
    # auto-resets draw from the same counter stream in both modes: identical trajectories and seed sequences
    a = torch.zeros(N, 3, device="cuda")
    for t in range(45):
        tab.step(a); per.step(a)
        assert torch.equal(tab.terminated, per.terminated)
    assert torch.equal(tab.terrain_seeds(), per.terrain_seeds())
    (q1, v1, _), (q2, v2, _) = tab.get_state(), per.get_state()
    assert torch.equal(q1, q2) and torch.equal(v1, v2)
    for i in range(N):
        assert torch.equal(tab.get_hfield(i), per.get_hfield(i))
    tab.close(); per.close()


def test_pcg64_seed_stream_matches_numpy():
    """seed_stream = pcg64: every (auto-)reset of env i draws `default_rng(seed_i).integers(0, 10000)` exactly like the
    reference's per-env self._np_random (ballbot_env.py:505-507), on the device."""
    N = 16
    env_seeds = [20 + i for i in range(N - 2)] + [2 ** 40 + 3, 0]
    eng = _engine(num_envs=N, precision=32, terrain="perlin", cameras=False, auto_reset=True, seed_stream="pcg64", max_ep_steps=5)
    eng.seed_pcg64(env_seeds)
    gens = [np.random.default_rng(s) for s in env_seeds]
    eng.reset()
    np.testing.assert_array_equal(eng.terrain_seeds().cpu().numpy(), [int(g.integers(0, 10000)) for g in gens])
    a = torch.zeros(N, 3, device="cuda")
    for ep in range(12):                                                 # 12 timeouts -> 12 more draws per env
        for t in range(5):
            eng.step(a)
        assert bool(eng.terminated.all())
        np.testing.assert_array_equal(eng.terrain_seeds().cpu().numpy(), [int(g.integers(0, 10000)) for g in gens])
    eng.close()


def test_late_perlin_policies_statistics_vs_reference_evals():
    """The perlin pairs whose terrain seeds cannot be replayed (later evaluations of the 5.2 M-step run): the reference's eight
    episode lengths per evaluation against the engine's distribution over 512 random terrains under the same policy.  Median of
    the reference inside the engine's 10..90 % band and within the stated relative error of the engine's median."""
    from openballbot_rl_b200.envs import BallbotVecEnv
    from openballbot_rl_b200.training.evaluate import evaluate_policy
    from tests import policy_pairs as P
    from tests.test_gpu_envs import ENV_CFG, PERLIN, REWARD
    for name, tol in (("perlin_5p2M_800k", 0.25), ("perlin_5p2M_1M", 0.25)):
        m = P.meta(name); pol = P.policy(name, "cuda")
        venv = BallbotVecEnv(512, terrain_config=PERLIN, reward_config=REWARD, env_config=ENV_CFG, precision=64, seed=3)
        out = evaluate_policy(venv, pol, max_steps=1200, deterministic=True)
        L = out["lengths"].float().cpu().numpy(); Lr = np.array(m["eval_lengths"], float)
        lo, hi = np.percentile(L, [10, 90])
        print(name, "engine median", np.median(L), "p10/p90", lo, hi, "reference", Lr)
        assert lo <= np.median(Lr) <= hi, (name, np.median(Lr), lo, hi)
        assert abs(np.median(L) - np.median(Lr)) < tol * np.median(Lr), (name, np.median(L), np.median(Lr))
        venv.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_engine_runs_on_its_own_device_whatever_the_current_device_is():
    """every entry point selects the engine's device and restores the caller's."""
    torch.cuda.set_device(0)
    e0 = _engine(num_envs=64, device=0, precision=64, terrain="flat", cameras=False)
    e1 = _engine(num_envs=64, device=1, precision=64, terrain="flat", cameras=False)
    assert torch.cuda.current_device() == 0
    e0.reset(); e1.reset()
    a = torch.rand(64, 3) * 2 - 1
    for t in range(20):
        e0.step(a.to("cuda:0")); e1.step(a.to("cuda:1"))
    assert torch.cuda.current_device() == 0
    (q0, v0, _), (q1, v1, _) = e0.get_state(), e1.get_state()
    assert q1.device.index == 1 and torch.equal(q0.cpu(), q1.cpu()) and torch.equal(v0.cpu(), v1.cpu())
    e0.close(); e1.close()


def test_cuda_engine_and_cpu_ref_backend_through_one_harness():
    """the same C ABI implemented twice -- libballbot_b200.so (CUDA) and oracle/_build/libballbot_cpu_ref.so (the fp64
    oracle) -- driven by ONE harness with the same script: perlin terrain, cameras, auto-reset, numpy PCG64 terrain seeds.
    Observations / rewards / flags / seeds / states are compared after every step."""
    from tests.backend_harness import Backend, pcg64_states
    N = 6
    kw = dict(num_envs=N, terrain_type=1, cameras=1, max_ep_steps=40, seed_stream=1, auto_reset=1, perlin_table=0)
    bc, br = Backend("cuda", **kw), Backend("cpu_ref", **kw)
    st = pcg64_states([100 + i for i in range(N)])
    bc.set_rng_state(st); br.set_rng_state(st)
    oc, orf = bc.reset(), br.reset()
    assert bc.terrain_seeds().tolist() == br.terrain_seeds().tolist()
    rng = np.random.default_rng(1)
    resets = 0
    for t in range(90):
        a = np.clip(rng.normal(size=(N, 3)), -1, 1).astype(np.float32)
        oc, orf = bc.step(a), br.step(a)
        np.testing.assert_array_equal(oc["terminated"], orf["terminated"]); np.testing.assert_array_equal(oc["failure"], orf["failure"])
        np.testing.assert_allclose(oc["reward"], orf["reward"], atol=1e-6)
        for k in ("orientation", "angular_vel", "vel", "motor_state", "actions", "rel_image_ts", "pos2d"):
            np.testing.assert_allclose(oc[k], orf[k], atol=2e-6, err_msg=f"{k} t={t}")
        d = oc["terminated"].astype(bool)
        if d.any():
            resets += int(d.sum())
            np.testing.assert_allclose(oc["terminal_obs"][d], orf["terminal_obs"][d], atol=2e-6)
            np.testing.assert_array_equal(oc["episode_length"][d], orf["episode_length"][d])
            np.testing.assert_allclose(oc["episode_return"][d], orf["episode_return"][d], atol=1e-5)
        assert bc.terrain_seeds().tolist() == br.terrain_seeds().tolist()
        assert (np.abs(oc["rgbd_0"] - orf["rgbd_0"]) > 1e-3).mean() < 0.01
    (qc, vc, _), (qr, vr, _) = bc.get_state(), br.get_state()
    assert _rel(qc, qr) < 1e-6 and _rel(vc, vr) < 1e-6 and resets >= N
    rows_c, rows_r = bc.contacts(0), br.contacts(0)
    assert rows_c.shape == rows_r.shape
    bc.close(); br.close()


@pytest.mark.parametrize("precision,tol", [(64, 1e-5), (32, 1e-3)])
def test_solver_mode_1_single_step_parity_through_perlin_landings(oracle_mod, precision, tol):
    """solver='fast' (strong-Wolfe line search with cone-apex candidates, analytic p0, warm start chained through the RK
    stages) on the production lane-group kernels: every step of a drop / landing / bounce sequence on Perlin terrain is
    re-done by the oracle's exact-line-search mj_step FROM THE ENGINE'S OWN PRE-STEP STATE; qpos / qvel after the step agree
    to BASELINE's single-step tolerance, with the same contact counts."""
    N, NO = 64, 8
    eng = _engine(num_envs=N, precision=precision, terrain="perlin", cameras=False, auto_reset=False, seed=21, solver="fast")
    eng.reset()
    oras = []
    for i in range(NO):
        o = oracle_mod.OracleEnv(); o.reset(eng.get_hfield(i).cpu().numpy()); oras.append(o)
    rng = np.random.default_rng(4)
    worst, ncmax, contact_steps = 0.0, 0, 0
    for t in range(160):
        a = rng.uniform(-1, 1, (N, 3)).astype(np.float32)
        q0, v0, w0 = [x.cpu().numpy().astype(np.float64) for x in eng.get_state()]
        eng.step(torch.from_numpy(a).cuda())
        q1, v1, _ = [x.cpu().numpy().astype(np.float64) for x in eng.get_state()]
        nc = (eng.status.cpu().numpy() >> 8) & 255
        term = eng.terminated.cpu().numpy().astype(bool)
        for i, o in enumerate(oras):
            if term[i] or not np.isfinite(q1[i]).all():
                continue
            o.set_state(q0[i], v0[i], w0[i]); o.mj_step(-10.0 * a[i].astype(np.float64))
            qo, vo, _, _ = o.get_state()
            worst = max(worst, _rel(q1[i], qo), _rel(v1[i], vo))
            ncmax = max(ncmax, int(nc[i])); contact_steps += int(nc[i] > 3)
    assert worst < tol and ncmax >= 8 and contact_steps > 100, (worst, ncmax, contact_steps)
    eng.close()
