import sys, numpy as np, torch
sys.path.insert(0, ".")
from openballbot_rl_b200.engine import BallbotEngine
from oracle import oracle as O
QPOS0 = np.array([0, 0, 0.24, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.26, 1, 0, 0, 0], np.float64)
N = 24
eng = BallbotEngine(num_envs=N, precision=64, terrain="perlin", cameras=False, auto_reset=False, seed=5)
eng.reset()
rng = np.random.default_rng(9)
hfs = [eng.get_hfield(i).cpu().numpy() for i in range(N)]
qpos = np.tile(QPOS0, (N, 1)); qvel = np.zeros((N, 15)); warm = np.zeros((N, 15))
for i in range(N):
    h = hfs[i].reshape(293, 293)
    ground = 2.0 * float(h[140:153, 140:153].max())
    sink = rng.uniform(0.02, 0.06)
    qpos[i, 12] = ground + 0.09 - 0.03 - sink; qpos[i, 2] = qpos[i, 12] - 0.02
    qvel[i, 11] = -rng.uniform(0.5, 2.0); qvel[i, 2] = qvel[i, 11]
eng.set_state(qpos, qvel, warm)
a = rng.uniform(-1, 1, (N, 3)).astype(np.float32)
e = O.OracleEnv()
for i in range(N):
    pr = eng.probe_forward(i, ctrl=tuple(-10.0 * a[i].astype(np.float64)))
    e.set_hfield(hfs[i]); e.set_state(qpos[i], qvel[i], warm[i])
    out = e.forward(-10.0 * a[i].astype(np.float64))
    d = np.abs(pr["qacc"] - out["qacc"]).max()
    print("env %2d: engine ncon %2d niter %3d | oracle ncon %2d niter %3d | max|dqacc| %.3e  |qacc| %.2e  dqas %.1e" % (i, pr["ncon"], pr["niter"], out["ncon"], out["niter"], d, np.abs(out["qacc"]).max(), np.abs(pr["qacc_smooth"] - out["qacc_smooth"]).max()))
# full step comparison
eng.step(torch.from_numpy(a).cuda())
q1, v1, w1 = [x.cpu().numpy() for x in eng.get_state()]
for i in range(N):
    e.set_hfield(hfs[i]); e.set_state(qpos[i], qvel[i], warm[i]); e.mj_step(-10.0 * a[i].astype(np.float64))
    qo, vo, wo, _ = e.get_state()
    print("env %2d step: dq %.2e dv %.2e dw %.2e" % (i, np.abs(q1[i] - qo).max(), np.abs(v1[i] - vo).max(), np.abs(w1[i] - wo).max()))
print("---- thread-per-env core on the GPU for the mismatching envs")
eng2 = BallbotEngine(num_envs=N, precision=64, terrain="external", cameras=False, auto_reset=False, seed=5, step_kernel="thread")
eng2.reset()
eng2.set_hfield(np.arange(N), np.stack(hfs))
eng2.set_state(qpos, qvel, warm)
for i in (2, 6, 19, 0):
    pr2 = eng2.probe_forward(i, ctrl=tuple(-10.0 * a[i].astype(np.float64)))
    e.set_hfield(hfs[i]); e.set_state(qpos[i], qvel[i], warm[i]); e.forward(-10.0 * a[i].astype(np.float64))
    cc = e.contacts(); dist, pos, frame, pair = cc["dist"], cc["pos"], cc["frame"], cc["pair"]
    print("env", i, "thread core ncon", pr2["ncon"], "oracle ncon", len(dist))
    print("  oracle dist", np.round(dist, 5), "pairs", pair)
    print("  core   dist", np.round(pr2["dist"], 5))
    print("  ball centre", qpos[i, 10:13] + np.array([0, 0, 0.0]), "oracle contact pos z", np.round(pos[:, 2], 4))
