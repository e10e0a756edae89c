"""Debug: first divergence between the thread-per-env kernel and the lane-group kernels (and the oracle) on a perlin rollout."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from openballbot_rl_b200.engine import BallbotEngine
from oracle import oracle as O
N = 16
e1 = BallbotEngine(num_envs=N, precision=64, terrain="perlin", cameras=False, auto_reset=False, seed=5, step_kernel="warp")
e2 = BallbotEngine(num_envs=N, precision=64, terrain="perlin", cameras=False, auto_reset=False, seed=5, step_kernel="thread")
e1.reset(); e2.reset()
g = torch.Generator(device="cuda"); g.manual_seed(0)
oras = []
for i in range(N):
    o = O.OracleEnv(); o.reset(e1.get_hfield(i).cpu().numpy()); oras.append(o)
    assert torch.equal(e1.get_hfield(i), e2.get_hfield(i))
for t in range(80):
    a = torch.rand(N, 3, device="cuda", generator=g) * 2 - 1
    q_prev = [x.clone() for x in e2.get_state()]
    pr2 = [e2.probe_forward(i, ctrl=tuple((-10.0 * a[i].double()).tolist())) for i in range(N)] if t >= 0 else None
    pr1 = [e1.probe_forward(i, ctrl=tuple((-10.0 * a[i].double()).tolist())) for i in range(N)]
    e1.step(a); e2.step(a)
    (q1, v1, _), (q2, v2, _) = e1.get_state(), e2.get_state()
    an = a.cpu().numpy()
    for i in range(N):
        oras[i].mj_step(-10.0 * an[i].astype(np.float64))
    qo = np.stack([o.get_state()[0] for o in oras])
    d12 = (q1 - q2).abs().max(dim=1).values.cpu().numpy(); d1o = np.abs(q1.cpu().numpy() - qo).max(1); d2o = np.abs(q2.cpu().numpy() - qo).max(1)
    bad = np.nonzero(d12 > 1e-9)[0]
    if len(bad):
        i = int(bad[0])
        print("step", t, "env", i, "warp-thread", d12[i], "warp-oracle", d1o[i], "thread-oracle", d2o[i])
        print(" warp   probe: ncon", pr1[i]["ncon"], "types", pr1[i]["type"], "niter", pr1[i]["niter"])
        print(" thread probe: ncon", pr2[i]["ncon"], "types", pr2[i]["type"], "niter", pr2[i]["niter"])
        print(" warp dist", np.round(pr1[i]["dist"], 6)); print(" thread dist", np.round(pr2[i]["dist"], 6))
        print(" qacc warp", np.round(pr1[i]["qacc"], 4)); print(" qacc thread", np.round(pr2[i]["qacc"], 4))
        hf = e1.get_hfield(i).cpu().numpy().reshape(293, 293)
        print(" qpos before", np.round(q_prev[0][i].cpu().numpy(), 4))
        for c in range(pr2[i]["ncon"]):
            P = pr2[i]["pos"][c]; cc, rr = (P[0] + 5) / 10 * 292, (P[1] + 5) / 10 * 292
            print("  thread contact", c, "type", pr2[i]["type"][c], "dist", round(float(pr2[i]["dist"][c]), 5), "pos", np.round(P, 4), "normal", np.round(pr2[i]["frame"][c][0], 4),
                  "cell", round(rr, 2), round(cc, 2), "terrain z there", round(2.0 * float(hf[int(rr), int(cc)]), 4))
        break
else:
    print("no divergence in 80 steps")
