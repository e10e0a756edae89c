import sys, torch
sys.path.insert(0, ".")
from openballbot_rl_b200.engine import BallbotEngine
N = 64
for solver in ("exact",):
    e1 = BallbotEngine(num_envs=N, precision=64, terrain="perlin", cameras=False, seed=9, step_kernel="split", solver=solver, auto_reset=False)
    e2 = BallbotEngine(num_envs=N, precision=64, terrain="perlin", cameras=False, seed=9, step_kernel="fused", solver=solver, auto_reset=False)
    e1.reset(); e2.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    for t in range(200):
        a = torch.rand(N, 3, device="cuda", generator=g) * 2 - 1
        e1.step(a); e2.step(a)
        (q1, v1, w1), (q2, v2, w2) = e1.get_state(), e2.get_state()
        dq, dv, dw = (q1 - q2).abs().max().item(), (v1 - v2).abs().max().item(), (w1 - w2).abs().max().item()
        if dq or dv or dw:
            env = int(((v1 - v2).abs().max(1).values > 0).nonzero()[0])
            print("first difference at step", t, "dq %.3e dv %.3e dw %.3e" % (dq, dv, dw), "env", env, "ncon", int(e1.status[env]) >> 8, int(e2.status[env]) >> 8)
            print(" dv per dof", (v1 - v2)[env].cpu().numpy())
            print(" dw per dof", (w1 - w2)[env].cpu().numpy())
            # make them equal again and continue
            e1.set_state(q2, v2, w2)
            if t > 120: break
