"""Newton iterations per env-step on the GPU engine (status word), by precision and solver mode."""
import sys, torch, numpy as np
sys.path.insert(0, ".")
from openballbot_rl_b200.engine import BallbotEngine
N = 32768
terrain = sys.argv[1] if len(sys.argv) > 1 else "perlin"
for prec in (64, 32):
    for solver in ("exact", "fast"):
        eng = BallbotEngine(num_envs=N, precision=prec, terrain=terrain, cameras=False, seed=0, solver=solver)
        eng.reset()
        g = torch.Generator(device="cuda"); g.manual_seed(0)
        act = torch.rand(16, N, 3, device="cuda", generator=g) * 2 - 1
        tot = np.zeros(N); mx = 0
        for t in range(330):
            eng.step(act[t % 16])
            if t >= 300:
                st = eng.status.cpu().numpy(); nit = st >> 16; tot += nit; mx = max(mx, nit.max())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(30): eng.step(act[t % 16])
        e1.record(); torch.cuda.synchronize()
        nit = tot / 30
        print(f"{terrain} fp{prec} {solver:5s}: newton iterations per step mean {nit.mean():.2f} p99 {np.percentile(nit, 99):.1f} max single step {mx}; {e0.elapsed_time(e1) / 30:.2f} ms / step")
        eng.close()
