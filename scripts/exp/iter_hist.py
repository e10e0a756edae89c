import sys, torch, numpy as np
sys.path.insert(0, ".")
from openballbot_rl_b200.engine import BallbotEngine
N = 32768
terrain = sys.argv[1] if len(sys.argv) > 1 else "perlin"
eng = BallbotEngine(num_envs=N, precision=64, terrain=terrain, cameras=(terrain == "perlin"), seed=0)
eng.reset()
g = torch.Generator(device="cuda"); g.manual_seed(0)
act = torch.rand(16, N, 3, device="cuda", generator=g) * 2 - 1
for t in range(330): eng.step(act[t % 16])
st = eng.status.cpu().numpy()
nit, ncon = st >> 16, (st >> 8) & 255
print(terrain, "newton iterations per step: mean %.1f median %d p90 %d p99 %d max %d; frac zero %.2f" % (nit.mean(), np.median(nit), np.percentile(nit, 90), np.percentile(nit, 99), nit.max(), (nit == 0).mean()))
print("ncon: mean %.2f, hist" % ncon.mean(), np.bincount(ncon)[:20])
for k in range(0, 16):
    m = ncon == k
    if m.sum(): print("  ncon=%2d: %6d envs, niter mean %.1f" % (k, m.sum(), nit[m].mean()))
