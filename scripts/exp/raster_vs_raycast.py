"""A/B: rasterising depth kernel against the ray-caster on the same states (perlin rollouts): pixel differences and timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from openballbot_rl_b200.engine import BallbotEngine
N = 4096
kw = dict(num_envs=N, precision=64, terrain="perlin", cameras=True, seed=3, perlin_table=True)
a_ = BallbotEngine(depth_kernel="raster", **kw); b_ = BallbotEngine(depth_kernel="raycast", **kw)
a_.reset(); b_.reset()
g = torch.Generator(device="cuda"); g.manual_seed(0)
worst = 0.0; nbad = 0; ntot = 0
for t in range(150):
    act = torch.rand(N, 3, device="cuda", generator=g) * 2 - 1
    a_.step(act); b_.step(act)
    if t % 6 == 5:
        for k in ("rgbd_0", "rgbd_1"):
            d = (a_.obs[k] - b_.obs[k]).abs()
            worst = max(worst, float(d.max())); nbad += int((d > 1e-6).sum()); ntot += d.numel()
print("max |raster - raycast|", worst, "pixels differing by > 1e-6:", nbad, "of", ntot, f"({100.0 * nbad / ntot:.4f} %)")
for name, e in (("raster", a_), ("raycast", b_)):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        e.render_depth()
    e1.record(); torch.cuda.synchronize()
    print(name, "full render of", N, "envs x 2 cameras:", e0.elapsed_time(e1) / 20, "ms")
