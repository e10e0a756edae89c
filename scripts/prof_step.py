"""Small deterministic workload for ncu: flat or perlin, N envs, S steps (the last few are the ones profiled)."""
import argparse, sys, torch
sys.path.insert(0, ".")
from openballbot_rl_b200.engine import BallbotEngine
ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=4096); ap.add_argument("--steps", type=int, default=70)
ap.add_argument("--precision", type=int, default=64); ap.add_argument("--terrain", default="flat")
ap.add_argument("--kernel", default="warp"); ap.add_argument("--solver", default="exact"); ap.add_argument("--lib", default=None)
a = ap.parse_args()
if a.lib:
    from openballbot_rl_b200 import _lib
    _lib.LIB_PATH = a.lib   # A/B experiments only: an alternative build of the same CUDA library
eng = BallbotEngine(num_envs=a.envs, precision=a.precision, terrain=a.terrain, cameras=(a.terrain == "perlin"), step_kernel=a.kernel, solver=a.solver, seed=0)
eng.reset()
g = torch.Generator(device="cuda"); g.manual_seed(0)
act = torch.rand(8, a.envs, 3, device="cuda", generator=g) * 2 - 1
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for t in range(a.steps):
    if t == a.steps - 5: e0.record()
    eng.step(act[t % 8])
e1.record(); torch.cuda.synchronize()
st = eng.status.cpu()
print(a.solver, a.precision, a.envs, "last 5 steps ms/step", e0.elapsed_time(e1) / 5, "ncon max", int(((st >> 8) & 255).max()), "mean", float(((st >> 8) & 255).float().mean()))
