"""BASELINE.json configs[4]: env-count sweep, fp32 vs fp64 physics, flat (proprio only) and perlin (+ depth, terrain regen).
Writes a markdown table (default gpurun_out/sweep.md).  Device-resident stepping, CUDA events, exact solver mode."""
import argparse, sys, torch
sys.path.insert(0, ".")
from openballbot_rl_b200.engine import BallbotEngine

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/sweep.md")
ap.add_argument("--steps", type=int, default=60)
ap.add_argument("--max-perlin", type=int, default=262144)
ap.add_argument("--max-flat", type=int, default=1048576)
a = ap.parse_args()
rows = []
for terrain, sizes, pre in (("flat", [1024, 4096, 16384, 65536, 262144, 1048576], 150), ("perlin", [1024, 4096, 16384, 65536, 262144], 300)):
    for N in sizes:
        if N > (a.max_flat if terrain == "flat" else a.max_perlin):
            continue
        for prec in (64, 32):
            eng = BallbotEngine(num_envs=N, precision=prec, terrain=terrain, cameras=(terrain == "perlin"), seed=0)
            eng.reset()
            g = torch.Generator(device="cuda"); g.manual_seed(0)
            act = torch.rand(8, N, 3, device="cuda", generator=g) * 2 - 1
            for t in range(pre):
                eng.step(act[t % 8])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for t in range(a.steps):
                eng.step(act[t % 8])
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            rows.append((terrain, N, prec, ms, N / ms * 1e3))
            print(rows[-1], flush=True)
            eng.close(); del eng, act
            torch.cuda.empty_cache()
with open(a.out, "w") as f:
    f.write("| terrain | envs | physics | ms / step | env-steps/s |\n|---|---:|---|---:|---:|\n")
    for r in rows:
        f.write(f"| {r[0]} | {r[1]:,} | fp{r[2]} | {r[3]:.2f} | {r[4] / 1e6:.2f} M |\n")
