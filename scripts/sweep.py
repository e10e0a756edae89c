"""BASELINE.json configs[4]: env-count sweep, fp32 vs fp64 physics, flat (proprio only) and perlin (+ depth cameras, Perlin
table), at 1 / 2 / 4 / 8 GPUs (run under torchrun for more than one: envs are sharded, no collective on the step path).
Appends markdown rows to --out (default gpurun_out/sweep.md).  Device-resident stepping, CUDA events, max over ranks."""
import argparse, os, sys, torch
sys.path.insert(0, ".")
from openballbot_rl_b200.engine import BallbotEngine

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/sweep.md")
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--sizes", default="1024,4096,16384,65536,262144,1048576", help="envs PER GPU")
ap.add_argument("--terrains", default="flat,perlin")
ap.add_argument("--solver", default="fast")
a = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
rows = []
for terrain in a.terrains.split(","):
    for N in [int(x) for x in a.sizes.split(",")]:
        pre = (150 if terrain == "flat" else 300) if N <= 65536 else 100
        for prec in (64, 32):
            eng = BallbotEngine(num_envs=N, device=local, precision=prec, terrain=terrain, cameras=(terrain == "perlin"), seed=0, env_offset=rank * N, solver=a.solver)
            eng.reset()
            g = torch.Generator(device=dev); g.manual_seed(rank)
            act = torch.rand(8, N, 3, device=dev, generator=g) * 2 - 1
            for t in range(pre):
                eng.step(act[t % 8])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            e0.record()
            for t in range(a.steps):
                eng.step(act[t % 8])
            e1.record(); torch.cuda.synchronize(dev)
            ms = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            ms = float(ms)
            rows.append((terrain, world, N, N * world, prec, ms, N * world / ms * 1e3))
            if rank == 0:
                print(rows[-1], flush=True)
            eng.close(); del eng, act
            torch.cuda.empty_cache()
if rank == 0:
    new = not os.path.exists(a.out)
    with open(a.out, "a") as f:
        if new:
            f.write("| terrain | GPUs | envs / GPU | envs total | physics | ms / step | env-steps/s |\n|---|---:|---:|---:|---|---:|---:|\n")
        for r in rows:
            f.write(f"| {r[0]} | {r[1]} | {r[2]:,} | {r[3]:,} | fp{r[4]} | {r[5]:.2f} | {r[6] / 1e6:.2f} M |\n")
if world > 1:
    dist.destroy_process_group()
