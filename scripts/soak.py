"""Soak run: many steps at full size, counting numerical-failure flags and tracking episode statistics / throughput drift."""
import sys, time, torch
sys.path.insert(0, ".")
from openballbot_rl_b200.engine import BallbotEngine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
prec = int(sys.argv[3]) if len(sys.argv) > 3 else 64
terrain = sys.argv[4] if len(sys.argv) > 4 else "perlin"
solver = sys.argv[5] if len(sys.argv) > 5 else "exact"
eng = BallbotEngine(num_envs=N, precision=prec, terrain=terrain, cameras=True, seed=0, solver=solver)
eng.reset()
g = torch.Generator(device="cuda"); g.manual_seed(0)
act = torch.rand(64, N, 3, device="cuda", generator=g) * 2 - 1
bad = torch.zeros((), dtype=torch.int64, device="cuda"); eps = torch.zeros((), dtype=torch.int64, device="cuda")
slen = torch.zeros((), dtype=torch.float64, device="cuda"); ncmax = torch.zeros((), dtype=torch.int32, device="cuda")
t0 = time.perf_counter()
for t in range(T):
    eng.step(act[t % 64])
    bad += (eng.status & 1).sum(); d = eng.terminated.bool(); eps += d.sum(); slen += eng.episode_length[d].sum()
    ncmax = torch.maximum(ncmax, ((eng.status >> 8) & 255).max())
    if (t + 1) % 1000 == 0:
        torch.cuda.synchronize()
        print(f"step {t+1}: {N*(t+1)/(time.perf_counter()-t0)/1e6:.2f} M env-steps/s wall, episodes {int(eps)}, mean len {float(slen)/max(1,int(eps)):.1f}, "
              f"numerical failures {int(bad)}, max contacts {int(ncmax)}, finite images {bool(torch.isfinite(eng.obs['rgbd_0']).all())}", flush=True)
q, v, w = eng.get_state()
print("final state finite:", bool(torch.isfinite(q).all() and torch.isfinite(v).all() and torch.isfinite(w).all()), "mem GB", torch.cuda.memory_allocated() / 1e9)
