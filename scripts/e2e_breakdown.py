"""Where the host-buffer path (bb_step_host) spends its time: device step vs copies vs host memcpy (experiment)."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from openballbot_rl_b200.engine import BallbotEngine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
eng = BallbotEngine(num_envs=N, precision=64, terrain="perlin", cameras=True, seed=0)
eng.reset()
g = torch.Generator(device="cuda"); g.manual_seed(0)
act = torch.rand(16, N, 3, device="cuda", generator=g) * 2 - 1
for t in range(300): eng.step(act[t % 16])
torch.cuda.synchronize()
t0 = time.perf_counter()
for t in range(30): eng.step(act[t % 16])
torch.cuda.synchronize(); t1 = time.perf_counter()
ah = act.cpu().numpy()
eng.step_host(ah[0], images=False)
t2 = time.perf_counter()
for t in range(30): eng.step_host(ah[t % 16], images=False)
torch.cuda.synchronize(); t3 = time.perf_counter()
print(f"device-resident step {1e3*(t1-t0)/30:.2f} ms, host-buffer step {1e3*(t3-t2)/30:.2f} ms")
