import sys, torch, numpy as np
sys.path.insert(0, ".")
from openballbot_rl_b200.engine import BallbotEngine
N = 65536
eng = BallbotEngine(num_envs=N, precision=32, terrain="perlin", cameras=False, seed=0)
eng.reset()
g = torch.Generator(device="cuda"); g.manual_seed(0)
act = torch.rand(64, N, 3, device="cuda", generator=g) * 2 - 1
prev = eng.status.clone(); prev_len = eng.episode_length.clone()
hist_nc, hist_nit, hist_len = [], [], []
step_count = torch.zeros(N, dtype=torch.int32, device="cuda")
for t in range(600):
    eng.step(act[t % 64])
    st = eng.status
    bad = (st & 1).bool()
    if bool(bad.any()):
        hist_nc += ((prev[bad] >> 8) & 255).tolist(); hist_nit += (prev[bad] >> 16).tolist(); hist_len += step_count[bad].tolist()
    step_count += 1; step_count[eng.terminated.bool()] = 0
    prev = st.clone()
print("failures", len(hist_nc))
print("prev-step max contacts of failing envs: ", np.bincount(hist_nc)[:45])
print("prev-step newton iterations:", np.percentile(hist_nit, [10, 50, 90]) if hist_nit else None)
print("episode step at failure:", np.percentile(hist_len, [10, 50, 90]) if hist_len else None)
