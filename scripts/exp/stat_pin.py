import sys, torch
sys.path.insert(0, ".")
from openballbot_rl_b200.engine import BallbotEngine
for terrain in ("flat", "perlin"):
    N = 4096
    eng = BallbotEngine(num_envs=N, precision=64, terrain=terrain, cameras=False, seed=4)
    eng.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(123)
    n_ep = torch.zeros((), device="cuda"); s_len = torch.zeros((), device="cuda"); s_ret = torch.zeros((), device="cuda"); s_len2 = torch.zeros((), device="cuda")
    for t in range(2048):
        a = torch.randn(N, 3, device="cuda", generator=g).clamp_(-1, 1)
        eng.step(a)
        d = eng.terminated.bool()
        l = eng.episode_length[d].float()
        n_ep += d.sum(); s_len += l.sum(); s_len2 += (l * l).sum(); s_ret += eng.episode_return[d].sum()
    n = float(n_ep); ml = float(s_len) / n
    print(terrain, "episodes", int(n), "ep_len_mean %.1f (std %.1f)" % (ml, (float(s_len2) / n - ml * ml) ** 0.5), "ep_rew_mean %.3f" % (float(s_ret) / n))
    eng.close()
