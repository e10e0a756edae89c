"""Where a PPO rollout step goes (one GPU): engine step alone, + VecEnv wrapper, + features (embedding cache), + policy heads."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from openballbot_rl_b200.training.policy import BallbotPolicy
from openballbot_rl_b200.training.ppo import PPOConfig, PPOLearner
from openballbot_rl_b200.training.utils import make_ballbot_vec_env
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
T = 48
dev = torch.device("cuda", 0)
venv = make_ballbot_vec_env(N, terrain_config={"type": "perlin", "config": {}}, seed=0, device=0)
pol = BallbotPolicy().to(dev)
L = PPOLearner(venv, pol, PPOConfig(n_steps=T, batch_size=max(256, N * T // 80)), total_timesteps=10 ** 12)
L.update(L.collect()[0])          # warm-up (lazy initialisation, desynchronised episodes)
acts = torch.rand(8, N, 3, device=dev) * 2 - 1


def timed(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in range(T):
        fn(t)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / T * 1e3
    print(f"{name:60s} {dt:7.3f} ms / step", flush=True)
    return dt


obs = [venv.reset()]
timed("engine.step", lambda t: venv.engine.step(acts[t % 8]))
timed("venv.step", lambda t: obs.__setitem__(0, venv.step(acts[t % 8])[0]))
def f(t):
    obs[0] = venv.step(acts[t % 8])[0]; L.features(obs[0])
timed("venv.step + features (embedding cache, nonzero sync)", f)
def g(t):
    obs[0] = venv.step(acts[t % 8])[0]; ft = L.features(obs[0]); m, ls = L._dist(ft); L._value(ft)
timed("venv.step + features + policy / value heads", g)
torch.cuda.synchronize(); t0 = time.perf_counter(); buf, _ = L.collect(); torch.cuda.synchronize(); t1 = time.perf_counter(); L.update(buf); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"collect {(t1 - t0) / T * 1e3:.3f} ms / step; update {(t2 - t1) * 1e3:.1f} ms per iteration")
