"""Rollout post-processing and the PPO learner's host logic on CPU (no GPU): GAE oracle properties, learner on a stub
VecEnv, and the 2-rank gradient all-reduce over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import gae_oracle
from openballbot_rl_b200.training.policy import BallbotPolicy
from openballbot_rl_b200.training.ppo import PPOConfig, PPOLearner, lr_schedule


def _gae_torch(rew, val, done, gamma, lam):
    a, r = gae_oracle.gae(rew.numpy(), val[:-1].numpy(), done.numpy(), val[-1].numpy(), gamma, lam)
    return torch.from_numpy(a), torch.from_numpy(r)


def test_gae_oracle_known_answers():
    # one env, no terminations, lambda = 1: advantage = discounted return + bootstrap - value
    r = np.array([[1.0], [2.0], [3.0]], np.float32); v = np.array([[0.5], [0.25], [0.125]], np.float32); lv = np.array([4.0], np.float32)
    a, ret = gae_oracle.gae(r, v, np.zeros((3, 1), bool), lv, gamma=0.5, gae_lambda=1.0)
    g2 = 3 + 0.5 * 4; g1 = 2 + 0.5 * g2; g0 = 1 + 0.5 * g1
    assert np.allclose(ret[:, 0], [g0, g1, g2]) and np.allclose(a[:, 0], [g0 - 0.5, g1 - 0.25, g2 - 0.125])
    # a done at step 1 cuts the bootstrap and the recursion
    d = np.array([[0], [1], [0]], bool)
    a, ret = gae_oracle.gae(r, v, d, lv, gamma=0.5, gae_lambda=1.0)
    assert np.allclose(ret[:, 0], [1 + 0.5 * 2, 2, g2])
    # lambda = 0: one-step TD error
    a, _ = gae_oracle.gae(r, v, np.zeros((3, 1), bool), lv, gamma=0.9, gae_lambda=0.0)
    assert np.allclose(a[:, 0], [1 + 0.9 * 0.25 - 0.5, 2 + 0.9 * 0.125 - 0.25, 3 + 0.9 * 4 - 0.125])


def test_lr_schedule_matches_reference_steps():
    assert lr_schedule(1.0) == 1e-4 and lr_schedule(0.71) == 1e-4 and lr_schedule(0.6) == 5e-5 and lr_schedule(0.7) == 1e-5 and lr_schedule(0.1) == 1e-5


class StubVecEnv:
    """Torch-output VecEnv stand-in: reward = -|a - 0.3|^2, episodes of 7 steps, cameras off."""
    def __init__(self, n, seed=0):
        self.num_envs = n
        self.g = torch.Generator().manual_seed(seed)
        self.t = torch.zeros(n, dtype=torch.int32)
        self.ret = torch.zeros(n)

    def _obs(self):
        return {k: torch.rand(self.num_envs, 3, generator=self.g) for k in ("orientation", "angular_vel", "vel", "motor_state", "actions")}

    def reset(self):
        return self._obs()

    def step(self, a):
        r = -((a - 0.3) ** 2).sum(-1)
        self.t += 1; self.ret += r
        d = self.t >= 7
        info = {"episode_r": self.ret.clone(), "episode_l": self.t.clone()}
        self.t[d] = 0; self.ret[d] = 0
        return self._obs(), r, d, info


def test_ppo_learner_improves_on_stub_env():
    torch.manual_seed(0)
    pol = BallbotPolicy(cameras=False, hidden=32)
    cfg = PPOConfig(n_steps=32, batch_size=64, n_epochs=4, clip_range=0.2, learning_rate=3e-3, target_kl=None)
    L = PPOLearner(StubVecEnv(16), pol, cfg, total_timesteps=16 * 32 * 12, gae_fn=_gae_torch)
    hist = []
    L.learn(callback=lambda d: hist.append(d))
    assert len(hist) == 12 and all(np.isfinite(h["policy_loss"]) and np.isfinite(h["value_loss"]) for h in hist)
    assert hist[0]["episodes"] > 0 and hist[-1]["ep_rew_mean"] > hist[0]["ep_rew_mean"]     # mean action moved towards 0.3


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                       # different initial weights: the learner must broadcast rank 0's
    pol = BallbotPolicy(cameras=False, hidden=16)
    cfg = PPOConfig(n_steps=8, batch_size=16, n_epochs=2, learning_rate=1e-3, clip_range=0.2)
    L = PPOLearner(StubVecEnv(8, seed=rank), pol, cfg, total_timesteps=10 ** 6, gae_fn=_gae_torch, seed=3)
    buf, stats = L.collect()
    L.update(buf)
    out[rank] = (torch.cat([p.detach().reshape(-1) for p in pol.parameters()]).numpy(), stats["env_steps"], L.num_timesteps)
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_keeps_replicas_identical():
    mgr = mp.Manager(); out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    (p0, s0, n0), (p1, s1, n1) = out[0], out[1]
    assert np.array_equal(p0, p1)                       # different shards, same averaged gradients => identical replicas
    assert s0 == s1 == 2 * 8 * 8 and n0 == n1 == 2 * 8 * 8


def _worker_uneven(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(7 + rank)
    pol = BallbotPolicy(cameras=False, hidden=16)
    cfg = PPOConfig(n_steps=8, batch_size=24, n_epochs=2, learning_rate=1e-3, clip_range=0.2)
    n_envs = 8 if rank == 0 else 5                      # 13 envs over 2 ranks: shard_envs gives 7 / 6, any uneven split must work
    L = PPOLearner(StubVecEnv(n_envs, seed=rank), pol, cfg, total_timesteps=10 ** 6, gae_fn=_gae_torch, seed=3)
    buf, stats = L.collect()
    info = L.update(buf)
    out[rank] = (torch.cat([p.detach().reshape(-1) for p in pol.parameters()]).numpy(), stats["env_steps"], L.num_timesteps, info["minibatches_per_epoch"], info["n_updates"])
    dist.destroy_process_group()


def test_uneven_env_shards_do_not_desynchronise_the_ranks():
    """ranks with different env counts must issue the same collectives (same minibatch count), count the same global
    timesteps and weight their gradients by sample count, so that the replicas stay identical."""
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker_uneven, args=(2, _free_port(), out), nprocs=2, join=True)
    (p0, s0, n0, m0, u0), (p1, s1, n1, m1, u1) = out[0], out[1]
    assert np.array_equal(p0, p1)
    assert s0 == s1 == n0 == n1 == 8 * (8 + 5)
    assert m0 == m1 == -(-8 * 13 // 24) and u0 == u1 == 2 * m0                 # ceil(104 / 24) = 5 minibatches incl. the remainder, 2 epochs


def test_folded_encoder_equals_the_eval_mode_encoder():
    """fold_encoder (BatchNorm folded into conv / linear for the frozen encoders of the rollout collector) is the same function."""
    from openballbot_rl_b200.training.policy import fold_encoder, make_depth_encoder
    torch.manual_seed(0)
    enc = make_depth_encoder()
    for m in enc:
        if hasattr(m, "running_mean"):
            m.running_mean.normal_(); m.running_var.uniform_(0.5, 2.0); m.weight.data.normal_(); m.bias.data.normal_()
    enc.eval()
    x = torch.rand(5, 1, 64, 64)
    with torch.no_grad():
        assert (fold_encoder(enc)(x) - enc(x)).abs().max() < 1e-5
