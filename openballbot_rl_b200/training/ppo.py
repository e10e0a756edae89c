"""Device-resident PPO for the GPU VecEnv: rollout collector with a frozen-encoder embedding cache, GAE on device
(``bb_gae``), clipped-surrogate updates with the gradient all-reduce over ``torch.distributed``.

Stands in for ``PPO("MultiInputPolicy", vec_env, ...).learn(...)`` of the reference (ballbot_rl/training/train.py:126-141,
284; hyper-parameters configs/train/ppo_directional.yaml:24-99; learning-rate steps ballbot_rl/training/schedules.py:4-19):
same objective (clip 0.015, vf 2.0, entropy 1e-3, target-KL early stop at 1.5 x 0.3, AdamW weight decay 0.01, GAE
0.99 / 0.95, no advantage normalisation), same policy architecture (``BallbotPolicy``).  Differences that the batch size
forces: the rollout buffer stores the 56 policy *features* (proprio 16 + 2 x 20 frozen depth embeddings) instead of raw
images -- the encoders are frozen in the reference too (mlp_policy.py:129-131), so the features are exactly what the MLPs
see -- and an embedding is recomputed only when its camera refreshed (every 6th step, ballbot_env.py:743-767).

Multi-GPU: every rank owns a shard of the envs and of the rollout; the only collectives are the flat-gradient all-reduce
per minibatch, one scalar all-reduce per epoch for the KL stop, and the rollout-statistics reduce.
"""
from dataclasses import dataclass
from typing import Callable, Dict, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from .policy import FEATURE_ORDER, BallbotPolicy
from .rollout import reduce_rollout_stats


def lr_schedule(progress_remaining: float) -> float:
    """Piece-wise constant learning rate of the reference (ballbot_rl/training/schedules.py:4-19)."""
    if progress_remaining > 0.7:
        return 1e-4
    if 0.5 < progress_remaining < 0.7:
        return 5e-5
    return 1e-5


@dataclass
class PPOConfig:
    n_steps: int = 2048
    batch_size: int = 256
    n_epochs: int = 5
    gamma: float = 0.99
    gae_lambda: float = 0.95
    clip_range: float = 0.015
    ent_coef: float = 0.001
    vf_coef: float = 2.0
    target_kl: Optional[float] = 0.3
    max_grad_norm: float = 0.5
    weight_decay: float = 0.01
    normalize_advantage: bool = False
    learning_rate: float = -1.0          # -1: reference schedule


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def allreduce_mean_(flat: torch.Tensor) -> torch.Tensor:
    if _world() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(_world())
    return flat


class PPOLearner:
    def __init__(self, venv, policy: BallbotPolicy, cfg: PPOConfig = PPOConfig(), total_timesteps: int = 10_000_000,
                 gae_fn: Optional[Callable] = None, seed: int = 0):
        self.venv, self.policy, self.cfg = venv, policy, cfg
        self.total_timesteps = int(total_timesteps)
        if gae_fn is None:
            from .gae import compute_gae as gae_fn   # CUDA kernel through the C ABI (no CPU fallback)
        self.gae_fn = gae_fn
        self.device = next(policy.parameters()).device
        for p in policy.encoders.parameters():      # frozen pre-trained encoders (mlp_policy.py:129-131)
            p.requires_grad_(False)
        policy.encoders.eval()
        self.params = [p for p in policy.parameters() if p.requires_grad]
        self.opt = torch.optim.AdamW(self.params, lr=lr_schedule(1.0) if cfg.learning_rate < 0 else cfg.learning_rate,
                                     weight_decay=cfg.weight_decay)
        self.gen = torch.Generator(device=self.device); self.gen.manual_seed(seed + 7919 * (dist.get_rank() if _world() > 1 else 0))
        self.num_timesteps = 0
        self._emb = None
        self._obs = None
        if _world() > 1:                            # identical initial weights on every rank
            for p in policy.parameters():
                dist.broadcast(p.data, src=0)

    # ------------------------------------------------------------------ features with the embedding cache
    @torch.no_grad()
    def features(self, obs: Dict[str, torch.Tensor]) -> torch.Tensor:
        pol = self.policy
        N = obs["actions"].shape[0]
        if pol.cameras:
            if self._emb is None:
                self._emb = {k: pol.encoders[k](obs[k]) for k in ("rgbd_0", "rgbd_1")}
            else:   # relative_image_timestamp == 0 <=> the cameras were rendered for this observation
                fresh = torch.nonzero(obs["relative_image_timestamp"].reshape(N) == 0).flatten()
                if fresh.numel() == 1:          # BatchNorm1d in eval mode accepts a batch of one; keep the indexing uniform
                    fresh = fresh.repeat(2)
                if fresh.numel():
                    for k in ("rgbd_0", "rgbd_1"):
                        self._emb[k][fresh] = pol.encoders[k](obs[k][fresh])
        parts = []
        for k in FEATURE_ORDER:
            if k.startswith("rgbd_"):
                if pol.cameras:
                    parts.append(self._emb[k])
            elif k in obs:
                parts.append(obs[k].flatten(1))
        return torch.cat(parts, dim=1)

    def _dist(self, feat: torch.Tensor):
        mean = self.policy.action_net(self.policy.policy_net(feat))
        return mean, self.policy.log_std.expand_as(mean)

    @staticmethod
    def _log_prob(a, mean, log_std):
        return (-0.5 * ((a - mean) / log_std.exp()) ** 2 - log_std - 0.9189385332046727).sum(-1)

    def _value(self, feat):
        return self.policy.value_net(self.policy.value_net_body(feat)).squeeze(-1)

    # ------------------------------------------------------------------ rollout
    @torch.no_grad()
    def collect(self):
        cfg, venv = self.cfg, self.venv
        N, T, dev = venv.num_envs, cfg.n_steps, self.device
        if self._obs is None:
            self._obs = venv.reset()
        obs = self._obs
        feat0 = self.features(obs)
        F = feat0.shape[1]
        buf = dict(feat=torch.empty(T, N, F, device=dev), act=torch.empty(T, N, 3, device=dev), logp=torch.empty(T, N, device=dev),
                   rew=torch.empty(T, N, device=dev), done=torch.empty(T, N, dtype=torch.uint8, device=dev), val=torch.empty(T + 1, N, device=dev))
        ep_r = torch.zeros(N, device=dev); ep_l = torch.zeros(N, dtype=torch.int32, device=dev); ep_d = torch.zeros(N, dtype=torch.bool, device=dev)
        feat = feat0
        for t in range(T):
            mean, log_std = self._dist(feat)
            a = mean + torch.randn(mean.shape, device=dev, generator=self.gen) * log_std.exp()
            buf["feat"][t], buf["act"][t], buf["logp"][t], buf["val"][t] = feat, a, self._log_prob(a, mean, log_std), self._value(feat)
            obs, r, d, info = venv.step(a.clamp(-1.0, 1.0))          # SB3 clips to the action space before env.step
            buf["rew"][t], buf["done"][t] = r, d
            ep_r = torch.where(d, info["episode_r"], ep_r); ep_l = torch.where(d, info["episode_l"], ep_l); ep_d |= d
            feat = self.features(obs)
        buf["val"][T] = self._value(feat)
        self._obs = obs
        buf["adv"], buf["ret"] = self.gae_fn(buf["rew"], buf["val"], buf["done"], cfg.gamma, cfg.gae_lambda)
        self.num_timesteps += T * N * _world()
        stats = reduce_rollout_stats(ep_r, ep_l, ep_d, steps=T * N)
        return buf, stats

    # ------------------------------------------------------------------ update
    def update(self, buf) -> Dict[str, float]:
        cfg = self.cfg
        T, N, F = buf["feat"].shape
        n = T * N
        feat = buf["feat"].reshape(n, F); act = buf["act"].reshape(n, 3); old_logp = buf["logp"].reshape(n)
        adv = buf["adv"].reshape(n); ret = buf["ret"].reshape(n)
        if cfg.learning_rate < 0:
            lr = lr_schedule(1.0 - min(1.0, self.num_timesteps / self.total_timesteps))
            for g in self.opt.param_groups:
                g["lr"] = lr
        acc = torch.zeros(4, device=self.device)   # policy loss, value loss, entropy, KL summed on the device
        n_updates = 0
        bs = min(cfg.batch_size, n)
        stop = False
        for epoch in range(cfg.n_epochs):
            perm = torch.randperm(n, device=self.device, generator=self.gen)
            for s in range(0, n - bs + 1, bs):
                idx = perm[s:s + bs]
                mean, log_std = self._dist(feat[idx])
                logp = self._log_prob(act[idx], mean, log_std)
                a = adv[idx]
                if cfg.normalize_advantage and bs > 1:
                    a = (a - a.mean()) / (a.std() + 1e-8)
                ratio = (logp - old_logp[idx]).exp()
                pl = -torch.min(a * ratio, a * ratio.clamp(1 - cfg.clip_range, 1 + cfg.clip_range)).mean()
                vl = nn.functional.mse_loss(self._value(feat[idx]), ret[idx])
                ent = (log_std + 1.4189385332046727).sum(-1).mean()
                loss = pl + cfg.vf_coef * vl - cfg.ent_coef * ent
                with torch.no_grad():
                    lr_ = logp - old_logp[idx]
                    kl = allreduce_mean_(((lr_.exp() - 1) - lr_).mean().reshape(1))
                if cfg.target_kl is not None and float(kl) > 1.5 * cfg.target_kl:   # same decision on every rank (reduced KL)
                    stop = True
                    break
                self.opt.zero_grad(set_to_none=True)
                loss.backward()
                flat = torch.cat([p.grad.reshape(-1) for p in self.params])
                allreduce_mean_(flat)                                               # PPO gradient all-reduce (NCCL over NVLink)
                nrm = flat.norm()
                flat.mul_(torch.clamp(cfg.max_grad_norm / (nrm + 1e-6), max=1.0))
                o = 0
                for p in self.params:
                    p.grad.copy_(flat[o:o + p.numel()].view_as(p)); o += p.numel()
                self.opt.step()
                acc += torch.stack([pl.detach(), vl.detach(), ent.detach(), kl[0]])
                n_updates += 1
            if stop:
                break
        m = (acc / max(1, n_updates)).tolist()
        return {"policy_loss": m[0], "value_loss": m[1], "entropy": m[2], "approx_kl": m[3], "n_updates": n_updates}

    def learn(self, total_timesteps: Optional[int] = None, callback: Optional[Callable] = None):
        if total_timesteps is not None:
            self.total_timesteps = int(total_timesteps)
        while self.num_timesteps < self.total_timesteps:
            buf, stats = self.collect()
            info = self.update(buf)
            if callback is not None:
                callback({**stats, **info, "num_timesteps": self.num_timesteps})
        return self
