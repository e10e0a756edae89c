"""Device-resident PPO for the GPU VecEnv: rollout collector with a frozen-encoder embedding cache, GAE on device
(``bb_gae``), clipped-surrogate updates with the gradient all-reduce over ``torch.distributed``.

Stands in for ``PPO("MultiInputPolicy", vec_env, ...).learn(...)`` of the reference (ballbot_rl/training/train.py:126-141,
284; hyper-parameters configs/train/ppo_directional.yaml:24-99; learning-rate steps ballbot_rl/training/schedules.py:4-19):
same objective (clip 0.015, vf 2.0, entropy 1e-3, target-KL early stop at 1.5 x 0.3, AdamW weight decay 0.01, GAE
0.99 / 0.95, no advantage normalisation), same policy architecture (``BallbotPolicy``).  Differences that the batch size
forces: the rollout buffer stores the 56 policy *features* (proprio 16 + 2 x 20 frozen depth embeddings) instead of raw
images -- the encoders are frozen in the reference too (mlp_policy.py:129-131), so the features are exactly what the MLPs
see -- and an embedding is recomputed only when its camera refreshed (every 6th step, ballbot_env.py:743-767).

Multi-GPU: every rank owns a shard of the envs and of the rollout (shards may be uneven: the ranks agree once per iteration
on the global sample count and the number of minibatches, and every gradient is weighted by its local sample count).  The only
collectives are ONE all-reduce per minibatch of a persistent flat buffer -- the gradients of every trainable parameter (which
are views of it) plus {KL x samples, samples} in its tail -- one scalar all-reduce per epoch for the early-stop flag, and the
rollout-statistics reduce.  On CUDA the step after the all-reduce is the fused kernel ``bb_adamw_step``
(csrc/bb_rollout.cu): mean, target-KL stop (sticky device flag, no host sync per minibatch), gradient clipping and AdamW.

Differences from SB3 that the batch size forces: ``batch_size`` is the GLOBAL minibatch size; the reference's 256 is for
10 envs x 2048 steps (80 minibatches per epoch) -- scale it with the env count (e.g. 65,536 envs x 128 steps -> 2^19) to keep
a comparable number of updates per iteration.
"""
from dataclasses import dataclass
from typing import Callable, Dict, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from .policy import FEATURE_ORDER, BallbotPolicy, fold_encoder
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
    cuda_graph: bool = True              # CUDA tensors: capture the minibatch forward / backward once and replay it (update loop is launch bound otherwise)


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
        if _world() > 1:                            # identical initial weights on every rank
            for p in policy.parameters():
                dist.broadcast(p.data, src=0)
        # ---- persistent flat buffers: parameters and gradients are views, so one all-reduce and one fused step serve them all
        n = sum(p.numel() for p in self.params)
        self.n_param = n
        self.flat_p = torch.empty(n, device=self.device); self.flat_g = torch.zeros(n + 2, device=self.device)
        self.adam_m = torch.zeros(n, device=self.device); self.adam_v = torch.zeros(n, device=self.device)
        o = 0
        for p in self.params:
            k = p.numel()
            self.flat_p[o:o + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[o:o + k].view_as(p)
            p.grad = self.flat_g[o:o + k].view_as(p)
            o += k
        self.ctrl = torch.zeros(8, dtype=torch.float64, device=self.device)       # stop flag, step count, grad norm, KL, updates
        self._scratch = torch.zeros(1, dtype=torch.float64, device=self.device)
        self.lr = lr_schedule(1.0) if cfg.learning_rate < 0 else cfg.learning_rate
        self.betas, self.eps = (0.9, 0.999), 1e-8
        self.gen = torch.Generator(device=self.device); self.gen.manual_seed(seed + 7919 * (dist.get_rank() if _world() > 1 else 0))
        self.num_timesteps = 0
        self._emb = None
        self._obs = None
        self._buf = None
        self._enc_folded = None
        self._heads = None           # captured policy step of the rollout
        self._graph = None           # (key, CUDAGraph, static index tensor) of the captured minibatch forward / backward
        self._acc = torch.zeros(3, device=self.device)     # policy loss, value loss, entropy summed on the device
        self.timing = {"collect_s": 0.0, "update_s": 0.0, "allreduce_s": 0.0}

    # ------------------------------------------------------------------ features with the embedding cache
    @torch.no_grad()
    def features(self, obs: Dict[str, torch.Tensor]) -> torch.Tensor:
        pol = self.policy
        N = obs["actions"].shape[0]
        if pol.cameras:
            enc = self._frozen_encoders()
            if self._emb is None:
                self._emb = {k: enc[k](obs[k]) for k in ("rgbd_0", "rgbd_1")}
            else:   # relative_image_timestamp == 0 <=> the cameras were rendered for this observation
                fresh = torch.nonzero(obs["relative_image_timestamp"].reshape(N) == 0).flatten()
                nf = fresh.numel()
                if nf:
                    # pad the batch to a multiple of 512 with repeats of its first env (recomputing an embedding from the
                    # same image is idempotent): a handful of batch shapes instead of a new one every step
                    pad = (-nf) % 512 if nf < N else 0
                    if pad:
                        fresh = torch.cat([fresh, fresh[:1].expand(pad)])
                    for k in ("rgbd_0", "rgbd_1"):
                        self._emb[k][fresh] = enc[k](obs[k][fresh])
        parts = []
        for k in FEATURE_ORDER:
            if k.startswith("rgbd_"):
                if pol.cameras:
                    parts.append(self._emb[k])
            elif k in obs:
                parts.append(obs[k].flatten(1))
        return torch.cat(parts, dim=1)

    def _frozen_encoders(self):
        """The frozen depth encoders with their eval-mode BatchNorms folded into the preceding conv / linear layers
        (same function, three elementwise passes over the activations fewer); rebuilt when the weights change."""
        pol = self.policy
        ver = tuple(p._version for p in pol.encoders.parameters()) + tuple(b._version for b in pol.encoders.buffers())
        if self._enc_folded is None or self._enc_folded[0] != ver:
            self._enc_folded = (ver, {k: fold_encoder(pol.encoders[k]) for k in pol.encoders})
        return self._enc_folded[1]

    @torch.no_grad()
    def _policy_heads_graph(self, N: int, F: int):
        """CUDA graph of the rollout's policy step: features -> (sampled action, clipped action, log-probability, value), static
        input / output tensors.  The parameters are views of the persistent flat buffer, so the optimiser's in-place updates are
        seen by the replays; the noise is drawn outside (generator state) into the static tensor."""
        if self._heads is not None and self._heads[0] == (N, F):
            return self._heads[1]
        dev = self.device
        f_in = torch.zeros(N, F, device=dev); noise = torch.zeros(N, 3, device=dev)

        def fn():
            mean, log_std = self._dist(f_in)
            a = mean + noise * log_std.exp()
            return a, self._log_prob(a, mean, log_std), self._value(f_in), a.clamp(-1.0, 1.0)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            a, logp, val, a_clip = fn()
        self._heads = ((N, F), (g, f_in, noise, a, logp, val, a_clip))
        return self._heads[1]

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
        if self._buf is None or self._buf["feat"].shape != (T, N, F):   # persistent rollout storage: fixed addresses for the captured update
            self._buf = dict(feat=torch.empty(T, N, F, device=dev), act=torch.empty(T, N, 3, device=dev), logp=torch.empty(T, N, device=dev),
                             rew=torch.empty(T, N, device=dev), done=torch.empty(T, N, dtype=torch.uint8, device=dev), val=torch.empty(T + 1, N, device=dev),
                             adv=torch.empty(T, N, device=dev), ret=torch.empty(T, N, device=dev))
            self._graph = None
        buf = self._buf
        ep_r = torch.zeros(N, device=dev); ep_l = torch.zeros(N, dtype=torch.int32, device=dev); ep_d = torch.zeros(N, dtype=torch.bool, device=dev)
        feat = feat0
        heads = self._policy_heads_graph(N, F) if (cfg.cuda_graph and feat0.is_cuda) else None
        for t in range(T):
            if heads is not None:      # captured: action / value heads, sampling arithmetic, log-probability (~45 launches -> 1)
                g, f_in, noise, a, logp, val, a_clip = heads
                f_in.copy_(feat); noise.normal_(generator=self.gen)
                g.replay()
                buf["feat"][t], buf["act"][t], buf["logp"][t], buf["val"][t] = feat, a, logp, val
                act_env = a_clip
            else:
                mean, log_std = self._dist(feat)
                a = mean + torch.randn(mean.shape, device=dev, generator=self.gen) * log_std.exp()
                buf["feat"][t], buf["act"][t], buf["logp"][t], buf["val"][t] = feat, a, self._log_prob(a, mean, log_std), self._value(feat)
                act_env = a.clamp(-1.0, 1.0)
            obs, r, d, info = venv.step(act_env)                      # SB3 clips to the action space before env.step
            buf["rew"][t], buf["done"][t] = r, d
            ep_r = torch.where(d, info["episode_r"], ep_r); ep_l = torch.where(d, info["episode_l"], ep_l); ep_d |= d
            feat = self.features(obs)
        buf["val"][T] = self._value(feat)
        self._obs = obs
        adv, ret = self.gae_fn(buf["rew"], buf["val"], buf["done"], cfg.gamma, cfg.gae_lambda)
        buf["adv"].copy_(adv); buf["ret"].copy_(ret)
        stats = reduce_rollout_stats(ep_r, ep_l, ep_d, steps=T * N)
        self.num_timesteps += stats["env_steps"]                     # global count (SUM over ranks): identical on every rank
        return buf, stats

    # ------------------------------------------------------------------ update
    def _optimizer_step(self, kl_limit: float):
        """mean over the global minibatch -> KL early stop -> clip -> AdamW on the flat buffers (fused kernel on CUDA)."""
        cfg, n = self.cfg, self.n_param
        if self.flat_p.is_cuda:
            import ctypes as C
            from .. import _lib
            vp = lambda t: C.c_void_p(t.data_ptr())
            with torch.cuda.device(self.device):
                rc = _lib.lib().bb_adamw_step(vp(self.flat_p), vp(self.flat_g), vp(self.adam_m), vp(self.adam_v), n, self.lr, self.betas[0], self.betas[1],
                                              self.eps, cfg.weight_decay, cfg.max_grad_norm, kl_limit, vp(self.ctrl), vp(self._scratch),
                                              C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
            if rc != 0:
                raise RuntimeError(f"bb_adamw_step failed ({rc})")
            return
        # host-logic path for CPU tensors (gloo tests of the learner): the same arithmetic with torch ops
        g, cnt = self.flat_g[:n], float(self.flat_g[n + 1]) or 1.0
        kl = float(self.flat_g[n]) / cnt
        if float(self.ctrl[0]) != 0.0 or (kl_limit > 0 and kl > kl_limit):
            self.ctrl[0] = 1.0; self.ctrl[3] = kl
            return
        gn = float(g.norm()) / cnt
        scale = (cfg.max_grad_norm / (gn + 1e-6) if cfg.max_grad_norm > 0 and gn > cfg.max_grad_norm else 1.0) / cnt
        step = float(self.ctrl[1]) + 1.0
        b1, b2 = self.betas
        gi = g * scale
        self.adam_m.mul_(b1).add_(gi, alpha=1 - b1); self.adam_v.mul_(b2).addcmul_(gi, gi, value=1 - b2)
        self.flat_p.mul_(1 - self.lr * cfg.weight_decay)
        self.flat_p.addcdiv_(self.adam_m / (1 - b1 ** step), (self.adam_v / (1 - b2 ** step)).sqrt() + self.eps, value=-self.lr)
        self.ctrl[1] = step; self.ctrl[2] = gn; self.ctrl[3] = kl; self.ctrl[4] += 1

    def update(self, buf) -> Dict[str, float]:
        cfg, dev = self.cfg, self.device
        T, N, F = buf["feat"].shape
        n = T * N
        feat = buf["feat"].reshape(n, F); act = buf["act"].reshape(n, 3); old_logp = buf["logp"].reshape(n)
        adv = buf["adv"].reshape(n); ret = buf["ret"].reshape(n)
        if cfg.learning_rate < 0:
            self.lr = lr_schedule(1.0 - min(1.0, self.num_timesteps / self.total_timesteps))
        # ---- the ranks agree on the minibatch count once per iteration; shards may be uneven: every rank cuts its
        # own samples into the same number of chunks, and gradients are summed weighted by chunk size, divided by the global size
        world = _world()
        n_glob = torch.tensor([float(n)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(n_glob, op=dist.ReduceOp.SUM)
        n_mb = max(1, int(-(-int(n_glob.item()) // max(1, cfg.batch_size))))       # ceil: the remainder minibatch is trained on, as in SB3
        n_mb = min(n_mb, n) if n > 0 else n_mb
        # equal chunks of b0 samples (the last one takes the remainder): one shape for the captured minibatch
        b0 = -(-n // n_mb) if n > 0 else 0
        bounds = [min(n, k * b0) for k in range(n_mb + 1)]
        kl_limit = 1.5 * cfg.target_kl if cfg.target_kl is not None else -1.0
        self.ctrl[0] = 0.0; self.ctrl[4] = 0.0
        acc = self._acc; acc.zero_()
        ar_events = []
        import time as _time

        def minibatch(idx):
            """zero the flat gradient, forward / backward of one minibatch (accumulates into the flat buffer's views), KL tail"""
            b = idx.numel()
            self.flat_g.zero_()
            mean, log_std = self._dist(feat[idx])
            logp = self._log_prob(act[idx], mean, log_std)
            a = adv[idx]
            if cfg.normalize_advantage and b > 1:
                a = (a - a.mean()) / (a.std() + 1e-8)
            lr_ = logp - old_logp[idx]
            ratio = lr_.exp()
            pl = -torch.min(a * ratio, a * ratio.clamp(1 - cfg.clip_range, 1 + cfg.clip_range)).mean()
            vl = nn.functional.mse_loss(self._value(feat[idx]), ret[idx])
            ent = (log_std + 1.4189385332046727).sum(-1).mean()
            loss = (pl + cfg.vf_coef * vl - cfg.ent_coef * ent) * float(b)         # weighted by the local sample count
            loss.backward()
            with torch.no_grad():
                d_ = lr_.detach()
                self.flat_g[self.n_param:self.n_param + 1].copy_(((d_.exp() - 1) - d_).sum().reshape(1))
                self.flat_g[self.n_param + 1:self.n_param + 2].fill_(float(b))        # (fill_: no host-to-device copy, capturable)
                acc.add_(torch.stack([pl.detach(), vl.detach(), ent.detach()]))

        graph = None
        if cfg.cuda_graph and self.flat_g.is_cuda and b0 > 0:
            key = (feat.data_ptr(), n, b0, F)
            if self._graph is None or self._graph[0] != key:
                idx_static = torch.zeros(b0, dtype=torch.long, device=dev)
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):                      # warm-up on a side stream (lazy cuBLAS / autograd initialisation)
                    for _ in range(3):
                        minibatch(idx_static)
                torch.cuda.current_stream(dev).wait_stream(side)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    minibatch(idx_static)
                self._graph = (key, g, idx_static)
                acc.zero_()
            graph = self._graph
        for epoch in range(cfg.n_epochs):
            perm = torch.randperm(n, device=dev, generator=self.gen)
            for k in range(n_mb):
                idx = perm[bounds[k]:bounds[k + 1]]
                b = idx.numel()
                if graph is not None and b == b0:
                    graph[2].copy_(idx)
                    graph[1].replay()
                elif b:
                    minibatch(idx)
                else:
                    self.flat_g.zero_()
                if world > 1:
                    if self.flat_g.is_cuda:      # device time of the collective: CUDA events on the stream NCCL synchronises with
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)   # the PPO gradient all-reduce (NCCL over NVLink): one per minibatch
                        e1.record(); ar_events.append((e0, e1))
                    else:
                        t0 = _time.perf_counter()
                        dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)
                        self.timing["allreduce_s"] += _time.perf_counter() - t0
                self._optimizer_step(kl_limit)
            if kl_limit > 0 and float(self.ctrl[0]) != 0.0:              # ONE host sync per epoch: the reduced KL is identical on every rank
                break
        c = self.ctrl.tolist()                       # (synchronises the stream: the events below are complete)
        self.timing["allreduce_s"] += sum(a.elapsed_time(b) for a, b in ar_events) * 1e-3
        n_updates = int(c[4])
        m = (acc / max(1, n_mb * (epoch + 1))).tolist()
        return {"policy_loss": m[0], "value_loss": m[1], "entropy": m[2], "approx_kl": c[3], "grad_norm": c[2], "n_updates": n_updates,
                "minibatches_per_epoch": n_mb, "early_stop": bool(c[0])}

    def learn(self, total_timesteps: Optional[int] = None, callback: Optional[Callable] = None):
        import time as _time
        if total_timesteps is not None:
            self.total_timesteps = int(total_timesteps)
        sync = (lambda: torch.cuda.synchronize(self.device)) if self.flat_p.is_cuda else (lambda: None)
        while self.num_timesteps < self.total_timesteps:
            sync(); t0 = _time.perf_counter()
            buf, stats = self.collect()
            sync(); t1 = _time.perf_counter()
            info = self.update(buf)
            sync(); t2 = _time.perf_counter()
            self.timing["collect_s"] += t1 - t0; self.timing["update_s"] += t2 - t1
            if callback is not None:
                callback({**stats, **info, "num_timesteps": self.num_timesteps})
        return self
