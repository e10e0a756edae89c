"""Depth-frame collection for encoder pre-training from the GPU VecEnv.

Stands in for ``ballbot_rl/data/collect.py:17-47`` (a policy drives ``SubprocVecEnv`` envs whose ``log_options={"cams": True}``
dump every rendered depth frame under ``/tmp/log_*``, ``ballbot_gym/utils/logging.py:52-117``).  Here the frames never pass
through per-env log directories: every step the freshly rendered images (relative_image_timestamp == 0, i.e. one in six
steps per env plus resets) are gathered on the device and flushed to ``.npy`` shards of ``[frames, 64, 64]`` float16.
"""
import os
from typing import Callable, Optional

import numpy as np
import torch


@torch.no_grad()
def collect_depth_frames(venv, policy: Optional[Callable], n_steps: int, out_dir: str, shard_frames: int = 65536,
                         deterministic: bool = True, dtype=np.float16) -> int:
    """Runs ``n_steps`` env steps (``policy=None``: uniform random actions) and writes the fresh depth frames of both
    cameras to ``out_dir/depth_XXXX.npy``.  Returns the number of frames written."""
    os.makedirs(out_dir, exist_ok=True)
    dev, N = venv.engine.device, venv.num_envs
    gen = torch.Generator(device=dev); gen.manual_seed(0)
    obs = venv.reset()
    pending, n_pending, n_written, shard = [], 0, 0, 0

    def flush(final=False):
        nonlocal pending, n_pending, n_written, shard
        while n_pending >= shard_frames or (final and n_pending > 0):
            buf = torch.cat(pending, 0)
            take = min(shard_frames, buf.shape[0])
            np.save(os.path.join(out_dir, f"depth_{shard:04d}.npy"), buf[:take].cpu().numpy().astype(dtype))
            n_written += take; shard += 1
            pending = [buf[take:]] if buf.shape[0] > take else []
            n_pending = buf.shape[0] - take

    for t in range(n_steps + 1):
        fresh = obs["relative_image_timestamp"].reshape(N) == 0
        if bool(fresh.any()):
            frames = torch.cat([obs["rgbd_0"][fresh, 0], obs["rgbd_1"][fresh, 0]], 0)
            pending.append(frames.clone()); n_pending += frames.shape[0]
            flush()
        if t == n_steps:
            break
        a = policy(obs, deterministic=deterministic) if policy is not None else torch.rand(N, 3, device=dev, generator=gen) * 2 - 1
        obs, _, _, _ = venv.step(a)
    flush(final=True)
    return n_written
