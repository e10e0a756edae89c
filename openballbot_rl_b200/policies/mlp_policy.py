"""``Extractor``: the per-key feature extractor the reference registers as policy plugin ``"mlp"``
(ballbot_rl/policies/mlp_policy.py:7-157, ballbot_rl/policies/__init__.py:8).

Same constructor (``observation_space``, ``frozen_encoder_path``), same ``extractors`` ModuleDict layout (state-dict keys of an
SB3 ``policy.pth`` load unchanged), same validation messages for a frozen encoder, same key order = the observation space's
(alphabetical) order.  It subclasses Stable-Baselines3's ``BaseFeaturesExtractor`` when SB3 is importable (so
``policy_kwargs=dict(features_extractor_class=Extractor)`` works, train.py:38-56) and ``torch.nn.Module`` otherwise; it consumes
the GPU VecEnv's device-resident observation dict directly.
"""
from pathlib import Path

import torch

from ..training.policy import make_depth_encoder

try:  # pragma: no cover - stable_baselines3 is not installed in the build image
    from stable_baselines3.common.torch_layers import BaseFeaturesExtractor as _Base
    _HAVE_SB3 = True
except Exception:
    _Base = torch.nn.Module
    _HAVE_SB3 = False


class Extractor(_Base):
    def __init__(self, observation_space, frozen_encoder_path: str = ""):
        if _HAVE_SB3:  # pragma: no cover
            super().__init__(observation_space, features_dim=1)
        else:
            super().__init__()
        extractors = {}
        total_concat_size = 0
        for key, subspace in observation_space.spaces.items():
            if "rgbd_" in key:
                C, H, W = subspace.shape
                if not frozen_encoder_path:
                    self.out_sz = 20
                    extractors[key] = make_depth_encoder(H, W, in_c=C, out_sz=self.out_sz)
                else:
                    encoder_path = Path(frozen_encoder_path).resolve()
                    enc = torch.load(str(encoder_path), map_location="cpu", weights_only=False)
                    first_conv = next((m for m in enc.modules() if isinstance(m, torch.nn.Conv2d)), None)
                    if first_conv is not None and C != first_conv.in_channels:
                        exp = first_conv.in_channels
                        raise ValueError(
                            f"Channel mismatch: Encoder expects {exp} channels (trained with {'depth-only' if exp == 1 else 'RGB-D'}), "
                            f"but environment provides {C} channels ({'depth-only' if C == 1 else 'RGB-D'}). "
                            f"Set camera.disable_rgb={'true' if exp == 1 else 'false'} in your environment config to match the encoder.")
                    linear = next((m for m in enc.modules() if isinstance(m, torch.nn.Linear)), None)
                    if linear is not None and 32 * H * W // 16 != linear.in_features:
                        hw = int((linear.in_features * 16 / 32) ** 0.5)
                        raise ValueError(
                            f"Image size mismatch: Encoder expects {hw}x{hw} images (produces {linear.in_features} features after conv layers), "
                            f"but environment provides {H}x{W} images (would produce {32 * H * W // 16} features). "
                            f"Set camera.height={hw} and camera.width={hw} in your environment config to match the encoder.")
                    self.out_sz = [m for m in enc.modules() if isinstance(m, torch.nn.Linear)][-1].out_features
                    for p in enc.parameters():            # kept frozen
                        p.requires_grad = False
                    extractors[key] = enc
                total_concat_size += self.out_sz
            else:
                extractors[key] = torch.nn.Flatten()
                total_concat_size += subspace.shape[0]
        self.extractors = torch.nn.ModuleDict(extractors)
        self._features_dim = total_concat_size

    @property
    def features_dim(self) -> int:
        return self._features_dim

    def forward(self, observations) -> torch.Tensor:
        return torch.cat([extractor(observations[key]) for key, extractor in self.extractors.items()], dim=1)
