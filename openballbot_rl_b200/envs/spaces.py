"""Observation / action spaces (ballbot_gym/envs/observation_spaces.py:9-82, ballbot_env.py:235-238).

gymnasium is used when it is importable; otherwise two minimal stand-ins with the attributes the callers read
(``shape``, ``dtype``, ``low``, ``high``, ``spaces``, ``sample``, ``contains``) keep the package self-contained.
"""
from collections import OrderedDict

import numpy as np

try:  # pragma: no cover - depends on the environment
    import gymnasium as _gym
    Box, Dict = _gym.spaces.Box, _gym.spaces.Dict
    HAVE_GYMNASIUM = True
except Exception:  # gymnasium is not installed in the build image
    HAVE_GYMNASIUM = False

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            self.shape = tuple(shape) if shape is not None else np.shape(low)
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Dict:
        def __init__(self, spaces):
            # gymnasium sorts the keys of a plain dict alphabetically: the policy's feature order
            self.spaces = OrderedDict(sorted(spaces.items()))

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def sample(self):
            return OrderedDict((k, s.sample()) for k, s in self.spaces.items())

        def contains(self, x):
            return set(x.keys()) == set(self.spaces.keys()) and all(s.contains(x[k]) for k, s in self.spaces.items())

_F32 = np.float32


def create_observation_space(im_shape: dict, num_channels: int, disable_cameras: bool):
    """Same keys, bounds, shapes and dtypes as the reference (incl. ``relative_image_timestamp`` being declared in the
    camera-less space although it is never emitted there)."""
    spaces = {
        "orientation": Box(low=-np.pi, high=np.pi, shape=(3,), dtype=_F32),
        "angular_vel": Box(low=-2, high=2, shape=(3,), dtype=_F32),
        "vel": Box(low=-2, high=2, shape=(3,), dtype=_F32),
        "motor_state": Box(-2.0, 2.0, shape=(3,), dtype=_F32),
        "actions": Box(-1.0, 1.0, shape=(3,), dtype=_F32),
        "relative_image_timestamp": Box(low=0.0, high=0.1, shape=(1,), dtype=_F32),
    }
    if not disable_cameras:
        for cam in ("rgbd_0", "rgbd_1"):
            spaces[cam] = Box(low=0.0, high=1.0, shape=(num_channels, im_shape["h"], im_shape["w"]), dtype=_F32)
    return Dict(spaces)


def create_action_space():
    return Box(-1.0, 1.0, shape=(3,), dtype=_F32)
