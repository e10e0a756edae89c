"""Terrain generators ``(n, **cfg) -> flat array in [0, 1]`` and their registration (ballbot_gym/terrain/__init__.py:18-36).

``perlin`` and ``flat`` are the two terrains on the hot path: the batched engine generates them on the GPU per reset
(k_terrain).  The registry callable ``perlin`` exposed here runs the same device code for arbitrary ``n`` through the
C ABI (bb_perlin_grid); it needs a CUDA device -- there is no CPU implementation in the product.
"""
import ctypes as C

import numpy as np

from ..core.registry import ComponentRegistry


def generate_perlin_terrain(n: int, scale: float = 25.0, octaves: int = 4, persistence: float = 0.2, lacunarity: float = 2.0,
                            amplitude: float = 1.0, seed: int = 0, device: int = 0) -> np.ndarray:
    """terrain/perlin.py:8-74: 4-octave simplex fBm (noise.snoise2 tiled at 1024) mapped to [0, 1]; shape (n*n,)."""
    assert n % 2 == 1, "n should be odd for heightfield symmetry"
    from .. import _lib
    L = _lib.lib()
    out = np.empty(n * n, np.float32)
    seeds = np.array([int(seed)], np.int32)
    rc = L.bb_perlin_grid(int(device), int(n), float(scale), int(octaves), float(persistence), float(lacunarity), float(amplitude),
                          C.c_void_p(seeds.ctypes.data), 1, C.c_void_p(out.ctypes.data))
    if rc != 0:
        raise _lib.EngineError(f"bb_perlin_grid failed ({rc}): {L.bb_last_error(None).decode()}")
    return out.astype(np.float64)


def generate_flat_terrain(n: int, **kwargs) -> np.ndarray:
    """terrain/__init__.py:32-36."""
    return np.zeros(n * n)


def register_builtin_terrains():
    from . import shapes
    table = {"perlin": generate_perlin_terrain, "flat": generate_flat_terrain}
    table.update(shapes.GENERATORS)
    for name, fn in table.items():
        if name not in ComponentRegistry.list_terrains():
            ComponentRegistry.register_terrain(name, fn)


register_builtin_terrains()

__all__ = ["generate_perlin_terrain", "generate_flat_terrain", "register_builtin_terrains"]
