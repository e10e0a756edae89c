"""Host-side (numpy) terrain plugins: the reference's non-Perlin generators (ballbot_gym/terrain/*.py).

These are plugin terrains off the hot path: they run on the host at reset and are uploaded with bb_set_hfield
(SURVEY.md section 2 #4).  Each function is a vectorised re-implementation checked against golden fixtures that were
generated from the reference's own modules (tests/golden/make_golden.py).
"""
from typing import Callable, Dict, Optional

import numpy as np


def _normalise(t: np.ndarray) -> np.ndarray:
    lo, hi = t.min(), t.max()
    return (t - lo) / (hi - lo) if hi > lo else np.zeros_like(t)


def _smoothstep01(x):
    x = np.clip(x, 0.0, 1.0)
    return x * x * (3.0 - 2.0 * x)


def generate_stepped_terrain(n: int, num_steps: int = 5, step_height: float = 0.1, seed: Optional[int] = None) -> np.ndarray:
    """terrain/stepped.py: diagonal staircase, one Gauss-Seidel-style smoothing sweep in row-major order, normalised."""
    assert n % 2 == 1, "n should be odd for heightfield symmetry"
    assert num_steps > 0, "num_steps must be positive"
    assert step_height > 0, "step_height must be positive"
    size = n // num_steps
    idx = np.arange(n) // size
    t = np.minimum(idx[:, None] + idx[None, :], num_steps - 1) * float(step_height)
    t = t.astype(np.float64)
    # the reference smooths IN PLACE while sweeping i, j upward: (i-1, j) and (i, j-1) are already updated values
    for i in range(1, n - 1):
        up, cur, down = t[i - 1], t[i], t[i + 1]
        for j in range(1, n - 1):
            cur[j] = 0.7 * cur[j] + 0.3 * np.mean([up[j], down[j], cur[j - 1], cur[j + 1]])
    return _normalise(t).flatten()


def generate_ramp_terrain(n: int, ramp_angle: float = 15.0, ramp_direction: str = "x", flat_ratio: float = 0.3, num_ramps: int = 1,
                          transition_smoothness: float = 0.5, seed: Optional[int] = None) -> np.ndarray:
    """terrain/ramp.py: smooth-stepped ramps along x, y or radially, normalised."""
    assert n % 2 == 1, "n should be odd for heightfield symmetry"
    assert 0 <= ramp_angle <= 45, "ramp_angle should be between 0 and 45 degrees"
    assert 0 <= flat_ratio <= 1.0, "flat_ratio should be between 0 and 1"
    assert num_ramps > 0, "num_ramps must be positive"
    assert ramp_direction in ["x", "y", "radial"], "ramp_direction must be 'x', 'y', or 'radial'"
    hmax = np.tan(np.radians(ramp_angle)) * 2.0
    c = n // 2
    u = (np.arange(n) - c) / c
    X, Y = np.meshgrid(u, u, indexing="ij")
    if ramp_direction in ("x", "y"):
        v = X if ramp_direction == "x" else Y
        if num_ramps == 1:
            fw = flat_ratio / 2.0
            with np.errstate(divide="ignore", invalid="ignore"):
                mid = _smoothstep01((v + fw) / (fw * 2)) * hmax
            t = np.where(v < -fw, 0.0, np.where(v < fw, mid, hmax))
        else:
            period = 2.0 / num_ramps
            ph = ((v + 1.0) % period) / period
            with np.errstate(divide="ignore", invalid="ignore"):
                mid = _smoothstep01((ph - flat_ratio / 2) / (1.0 - flat_ratio)) * hmax
            t = np.where(ph < flat_ratio / 2, 0.0, np.where(ph < 1.0 - flat_ratio / 2, mid, hmax))
    else:
        R = np.sqrt(X ** 2 + Y ** 2)
        rmax = np.sqrt(2.0)
        rflat = flat_ratio * rmax / np.sqrt(2.0)
        t = np.where(R < rflat, 0.0, _smoothstep01(np.clip((R - rflat) / (rmax - rflat), 0.0, 1.0)) * hmax)
    return _normalise(t.astype(np.float64)).flatten()


def generate_sinusoidal_terrain(n: int, amplitude: float = 0.5, frequency: float = 0.1, direction: str = "both", phase: float = 0.0,
                                seed: Optional[int] = None) -> np.ndarray:
    """terrain/sinusoidal.py."""
    assert n % 2 == 1, "n should be odd for heightfield symmetry"
    assert 0 <= amplitude <= 1.0, "amplitude should be between 0 and 1"
    assert frequency > 0, "frequency must be positive"
    assert direction in ["x", "y", "both"], "direction must be 'x', 'y', or 'both'"
    g = np.linspace(0, 2 * np.pi * frequency * n, n)
    X, Y = np.meshgrid(g, g, indexing="ij")
    if direction == "x":
        t = amplitude * np.sin(X + phase)
    elif direction == "y":
        t = amplitude * np.sin(Y + phase)
    else:
        t = amplitude * (np.sin(X + phase) + np.sin(Y + phase)) / 2.0
    return _normalise(t).flatten()


def generate_hills_terrain(n: int, num_hills: int = 5, hill_height: float = 0.7, hill_radius: float = 0.15, flat_ratio: float = 0.4,
                           seed: Optional[int] = None) -> np.ndarray:
    """terrain/hills.py: rejection-sampled Gaussian bumps with a smooth-step cut-off, clipped to [0, 1] (not normalised)."""
    assert n % 2 == 1, "n should be odd for heightfield symmetry"
    assert num_hills > 0, "num_hills must be positive"
    assert 0 <= hill_height <= 1.0, "hill_height should be between 0 and 1"
    assert 0 < hill_radius <= 0.5, "hill_radius should be between 0 and 0.5"
    rng = np.random.RandomState(seed if seed is not None else 0)
    centres = []
    tries = 0
    while len(centres) < num_hills and tries < num_hills * 100:
        tries += 1
        x = rng.uniform(hill_radius, 1.0 - hill_radius)
        y = rng.uniform(hill_radius, 1.0 - hill_radius)
        if all(np.sqrt((x - cx) ** 2 + (y - cy) ** 2) >= hill_radius * 2.0 for cx, cy in centres):
            centres.append((x, y))
    g = np.linspace(0, 1, n)
    X, Y = np.meshgrid(g, g, indexing="ij")
    t = np.zeros((n, n))
    sigma = hill_radius / 3.0
    for cx, cy in centres:
        r = np.sqrt((X - cx) ** 2 + (Y - cy) ** 2)
        cut = np.clip(1.0 - (r / hill_radius), 0.0, 1.0)
        t += hill_height * np.exp(-(r ** 2) / (2 * sigma ** 2)) * (cut * cut * (3.0 - 2.0 * cut))
    return np.clip(t, 0.0, 1.0).flatten()


def _grid01(n):
    g = np.linspace(0, 1, n)
    return np.meshgrid(g, g, indexing="ij")


def generate_ridge_valley_terrain(n: int, ridge_height: float = 0.6, valley_depth: float = 0.4, spacing: float = 0.2, orientation: str = "x",
                                  smoothness: float = 0.3, seed: Optional[int] = None) -> np.ndarray:
    """terrain/ridge_valley.py: cosine ridges blended with a box-filtered copy (edge padding), clipped."""
    assert n % 2 == 1, "n should be odd for heightfield symmetry"
    assert 0 <= ridge_height <= 1.0, "ridge_height should be between 0 and 1"
    assert 0 <= valley_depth <= 1.0, "valley_depth should be between 0 and 1"
    assert spacing > 0, "spacing must be positive"
    assert orientation in ["x", "y", "diagonal"], "orientation must be 'x', 'y', or 'diagonal'"
    X, Y = _grid01(n)
    arg = X if orientation == "x" else (Y if orientation == "y" else X + Y)
    t = valley_depth + (ridge_height - valley_depth) * (np.cos(2 * np.pi * spacing * arg) + 1.0) / 2.0
    if smoothness > 0:
        k = int(smoothness * 5) + 1
        if k > 1:
            padded = np.pad(t, k // 2, mode="edge")
            win = np.lib.stride_tricks.sliding_window_view(padded, (k, k))[:n, :n]
            smoothed = win.reshape(n, n, -1).mean(axis=-1)
            t = t * (1.0 - smoothness) + smoothed * smoothness
    return np.clip(t, 0.0, 1.0).flatten()


def generate_bowl_terrain(n: int, depth: float = 0.6, radius: float = 0.4, center_x: float = 0.5, center_y: float = 0.5,
                          smoothness: float = 0.5, seed: Optional[int] = None) -> np.ndarray:
    """terrain/bowl.py: radial smooth-step depression below a unit plateau."""
    assert n % 2 == 1, "n should be odd for heightfield symmetry"
    assert 0 <= depth <= 1.0, "depth should be between 0 and 1"
    assert 0 < radius <= 1.0, "radius should be between 0 and 1"
    assert 0 <= center_x <= 1.0, "center_x should be between 0 and 1"
    assert 0 <= center_y <= 1.0, "center_y should be between 0 and 1"
    X, Y = _grid01(n)
    r = np.sqrt((X - center_x) ** 2 + (Y - center_y) ** 2)
    t = np.ones((n, n)) - depth * (1.0 - _smoothstep01(np.clip(r / radius, 0.0, 1.0)))
    return np.clip(t, 0.0, 1.0).flatten()


def generate_gradient_terrain(n: int, max_slope: float = 20.0, gradient_type: str = "linear", smoothness: float = 0.5, direction: str = "x",
                              seed: Optional[int] = None) -> np.ndarray:
    """terrain/gradient.py:7-93. The 'perlin' variant modulates the linear ramp with the untiled 2-D ``noise.snoise2``
    (3 octaves, persistence 0.3, base = seed), evaluated by the engine's device code (``bb_snoise2_grid``; needs a CUDA device)."""
    assert n % 2 == 1, "n should be odd for heightfield symmetry"
    assert 0 <= max_slope <= 45, "max_slope should be between 0 and 45 degrees"
    assert gradient_type in ["linear", "radial", "perlin"], "gradient_type must be 'linear', 'radial', or 'perlin'"
    assert direction in ["x", "y"], "direction must be 'x' or 'y'"
    hmax = np.tan(np.radians(max_slope)) * 2.0
    c = n // 2
    u = (np.arange(n) - c) / c
    X, Y = np.meshgrid(u, u, indexing="ij")
    if gradient_type == "linear":
        t = hmax * ((X if direction == "x" else Y) + 1.0) / 2.0
    elif gradient_type == "radial":
        t = hmax * np.clip(np.sqrt(X ** 2 + Y ** 2) / np.sqrt(2.0), 0.0, 1.0)
    else:
        import ctypes as C
        from .. import _lib
        noise = np.empty(n * n, np.float32)
        rc = _lib.lib().bb_snoise2_grid(0, int(n), 25.0, 3, 0.3, 2.0, int(seed) if seed is not None else 0, C.c_void_p(noise.ctypes.data))
        if rc != 0:
            raise _lib.EngineError(f"bb_snoise2_grid failed ({rc}): {_lib.lib().bb_last_error(None).decode()}")
        t = hmax * (((X if direction == "x" else Y) + 1.0) / 2.0 + noise.reshape(n, n).astype(np.float64) * smoothness)
    return _normalise(t).flatten()


def generate_terraced_terrain(n: int, num_terraces: int = 5, terrace_height: float = 0.15, transition_width: float = 0.1, smoothness: float = 0.7,
                              direction: str = "x", seed: Optional[int] = None) -> np.ndarray:
    """terrain/terraced.py: flat terraces joined by smooth-step ramps of relative width ``transition_width``."""
    assert n % 2 == 1, "n should be odd for heightfield symmetry"
    assert num_terraces > 0, "num_terraces must be positive"
    assert 0 < terrace_height <= 1.0, "terrace_height should be between 0 and 1"
    assert 0 < transition_width < 1.0, "transition_width should be between 0 and 1"
    assert direction in ["x", "y"], "direction must be 'x' or 'y'"
    X, Y = _grid01(n)
    coord = X if direction == "x" else Y
    width = 1.0 / num_terraces
    ts = width * transition_width
    idx = np.minimum((coord / width).astype(int), num_terraces - 1)
    pos = (coord % width) / width
    base = idx * terrace_height
    with np.errstate(divide="ignore", invalid="ignore"):
        up_from_prev = (idx - 1) * terrace_height + terrace_height * _smoothstep01(pos / ts)
        up_to_next = base + terrace_height * _smoothstep01((pos - (1.0 - ts)) / ts)
    t = np.where((pos < ts) & (idx > 0), up_from_prev, np.where((pos > 1.0 - ts) & (idx < num_terraces - 1) & ~(pos < ts), up_to_next, base))
    return np.clip(t, 0.0, 1.0).flatten()


def generate_wavy_terrain(n: int, wave_amplitudes=None, wave_frequencies=None, wave_directions=None, phase_offsets=None,
                          seed: Optional[int] = None) -> np.ndarray:
    """terrain/wavy.py: sum of plane sine waves around 0.5, clipped."""
    assert n % 2 == 1, "n should be odd for heightfield symmetry"
    amps = [0.3, 0.2, 0.1] if wave_amplitudes is None else wave_amplitudes
    freqs = [0.05, 0.1, 0.2] if wave_frequencies is None else wave_frequencies
    dirs = [0.0, 45.0, 90.0] if wave_directions is None else wave_directions
    phases = [0.0, 0.5, 1.0] if phase_offsets is None else phase_offsets
    assert len(freqs) == len(amps), "wave_frequencies must match wave_amplitudes length"
    assert len(dirs) == len(amps), "wave_directions must match wave_amplitudes length"
    assert len(phases) == len(amps), "phase_offsets must match wave_amplitudes length"
    g = np.linspace(0, 2 * np.pi, n)
    X, Y = np.meshgrid(g, g, indexing="ij")
    t = np.zeros((n, n))
    for a, f, d, ph in zip(amps, freqs, dirs, phases):
        rad = np.radians(d)
        t += a * np.sin(f * (X * np.cos(rad) + Y * np.sin(rad)) + ph)
    return np.clip(t + 0.5, 0.0, 1.0).flatten()


def generate_spiral_terrain(n: int, spiral_tightness: float = 0.1, height_variation: float = 0.5, direction: str = "cw", center_x: float = 0.5,
                            center_y: float = 0.5, seed: Optional[int] = None) -> np.ndarray:
    """terrain/spiral.py."""
    assert n % 2 == 1, "n should be odd for heightfield symmetry"
    assert spiral_tightness > 0, "spiral_tightness must be positive"
    assert 0 <= height_variation <= 1.0, "height_variation should be between 0 and 1"
    assert direction in ["cw", "ccw"], "direction must be 'cw' or 'ccw'"
    X, Y = _grid01(n)
    dx, dy = X - center_x, Y - center_y
    r = np.sqrt(dx ** 2 + dy ** 2)
    th = (np.arctan2(dy, dx) + 2 * np.pi) % (2 * np.pi)
    if direction == "cw":
        th = 2 * np.pi - th
    t = height_variation * np.sin(spiral_tightness * th + r) * (1.0 - np.clip(r / (np.sqrt(2.0) / 2.0), 0.0, 1.0) * 0.3)
    return np.clip(0.5 + t * 0.5, 0.0, 1.0).flatten()


def generate_mixed_terrain(n: int, components, blend_mode: str = "additive", seed: Optional[int] = None) -> np.ndarray:
    """terrain/mixed.py: weighted blend of other registered terrains (built through create_terrain)."""
    from ..core.factories import create_terrain
    assert n % 2 == 1, "n should be odd for heightfield symmetry"
    assert len(components) > 0, "components list cannot be empty"
    assert blend_mode in ["additive", "max", "weighted"], "blend_mode must be 'additive', 'max', or 'weighted'"
    gens, weights = [], []
    for comp in components:
        if not isinstance(comp, dict):
            raise ValueError(f"Component must be a dict, got {type(comp)}")
        if comp.get("type") is None:
            raise ValueError("Component must have 'type' key")
        cfg = {"type": comp["type"], "config": comp.get("config", {})}
        if "seed" not in cfg["config"] and seed is not None:
            cfg["config"]["seed"] = seed          # mutates the caller's dict exactly like the reference does
        gens.append(create_terrain(cfg))
        weights.append(comp.get("weight", 1.0))
    parts = [g(n, seed=seed).reshape(n, n) for g in gens]
    t = np.zeros((n, n))
    total = sum(weights)
    if blend_mode == "additive":
        for p, w in zip(parts, weights):
            t += p * (w / total)
    elif blend_mode == "max":
        for p, w in zip(parts, weights):
            t = np.maximum(t, p * w)
    else:
        for p, w in zip(parts, weights):
            t += p * w
        t = t / total
    return np.clip(t, 0.0, 1.0).flatten()


GENERATORS: Dict[str, Callable] = {
    "stepped": generate_stepped_terrain,
    "ramp": generate_ramp_terrain,
    "sinusoidal": generate_sinusoidal_terrain,
    "ridge_valley": generate_ridge_valley_terrain,
    "hills": generate_hills_terrain,
    "bowl": generate_bowl_terrain,
    "gradient": generate_gradient_terrain,
    "terraced": generate_terraced_terrain,
    "wavy": generate_wavy_terrain,
    "spiral": generate_spiral_terrain,
    "mixed": generate_mixed_terrain,
}
