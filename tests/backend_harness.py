"""One harness for both implementations of the C ABI in include/ballbot_b200.h:
the CUDA engine (openballbot_rl_b200/libballbot_b200.so, device pointers = torch CUDA tensors) and the CPU oracle behind the
same symbols (oracle/_build/libballbot_cpu_ref.so, host pointers = numpy arrays).  Test infrastructure only."""
import ctypes as C

import numpy as np

from openballbot_rl_b200 import _lib

IO_FIELDS = [("orientation", np.float32, 3), ("angular_vel", np.float32, 3), ("vel", np.float32, 3), ("motor_state", np.float32, 3),
             ("actions", np.float32, 3), ("rel_image_ts", np.float32, 1), ("rgbd_0", np.float32, None), ("rgbd_1", np.float32, None),
             ("reward", np.float32, 0), ("terminated", np.uint8, 0), ("failure", np.uint8, 0), ("pos2d", np.float32, 2),
             ("terminal_obs", np.float32, 16), ("episode_return", np.float32, 0), ("episode_length", np.int32, 0), ("status", np.int32, 0)]


class Backend:
    """kind = "cuda" (the product library) or "cpu_ref" (the oracle library).  All inputs / outputs are numpy arrays."""

    def __init__(self, kind, **cfg_kw):
        self.kind = kind
        if kind == "cuda":
            import torch
            self.torch = torch
            self.L = _lib.lib()
        else:
            from oracle import oracle as O
            O.build()
            self.L = C.CDLL(O.CPU_REF_SO)
            vp = C.c_void_p
            self.L.bb_create.argtypes = [C.POINTER(_lib.Config), C.POINTER(vp)]
            self.L.bb_default_config.argtypes = [C.POINTER(_lib.Config)]; self.L.bb_default_config.restype = None
            self.L.bb_last_error.argtypes = [vp]; self.L.bb_last_error.restype = C.c_char_p
            self.L.bb_destroy.argtypes = [vp]
            self.L.bb_reset.argtypes = [vp, vp, vp, C.POINTER(_lib.IO), vp]
            self.L.bb_step.argtypes = [vp, vp, C.POINTER(_lib.IO), vp]
            self.L.bb_set_state.argtypes = [vp, vp, vp, vp, vp]; self.L.bb_get_state.argtypes = [vp, vp, vp, vp, vp]
            self.L.bb_set_hfield.argtypes = [vp, vp, C.c_int32, vp, vp]; self.L.bb_get_hfield.argtypes = [vp, C.c_int32, vp, vp]
            self.L.bb_get_terrain_seeds.argtypes = [vp, vp, vp]; self.L.bb_set_rng_state.argtypes = [vp, vp, vp]
            self.L.bb_get_contacts.argtypes = [vp, C.c_int32, vp, vp, vp]
        cfg = _lib.Config(); self.L.bb_default_config(C.byref(cfg))
        for k, v in cfg_kw.items():
            if k in ("target_direction", "goal_position"):
                getattr(cfg, k)[0], getattr(cfg, k)[1] = v
            else:
                setattr(cfg, k, v)
        self.cfg, self.N = cfg, cfg.num_envs
        self.h = C.c_void_p()
        rc = self.L.bb_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            raise RuntimeError(f"bb_create({kind}) failed ({rc}): {self.L.bb_last_error(None).decode()}")
        N, npix = self.N, cfg.im_h * cfg.im_w
        self.buf, self.io = {}, _lib.IO()
        for name, dt, w in IO_FIELDS:
            if w is None:
                if not cfg.cameras:
                    continue
                shape = (N, 1, cfg.im_h, cfg.im_w)
            else:
                shape = (N, w) if w else (N,)
            self.buf[name] = self._alloc(shape, dt)
            setattr(self.io, name, self._ptr(self.buf[name]))

    # ---- storage: torch CUDA tensors for the CUDA library, numpy arrays for cpu_ref
    def _alloc(self, shape, dt):
        if self.kind == "cuda":
            t = {np.float32: self.torch.float32, np.uint8: self.torch.uint8, np.int32: self.torch.int32, np.float64: self.torch.float64, np.uint64: self.torch.int64}[dt]
            return self.torch.zeros(shape, dtype=t, device=f"cuda:{self.cfg.device}")
        return np.zeros(shape, dt)

    def _up(self, a, dt):
        a = np.ascontiguousarray(a, dt)
        if self.kind == "cuda":
            return self.torch.from_numpy(a.view(np.int64) if dt == np.uint64 else a).to(f"cuda:{self.cfg.device}")
        return a

    def _ptr(self, x):
        if x is None:
            return None
        return C.c_void_p(x.data_ptr()) if self.kind == "cuda" else C.c_void_p(x.ctypes.data)

    def _np(self, x):
        return x.cpu().numpy() if self.kind == "cuda" else np.array(x)

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.L.bb_last_error(self.h).decode()}")

    def _sync(self):
        if self.kind == "cuda":
            self.torch.cuda.synchronize()

    # ---- the ABI
    def out(self):
        self._sync()
        return {k: self._np(v) for k, v in self.buf.items()}

    def reset(self, mask=None, seeds=None):
        m = self._up(mask, np.uint8) if mask is not None else None
        s = self._up(seeds, np.int32) if seeds is not None else None
        self._check(self.L.bb_reset(self.h, self._ptr(m), self._ptr(s), C.byref(self.io), None), "bb_reset")
        return self.out()

    def step(self, actions):
        a = self._up(actions, np.float32)
        self._check(self.L.bb_step(self.h, self._ptr(a), C.byref(self.io), None), "bb_step")
        return self.out()

    def set_rng_state(self, state):
        self._check(self.L.bb_set_rng_state(self.h, self._ptr(self._up(state, np.uint64)), None), "bb_set_rng_state"); self._sync()

    def get_state(self):
        q, v, w = self._alloc((self.N, 17), np.float64), self._alloc((self.N, 15), np.float64), self._alloc((self.N, 15), np.float64)
        self._check(self.L.bb_get_state(self.h, self._ptr(q), self._ptr(v), self._ptr(w), None), "bb_get_state"); self._sync()
        return self._np(q), self._np(v), self._np(w)

    def set_state(self, qpos, qvel, warm):
        q, v, w = self._up(qpos, np.float64), self._up(qvel, np.float64), self._up(warm, np.float64)
        self._check(self.L.bb_set_state(self.h, self._ptr(q), self._ptr(v), self._ptr(w), None), "bb_set_state"); self._sync()

    def set_hfield(self, ids, hf):
        i, f = self._up(ids, np.int32), self._up(hf, np.float32)
        self._check(self.L.bb_set_hfield(self.h, self._ptr(i), len(ids), self._ptr(f), None), "bb_set_hfield"); self._sync()

    def get_hfield(self, env):
        f = self._alloc((293 * 293,), np.float32)
        self._check(self.L.bb_get_hfield(self.h, env, self._ptr(f), None), "bb_get_hfield"); self._sync()
        return self._np(f)

    def terrain_seeds(self):
        s = self._alloc((self.N,), np.int32)
        self._check(self.L.bb_get_terrain_seeds(self.h, self._ptr(s), None), "bb_get_terrain_seeds"); self._sync()
        return self._np(s)

    def contacts(self, env):
        rows, n = self._alloc((_lib.PROBE_MAXCON, _lib.CONTACT_STRIDE), np.float64), self._alloc((1,), np.int32)
        self._check(self.L.bb_get_contacts(self.h, env, self._ptr(rows), self._ptr(n), None), "bb_get_contacts"); self._sync()
        return self._np(rows)[:int(self._np(n)[0])]

    def close(self):
        if self.h:
            self._sync(); self.L.bb_destroy(self.h); self.h = None


def pcg64_states(seeds):
    st = np.zeros((len(seeds), 5), np.uint64); m64 = (1 << 64) - 1
    for i, sd in enumerate(seeds):
        s = np.random.PCG64(int(sd)).state
        v, inc = s["state"]["state"], s["state"]["inc"]
        st[i] = (v >> 64, v & m64, inc >> 64, inc & m64, int(s["has_uint32"]) | (int(s["uinteger"]) << 32))
    return st
