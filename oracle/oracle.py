"""ctypes binding of the fp64 CPU ORACLE (oracle/ballbot_oracle.cpp).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package. PARITY UNPINNED (see
oracle/ballbot_oracle.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libballbot_oracle.so")
NQ, NV, HN = 17, 15, 293


class Config(C.Structure):
    _fields_ = [("max_ep_steps", C.c_int), ("max_allowed_tilt", C.c_double), ("max_wheel_velocity", C.c_double),
                ("camera_frame_rate", C.c_double), ("reward_scale", C.c_double), ("action_reg_coef", C.c_double),
                ("survival_bonus", C.c_double), ("target_dir", C.c_double * 2), ("hfield_zscale", C.c_double),
                ("cameras", C.c_int), ("im_h", C.c_int), ("im_w", C.c_int),
                ("reward_type", C.c_int), ("goal", C.c_double * 2), ("distance_scale", C.c_double)]


def build(force=False):
    """make -C oracle: libballbot_oracle.so (bbo_* API) and libballbot_cpu_ref.so (the oracle behind the engine's bb_* ABI)."""
    if force:
        subprocess.check_call(["make", "-C", _HERE, "-s", "clean"])
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


CPU_REF_SO = os.path.join(_HERE, "_build", "libballbot_cpu_ref.so")


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        dp, fp, ip, u8p = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_uint8)
        L.bbo_create.restype = C.c_void_p
        L.bbo_create.argtypes = [C.POINTER(Config)]
        L.bbo_destroy.argtypes = [C.c_void_p]
        L.bbo_reset.argtypes = [C.c_void_p, fp, fp]
        L.bbo_step.argtypes = [C.c_void_p, fp, fp, fp, u8p, u8p, fp]
        L.bbo_get_depth.argtypes = [C.c_void_p, fp, fp]
        L.bbo_get_state.argtypes = [C.c_void_p, dp, dp, dp, dp]
        L.bbo_set_state.argtypes = [C.c_void_p, dp, dp, dp, C.c_double]
        L.bbo_set_hfield.argtypes = [C.c_void_p, fp]
        L.bbo_mj_step.argtypes = [C.c_void_p, dp]
        L.bbo_forward.argtypes = [C.c_void_p, dp, dp, dp, dp, dp, ip, ip]
        L.bbo_get_contacts.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, ip]
        L.bbo_get_efc.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, dp]
        L.bbo_get_model.argtypes = [dp, dp, dp, dp, dp]
        L.bbo_get_kin.argtypes = [C.c_void_p, dp, dp, dp]
        L.bbo_snoise2_tiled.restype = C.c_float
        L.bbo_snoise2_tiled.argtypes = [C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int]
        L.bbo_snoise2.restype = C.c_float
        L.bbo_snoise2.argtypes = [C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_int]
        L.bbo_perlin_terrain.argtypes = [C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, fp]
        L.bbo_spawn_offset.restype = C.c_double
        L.bbo_spawn_offset.argtypes = [fp, C.c_double]
        L.bbo_render_depth.argtypes = [C.c_void_p, C.c_int, fp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def perlin_terrain(n=HN, scale=25.0, octaves=4, persistence=0.2, lacunarity=2.0, amplitude=1.0, seed=0):
    out = np.empty(n * n, np.float32)
    lib().bbo_perlin_terrain(n, scale, octaves, persistence, lacunarity, amplitude, int(seed), _fp(out))
    return out


def spawn_offset(hfield, zscale=2.0):
    h = np.ascontiguousarray(hfield, np.float32)
    return lib().bbo_spawn_offset(_fp(h), zscale)


def model_constants():
    mass = np.zeros(8); ipos = np.zeros((8, 3)); inertia = np.zeros((8, 9)); invw = np.zeros((8, 2)); mi = C.c_double()
    lib().bbo_get_model(_dp(mass), _dp(ipos), _dp(inertia), _dp(invw), C.byref(mi))
    return dict(mass=mass, ipos=ipos, inertia=inertia.reshape(8, 3, 3), invweight0=invw, meaninertia=mi.value)


class OracleEnv:
    """Single fp64 env with the reference env semantics (ballbot_env.py:567-1036)."""

    def __init__(self, cameras=False, im=64, **kw):
        cfg = Config()
        lib().bbo_default_config(C.byref(cfg))
        cfg.cameras = int(cameras); cfg.im_h = cfg.im_w = im
        for k, v in kw.items():
            if k in ("target_dir", "goal"):
                getattr(cfg, k)[0], getattr(cfg, k)[1] = v
            else:
                setattr(cfg, k, v)
        self.cfg = cfg
        self.h = lib().bbo_create(C.byref(cfg))
        self.im = im

    def close(self):
        if self.h:
            lib().bbo_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, hfield=None):
        obs = np.zeros(16, np.float32)
        hp = None
        if hfield is not None:
            hf = np.ascontiguousarray(hfield, np.float32).ravel(); assert hf.size == HN * HN
            hp = _fp(hf)
        lib().bbo_reset(self.h, hp, _fp(obs))
        return obs

    def step(self, action):
        a = np.ascontiguousarray(action, np.float32)
        obs = np.zeros(16, np.float32); r = C.c_float(); t = C.c_uint8(); f = C.c_uint8(); info = np.zeros(4, np.float32)
        lib().bbo_step(self.h, _fp(a), _fp(obs), C.byref(r), C.byref(t), C.byref(f), _fp(info))
        return obs, r.value, bool(t.value), bool(f.value), info

    def depth(self):
        a = np.zeros((self.im, self.im), np.float32); b = np.zeros((self.im, self.im), np.float32)
        lib().bbo_get_depth(self.h, _fp(a), _fp(b))
        return a, b

    def render_depth(self, cam):
        a = np.zeros((self.im, self.im), np.float32)
        lib().bbo_render_depth(self.h, cam, _fp(a))
        return a

    def get_state(self):
        qpos = np.zeros(NQ); qvel = np.zeros(NV); warm = np.zeros(NV); t = C.c_double()
        lib().bbo_get_state(self.h, _dp(qpos), _dp(qvel), _dp(warm), C.byref(t))
        return qpos, qvel, warm, t.value

    def set_state(self, qpos, qvel, warm=None, time=0.0):
        qpos = np.ascontiguousarray(qpos, np.float64); qvel = np.ascontiguousarray(qvel, np.float64)
        warm = np.zeros(NV) if warm is None else np.ascontiguousarray(warm, np.float64)
        lib().bbo_set_state(self.h, _dp(qpos), _dp(qvel), _dp(warm), float(time))

    def set_hfield(self, hfield):
        hf = np.ascontiguousarray(hfield, np.float32).ravel()
        lib().bbo_set_hfield(self.h, _fp(hf))

    def mj_step(self, ctrl):
        c = np.ascontiguousarray(ctrl, np.float64)
        lib().bbo_mj_step(self.h, _dp(c))

    def forward(self, ctrl=None):
        qM = np.zeros((NV, NV)); bias = np.zeros(NV); qas = np.zeros(NV); qacc = np.zeros(NV); nc = C.c_int(); ni = C.c_int()
        cp = None
        if ctrl is not None:
            c = np.ascontiguousarray(ctrl, np.float64); cp = _dp(c)
        lib().bbo_forward(self.h, cp, _dp(qM), _dp(bias), _dp(qas), _dp(qacc), C.byref(nc), C.byref(ni))
        return dict(qM=qM, qfrc_bias=bias, qacc_smooth=qas, qacc=qacc, ncon=nc.value, niter=ni.value)

    def contacts(self, maxcon=64):
        dist = np.zeros(maxcon); pos = np.zeros((maxcon, 3)); frame = np.zeros((maxcon, 9)); pair = np.zeros(maxcon, np.int32)
        n = lib().bbo_get_contacts(self.h, maxcon, _dp(dist), _dp(pos), _dp(frame), pair.ctypes.data_as(C.POINTER(C.c_int)))
        n = min(n, maxcon)
        return dict(n=n, dist=dist[:n], pos=pos[:n], frame=frame[:n].reshape(n, 3, 3), pair=pair[:n])

    def efc(self, maxefc=192):
        J = np.zeros((maxefc, NV)); aref = np.zeros(maxefc); D = np.zeros(maxefc); force = np.zeros(maxefc)
        n = lib().bbo_get_efc(self.h, maxefc, _dp(J), _dp(aref), _dp(D), _dp(force))
        return dict(n=n, J=J[:n], aref=aref[:n], D=D[:n], force=force[:n])

    def kin(self):
        xpos = np.zeros(3); xquat = np.zeros(4); cvel = np.zeros(6)
        lib().bbo_get_kin(self.h, _dp(xpos), _dp(xquat), _dp(cvel))
        return xpos, xquat, cvel
