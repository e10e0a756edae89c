"""ctypes binding of tests/hostcore/libbb_hostcore.so (engine physics core compiled for the CPU; TEST TOOL ONLY)."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbb_hostcore.so")
_SRC = [os.path.join(_HERE, "hostcore.cpp"),
        os.path.join(_HERE, "..", "..", "openballbot_rl_b200", "csrc", "bb_core.cuh"),
        os.path.join(_HERE, "..", "..", "openballbot_rl_b200", "csrc", "bb_model.h")]
NQ, NV, NC = 17, 15, 64


def build(force=False):
    stale = not os.path.exists(_SO) or any(os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_SO) for s in _SRC)
    if force or stale:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-o", _SO, _SRC[0]])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
    return _lib


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def forward(qpos, qvel, ctrl, warm, hfield=None, zscale=2.0, prec=64):
    qpos = np.ascontiguousarray(qpos, np.float64); qvel = np.ascontiguousarray(qvel, np.float64)
    ctrl = np.ascontiguousarray(ctrl, np.float64); warm = np.ascontiguousarray(warm, np.float64)
    hf = np.zeros(293 * 293, np.float32) if hfield is None else np.ascontiguousarray(hfield, np.float32).ravel()
    qacc = np.zeros(NV); M = np.zeros((NV, NV)); qfs = np.zeros(NV); kin = np.zeros(13)
    nc = C.c_int(); ni = C.c_int(); cd = np.zeros(NC); cp = np.zeros((NC, 3)); cf = np.zeros((NC, 9)); ct = np.zeros(NC, np.int32)
    n = lib().hc_forward(prec, _p(qpos), _p(qvel), _p(ctrl), _p(warm), _p(hf, C.c_float), C.c_double(zscale), _p(qacc), _p(M), _p(qfs),
                         _p(kin), C.byref(nc), C.byref(ni), _p(cd), _p(cp), _p(cf), _p(ct, C.c_int))
    return dict(qacc=qacc, qM=M, qfs=qfs, kin=kin, ncon=n, niter=ni.value, dist=cd[:n], pos=cp[:n], frame=cf[:n].reshape(n, 3, 3), type=ct[:n])


def step(qpos, qvel, warm, ctrl, hfield=None, zscale=2.0, prec=64):
    qpos = np.array(qpos, np.float64); qvel = np.array(qvel, np.float64); warm = np.array(warm, np.float64)
    ctrl = np.ascontiguousarray(ctrl, np.float64)
    hf = np.zeros(293 * 293, np.float32) if hfield is None else np.ascontiguousarray(hfield, np.float32).ravel()
    kin = np.zeros(13); nc = C.c_int(); ni = C.c_int()
    lib().hc_step(prec, _p(qpos), _p(qvel), _p(warm), _p(ctrl), _p(hf, C.c_float), C.c_double(zscale), _p(kin), C.byref(nc), C.byref(ni))
    return qpos, qvel, warm, kin, nc.value, ni.value


def model():
    dA = np.zeros(12); mi = C.c_double(); ms = np.zeros(3); c0 = np.zeros(3)
    lib().hc_model(_p(dA), C.byref(mi), _p(ms), _p(c0))
    return dict(dA=dA, meaninertia=mi.value, masses=ms, c0=c0)


def set_solver(fast):
    """0: reference iteration path (exact line search, every RK stage warm-started from qacc_warmstart); 1: the engine's
    solver_mode 1 (strong-Wolfe line search with cone-apex candidates, warm start chained through the stages)."""
    lib().hc_set_solver(int(bool(fast)))
