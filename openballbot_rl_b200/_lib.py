"""ctypes loader for the C-ABI engine library (include/ballbot_b200.h).

The product path is CUDA only: if ``libballbot_b200.so`` is missing or cannot be loaded this module raises -- there is
no CPU or PyTorch fallback.  ``build()`` compiles the library in-tree with nvcc for sm_100a.
"""
import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
LIB_PATH = os.path.join(_PKG, "libballbot_b200.so")
_SOURCES = ["bb_engine.cu", "bb_rollout.cu", "bb_core.cuh", "bb_group.cuh", "bb_model.h"]
_UNITS = ["bb_engine.cu", "bb_rollout.cu"]   # translation units of the library
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]

HFIELD_N = 293
NQ, NV = 17, 15
PROBE_MAXCON, CONTACT_STRIDE = 80, 14      # BB_PROBE_MAXCON, BB_CONTACT_STRIDE


class EngineError(RuntimeError):
    pass


class Config(C.Structure):
    """bb_config (include/ballbot_b200.h)."""
    _fields_ = [
        ("abi_version", C.c_int32), ("num_envs", C.c_int32), ("env_offset", C.c_int64), ("device", C.c_int32),
        ("precision", C.c_int32), ("terrain_type", C.c_int32), ("terrain_seed", C.c_int32),
        ("perlin_scale", C.c_float), ("perlin_persistence", C.c_float), ("perlin_lacunarity", C.c_float),
        ("perlin_amplitude", C.c_float), ("perlin_octaves", C.c_int32), ("hfield_zscale", C.c_float),
        ("cameras", C.c_int32), ("im_h", C.c_int32), ("im_w", C.c_int32), ("camera_frame_rate", C.c_float),
        ("max_ep_steps", C.c_int32), ("max_allowed_tilt", C.c_float), ("max_wheel_velocity", C.c_float),
        ("reward_type", C.c_int32), ("reward_scale", C.c_float), ("action_reg_coef", C.c_float),
        ("survival_bonus", C.c_float), ("target_direction", C.c_float * 2), ("goal_position", C.c_float * 2),
        ("distance_scale", C.c_float), ("seed", C.c_uint64), ("auto_reset", C.c_int32), ("step_kernel", C.c_int32), ("solver_mode", C.c_int32),
        ("perlin_table", C.c_int32), ("seed_stream", C.c_int32), ("depth_kernel", C.c_int32),
    ]


class IO(C.Structure):
    """bb_io: caller-owned device output pointers."""
    _fields_ = [(n, C.c_void_p) for n in (
        "orientation", "angular_vel", "vel", "motor_state", "actions", "rel_image_ts", "rgbd_0", "rgbd_1", "reward",
        "terminated", "failure", "pos2d", "terminal_obs", "episode_return", "episode_length", "status")]


class HostIO(C.Structure):
    """bb_host_io: host output pointers of the numpy path."""
    _fields_ = [(n, C.c_void_p) for n in (
        "obs16", "reward", "terminated", "failure", "pos2d", "terminal_obs", "episode_return", "episode_length",
        "img_0", "img_1")]


EXPORTED = ["bb_create", "bb_destroy", "bb_default_config", "bb_last_error", "bb_num_envs", "bb_reset", "bb_step",
            "bb_add_reward", "bb_set_state", "bb_get_state", "bb_set_hfield", "bb_get_hfield", "bb_get_terrain_seeds",
            "bb_perlin_terrain", "bb_render_depth", "bb_step_host", "bb_reset_host", "bb_launch_count",
            "bb_model_constants", "bb_probe_forward", "bb_profile_begin", "bb_profile_end", "bb_perlin_grid", "bb_gae", "bb_host_buffers",
            "bb_set_rng_state", "bb_get_contacts", "bb_build_info", "bb_fp64_peak", "bb_adamw_step", "bb_snoise2_grid"]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    srcs = [os.path.join(_CSRC, s) for s in _SOURCES] + [os.path.join(_PKG, "..", "include", "ballbot_b200.h")]
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in srcs)


def source_hash():
    """Short hash of every source of the library (what bb_build_info() reports: tells a stale .so from a fresh one)."""
    import hashlib
    h = hashlib.sha1()
    for f in [os.path.join(_CSRC, s) for s in _SOURCES] + [os.path.join(_PKG, "..", "include", "ballbot_b200.h")]:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:12]


class _BuildLock:
    """Inter-process lock around an in-tree build: the ranks of a torchrun job (one process per GPU) may all find the library
    stale at import time; one of them compiles (into a temporary file, renamed atomically), the others wait and reuse it."""

    def __init__(self, target):
        self.path = target + ".lock"

    def __enter__(self):
        import fcntl
        self.f = open(self.path, "w")
        fcntl.flock(self.f, fcntl.LOCK_EX)
        return self

    def __exit__(self, *exc):
        import fcntl
        fcntl.flock(self.f, fcntl.LOCK_UN)
        self.f.close()


def build(force=False, verbose=False):
    """Compile csrc/bb_engine.cu for sm_100a into libballbot_b200.so (nvcc cross-compiles without a GPU)."""
    stale = needs_build()
    if not force and not stale:
        return LIB_PATH
    with _BuildLock(LIB_PATH):
        if stale and not needs_build():          # another process built it while this one waited for the lock
            return LIB_PATH
        tmp = f"{LIB_PATH}.{os.getpid()}.tmp"
        cmd = ["nvcc"] + NVCC_FLAGS + [f'-DBB_SRC_HASH="{source_hash()}"'] + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + [os.path.join(_CSRC, u) for u in _UNITS]
        try:
            subprocess.check_call(cmd)
            os.replace(tmp, LIB_PATH)
        finally:
            if os.path.exists(tmp):
                os.remove(tmp)
    return LIB_PATH


TORCH_OPS_PATH = os.path.join(_PKG, "libballbot_torch_ops.so")
_TORCH_OPS_SRC = os.path.join(_CSRC, "bb_torch_ops.cpp")


def build_torch_ops(force=False):
    """Compile csrc/bb_torch_ops.cpp (host code: TORCH_LIBRARY(ballbot) over the C ABI) in-tree against the installed torch."""
    srcs = [_TORCH_OPS_SRC, os.path.join(_PKG, "..", "include", "ballbot_b200.h")]
    if not force and os.path.exists(TORCH_OPS_PATH) and all(os.path.getmtime(s) <= os.path.getmtime(TORCH_OPS_PATH) for s in srcs):
        return TORCH_OPS_PATH
    build()
    import torch
    from torch.utils import cpp_extension as X
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    with _BuildLock(TORCH_OPS_PATH):
        if not force and os.path.exists(TORCH_OPS_PATH) and all(os.path.getmtime(s) <= os.path.getmtime(TORCH_OPS_PATH) for s in srcs):
            return TORCH_OPS_PATH
        return _compile_torch_ops(torch, X, cuda_home, tlib)


def _compile_torch_ops(torch, X, cuda_home, tlib):
    tmp = f"{TORCH_OPS_PATH}.{os.getpid()}.tmp"
    cmd = (["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", tmp, _TORCH_OPS_SRC,
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch.compiled_with_cxx11_abi())}"]
           + [f"-I{d}" for d in X.include_paths()] + [f"-I{cuda_home}/include"]
           + [f"-L{tlib}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", f"-L{_PKG}", "-lballbot_b200",
              "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tlib}"])
    try:
        subprocess.check_call(cmd)
        os.replace(tmp, TORCH_OPS_PATH)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    return TORCH_OPS_PATH


_torch_ops = None


def torch_ops():
    """torch.ops.ballbot (loads libballbot_torch_ops.so, building it on first use); raises EngineError when it cannot be had."""
    global _torch_ops
    if _torch_ops is None:
        import torch
        try:
            lib()                                   # libballbot_b200.so first: the ops library links against it
            torch.ops.load_library(build_torch_ops())
        except Exception as exc:
            raise EngineError(f"torch.ops.ballbot is unavailable ({exc}); the ctypes binding of the same C ABI still works") from exc
        _torch_ops = torch.ops.ballbot
    return _torch_ops


_lib = None


def lib():
    """Load libballbot_b200.so; raises EngineError when the CUDA library is absent (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if needs_build():
        try:                                   # a fresh checkout or edited sources: (re)compile if the toolchain is here
            build(force=True)
        except Exception as exc:
            if not os.path.exists(LIB_PATH):   # ... else fail loudly; a stale library is reported by bb_build_info() (source hash)
                raise EngineError(f"{LIB_PATH} not found and could not be built ({exc}): run `python -c 'import __graft_entry__ "
                                  "as g; g.build()'` (nvcc, sm_100a). The ballbot engine has no CPU fallback.") from exc
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.bb_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.bb_destroy.argtypes = [vp]
    L.bb_default_config.argtypes = [C.POINTER(Config)]
    L.bb_default_config.restype = None
    L.bb_last_error.argtypes = [vp]
    L.bb_last_error.restype = C.c_char_p
    L.bb_num_envs.argtypes = [vp]
    L.bb_reset.argtypes = [vp, vp, vp, C.POINTER(IO), vp]
    L.bb_set_rng_state.argtypes = [vp, vp, vp]
    L.bb_get_contacts.argtypes = [vp, C.c_int32, vp, vp, vp]
    L.bb_snoise2_grid.argtypes = [C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_float, C.c_float, C.c_int32, vp]
    L.bb_adamw_step.argtypes = [vp, vp, vp, vp, C.c_int32] + [C.c_float] * 7 + [vp, vp, vp]
    L.bb_fp64_peak.argtypes = [C.c_int32, C.POINTER(C.c_double)]
    L.bb_build_info.argtypes = []
    L.bb_build_info.restype = C.c_char_p
    L.bb_step.argtypes = [vp, vp, C.POINTER(IO), vp]
    L.bb_add_reward.argtypes = [vp, vp, C.POINTER(IO), vp]
    L.bb_set_state.argtypes = [vp, vp, vp, vp, vp]
    L.bb_get_state.argtypes = [vp, vp, vp, vp, vp]
    L.bb_set_hfield.argtypes = [vp, vp, C.c_int32, vp, vp]
    L.bb_get_hfield.argtypes = [vp, C.c_int32, vp, vp]
    L.bb_get_terrain_seeds.argtypes = [vp, vp, vp]
    L.bb_perlin_terrain.argtypes = [vp, vp, C.c_int32, vp, vp]
    L.bb_render_depth.argtypes = [vp, vp, vp, vp]
    L.bb_host_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(HostIO)]
    L.bb_step_host.argtypes = [vp, vp, C.POINTER(HostIO)]
    L.bb_reset_host.argtypes = [vp, vp, C.POINTER(HostIO)]
    L.bb_launch_count.argtypes = [vp]
    L.bb_launch_count.restype = C.c_int64
    L.bb_perlin_grid.argtypes = [C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_float, C.c_float, C.c_float, vp, C.c_int32, vp]
    L.bb_gae.argtypes = [vp, vp, vp, C.c_int32, C.c_int32, C.c_float, C.c_float, vp, vp, vp]
    L.bb_profile_begin.argtypes = [vp, C.c_int32]
    L.bb_profile_end.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int32)]
    L.bb_probe_forward.argtypes = [vp, C.c_int32, vp, vp, vp, vp]
    L.bb_model_constants.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    _lib = L
    return L


def default_config():
    cfg = Config()
    lib().bb_default_config(C.byref(cfg))
    return cfg
