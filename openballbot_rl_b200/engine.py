"""BallbotEngine: thin Python handle over the C-ABI CUDA engine, all I/O as torch CUDA tensors (no copies).

Replaces the N worker processes each holding a patched-MuJoCo ``MjModel/MjData`` in the reference
(ballbot_rl/training/train.py:82-97 -> ballbot_gym/envs/ballbot_env.py:261-262).

The hot-path calls (``step`` / ``reset`` / ``add_reward``) go through the PyTorch extension ``torch.ops.ballbot.*``
(csrc/bb_torch_ops.cpp: dispatcher ops over the C ABI, current-stream and device-guard handling inside the op) when that
library is available, and through ctypes on the very same C entry points otherwise (``binding="ctypes"`` forces it; non-torch
hosts bind the C ABI directly).  Either way the work is done by the CUDA kernels of libballbot_b200.so.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import EngineError, HFIELD_N, NQ, NV

TERRAIN_FLAT, TERRAIN_PERLIN, TERRAIN_EXTERNAL, TERRAIN_SHARED, TERRAIN_TABLE = 0, 1, 2, 3, 4
REWARD_DIRECTIONAL, REWARD_DISTANCE, REWARD_EXTERNAL = 0, 1, 2
OBS_KEYS = ("orientation", "angular_vel", "vel", "motor_state", "actions")


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class BallbotEngine:
    def __init__(self, num_envs, device=0, precision=64, terrain="perlin", terrain_seed=None, perlin=None, hfield_zscale=2.0,
                 cameras=True, im_h=64, im_w=64, camera_frame_rate=90.0, max_ep_steps=4000, max_allowed_tilt=20.0,
                 max_wheel_velocity=10.0, reward="directional", reward_scale=0.01, action_reg_coef=-0.0001,
                 survival_bonus=0.02, target_direction=(0.0, 1.0), goal_position=(0.0, 0.0), distance_scale=1.0, seed=0,
                 auto_reset=True, env_offset=0, step_kernel="warp", solver="exact", perlin_table=None, seed_stream="counter",
                 binding="auto", depth_kernel="raycast"):
        if not torch.cuda.is_available():
            raise EngineError("BallbotEngine needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
        L = _lib.lib()
        cfg = _lib.default_config()
        cfg.num_envs = int(num_envs); cfg.device = int(device); cfg.precision = int(precision); cfg.env_offset = int(env_offset)
        cfg.terrain_type = {"flat": TERRAIN_FLAT, "perlin": TERRAIN_PERLIN, "external": TERRAIN_EXTERNAL, "shared": TERRAIN_SHARED, "table": TERRAIN_TABLE}[terrain]
        cfg.terrain_seed = -1 if terrain_seed is None else int(terrain_seed)
        perlin = perlin or {}
        cfg.perlin_scale = float(perlin.get("scale", 25.0)); cfg.perlin_octaves = int(perlin.get("octaves", 4))
        cfg.perlin_persistence = float(perlin.get("persistence", 0.2)); cfg.perlin_lacunarity = float(perlin.get("lacunarity", 2.0))
        cfg.perlin_amplitude = float(perlin.get("amplitude", 1.0))
        cfg.hfield_zscale = float(hfield_zscale)
        cfg.cameras = int(bool(cameras)); cfg.im_h = int(im_h); cfg.im_w = int(im_w); cfg.camera_frame_rate = float(camera_frame_rate)
        cfg.max_ep_steps = int(max_ep_steps); cfg.max_allowed_tilt = float(max_allowed_tilt); cfg.max_wheel_velocity = float(max_wheel_velocity)
        cfg.reward_type = {"directional": REWARD_DIRECTIONAL, "distance": REWARD_DISTANCE, "external": REWARD_EXTERNAL}[reward]
        cfg.reward_scale = float(reward_scale); cfg.action_reg_coef = float(action_reg_coef); cfg.survival_bonus = float(survival_bonus)
        cfg.target_direction[0], cfg.target_direction[1] = float(target_direction[0]), float(target_direction[1])
        cfg.goal_position[0], cfg.goal_position[1] = float(goal_position[0]), float(goal_position[1])
        cfg.distance_scale = float(distance_scale); cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF; cfg.auto_reset = int(bool(auto_reset))
        cfg.step_kernel = {"warp": 0, "split": 0, "thread": 1, "fused": 2}[step_kernel]
        cfg.solver_mode = {"exact": 0, "fast": 1}[solver]
        cfg.perlin_table = -1 if perlin_table is None else int(bool(perlin_table))
        cfg.seed_stream = {"counter": 0, "pcg64": 1}[seed_stream]
        cfg.depth_kernel = {"raster": 0, "raycast": 1}[depth_kernel]
        self.cfg = cfg
        self._L = L
        self._h = C.c_void_p()
        rc = L.bb_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise EngineError(f"bb_create failed ({rc}): {L.bb_last_error(None).decode()}")
        self.num_envs = int(num_envs)
        self.device = torch.device("cuda", int(device))
        self.cameras = bool(cameras)
        self.im_h, self.im_w = int(im_h), int(im_w)
        self.precision = int(precision)
        N, dev = self.num_envs, self.device      # every tensor names its device explicitly: the caller's current device is irrelevant
        f32 = dict(dtype=torch.float32, device=dev)
        self.obs = {k: torch.zeros(N, 3, **f32) for k in OBS_KEYS}
        self.obs["relative_image_timestamp"] = torch.zeros(N, 1, **f32)
        if self.cameras:
            self.obs["rgbd_0"] = torch.ones(N, 1, im_h, im_w, **f32)
            self.obs["rgbd_1"] = torch.ones(N, 1, im_h, im_w, **f32)
        self.reward = torch.zeros(N, **f32)
        self.terminated = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.failure = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.pos2d = torch.zeros(N, 2, **f32)
        self.terminal_obs = torch.zeros(N, 16, **f32)
        self.episode_return = torch.zeros(N, **f32)
        self.episode_length = torch.zeros(N, dtype=torch.int32, device=dev)
        self.status = torch.zeros(N, dtype=torch.int32, device=dev)
        io = _lib.IO()
        io.orientation = self.obs["orientation"].data_ptr(); io.angular_vel = self.obs["angular_vel"].data_ptr()
        io.vel = self.obs["vel"].data_ptr(); io.motor_state = self.obs["motor_state"].data_ptr()
        io.actions = self.obs["actions"].data_ptr(); io.rel_image_ts = self.obs["relative_image_timestamp"].data_ptr()
        io.rgbd_0 = self.obs["rgbd_0"].data_ptr() if self.cameras else None
        io.rgbd_1 = self.obs["rgbd_1"].data_ptr() if self.cameras else None
        io.reward = self.reward.data_ptr(); io.terminated = self.terminated.data_ptr(); io.failure = self.failure.data_ptr()
        io.pos2d = self.pos2d.data_ptr(); io.terminal_obs = self.terminal_obs.data_ptr()
        io.episode_return = self.episode_return.data_ptr(); io.episode_length = self.episode_length.data_ptr()
        io.status = self.status.data_ptr()
        self._io = io
        empty = torch.empty(0, **f32)
        self._io_list = [self.obs["orientation"], self.obs["angular_vel"], self.obs["vel"], self.obs["motor_state"], self.obs["actions"],
                         self.obs["relative_image_timestamp"], self.obs.get("rgbd_0", empty), self.obs.get("rgbd_1", empty), self.reward,
                         self.terminated, self.failure, self.pos2d, self.terminal_obs, self.episode_return, self.episode_length, self.status]
        self._ops = None
        if binding not in ("auto", "torch", "ctypes"):
            raise ValueError("binding must be 'auto', 'torch' or 'ctypes'")
        if binding != "ctypes":
            try:
                self._ops = _lib.torch_ops()
            except EngineError:
                if binding == "torch":
                    raise
        self.binding = "torch" if self._ops is not None else "ctypes"

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_h", None):
            torch.cuda.synchronize(self.device)
            self._L.bb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise EngineError(f"{what} failed ({rc}): {self._L.bb_last_error(self._h).decode()}")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ hot path
    def reset(self, mask=None, seeds=None):
        """Reset the envs selected by ``mask`` (uint8/bool CUDA tensor [N]; None = all). ``seeds`` (int32 [N], optional) fixes the
        terrain seed ``r_seed`` of the selected envs (ballbot_env.py:505-510) instead of drawing it. Returns the obs dict."""
        if seeds is not None:
            seeds = torch.as_tensor(seeds, dtype=torch.int32, device=self.device).reshape(self.num_envs).contiguous()
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            keep = (mask == 0)
            self.reward.mul_(keep); self.terminated.mul_(keep.to(torch.uint8)); self.failure.mul_(keep.to(torch.uint8))
            self.pos2d.mul_(keep[:, None])
        else:
            self.reward.zero_(); self.terminated.zero_(); self.failure.zero_(); self.pos2d.zero_()
        if self._ops is not None:
            self._ops.reset(self._h.value, mask, seeds, self._io_list)
        else:
            self._check(self._L.bb_reset(self._h, _ptr(mask), _ptr(seeds), C.byref(self._io), self._stream()), "bb_reset")
        return self.obs

    def step(self, actions):
        """One env step for every env. ``actions``: float32 CUDA tensor [N,3]. Outputs are written into the
        persistent tensors ``obs`` / ``reward`` / ``terminated`` / ``failure`` / ``pos2d`` (stream-ordered, no sync)."""
        if actions.dtype != torch.float32 or not actions.is_cuda or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.float32).contiguous()
        if actions.shape != (self.num_envs, 3):
            raise ValueError(f"actions must have shape ({self.num_envs}, 3), got {tuple(actions.shape)}")
        if self._ops is not None:
            self._ops.step(self._h.value, actions, self._io_list)
        else:
            self._check(self._L.bb_step(self._h, _ptr(actions), C.byref(self._io), self._stream()), "bb_step")
        return self.obs, self.reward, self.terminated, self.failure

    def add_reward(self, term):
        term = term.to(device=self.device, dtype=torch.float32).contiguous()
        if self._ops is not None:
            self._ops.add_reward(self._h.value, term, self._io_list)
        else:
            self._check(self._L.bb_add_reward(self._h, _ptr(term), C.byref(self._io), self._stream()), "bb_add_reward")

    # ------------------------------------------------------------------ state / terrain access (parity tests, plugins)
    def set_state(self, qpos=None, qvel=None, warm=None):
        def prep(x, n):
            if x is None:
                return None
            x = torch.as_tensor(x, dtype=torch.float64, device=self.device).reshape(self.num_envs, n).contiguous()
            return x
        qpos, qvel, warm = prep(qpos, NQ), prep(qvel, NV), prep(warm, NV)
        self._check(self._L.bb_set_state(self._h, _ptr(qpos), _ptr(qvel), _ptr(warm), self._stream()), "bb_set_state")
        torch.cuda.current_stream(self.device).synchronize()

    def get_state(self):
        N, dev = self.num_envs, self.device
        qpos = torch.empty(N, NQ, dtype=torch.float64, device=dev); qvel = torch.empty(N, NV, dtype=torch.float64, device=dev)
        warm = torch.empty(N, NV, dtype=torch.float64, device=dev)
        self._check(self._L.bb_get_state(self._h, _ptr(qpos), _ptr(qvel), _ptr(warm), self._stream()), "bb_get_state")
        return qpos, qvel, warm

    def set_hfield(self, env_ids, hfield):
        ids = torch.as_tensor(env_ids, dtype=torch.int32, device=self.device).contiguous()
        hf = torch.as_tensor(hfield, dtype=torch.float32, device=self.device).reshape(ids.numel(), HFIELD_N * HFIELD_N).contiguous()
        self._check(self._L.bb_set_hfield(self._h, _ptr(ids), int(ids.numel()), _ptr(hf), self._stream()), "bb_set_hfield")
        torch.cuda.current_stream(self.device).synchronize()

    def get_hfield(self, env):
        out = torch.empty(HFIELD_N * HFIELD_N, dtype=torch.float32, device=self.device)
        self._check(self._L.bb_get_hfield(self._h, int(env), _ptr(out), self._stream()), "bb_get_hfield")
        return out

    def terrain_seeds(self):
        out = torch.empty(self.num_envs, dtype=torch.int32, device=self.device)
        self._check(self._L.bb_get_terrain_seeds(self._h, _ptr(out), self._stream()), "bb_get_terrain_seeds")
        return out

    def perlin_terrain(self, seeds):
        seeds = torch.as_tensor(seeds, dtype=torch.int32, device=self.device).contiguous()
        out = torch.empty(seeds.numel(), HFIELD_N * HFIELD_N, dtype=torch.float32, device=self.device)
        self._check(self._L.bb_perlin_terrain(self._h, _ptr(seeds), int(seeds.numel()), _ptr(out), self._stream()), "bb_perlin_terrain")
        return out

    def render_depth(self):
        N = self.num_envs
        a = torch.empty(N, 1, self.im_h, self.im_w, dtype=torch.float32, device=self.device); b = torch.empty_like(a)
        self._check(self._L.bb_render_depth(self._h, _ptr(a), _ptr(b), self._stream()), "bb_render_depth")
        return a, b

    def seed_pcg64(self, seeds):
        """numpy-compatible terrain-seed streams (``seed_stream="pcg64"``): env i continues ``np.random.default_rng(seeds[i])``
        exactly like the reference's per-env ``self._np_random`` (ballbot_env.py:378-384, 505-507)."""
        seeds = np.atleast_1d(np.asarray(seeds)).astype(np.int64)
        if seeds.shape != (self.num_envs,):
            raise ValueError(f"seeds must have shape ({self.num_envs},)")
        st = np.zeros((self.num_envs, 5), np.uint64)
        m64 = (1 << 64) - 1
        for i, sd in enumerate(seeds):
            s = np.random.PCG64(int(sd)).state
            v, inc = s["state"]["state"], s["state"]["inc"]
            st[i] = (v >> 64, v & m64, inc >> 64, inc & m64, int(s["has_uint32"]) | (int(s["uinteger"]) << 32))
        t = torch.from_numpy(st.view(np.int64)).to(self.device)
        self._check(self._L.bb_set_rng_state(self._h, _ptr(t), self._stream()), "bb_set_rng_state")
        torch.cuda.current_stream(self.device).synchronize()

    @staticmethod
    def _contact_rows(rows, n):
        r = rows.cpu().numpy()[:n]
        return dict(type=r[:, 0].astype(np.int32), dist=r[:, 1], pos=r[:, 2:5], frame=r[:, 5:14].reshape(n, 3, 3))

    def probe_forward(self, env, ctrl=(0.0, 0.0, 0.0)):
        """One forward-dynamics evaluation of env ``env`` at its current state through the production device code
        (solver / contact parity probe): qacc, qacc_smooth, qfrc_smooth, iteration count and the contact set."""
        dev = self.device
        c = torch.tensor(ctrl, dtype=torch.float64, device=dev)
        out = torch.zeros(64, dtype=torch.float64, device=dev)
        rows = torch.zeros(_lib.PROBE_MAXCON, _lib.CONTACT_STRIDE, dtype=torch.float64, device=dev)
        self._check(self._L.bb_probe_forward(self._h, int(env), _ptr(c), _ptr(out), _ptr(rows), self._stream()), "bb_probe_forward")
        o = out.cpu().numpy(); n = int(o[45])
        return dict(qacc=o[:15], qacc_smooth=o[15:30], qfrc_smooth=o[30:45], ncon=n, niter=int(o[46]), consts=o[47:54], **self._contact_rows(rows, n))

    def get_contacts(self, env):
        """mjData.contact of env ``env`` at its current state: dict(type, dist, pos, frame)."""
        rows = torch.zeros(_lib.PROBE_MAXCON, _lib.CONTACT_STRIDE, dtype=torch.float64, device=self.device)
        n = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._check(self._L.bb_get_contacts(self._h, int(env), _ptr(rows), _ptr(n), self._stream()), "bb_get_contacts")
        return self._contact_rows(rows, int(n.item()))

    def profile_begin(self, max_steps):
        self._check(self._L.bb_profile_begin(self._h, int(max_steps)), "bb_profile_begin")

    def profile_end(self):
        ms = (C.c_double * 4)(); n = C.c_int32()
        self._check(self._L.bb_profile_end(self._h, ms, C.byref(n)), "bb_profile_end")
        return dict(step_ms=ms[0], terrain_ms=ms[1], reset_ms=ms[2], depth_ms=ms[3], steps=n.value)

    @property
    def launch_count(self):
        return int(self._L.bb_launch_count(self._h))

    # ------------------------------------------------------------------ host-buffer path (numpy in / numpy out)
    def _host_io(self, images):
        N = self.num_envs
        if not hasattr(self, "_hbuf"):
            # numpy views of the engine's page-locked staging buffers: results land in place (no second host copy)
            h = _lib.HostIO(); act = C.c_void_p()
            self._check(self._L.bb_host_buffers(self._h, C.byref(act), C.byref(h)), "bb_host_buffers")

            def view(ptr, ctype, shape):
                return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=shape)
            self._hact = view(act, C.c_float, (N, 3))
            self._hbuf = dict(obs16=view(h.obs16, C.c_float, (N, 16)), reward=view(h.reward, C.c_float, (N,)),
                              terminated=view(h.terminated, C.c_uint8, (N,)), failure=view(h.failure, C.c_uint8, (N,)),
                              pos2d=view(h.pos2d, C.c_float, (N, 2)), terminal_obs=view(h.terminal_obs, C.c_float, (N, 16)),
                              episode_return=view(h.episode_return, C.c_float, (N,)), episode_length=view(h.episode_length, C.c_int32, (N,)))
            if self.cameras:
                self._hbuf["img_0"] = np.ones((N, 1, self.im_h, self.im_w), np.float32)
                self._hbuf["img_1"] = np.ones((N, 1, self.im_h, self.im_w), np.float32)
        h = _lib.HostIO()
        for k, v in self._hbuf.items():
            if k.startswith("img") and not images:
                continue
            setattr(h, k, v.ctypes.data)
        return h

    def step_host(self, actions, images=True):
        a = np.ascontiguousarray(actions, np.float32)
        if a.shape != (self.num_envs, 3):
            raise ValueError(f"actions must have shape ({self.num_envs}, 3)")
        h = self._host_io(images and self.cameras)
        self._check(self._L.bb_step_host(self._h, C.c_void_p(a.ctypes.data), C.byref(h)), "bb_step_host")
        return self._hbuf

    def reset_host(self, mask=None, images=True):
        m = None
        if mask is not None:
            mk = np.ascontiguousarray(mask, np.uint8); m = C.c_void_p(mk.ctypes.data)
        h = self._host_io(images and self.cameras)
        self._check(self._L.bb_reset_host(self._h, m, C.byref(h)), "bb_reset_host")
        return self._hbuf


def fp64_peak_tflops(device=0):
    """Measured fp64 FMA throughput of the device (bb_fp64_peak), TFLOP/s."""
    v = C.c_double()
    rc = _lib.lib().bb_fp64_peak(int(device), C.byref(v))
    if rc != 0:
        raise EngineError(f"bb_fp64_peak failed ({rc})")
    return v.value


def build_info():
    return _lib.lib().bb_build_info().decode()


def model_constants():
    dA = (C.c_double * 12)(); mi = C.c_double(); ms = (C.c_double * 3)()
    _lib.lib().bb_model_constants(dA, C.byref(mi), ms)
    return dict(dA=np.array(dA[:]), meaninertia=mi.value, masses=np.array(ms[:]))
