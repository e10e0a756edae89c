"""BallbotVecEnv: the GPU-resident replacement of ``SubprocVecEnv([make_ballbot_env(...)] * N)``
(ballbot_rl/training/train.py:82-97) with the Stable-Baselines3 ``VecEnv`` protocol surface.

All N environments live in one CUDA engine (one warp per env); there are no worker processes and no pipes.
``output="torch"`` keeps observations, rewards and dones on the device (the PPO policy consumes them there);
``output="numpy"`` copies the step results to host arrays exactly like SubprocVecEnv returns them (obs dict of stacked
arrays, rewards float32[N], dones bool[N], infos list[dict] with ``terminal_observation`` / Monitor-style ``episode``).
"""
import math
import time
from typing import Any, Dict, List, Optional

import numpy as np

from ..core.factories import create_reward, create_terrain
from ..engine import BallbotEngine, OBS_KEYS
from .spaces import create_action_space, create_observation_space

try:  # pragma: no cover - stable_baselines3 is not installed in the build image
    from stable_baselines3.common.vec_env import VecEnv as _VecEnvBase     # isinstance(env, VecEnv) must hold for PPO(..., env)
    HAVE_SB3 = True
except Exception:
    _VecEnvBase = object
    HAVE_SB3 = False

BUILTIN_TERRAINS = ("perlin", "flat")
BUILTIN_REWARDS = ("directional", "distance")
HFIELD_HALF_EXTENT = 5.0      # ballbot.xml:23 size[0]
DEFAULT_ZSCALE = 2.0          # ballbot.xml:23 size[2]


def _generate_fields(args):
    """Pool worker: heightfields of a plugin terrain for a list of seeds (the generators are pure functions of (config, seed))."""
    terrain_config, seeds = args
    import openballbot_rl_b200.terrain  # noqa: F401
    gen = create_terrain(terrain_config)
    return np.stack([np.asarray(gen(293, seed=int(sd)), np.float32).reshape(-1) for sd in seeds])


def resolve_zscale(terrain_config: dict) -> float:
    """ramp / gradient terrains rescale the heightfield so the physical slope matches the angle (ballbot_env.py:486-495)."""
    kind = terrain_config.get("type")
    cfg = terrain_config.get("config", {}) or {}
    if kind == "ramp":
        return float(2 * HFIELD_HALF_EXTENT * math.tan(math.radians(cfg.get("ramp_angle", 15.0))))
    if kind == "gradient":
        return float(2 * HFIELD_HALF_EXTENT * math.tan(math.radians(cfg.get("max_slope", 20.0))))
    return DEFAULT_ZSCALE


class BallbotVecEnv(_VecEnvBase):
    """``output="numpy"`` returns host arrays with SubprocVecEnv's conventions and, when stable_baselines3 is importable, the
    class IS an SB3 ``VecEnv`` (``PPO("MultiInputPolicy", BallbotVecEnv(N, output="numpy"))`` runs unmodified).
    ``output="torch"`` hands out the engine's persistent device tensors: ``reward`` / ``infos`` values are clones, but the
    observation dict aliases buffers that the next ``step`` overwrites (clone what must outlive a step); its
    ``terminal_observation`` is the [N,16] proprio block (no images)."""

    def __init__(self, num_envs: int, terrain_config: Optional[dict] = None, reward_config: Optional[dict] = None,
                 env_config: Optional[dict] = None, seed: int = 0, disable_cams: bool = False, device: int = 0, precision: int = 64,
                 output: str = "torch", solver: str = "fast", env_offset: int = 0, terrain_type: Optional[str] = None,
                 env_seeds=None, perlin_table: Optional[bool] = None, terrain_bank: Optional[bool] = None):
        import openballbot_rl_b200.rewards  # noqa: F401  (registers the built-ins)
        import openballbot_rl_b200.terrain  # noqa: F401
        if terrain_config is None:
            terrain_config = {"type": terrain_type or "perlin", "config": {}}
        if reward_config is None:
            reward_config = {"type": "directional", "config": {"target_direction": [0.0, 1.0]}}
        env_config = env_config or {}
        env_s, cam_s = env_config.get("env", {}) or {}, env_config.get("camera", {}) or {}
        self.terrain_config, self.reward_config, self.env_config = terrain_config, reward_config, env_config
        self.num_envs = int(num_envs)
        self.output = output
        self.disable_cameras = bool(disable_cams)
        self.max_ep_steps = int(env_s.get("max_ep_steps", 4000))
        rcfg = reward_config.get("config", {}) or {}
        tcfg = terrain_config.get("config", {}) or {}
        ttype, rtype = terrain_config.get("type", "perlin"), reward_config.get("type", "directional")
        # plugin objects are built through the same factories as in the reference (also validates the configs)
        self.reward_obj = create_reward(reward_config)
        self.terrain_gen = create_terrain(terrain_config)
        self._terrain_plugin = ttype not in BUILTIN_TERRAINS
        self._reward_plugin = rtype not in BUILTIN_REWARDS
        # A plugin terrain that does not depend on the per-reset seed (fixed `seed` in its config, or a generator that ignores
        # it: ramp, stairs, bowl, ...) is the same heightfield at every reset of every env (ballbot_env.py:505-513 would
        # regenerate an identical array): upload it once, share it, and keep auto-reset on the device.
        self._terrain_shared = False
        if self._terrain_plugin:
            if tcfg.get("seed") is not None:
                self._terrain_shared = True
            else:
                a, b = (np.asarray(self.terrain_gen(293, seed=sd), np.float32) for sd in (0, 1))
                self._terrain_shared = bool(np.array_equal(a, b))
        # A seed-DEPENDENT plugin terrain (hills, mixed, gradient/perlin, user callables) still has only 10,000 possible fields
        # (r_seed = integers(0, 10000), ballbot_env.py:506).  terrain_bank: all of them are generated once on the host cores and
        # uploaded as a device table, after which resets never leave the device; None = automatic from 2048 envs up.
        self._terrain_bank = bool(self._terrain_plugin and not self._terrain_shared and
                                  (terrain_bank if terrain_bank is not None else int(num_envs) >= 2048))
        self._manual_reset = (self._terrain_plugin and not self._terrain_shared and not self._terrain_bank) or self._reward_plugin
        im_h, im_w = int(cam_s.get("height", 64)), int(cam_s.get("width", 64))
        if not cam_s.get("disable_rgb", True) and not disable_cams:
            raise NotImplementedError("RGB channels need the OpenGL rasteriser; the B200 engine provides depth only (camera.disable_rgb: true)")
        self.engine = BallbotEngine(
            num_envs=self.num_envs, device=device, precision=precision,
            terrain=("shared" if self._terrain_shared else ("table" if self._terrain_bank else "external")) if self._terrain_plugin else ttype,
            terrain_seed=tcfg.get("seed"),
            perlin={k: tcfg[k] for k in ("scale", "octaves", "persistence", "lacunarity", "amplitude") if k in tcfg},
            hfield_zscale=resolve_zscale(terrain_config), cameras=not disable_cams, im_h=im_h, im_w=im_w,
            camera_frame_rate=cam_s.get("frame_rate", 90), max_ep_steps=self.max_ep_steps,
            max_allowed_tilt=env_s.get("max_allowed_tilt", 20.0), max_wheel_velocity=env_s.get("max_wheel_velocity", 10.0),
            reward="external" if self._reward_plugin else rtype, reward_scale=rcfg.get("scale", 0.01),
            action_reg_coef=rcfg.get("action_reg_coef", -0.0001), survival_bonus=rcfg.get("survival_bonus", 0.02),
            target_direction=rcfg.get("target_direction", [0.0, 1.0]), goal_position=rcfg.get("goal_position", [0.0, 0.0]),
            distance_scale=rcfg.get("scale", 1.0) if rtype == "distance" else 1.0,
            seed=seed, auto_reset=not self._manual_reset, env_offset=env_offset, solver=solver, perlin_table=perlin_table,
            seed_stream="pcg64" if env_seeds is not None else "counter")
        if self._terrain_bank:
            self._upload_terrain_bank()
        if env_seeds is not None:
            # the reference's per-env generators: training env i is seeded with seed + i by SB3, eval env i is built with
            # eval_env=[True, seed + N + i] (train.py:83-97); terrain seeds are then integers(0, 10000) draws of that PCG64
            self.engine.seed_pcg64(env_seeds)
        self.observation_space = create_observation_space({"h": im_h, "w": im_w}, 1, disable_cams)
        self.action_space = create_action_space()
        self._rng = np.random.default_rng(seed)       # terrain seeds of plugin terrains (ballbot_env.py:505-510)
        self._terrain_cache = {}                      # host-generated plugin heightfields by seed
        self._actions = None
        self._t0 = time.time()
        self.last_terrain_seeds = np.zeros(self.num_envs, np.int64)
        self.render_mode = None
        self._closed = False
        self._logger = None
        if HAVE_SB3:   # pragma: no cover
            _VecEnvBase.__init__(self, self.num_envs, self.observation_space, self.action_space)

    # ------------------------------------------------------------------ helpers
    def attach_episode_logger(self, log_dir: str, log_options=None, max_envs: int = 4):
        """Reward-term histories and ``terrain_seed_history`` like the reference's env-side logger (utils/logging.py:52-117)."""
        from ..training.episode_logs import EpisodeLogger
        self._logger = EpisodeLogger(self, log_dir, log_options, max_envs)
        return self._logger

    @property
    def opt_timestep(self):
        return 0.002

    def _obs_view(self):
        obs = self.engine.obs
        if self.disable_cameras:
            obs = {k: obs[k] for k in OBS_KEYS}       # the reference emits no timestamp without cameras (App. C #6)
        if self.output == "numpy":
            return {k: v.detach().cpu().numpy() for k, v in obs.items()}
        return obs

    def _upload_plugin_terrain(self, env_ids: np.ndarray):
        if len(env_ids) == 0:
            return
        cfg = self.terrain_config.get("config", {}) or {}
        fields = np.empty((len(env_ids), 293 * 293), np.float32)
        for k, e in enumerate(env_ids):
            r_seed = cfg["seed"] if cfg.get("seed") is not None else int(self._rng.integers(0, 10000))
            self.last_terrain_seeds[e] = r_seed
            if r_seed not in self._terrain_cache:      # generators are pure functions of (config, seed)
                if len(self._terrain_cache) >= 64:
                    self._terrain_cache.clear()
                self._terrain_cache[r_seed] = np.asarray(self.terrain_gen(293, seed=r_seed), np.float32).reshape(-1)
            fields[k] = self._terrain_cache[r_seed]
        self.engine.set_hfield(env_ids.astype(np.int32), fields)

    def _upload_terrain_bank(self, chunk: int = 250):
        """Generate the heightfield of every possible terrain seed with a process pool and fill the engine's table (BB_TERRAIN_TABLE)."""
        import multiprocessing as mp
        import os
        seeds = np.arange(10000)
        jobs = [(self.terrain_config, seeds[k:k + chunk]) for k in range(0, len(seeds), chunk)]
        procs = min(len(jobs), os.cpu_count() or 1)
        with mp.get_context("fork").Pool(procs) as pool:          # fork: custom plugins registered in this process exist in the workers
            for (cfg, sd), fields in zip(jobs, pool.imap(_generate_fields, jobs)):
                self.engine.set_hfield(sd.astype(np.int32), fields)

    # ------------------------------------------------------------------ VecEnv protocol
    def reset(self, terrain_seeds=None):
        """``terrain_seeds`` (int[N], optional): replay recorded ``r_seed`` values instead of drawing them (built-in perlin)."""
        if terrain_seeds is not None:
            if self._terrain_plugin:
                raise ValueError("terrain_seeds applies to the built-in perlin terrain")
            self.engine.reset(seeds=terrain_seeds)
            if self._logger is not None:
                self._logger.after_reset()
            return self._obs_view()
        if self._terrain_shared:
            if not getattr(self, "_shared_uploaded", False):
                cfg = self.terrain_config.get("config", {}) or {}
                r_seed = cfg["seed"] if cfg.get("seed") is not None else 0
                self.last_terrain_seeds[:] = r_seed
                self.engine.set_hfield(np.zeros(1, np.int32), np.asarray(self.terrain_gen(293, seed=r_seed), np.float32).reshape(1, -1))
                self._shared_uploaded = True
        elif self._terrain_plugin and not self._terrain_bank:
            self._upload_plugin_terrain(np.arange(self.num_envs))
        self.engine.reset()
        if self._logger is not None:
            self._logger.after_reset()
        return self._obs_view()

    def step_async(self, actions):
        self._actions = actions

    def step_wait(self):
        import torch
        eng = self.engine
        a = self._actions
        if not hasattr(a, "is_cuda"):
            a = torch.as_tensor(np.asarray(a, np.float32), device=eng.device)
        eng.step(a)
        if self._reward_plugin:   # custom BaseReward evaluated on the device-resident batch (pre-reset observation)
            state = dict(eng.obs); state["pos2d"] = eng.pos2d
            eng.add_reward(torch.as_tensor(self.reward_obj(state), device=eng.device, dtype=torch.float32).reshape(self.num_envs))
        if self._logger is not None and not self._manual_reset:
            self._logger.after_step(a)
        term_obs = None
        if self._manual_reset:
            done_mask = eng.terminated.clone()
            if bool(done_mask.any()):
                term_obs = eng.terminal_obs.clone()
                if self._terrain_plugin and not self._terrain_shared and not self._terrain_bank:
                    self._upload_plugin_terrain(torch.nonzero(done_mask).flatten().cpu().numpy())
                rew, term, fail, pos = eng.reward.clone(), eng.terminated.clone(), eng.failure.clone(), eng.pos2d.clone()
                eng.reset(done_mask)
                eng.reward.copy_(rew); eng.terminated.copy_(term); eng.failure.copy_(fail); eng.pos2d.copy_(pos)
        if self.output == "torch":
            # [N]-sized results are handed out as fresh tensors (SubprocVecEnv returns new arrays every step); only the large
            # observation buffers alias engine memory
            infos = {"failure": eng.failure.clone(), "pos2d": eng.pos2d.clone(),
                     "terminal_observation": eng.terminal_obs.clone() if term_obs is None else term_obs,
                     "episode_r": eng.episode_return.clone(), "episode_l": eng.episode_length.clone()}
            return self._obs_view(), eng.reward.clone(), eng.terminated.bool(), infos
        obs = self._obs_view()
        rewards = eng.reward.cpu().numpy()
        dones = eng.terminated.cpu().numpy().astype(bool)
        fail = eng.failure.cpu().numpy().astype(bool)
        pos2d = eng.pos2d.cpu().numpy()
        infos: List[Dict[str, Any]] = [{"success": False, "failure": bool(fail[i]), "pos2d": pos2d[i]} for i in range(self.num_envs)]
        if dones.any():
            tobs = (eng.terminal_obs if term_obs is None else term_obs).cpu().numpy()
            er, el = eng.episode_return.cpu().numpy(), eng.episode_length.cpu().numpy()
            for i in np.nonzero(dones)[0]:
                t = tobs[i]
                infos[i]["terminal_observation"] = {"orientation": t[0:3], "angular_vel": t[3:6], "vel": t[6:9], "motor_state": t[9:12],
                                                    "actions": t[12:15], "relative_image_timestamp": t[15:16]}
                infos[i]["episode"] = {"r": round(float(er[i]), 6), "l": int(el[i]), "t": round(time.time() - self._t0, 6)}
        return obs, rewards, dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        if not self._closed:
            self.engine.close()
            self._closed = True

    def seed(self, seed: Optional[int] = None):
        self._rng = np.random.default_rng(seed)
        return [None if seed is None else seed + i for i in range(self.num_envs)]

    def get_attr(self, attr_name: str, indices=None):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [getattr(self, attr_name)] * n

    def set_attr(self, attr_name: str, value, indices=None):
        setattr(self, attr_name, value)

    def env_method(self, method_name: str, *args, indices=None, **kwargs):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [getattr(self, method_name)(*args, **kwargs)] * n

    def env_is_wrapped(self, wrapper_class, indices=None):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [False] * n

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
