"""BBotSimulation: single-environment view with the reference's gym.Env surface (ballbot_gym/envs/ballbot_env.py:60).

It is a one-env CUDA engine behind ``reset(seed=None, goal="random") -> (obs, info)`` and
``step(a) -> (obs, reward, terminated, truncated, info)`` with numpy observations, the same constructor keywords and
the same per-env RNG law for terrain seeds (``np.random.default_rng(seed).integers(0, 10000)``, ballbot_env.py:505-510).
GUI / video rendering are out of scope and raise.
"""
from typing import Optional

import numpy as np

from ..core.factories import create_reward, create_terrain
from ..engine import BallbotEngine
from .spaces import create_action_space, create_observation_space
from .vec_env import BUILTIN_TERRAINS, resolve_zscale

_default_dtype = np.float32

try:  # pragma: no cover - gymnasium is not installed in the build image
    import gymnasium as _gym
    _EnvBase = _gym.Env          # gym.make("ballbot-v0.1") / Monitor / SB3 wrappers check isinstance(env, gymnasium.Env)
except Exception:
    _EnvBase = object


class BBotSimulation(_EnvBase):
    metadata = {"render_modes": ["rgb_array"], "render_fps": 30}
    if _EnvBase is object:   # gymnasium.Env provides the same property over the same `_np_random` attribute
        np_random = property(lambda self: self._np_random)

    def __init__(self, xml_path=None, GUI=False, im_shape={"h": 64, "w": 64}, disable_cameras=False, depth_only=True,
                 log_options={"cams": False, "reward_terms": False}, max_ep_steps=None, terrain_type: str = "perlin",
                 eval_env=[False, None], reward_config=None, terrain_config=None, env_config=None,
                 render_mode: Optional[str] = None, viewer_title: Optional[str] = None, device: int = 0, precision: int = 64):
        import openballbot_rl_b200.rewards  # noqa: F401
        import openballbot_rl_b200.terrain  # noqa: F401
        if render_mode is not None and render_mode not in self.metadata["render_modes"]:
            raise ValueError(f"Invalid render_mode: {render_mode}. Supported modes: {self.metadata['render_modes']}")
        if GUI:
            raise NotImplementedError("the MuJoCo passive viewer is not part of the B200 engine (GUI=False only)")
        if terrain_config is None:
            terrain_config = {"type": terrain_type, "config": {}}
        self.terrain_config = terrain_config
        self.terrain_type = terrain_config.get("type", terrain_type)
        if reward_config is None:
            reward_config = {"type": "directional", "config": {"target_direction": [0.0, 1.0]}}
        self.reward_config = reward_config
        env_config = env_config or {}
        env_s, cam_s, log_s = env_config.get("env", {}) or {}, env_config.get("camera", {}) or {}, env_config.get("logging", {}) or {}
        self.log_options = {**log_options, **log_s} if log_s else log_options
        self.xml_path = xml_path                                    # accepted for API compatibility; the model is compiled in
        self.max_ep_steps = env_s.get("max_ep_steps", max_ep_steps if max_ep_steps is not None else 4000)
        self.max_allowed_tilt = env_s.get("max_allowed_tilt", 20.0)
        self.max_wheel_velocity = env_s.get("max_wheel_velocity", 10.0)
        self.camera_frame_rate = cam_s.get("frame_rate", 90)
        self.depth_only = depth_only
        rcfg = reward_config.get("config", {}) or {}
        self.reward_scale = rcfg.get("scale", 0.01)
        self.action_reg_coef = rcfg.get("action_reg_coef", -0.0001)
        self.survival_bonus = rcfg.get("survival_bonus", 0.02)
        self.action_space = create_action_space()
        actual_depth_only = cam_s.get("disable_rgb", depth_only) if cam_s else depth_only
        if not actual_depth_only and not disable_cameras:
            raise NotImplementedError("RGB channels need the OpenGL rasteriser; the B200 engine provides depth only")
        h = cam_s.get("height", im_shape["h"]) if cam_s else im_shape["h"]
        w = cam_s.get("width", im_shape["w"]) if cam_s else im_shape["w"]
        self.observation_space = create_observation_space({"h": h, "w": w}, 1, disable_cameras)
        self.disable_cameras = disable_cameras
        self.render_mode = render_mode if render_mode is not None else "rgb_array"
        self.passive_viewer = None
        self.log_dir = None
        self.num_episodes = -1
        self.eval_env = eval_env[0]
        self._np_random = np.random.default_rng(eval_env[1]) if self.eval_env else None
        self.verbose = False
        self.step_counter = 0
        self.G_tau = 0.0
        tcfg = terrain_config.get("config", {}) or {}
        self._terrain_builtin = self.terrain_type in BUILTIN_TERRAINS
        rtype = reward_config.get("type", "directional")
        # distance cannot work through the env in the reference either: obs has no 'pos2d'  -> host reward
        self._reward_on_device = rtype == "directional"
        self.engine = BallbotEngine(
            num_envs=1, device=device, precision=precision, terrain="external", hfield_zscale=resolve_zscale(terrain_config),
            perlin={k: tcfg[k] for k in ("scale", "octaves", "persistence", "lacunarity", "amplitude") if k in tcfg},
            cameras=not disable_cameras, im_h=h, im_w=w, camera_frame_rate=self.camera_frame_rate, max_ep_steps=self.max_ep_steps,
            max_allowed_tilt=self.max_allowed_tilt, max_wheel_velocity=self.max_wheel_velocity,
            reward="directional" if self._reward_on_device else "external", reward_scale=self.reward_scale,
            action_reg_coef=self.action_reg_coef, survival_bonus=self.survival_bonus,
            target_direction=rcfg.get("target_direction", [0.0, 1.0]), auto_reset=False)
        self._reset_goal_and_reward_objs()

    # ------------------------------------------------------------------ reference helpers
    @property
    def opt_timestep(self):
        return 0.002

    def effective_camera_frame_rate(self):
        n = np.ceil((1 / self.camera_frame_rate) / self.opt_timestep)
        return 1.0 / (n * self.opt_timestep)

    def _reset_goal_and_reward_objs(self):
        self.reward_obj = create_reward(self.reward_config)
        td = getattr(self.reward_obj, "target_direction", None)
        self.goal_2d = [0.0, 1.0] if td is None else (td.tolist() if hasattr(td, "tolist") else list(td))

    def _obs_numpy(self, host):
        o = host["obs16"][0]
        obs = {"orientation": o[0:3].copy(), "angular_vel": o[3:6].copy(), "vel": o[6:9].copy(), "motor_state": o[9:12].copy(),
               "actions": o[12:15].copy()}
        if not self.disable_cameras:
            obs["rgbd_0"] = host["img_0"][0].copy(); obs["rgbd_1"] = host["img_1"][0].copy()
            obs["relative_image_timestamp"] = o[15:16].copy()
        return obs

    def _info(self, host):
        return {"success": False, "failure": False, "step_counter": self.step_counter, "pos2d": host["pos2d"][0].copy()}

    # ------------------------------------------------------------------ gym.Env API
    def reset(self, seed=None, goal: str = "random", **kwargs):
        # gymnasium's Env.reset(seed=s) re-creates `_np_random` -- the very attribute the reference draws terrain seeds from
        # (ballbot_env.py:378-384, 596-599): an explicit seed replaces the stream even for eval envs, no seed continues it
        if seed is not None or self._np_random is None:
            self._np_random = np.random.default_rng(seed)
        self._reset_goal_and_reward_objs()
        self.step_counter = 0
        tcfg = self.terrain_config.get("config", {}) or {}
        r_seed = tcfg["seed"] if tcfg.get("seed") is not None else self._np_random.integers(0, 10000)
        self.last_r_seed = r_seed
        if self.terrain_type == "perlin":
            hf = self.engine.perlin_terrain([int(r_seed)])
        else:
            gen = create_terrain(self.terrain_config)
            hf = np.asarray(gen(293, seed=r_seed), np.float32)
        self.engine.set_hfield([0], hf)
        host = self.engine.reset_host()
        self.G_tau = 0.0
        self.num_episodes += 1
        return self._obs_numpy(host), self._info(host)

    def step(self, omniwheel_commands):
        a = np.asarray(omniwheel_commands, dtype=np.float32).reshape(1, 3)
        host = self.engine.step_host(a)
        obs = self._obs_numpy(host)
        self.step_counter += 1
        info = self._info(host)
        reward = host["reward"][0]
        if not self._reward_on_device:     # plugin / distance reward on the host, with the reference's exact call
            reward = np.float32(reward + np.float32(self.reward_obj(obs) * self.reward_scale))
        terminated = bool(host["terminated"][0])
        if bool(host["failure"][0]):
            info["success"] = False; info["failure"] = True
        self.G_tau += float(reward)
        return obs, reward, terminated, False, info

    def render(self):
        raise NotImplementedError("video rendering (world_view RGB camera) needs OpenGL and is out of the hot-path scope")

    def close(self):
        if getattr(self, "engine", None) is not None:
            self.engine.close()
            self.engine = None
