"""Episode data products of the reference's env-side logger on the GPU VecEnv (ballbot_gym/utils/logging.py:52-117).

The reference keeps, per (non-eval) env, the per-step reward terms of the running episode -- term 1 = reward_obj(obs) x scale,
term 2 = action regularisation (ballbot_env.py:929-937) -- saves them as ``term_1.npy`` / ``term_2.npy`` when the episode ends
and appends the episode's terrain seed to ``terrain_seed_history`` (perlin only).  ``EpisodeLogger`` produces the same files for
the first ``max_envs`` envs of a batch, one ``env_<i>`` directory each (the reference has one log_dir per env process).
It is opt-in: reading a few scalars per step costs a device synchronisation.
"""
import os

import numpy as np
import torch


class EpisodeLogger:
    def __init__(self, venv, log_dir: str, log_options=None, max_envs: int = 4):
        self.venv, self.log_dir = venv, log_dir
        self.log_options = {"cams": False, "reward_terms": True, **(log_options or {})}
        self.ids = list(range(min(int(max_envs), venv.num_envs)))
        self.num_episodes = [0] * len(self.ids)
        self._t1 = [[] for _ in self.ids]; self._t2 = [[] for _ in self.ids]
        self._seed = None
        rc = venv.reward_config.get("config", {}) or {}
        self.scale, self.reg = float(rc.get("scale", 0.01)), float(rc.get("action_reg_coef", -0.0001))
        self.tdir = np.asarray(rc.get("target_direction", [0.0, 1.0]), np.float32)
        self.perlin = venv.terrain_config.get("type") == "perlin"
        for i in self.ids:
            os.makedirs(self._dir(i), exist_ok=True)

    def _dir(self, i):
        return os.path.join(self.log_dir, f"env_{i}")

    def after_reset(self):
        self._seed = self.venv.engine.terrain_seeds()[self.ids].cpu().numpy()

    def after_step(self, actions):
        eng = self.venv.engine
        ids = torch.as_tensor(self.ids, device=eng.device)
        done = eng.terminated[ids].bool()
        vel = torch.where(done[:, None], eng.terminal_obs[ids, 6:8], eng.obs["vel"][ids, :2]).cpu().numpy()   # the obs the reward saw
        a = torch.as_tensor(actions, device=eng.device, dtype=torch.float32)[ids].cpu().numpy()
        done = done.cpu().numpy()
        new_seed = eng.terrain_seeds()[ids].cpu().numpy()
        if self._seed is None:
            self._seed = new_seed
        for k, i in enumerate(self.ids):
            self._t1[k].append(np.float32(np.float32(vel[k] @ self.tdir) * np.float32(self.scale)))
            self._t2[k].append(np.float32(self.reg) * np.float32(np.linalg.norm(a[k]) ** 2))
            if done[k]:
                if self.log_options.get("reward_terms", False):
                    np.save(os.path.join(self._dir(i), "term_1"), np.array(self._t1[k]))
                    np.save(os.path.join(self._dir(i), "term_2"), np.array(self._t2[k]))
                if self.perlin:
                    with open(os.path.join(self._dir(i), "terrain_seed_history"), "a") as fl:
                        fl.write(f"{int(self._seed[k])}\n")
                self._t1[k], self._t2[k] = [], []
                self.num_episodes[k] += 1
        self._seed = new_seed
