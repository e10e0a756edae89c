import sys, os, numpy as np, torch
sys.path.insert(0, ".")
from openballbot_rl_b200.envs import BallbotVecEnv
from openballbot_rl_b200.training.policy import BallbotPolicy
from openballbot_rl_b200.training.evaluate import evaluate_policy
z = np.load("tests/golden/policy_flat_10M.npz")
pol = BallbotPolicy().load_sb3_state({k: z[k] for k in z.files if not k.startswith(("eval_", "train_"))}).eval().cuda()
print("log_std", z["log_std"])
ENV_CFG = {"camera": {"height": 64, "width": 64, "frame_rate": 90, "disable_rgb": True}, "env": {"max_ep_steps": 4000, "max_allowed_tilt": 20, "max_wheel_velocity": 10.0}}
REWARD = {"type": "directional", "config": {"target_direction": [0.0, 1.0], "scale": 0.01, "action_reg_coef": -0.0001, "survival_bonus": 0.02}}
venv = BallbotVecEnv(1024, terrain_config={"type": "flat", "config": {}}, reward_config=REWARD, env_config=ENV_CFG, precision=64)
torch.manual_seed(0)
out = evaluate_policy(venv, pol, max_steps=700, deterministic=False)
L, G = out["lengths"].float().cpu().numpy(), out["returns"].cpu().numpy()
print("stochastic: len %.1f +- %.1f [%d..%d]  ret %.3f +- %.3f  | reference: len %.1f +- %.1f  ret %.3f +- %.3f" % (L.mean(), L.std(), L.min(), L.max(), G.mean(), G.std(), z["train_ep_lengths"].mean(), z["train_ep_lengths"].std(), z["train_ep_returns"].mean(), z["train_ep_returns"].std()))
