"""Extracts the archived reference PPO policy (flat terrain, 10 M steps) into tests/golden/policy_flat_10M.npz.

Source: /root/reference/outputs/experiments/archived_models/2025-12-04_ppo-flat-directional-seed10/checkpoints/
ppo_agent_10000000_steps.zip (SB3 zip: policy.pth state dict).  Recorded deterministic evaluation of this checkpoint:
return 9.198632, episode length 378, 8/8 episodes identical (results/evaluations.npz) -- the only reference-pinned
deterministic end-to-end number for the hot path -- plus the checkpoint's Monitor buffer: the last 100 training episodes under
the stochastic policy (return 8.005 +- 0.712, length 318.7 +- 36.0).  Run in the build container only; the .npz is committed.
"""
import io
import os
import zipfile

import numpy as np
import torch

ZIP = "/root/reference/outputs/experiments/archived_models/2025-12-04_ppo-flat-directional-seed10/checkpoints/ppo_agent_10000000_steps.zip"
EVAL = "/root/reference/outputs/experiments/archived_models/2025-12-04_ppo-flat-directional-seed10/results/evaluations.npz"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "policy_flat_10M.npz")

sd = torch.load(io.BytesIO(zipfile.ZipFile(ZIP).read("policy.pth")), map_location="cpu", weights_only=True)
keep = {}
for k, v in sd.items():
    if k.startswith(("pi_features_extractor.", "vf_features_extractor.", "mlp_extractor.value_net.", "value_net.")) or k.endswith("num_batches_tracked"):
        continue
    keep[k] = v.numpy().astype(np.float32)
# last 100 training episodes of this checkpoint under the stochastic policy (Monitor ep_info_buffer inside the SB3 zip)
import base64
import json
import cloudpickle
dq = cloudpickle.loads(base64.b64decode(json.loads(zipfile.ZipFile(ZIP).read("data"))["ep_info_buffer"][":serialized:"]))
keep["train_ep_returns"] = np.array([e["r"] for e in dq], np.float64)
keep["train_ep_lengths"] = np.array([e["l"] for e in dq], np.int64)
ev = np.load(EVAL)
keep["eval_timesteps"] = ev["timesteps"][-1:]
keep["eval_return"] = ev["results"][-1]
keep["eval_length"] = ev["ep_lengths"][-1]
np.savez_compressed(OUT, **keep)
print(OUT, os.path.getsize(OUT), "bytes;", len(keep), "arrays; recorded eval:", keep["eval_return"][0], keep["eval_length"][0])
