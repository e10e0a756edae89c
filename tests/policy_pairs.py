"""Loader for tests/golden/policy_pairs.npz (made by tests/golden/make_policy_pairs.py): the reference's archived
(policy checkpoint, deterministic evaluation) pairs."""
import json
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "policy_pairs.npz")
_cache = {}


def _load():
    if not _cache:
        z = np.load(_PATH)
        _cache["z"] = z
        _cache["meta"] = json.loads(bytes(z["meta_json"]).decode())
    return _cache["z"], _cache["meta"]


def names():
    return list(_load()[1].keys())


def meta(name):
    return _load()[1][name]


def policy(name, device="cpu"):
    """BallbotPolicy (eval mode) with the weights of the archived SB3 zip `name`."""
    from openballbot_rl_b200.training.policy import BallbotPolicy
    z, _ = _load()
    arrays = {k[len("common/"):]: z[k] for k in z.files if k.startswith("common/")}
    arrays.update({k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "/")})
    return BallbotPolicy().load_sb3_state(arrays).eval().to(device)


def oracle_episode(oracle_mod, pol, hfield=None, max_steps=4000):
    """One deterministic closed-loop episode of `pol` on the CPU oracle; returns (length, return, failed)."""
    import torch
    e = oracle_mod.OracleEnv(cameras=True)
    o = e.reset(hfield); d0, d1 = e.depth()
    G, n = 0.0, 0
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))[None]
    with torch.no_grad():
        while n < max_steps:
            obs = {"orientation": t(o[0:3]), "angular_vel": t(o[3:6]), "vel": t(o[6:9]), "motor_state": t(o[9:12]), "actions": t(o[12:15]),
                   "relative_image_timestamp": t(o[15:16]), "rgbd_0": t(d0)[None], "rgbd_1": t(d1)[None]}
            o, r, term, fail, _ = e.step(pol(obs)[0].numpy())
            d0, d1 = e.depth(); G += r; n += 1
            if term:
                break
    e.close()
    return n, G, fail
