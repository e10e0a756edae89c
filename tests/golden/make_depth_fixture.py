"""Extracts real depth observations of the reference (OpenGL `mujoco.Renderer`, sensors/rgbd.py:46-82) from the archived
flat-terrain SB3 checkpoints into tests/golden/ref_depth_samples.npz.

Source: `_last_obs` inside /root/reference/outputs/experiments/archived_models/*flat*/{checkpoints,results}/*.zip -- the
observation dict of the 10 training envs at save time.  Only samples with relative_image_timestamp == 0 are kept (image and
proprioceptive observation belong to the same step), with the base orientation (rotation vector of xquat[base],
ballbot_env.py:778-779).  On flat terrain the robot's pose follows from the orientation alone up to small dynamic effects
(base rotating about the ball centre), which is what the depth tests reconstruct.  Run in the build container only.
"""
import base64
import glob
import json
import os
import zipfile

import cloudpickle
import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_depth_samples.npz")
rot, im0, im1, src = [], [], [], []
for z in sorted(glob.glob("/root/reference/outputs/experiments/archived_models/*flat*/*/*.zip")):
    lo = cloudpickle.loads(base64.b64decode(json.loads(zipfile.ZipFile(z).read("data"))["_last_obs"][":serialized:"]))
    if "rgbd_0" not in lo:
        continue
    for i in range(lo["rgbd_0"].shape[0]):
        if lo["relative_image_timestamp"][i, 0] != 0:
            continue
        key = lo["rgbd_0"][i].tobytes()
        if any(key == k for k in src):      # best_model.zip duplicates a checkpoint
            continue
        src.append(key)
        rot.append(lo["orientation"][i]); im0.append(lo["rgbd_0"][i, 0]); im1.append(lo["rgbd_1"][i, 0])
np.savez_compressed(OUT, orientation=np.array(rot, np.float32), rgbd_0=np.array(im0, np.float16), rgbd_1=np.array(im1, np.float16))
print(OUT, os.path.getsize(OUT), "bytes;", len(rot), "samples")
