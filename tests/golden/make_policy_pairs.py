"""Extracts EVERY archived (policy checkpoint, deterministic evaluation) pair of the reference into tests/golden/policy_pairs.npz.

The reference's EvalCallback (ballbot_rl/training/callbacks.py:607-613: 8 episodes, deterministic=True, every 5000 calls = 50,000
steps with 10 envs) wrote results/evaluations.npz; CheckpointCallback / "new best" saved the policy at the same step, so each zip
below pairs a fixed policy with the reference's own closed-loop episode returns / lengths (Monitor values, rounded to 1e-6).

Eval envs are built with eval_env=[True, seed + N_ENVS + i] (ballbot_rl/training/train.py:90-97), i.e. env i owns
numpy PCG64(seed + 10 + i); evaluate_policy with 8 episodes over 10 envs takes the first episode of envs 2..9
(SB3 2.6.0: episode_count_targets = (8 + i) // 10) and appends them in completion order (so lengths are sorted).  For the FIRST
evaluation of a run (50 k steps) the terrain seed of env i is therefore the first `integers(0, 10000)` draw of PCG64(20 + i)
(ballbot_env.py:505-507) -- exactly replayable.  Later evaluations depend on how many auto-resets every eval env has seen, which
is not recorded: those pairs are compared statistically.

Policies share the frozen depth encoders (outputs/encoders/encoder_epoch_53) except for BatchNorm running statistics, so tensors
that are identical in all zips are stored once under "common/".  Run in the build container only; the .npz is committed.
"""
import base64
import io
import json
import os
import zipfile

import numpy as np
import torch

ROOT = "/root/reference/outputs/experiments/archived_models/"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "policy_pairs.npz")
# name -> (zip, evaluations.npz, eval timestep, terrain, replayable first evaluation?)
PAIRS = {
    "flat_seed10_10M": ("2025-12-04_ppo-flat-directional-seed10/checkpoints/ppo_agent_10000000_steps.zip", "2025-12-04_ppo-flat-directional-seed10", 10000000, "flat"),
    "flat_seed10_9p8M": ("2025-12-04_ppo-flat-directional-seed10/checkpoints/ppo_agent_9800000_steps.zip", "2025-12-04_ppo-flat-directional-seed10", 9800000, "flat"),
    "flat_seed10_best150k": ("2025-12-04_ppo-flat-directional-seed10/best_model.zip", "2025-12-04_ppo-flat-directional-seed10", 150000, "flat"),
    "flat_1M_800k": ("2025-12-04_ppo-flat-directional-1M-steps/checkpoints/ppo_agent_800000_steps.zip", "2025-12-04_ppo-flat-directional-1M-steps", 800000, "flat"),
    "flat_1M_best100k": ("2025-12-04_ppo-flat-directional-1M-steps/best_model.zip", "2025-12-04_ppo-flat-directional-1M-steps", 100000, "flat"),
    "perlin_seed10_best50k": ("2025-12-04_ppo-perlin-directional-seed10/best_model.zip", "2025-12-04_ppo-perlin-directional-seed10", 50000, "perlin"),
    "perlin_5p2M_800k": ("2025-12-03_ppo-perlin-directional-5.2M-steps/checkpoints/ppo_agent_800000_steps.zip", "2025-12-03_ppo-perlin-directional-5.2M-steps", 800000, "perlin"),
    "perlin_5p2M_1M": ("2025-12-03_ppo-perlin-directional-5.2M-steps/checkpoints/ppo_agent_1000000_steps.zip", "2025-12-03_ppo-perlin-directional-5.2M-steps", 1000000, "perlin"),
    "perlin_5p2M_best1p25M": ("2025-12-03_ppo-perlin-directional-5.2M-steps/best_model.zip", "2025-12-03_ppo-perlin-directional-5.2M-steps", 1250000, "perlin"),
}
SKIP = ("pi_features_extractor.", "vf_features_extractor.", "mlp_extractor.value_net.", "value_net.")


def main():
    import cloudpickle
    pols, meta = {}, {}
    for name, (zp, run, ts, terrain) in PAIRS.items():
        zf = zipfile.ZipFile(ROOT + zp)
        sd = torch.load(io.BytesIO(zf.read("policy.pth")), map_location="cpu", weights_only=True)
        pols[name] = {k: v.numpy().astype(np.float32) for k, v in sd.items() if not k.startswith(SKIP) and not k.endswith("num_batches_tracked")}
        data = json.loads(zf.read("data"))
        ev = np.load(ROOT + run + "/results/evaluations.npz")
        idx = int(np.nonzero(ev["timesteps"] == ts)[0][0])
        dq = cloudpickle.loads(base64.b64decode(data["ep_info_buffer"][":serialized:"]))
        assert int(data["num_timesteps"]) == ts or "best" in name, (name, data["num_timesteps"])
        meta[name] = dict(terrain=terrain, timestep=ts, eval_index=idx, num_timesteps=int(data["num_timesteps"]),
                          eval_returns=ev["results"][idx].tolist(), eval_lengths=ev["ep_lengths"][idx].tolist(),
                          train_ep_returns=[float(e["r"]) for e in dq], train_ep_lengths=[int(e["l"]) for e in dq])
        print(f"{name:24s} t={ts:9d} idx={idx:3d} zip_steps={data['num_timesteps']:9d} eval len {ev['ep_lengths'][idx]} ret {np.round(ev['results'][idx], 4)}")
    # the perlin seed-10 runs of 12-03 and 12-04 and the 5.2 M run share seed 10: their first evaluation is identical
    for other in ("2025-12-03_ppo-perlin-directional-seed10", "2025-12-03_ppo-perlin-directional-5.2M-steps"):
        e2 = np.load(ROOT + other + "/results/evaluations.npz")
        assert np.array_equal(e2["ep_lengths"][0], np.asarray(meta["perlin_seed10_best50k"]["eval_lengths"]))
    # terrain seeds of the replayable first perlin evaluation: env i of the eval VecEnv owns PCG64(seed + N_ENVS + i), seed = N_ENVS = 10
    meta["perlin_seed10_best50k"]["terrain_seeds_env2to9"] = [int(np.random.default_rng(20 + i).integers(0, 10000)) for i in range(2, 10)]
    keys = list(next(iter(pols.values())).keys())
    out = {}
    for k in keys:
        vals = [p[k] for p in pols.values()]
        if all(np.array_equal(vals[0], v) for v in vals[1:]):
            out["common/" + k] = vals[0]
        else:
            for name, p in pols.items():
                out[name + "/" + k] = p[k]
    out["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT), "bytes;", sum(k.startswith("common/") for k in out), "common tensors,", len(out), "arrays")


if __name__ == "__main__":
    main()
