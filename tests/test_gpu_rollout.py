"""GPU tests of the rollout post-processing (bb_gae through the C ABI) and of the PPO learner on the real engine."""
import numpy as np
import pytest
import torch

from oracle import gae_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,N", [(1, 1), (7, 3), (8, 128), (37, 1000), (2048, 257), (0, 5)])
def test_gae_kernel_matches_sb3_restatement(T, N):
    from openballbot_rl_b200.training.gae import compute_gae
    rng = np.random.default_rng(T * 1000 + N)
    rew = rng.normal(size=(T, N)).astype(np.float32); val = rng.normal(size=(T + 1, N)).astype(np.float32)
    done = rng.random((T, N)) < 0.05
    adv, ret = compute_gae(torch.from_numpy(rew).cuda(), torch.from_numpy(val).cuda(), torch.from_numpy(done).cuda(), 0.99, 0.95)
    a_ref, r_ref = gae_oracle.gae(rew, val[:-1], done, val[-1], 0.99, 0.95)
    # float32 recursion; the kernel may contract a*b+c into FMAs, so compare with a tolerance scaled to the horizon sum
    tol = 1e-5 * max(1.0, float(np.abs(a_ref).max()) if T else 1.0)
    assert adv.shape == (T, N) and np.abs(adv.cpu().numpy() - a_ref).max(initial=0.0) < tol
    assert np.abs(ret.cpu().numpy() - r_ref).max(initial=0.0) < tol


def test_gae_rejects_cpu_tensors():
    from openballbot_rl_b200._lib import EngineError
    from openballbot_rl_b200.training.gae import compute_gae
    with pytest.raises(EngineError):
        compute_gae(torch.zeros(2, 2), torch.zeros(3, 2), torch.zeros(2, 2, dtype=torch.bool))


def test_ppo_iteration_on_engine_with_embedding_cache():
    from openballbot_rl_b200.training.policy import BallbotPolicy
    from openballbot_rl_b200.training.ppo import PPOConfig, PPOLearner
    from openballbot_rl_b200.training.utils import make_ballbot_vec_env
    torch.manual_seed(0)
    venv = make_ballbot_vec_env(64, terrain_config={"type": "perlin", "config": {}}, seed=3)
    pol = BallbotPolicy().cuda()
    L = PPOLearner(venv, pol, PPOConfig(n_steps=24, batch_size=256, n_epochs=2), total_timesteps=64 * 24)
    before = torch.cat([p.detach().reshape(-1).clone() for p in L.params])
    buf, stats = L.collect()
    # the cached embeddings must equal a full re-encode of the current images
    obs = venv._obs_view()
    with torch.no_grad():
        for k in ("rgbd_0", "rgbd_1"):
            assert torch.allclose(L._emb[k], pol.encoders[k](obs[k]), atol=2e-3)   # BatchNorm-folded copy, cuDNN TF32 convolutions
    assert buf["feat"].shape == (24, 64, 56) and torch.isfinite(buf["adv"]).all() and stats["env_steps"] == 64 * 24
    info = L.update(buf)
    after = torch.cat([p.detach().reshape(-1) for p in L.params])
    assert info["n_updates"] > 0 and np.isfinite(info["value_loss"]) and not torch.equal(before, after)
    venv.close()


def test_depth_frame_collection(tmp_path):
    """Encoder pre-training data: fresh frames only (one in six steps per env + the reset frames), both cameras, sharded."""
    from openballbot_rl_b200.training.collect import collect_depth_frames
    from openballbot_rl_b200.training.utils import make_ballbot_vec_env
    venv = make_ballbot_vec_env(32, terrain_config={"type": "perlin", "config": {}}, seed=1)
    n = collect_depth_frames(venv, None, n_steps=24, out_dir=str(tmp_path), shard_frames=100)
    files = sorted(p.name for p in tmp_path.iterdir())
    frames = np.concatenate([np.load(tmp_path / f) for f in files])
    assert n == frames.shape[0] == 2 * 32 * (1 + 24 // 6) and frames.shape[1:] == (64, 64) and frames.dtype == np.float16
    assert files[0] == "depth_0000.npy" and len(files) == (n + 99) // 100
    assert 0.0 < frames.min() and frames.max() <= 1.0
    venv.close()


def test_fused_adamw_step_matches_torch_adamw_with_clipping_and_kl_stop():
    """bb_adamw_step (mean of the all-reduced flat gradient, clip_grad_norm_, AdamW, sticky target-KL stop on the device) against
    torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW on the same gradients."""
    import ctypes as C
    from openballbot_rl_b200 import _lib
    torch.manual_seed(0)
    n = 114_000
    p_ref = torch.nn.Parameter(torch.randn(n, device="cuda") * 0.1)
    opt = torch.optim.AdamW([p_ref], lr=1e-3, weight_decay=0.01)
    p = p_ref.detach().clone(); m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    g = torch.zeros(n + 2, device="cuda"); ctrl = torch.zeros(8, dtype=torch.float64, device="cuda"); scr = torch.zeros(1, dtype=torch.float64, device="cuda")
    vp = lambda t: C.c_void_p(t.data_ptr())
    L = _lib.lib()
    for step in range(6):
        cnt = 512.0
        grad_mean = torch.randn(n, device="cuda") * (0.02 if step % 2 else 0.0005)       # every other step exceeds max_grad_norm = 0.5
        g[:n] = grad_mean * cnt; g[n] = 0.01 * cnt; g[n + 1] = cnt                          # SUM over the minibatch, KL below the limit
        rc = L.bb_adamw_step(vp(p), vp(g), vp(m), vp(v), n, 1e-3, 0.9, 0.999, 1e-8, 0.01, 0.5, 0.45, vp(ctrl), vp(scr), None)
        assert rc == 0
        p_ref.grad = grad_mean.clone()
        torch.nn.utils.clip_grad_norm_([p_ref], 0.5)
        opt.step()
        assert torch.allclose(p, p_ref.detach(), rtol=2e-5, atol=2e-7), step
    c = ctrl.cpu().numpy()
    assert c[0] == 0 and c[1] == 6 and c[4] == 6 and abs(c[3] - 0.01) < 1e-7
    # KL above 1.5 x target: the step is skipped, the flag sticks for the rest of the iteration
    before = p.clone()
    g[n] = 0.5 * 512.0
    assert L.bb_adamw_step(vp(p), vp(g), vp(m), vp(v), n, 1e-3, 0.9, 0.999, 1e-8, 0.01, 0.5, 0.45, vp(ctrl), vp(scr), None) == 0
    g[n] = 0.0
    assert L.bb_adamw_step(vp(p), vp(g), vp(m), vp(v), n, 1e-3, 0.9, 0.999, 1e-8, 0.01, 0.5, 0.45, vp(ctrl), vp(scr), None) == 0
    c = ctrl.cpu().numpy()
    assert torch.equal(p, before) and c[0] == 1 and c[1] == 6
