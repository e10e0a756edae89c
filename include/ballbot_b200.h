/*
 * ballbot_b200.h -- C ABI of the B200 batched ballbot engine (libballbot_b200.so).
 *
 * This is the drop-in boundary for the reference's hot path: patched-MuJoCo mj_step driven through
 * Stable-Baselines3 SubprocVecEnv.  Every entry point names the reference interface
 * it replaces (paths relative to the reference repo root).  Plain C types only; device pointers are owned
 * by the caller (PyTorch tensors) for I/O, the engine owns the persistent simulation state and terrains.
 * All functions return 0 on success or a negative bb_status; the message is available via bb_last_error().
 * No function allocates or synchronises inside bb_step/bb_reset (stream-ordered, CUDA-graph capturable).
 * An engine is not thread-safe (like mjData); one engine per device.
 */
#ifndef BALLBOT_B200_H
#define BALLBOT_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define BB_ABI_VERSION 2
#define BB_HFIELD_N 293 /* ballbot_gym/models/ballbot.xml:23 (nrow = ncol = 293) */
#define BB_NQ 17
#define BB_NV 15
#define BB_PERLIN_SEEDS 10000   /* ballbot_env.py:506: r_seed = integers(0, 10000) => only this many distinct Perlin fields exist */
#define BB_PROBE_MAXCON 80      /* rows of the contact array written by bb_probe_forward / bb_get_contacts */
#define BB_CONTACT_STRIDE 14    /* one contact row: type, dist, pos[3], frame[9] (mjContact.dist / pos / frame) */

typedef enum bb_status {
  BB_OK = 0,
  BB_ERR_INVALID = -1,   /* bad argument / config */
  BB_ERR_CUDA = -2,      /* CUDA runtime error */
  BB_ERR_NO_DEVICE = -3, /* no usable CUDA device: the engine has NO CPU fallback */
  BB_ERR_STATE = -4
} bb_status;

/* BB_TERRAIN_SHARED: one caller-provided heightfield for all envs (plugin terrains that do not depend on the per-reset seed,
 * e.g. ramp / stairs / bowl with a fixed config): uploaded once with bb_set_hfield(env_ids = {0}, n = 1), auto-reset stays on
 * the device like for BB_TERRAIN_FLAT. */
/* BB_TERRAIN_TABLE: caller-provided heightfields for EVERY possible terrain seed (slot = r_seed, BB_PERLIN_SEEDS slots, or one slot
 * with a fixed terrain_seed), uploaded with bb_set_hfield(env_ids = slots): seed-dependent plugin terrains (hills, mixed, user
 * callables) then reset on the device like the built-in Perlin table -- the law r_seed ~ U{0..9999} (ballbot_env.py:506) is kept. */
typedef enum bb_terrain { BB_TERRAIN_FLAT = 0, BB_TERRAIN_PERLIN = 1, BB_TERRAIN_EXTERNAL = 2, BB_TERRAIN_SHARED = 3, BB_TERRAIN_TABLE = 4 } bb_terrain;
typedef enum bb_reward { BB_REWARD_DIRECTIONAL = 0, BB_REWARD_DISTANCE = 1, BB_REWARD_EXTERNAL = 2 } bb_reward;

/* Replaces the constructor arguments / YAML knobs of BBotSimulation.__init__ (ballbot_gym/envs/ballbot_env.py:157-231)
 * plus the terrain generator arguments of generate_perlin_terrain (ballbot_gym/terrain/perlin.py:8-16). */
typedef struct bb_config {
  int32_t abi_version;       /* BB_ABI_VERSION */
  int32_t num_envs;          /* envs simulated by this engine (this rank's shard) */
  int64_t env_offset;        /* global index of env 0 (multi-GPU sharding: rank * num_envs) */
  int32_t device;            /* CUDA device ordinal */
  int32_t precision;         /* 64 (parity mode, MuJoCo mjtNum=double) or 32 */
  int32_t terrain_type;      /* bb_terrain */
  int32_t terrain_seed;      /* >=0: fixed seed from terrain config; <0: per-reset U{0..9999} (ballbot_env.py:505-510) */
  float perlin_scale, perlin_persistence, perlin_lacunarity, perlin_amplitude;
  int32_t perlin_octaves;
  float hfield_zscale;       /* hfield_size[0,2] (2.0; ramp/gradient change it, ballbot_env.py:486-495) */
  int32_t cameras;           /* 0 == disable_cameras=True */
  int32_t im_h, im_w;        /* camera.height / camera.width (64 x 64) */
  float camera_frame_rate;   /* ballbot_env.py:224 (90 Hz) */
  int32_t max_ep_steps;      /* ballbot_env.py:221 (4000) */
  float max_allowed_tilt;    /* ballbot_env.py:222 (20 deg) */
  float max_wheel_velocity;  /* ballbot_env.py:223 (10) */
  int32_t reward_type;       /* bb_reward */
  float reward_scale;        /* ballbot_env.py:229 (0.01) */
  float action_reg_coef;     /* ballbot_env.py:230 (-1e-4) */
  float survival_bonus;      /* ballbot_env.py:231 (0.02) */
  float target_direction[2]; /* rewards/directional.py:33 */
  float goal_position[2];    /* rewards/distance.py:33 */
  float distance_scale;
  uint64_t seed;             /* base seed of the counter-based per-env terrain-seed generator */
  int32_t auto_reset;        /* 1: VecEnv semantics (done envs are reset inside bb_step) */
  int32_t step_kernel;       /* 0: lane-group kernels, split-phase (default); 1: thread-per-env reference mapping (cross-check); 2: lane-group kernel, fused RK4 step (cross-check) */
  int32_t solver_mode;       /* 0: mj_solNewton's own iteration path (exact line search, every RK4 stage warm-starts from qacc_warmstart);
                                1: same cost, Newton direction, tolerance and termination rule; strong-Wolfe line search with cone-apex
                                candidates and analytic p0, stages 2..4 warm-start from the previous stage (same minimiser: single
                                steps agree to ~1e-8 relative) */
  int32_t perlin_table;      /* BB_TERRAIN_PERLIN storage: 1 = all BB_PERLIN_SEEDS possible fields are generated once at bb_create
                                (3.4 GB, independent of num_envs) and a reset only selects one; 0 = one regenerated field per env
                                (343 KB per env, any seed value); -1 = auto (table from 2048 envs up) */
  int32_t seed_stream;       /* terrain-seed draws at (auto-)reset: 0 = counter-based hash of (seed, env, episode) with the law
                                U{0..9999}; 1 = numpy PCG64 per env, bit-compatible with self._np_random.integers(0, 10000)
                                (ballbot_env.py:506); states are uploaded with bb_set_rng_state */
  int32_t depth_kernel;      /* depth observation: 1 (bb_default_config) = ray-caster, one ray per pixel with a heightfield DDA; 0 = the
                                experimental rasteriser with a shared-memory z-buffer (images up to 64 x 64, else the ray-caster): same
                                images up to 0.013 % of the pixels, measured 1.6x slower on the rough Perlin terrain */
} bb_config;

/* Caller-owned device buffers written by bb_step / bb_reset.  Layouts follow the observation dict of
 * BBotSimulation._get_obs (ballbot_env.py:803-827) stacked per key like SubprocVecEnv does. */
typedef struct bb_io {
  float* orientation;      /* [N,3]  obs["orientation"] */
  float* angular_vel;      /* [N,3]  obs["angular_vel"] */
  float* vel;              /* [N,3]  obs["vel"] */
  float* motor_state;      /* [N,3]  obs["motor_state"] */
  float* actions;          /* [N,3]  obs["actions"] */
  float* rel_image_ts;     /* [N,1]  obs["relative_image_timestamp"] */
  float* rgbd_0;           /* [N,1,H,W] obs["rgbd_0"] (NULL when cameras are disabled) */
  float* rgbd_1;           /* [N,1,H,W] obs["rgbd_1"] */
  float* reward;           /* [N]    step() reward (for BB_REWARD_EXTERNAL: everything but the plugin term) */
  uint8_t* terminated;     /* [N]    step() terminated */
  uint8_t* failure;        /* [N]    info["failure"] */
  float* pos2d;            /* [N,2]  info["pos2d"] */
  float* terminal_obs;     /* [N,16] proprio obs before auto-reset (info["terminal_observation"]), valid where terminated */
  float* episode_return;   /* [N]    Monitor info["episode"]["r"], valid where terminated */
  int32_t* episode_length; /* [N]    Monitor info["episode"]["l"], valid where terminated */
  int32_t* status;         /* [N]    bit0: numerical failure (NaN / |x|>1e10) -> env was reset; bits 8..15: max contacts of the RK stages; bits 16..: Newton iterations of the step */
} bb_io;

typedef struct bb_engine bb_engine;

/* replaces gym.make("ballbot-v0.1", ...) x N inside SubprocVecEnv (ballbot_rl/training/train.py:82-97) */
int bb_create(const bb_config* cfg, bb_engine** out);
int bb_destroy(bb_engine* e);
void bb_default_config(bb_config* cfg);
const char* bb_last_error(const bb_engine* e); /* e may be NULL: error of the last failed bb_create */
int bb_num_envs(const bb_engine* e);

/* replaces VecEnv.reset() / BBotSimulation.reset (ballbot_env.py:567-671) for the envs selected by mask
 * (device uint8[N], NULL = all). seeds_dev (device int32[N], NULL = draw from the engine's seed stream) fixes the terrain seed
 * r_seed of every selected env (ballbot_env.py:505-510) -- replay of recorded episodes. Writes the reset observation into io. */
int bb_reset(bb_engine* e, const uint8_t* mask_dev, const int32_t* seeds_dev, const bb_io* io, void* cuda_stream);
/* numpy-compatible terrain-seed stream (seed_stream = 1): per-env PCG64 state as np.random.PCG64(seed).state gives it,
 * device uint64[N][5] = {state_hi, state_lo, inc_hi, inc_lo, has_uint32 | uinteger << 32}; replaces gymnasium's
 * seeding.np_random(seed) behind reset(seed=...) / eval_env=[True, seed] (ballbot_env.py:378-384,596-599) */
int bb_set_rng_state(bb_engine* e, const uint64_t* state_dev, void* cuda_stream);

/* replaces VecEnv.step_async+step_wait / BBotSimulation.step (ballbot_env.py:854-1036): one patched-MuJoCo
 * mj_step (RK4, 2 ms) per env + observation + reward + termination (+ auto-reset). actions_dev: float[N,3]. */
int bb_step(bb_engine* e, const float* actions_dev, const bb_io* io, void* cuda_stream);

/* adds a caller-computed plugin reward term (custom BaseReward, rewards/base.py:7-21) AFTER bb_step when
 * reward_type == BB_REWARD_EXTERNAL: reward += scale*term and the episode-return accumulators are fixed up. */
int bb_add_reward(bb_engine* e, const float* term_dev, const bb_io* io, void* cuda_stream);

/* state injection / readback for parity tests (mjData.qpos/qvel/qacc_warmstart/time); double[N,17|15|15], all
 * pointers are DEVICE pointers, any may be NULL */
int bb_set_state(bb_engine* e, const double* qpos, const double* qvel, const double* warm, void* cuda_stream);
int bb_get_state(bb_engine* e, double* qpos, double* qvel, double* warm, void* cuda_stream);

/* replaces `model.hfield_data = terrain_gen(nrows, seed=r_seed)` (ballbot_env.py:513) for host-generated plugin
 * terrains: env_ids int32[n] (device), hfield float[n,293*293] (device). BB_TERRAIN_EXTERNAL (per env) or
 * BB_TERRAIN_SHARED (n = 1, the single field every env uses). */
int bb_set_hfield(bb_engine* e, const int32_t* env_ids_dev, int32_t n, const float* hfield_dev, void* cuda_stream);
/* copies the heightfield of one env to a device buffer float[293*293] */
int bb_get_hfield(bb_engine* e, int32_t env, float* hfield_dev, void* cuda_stream);
/* terrain seeds currently in use, int32[N] (ballbot_env.py:507 last_r_seed) */
int bb_get_terrain_seeds(bb_engine* e, int32_t* seeds_dev, void* cuda_stream);

/* on-device replacement of generate_perlin_terrain (terrain/perlin.py:8-74): out float[n_seeds,293*293] */
int bb_perlin_terrain(bb_engine* e, const int32_t* seeds_dev, int32_t n_seeds, float* out_dev, void* cuda_stream);
/* engine-independent, host-in/host-out variant for arbitrary n == the registry callable `perlin`
 * (ComponentRegistry.get_terrain("perlin")(n, **cfg), terrain/__init__.py:19): out_host float[nseeds, n*n] */
int bb_perlin_grid(int32_t device, int32_t n, float scale, int32_t octaves, float persistence, float lacunarity, float amplitude,
                   const int32_t* seeds_host, int32_t nseeds, float* out_host);
/* raw untiled 2-D simplex fBm == noise.snoise2(i / scale, j / scale, octaves, persistence, lacunarity, base=seed) for i, j < n
 * (terrain/gradient.py:74-80, gradient_type="perlin"): out_host float[n*n] */
int bb_snoise2_grid(int32_t device, int32_t n, float scale, int32_t octaves, float persistence, float lacunarity, int32_t base, float* out_host);
/* depth ray-cast of both cameras for every env at its CURRENT state (sensors/rgbd.py:46-82), ignoring the cadence */
int bb_render_depth(bb_engine* e, float* rgbd_0, float* rgbd_1, void* cuda_stream);

/* parity probe == inspecting mjData.qacc / qacc_smooth / qfrc_smooth / ncon / solver_niter / contact[] after one
 * mj_forward (reference call sites ballbot_env.py:525,620) through the SAME device code the step kernels run:
 * out_dev double[64] = qacc[0:15], qacc_smooth[15:30], qfrc_smooth[30:45], ncon[45], niter[46], then model constants;
 * contacts_dev double[BB_PROBE_MAXCON][BB_CONTACT_STRIDE], row = {type, dist, pos[3], frame[9]} with type
 * 0..2 ball x wheel_i, 3 heightfield x ball, 4/5 heightfield x camera stick, 6..8 heightfield x wheel_i, 9 ball x tower,
 * 10/11 ball x camera stick (order: robot pairs first, then the ball x heightfield prisms in MuJoCo's scan order). */
int bb_probe_forward(bb_engine* e, int32_t env, const double* ctrl3_dev, double* out_dev, double* contacts_dev, void* cuda_stream);
/* mjData.contact[] / ncon of env `env` at its current state (contact-set parity): contacts_dev as above, ncon_dev int32[1] */
int bb_get_contacts(bb_engine* e, int32_t env, double* contacts_dev, int32_t* ncon_dev, void* cuda_stream);

/* Host-buffer convenience path == what SubprocVecEnv.step does for numpy callers: H2D copy of actions, bb_step,
 * D2H copy of the proprio observation block [N,16] (orientation, angular_vel, vel, motor_state, actions, rel ts),
 * reward, terminated, failure, pos2d (+ images when img_0/img_1 are non-NULL), then a stream synchronise. */
typedef struct bb_host_io {
  float* obs16;        /* [N,16] */
  float* reward;       /* [N] */
  uint8_t* terminated; /* [N] */
  uint8_t* failure;    /* [N] */
  float* pos2d;        /* [N,2] */
  float* terminal_obs; /* [N,16] */
  float* episode_return;   /* [N] */
  int32_t* episode_length; /* [N] */
  float* img_0;        /* [N,H*W] or NULL */
  float* img_1;
} bb_host_io;
/* The engine's own page-locked staging buffers (valid until bb_destroy): a caller that passes these pointers to
 * bb_step_host / bb_reset_host gets the results in place, without the extra host-to-host copy (like SubprocVecEnv's
 * reused observation buffers, the contents are overwritten by the next call). img_0 / img_1 stay NULL. */
int bb_host_buffers(bb_engine* e, float** actions_host, bb_host_io* out);
int bb_step_host(bb_engine* e, const float* actions_host, const bb_host_io* out);
int bb_reset_host(bb_engine* e, const uint8_t* mask_host, const bb_host_io* out);

/* per-kernel device timing (CUDA events on the caller's stream) of the next max_steps bb_step calls; bb_profile_end
 * synchronises and returns the summed milliseconds of {step, terrain, reset, depth} kernels and the step count */
int bb_profile_begin(bb_engine* e, int32_t max_steps);
int bb_profile_end(bb_engine* e, double* ms4, int32_t* nsteps);

/* Generalised advantage estimation on device-resident rollout tensors (engine-independent).  Replaces
 * RolloutBuffer.compute_returns_and_advantage of the reference's learner (SB3 PPO.collect_rollouts, driven from
 * ballbot_rl/training/train.py:126-141,284): rewards float[T,N], values float[T+1,N] (row T = value of the observation
 * after the last step), dones uint8[T,N] (done[t] = episode ended by the transition of step t, i.e. SB3's
 * episode_starts[t+1]); outputs advantages / returns float[T,N] with returns = advantages + values[:T]. */
int bb_gae(const float* rewards_dev, const float* values_dev, const uint8_t* dones_dev, int32_t T, int32_t N, float gamma,
           float gae_lambda, float* advantages_dev, float* returns_dev, void* cuda_stream);

/* Fused PPO optimiser step on flat device buffers: replaces, for the reference's learner (SB3 PPO.train driven from
 * ballbot_rl/training/train.py:126-141,284; hyper-parameters configs/train/ppo_directional.yaml:73-99), the host-side sequence
 * "mean of the all-reduced gradient -> approx-KL early stop (> 1.5 target_kl) -> clip_grad_norm_ -> AdamW.step" with no host
 * synchronisation.  grad_dev float[n + 2]: the (all-reduced, SUM) flat gradient followed by {sum of KL * samples, samples};
 * ctrl_dev double[8]: [0] sticky stop flag, [1] optimiser step count, [2] last gradient norm, [3] last KL, [4] updates applied;
 * scratch_dev double[1].  kl_limit <= 0 disables the early stop, max_grad_norm <= 0 the clipping. */
int bb_adamw_step(float* param_dev, const float* grad_dev, float* m_dev, float* v_dev, int32_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, float max_grad_norm, float kl_limit, double* ctrl_dev, double* scratch_dev, void* cuda_stream);

/* profiling aid: measured fp64 FMA throughput of `device` in TFLOP/s (8 independent DFMA chains per thread, 148 x 8 CTAs);
 * the denominator of bench.py's roofline.compute (the step kernels are fp64-latency bound, not HBM bound) */
int bb_fp64_peak(int32_t device, double* tflops);

/* number of kernel launches issued by this engine so far (bench.py's gpu_launches claim) */
int64_t bb_launch_count(const bb_engine* e);
/* engine model constants for tests: dA[12] (diagApprox per contact type), meaninertia, masses (m0, mw, mL) */
int bb_model_constants(double* dA12, double* meaninertia, double* masses3);
/* "libballbot_b200 abi <n> built <date> <time> src <hash of the translation unit's sources>": lets a test log tell a stale .so from a fresh one */
const char* bb_build_info(void);

#ifdef __cplusplus
}
#endif
#endif
