/*
 * bb_cpu_ref.cpp -- the CPU ORACLE behind the engine's own C ABI (include/ballbot_b200.h), "backend = cpu_ref".
 *
 * TEST INFRASTRUCTURE ONLY (last sentence: "Same symbols implemented by the CPU oracle library so one test
 * harness drives both").  N independent fp64 oracle envs (ballbot_oracle.cpp, PARITY UNPINNED -- see ORACLE_ASSUMPTIONS.md) are
 * stepped one after the other; every pointer the CUDA library documents as a DEVICE pointer is a HOST pointer here and the
 * stream argument is ignored.  Batch semantics restated from the engine's contract: auto-reset inside bb_step with terminal
 * observation and Monitor statistics, terrain-seed draws (counter hash or numpy PCG64, both bit-compatible with the engine),
 * explicit seeds in bb_reset, camera cadence.  Entry points that only make sense on the GPU (host staging path, per-kernel
 * profiling, GAE / AdamW kernels, fp64 peak) return BB_ERR_STATE.
 * Nothing in the product package may load this library (tests/test_abi.py checks).
 */
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../include/ballbot_b200.h"
#include "ballbot_oracle.h"

namespace {
const int HF = BB_HFIELD_N * BB_HFIELD_N;
char g_err[256] = "";

unsigned long long splitmix(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
unsigned pcg64Next32(unsigned long long* st) {   // numpy PCG64 + pcg64_next32 buffering (same restatement as csrc/bb_engine.cu)
  const unsigned long long buf = st[4];
  if (buf & 1ull) { st[4] = 0; return (unsigned)(buf >> 32); }
  const unsigned __int128 mult = ((unsigned __int128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull;
  unsigned __int128 state = ((unsigned __int128)st[0] << 64) | st[1];
  const unsigned __int128 inc = ((unsigned __int128)st[2] << 64) | st[3];
  state = state * mult + inc;
  const unsigned long long hi = (unsigned long long)(state >> 64), lo = (unsigned long long)state;
  st[0] = hi; st[1] = lo;
  const unsigned long long x = hi ^ lo; const unsigned r = (unsigned)(hi >> 58);
  const unsigned long long out = (x >> r) | (x << ((64u - r) & 63u));
  st[4] = 1ull | ((out >> 32) << 32);
  return (unsigned)out;
}
int pcg64Integers10000(unsigned long long* st) {
  const unsigned rng = 9999u, excl = 10000u;
  unsigned long long m = (unsigned long long)pcg64Next32(st) * excl;
  unsigned left = (unsigned)m;
  if (left < excl) { const unsigned thr = (0xFFFFFFFFu - rng) % excl; while (left < thr) { m = (unsigned long long)pcg64Next32(st) * excl; left = (unsigned)m; } }
  return (int)(m >> 32);
}
}  // namespace

struct bb_engine {
  bb_config cfg;
  int N;
  std::vector<bbo_env*> env;
  std::vector<std::vector<float>> hf;       // per env (perlin / external) or one entry (shared); flat: none
  std::vector<unsigned> episode;
  std::vector<int> tseed, ep_len;
  std::vector<float> ep_ret;
  std::vector<unsigned long long> rng;      // [N][5]
  int64_t calls;
  char err[256];
};

namespace {
int fail(bb_engine* e, int code, const char* msg) { snprintf(e->err, sizeof(e->err), "%s", msg); return code; }
int drawSeed(bb_engine* e, int i) {
  const bb_config& c = e->cfg;
  if (c.terrain_seed >= 0) return c.terrain_seed;
  if (c.seed_stream == 1) return pcg64Integers10000(&e->rng[5 * (size_t)i]);
  const unsigned long long h = splitmix(c.seed ^ splitmix((unsigned long long)(c.env_offset + i) * 0x100000001B3ull + e->episode[i]));
  return (int)(h % 10000ull);
}
const float* fieldOf(bb_engine* e, int i) {
  if (e->cfg.terrain_type == BB_TERRAIN_FLAT) return nullptr;
  if (e->cfg.terrain_type == BB_TERRAIN_SHARED) return e->hf[0].data();
  return e->hf[i].data();
}
void writeObs(const bb_io* io, int i, const float* o) {
  for (int k = 0; k < 3; k++) {
    io->orientation[3 * i + k] = o[k]; io->angular_vel[3 * i + k] = o[3 + k]; io->vel[3 * i + k] = o[6 + k];
    io->motor_state[3 * i + k] = o[9 + k]; io->actions[3 * i + k] = o[12 + k];
  }
  io->rel_image_ts[i] = o[15];
}
void writeImages(bb_engine* e, const bb_io* io, int i) {
  if (!e->cfg.cameras) return;
  const size_t npix = (size_t)e->cfg.im_h * e->cfg.im_w;
  bbo_get_depth(e->env[i], io->rgbd_0 + npix * i, io->rgbd_1 + npix * i);
}
// terrain of the new episode + oracle reset + reset observation (ballbot_env.py:567-671)
void resetEnv(bb_engine* e, const bb_io* io, int i, int seed) {
  e->tseed[i] = seed;
  if (e->cfg.terrain_type == BB_TERRAIN_PERLIN)
    bbo_perlin_terrain(BB_HFIELD_N, e->cfg.perlin_scale, e->cfg.perlin_octaves, e->cfg.perlin_persistence, e->cfg.perlin_lacunarity,
                       e->cfg.perlin_amplitude, seed, e->hf[i].data());
  float obs[16];
  bbo_reset(e->env[i], fieldOf(e, i), obs);
  e->ep_ret[i] = 0.f; e->ep_len[i] = 0;
  writeObs(io, i, obs);
  writeImages(e, io, i);
}
int checkIo(bb_engine* e, const bb_io* io) {
  if (!io || !io->orientation || !io->angular_vel || !io->vel || !io->motor_state || !io->actions || !io->rel_image_ts || !io->reward ||
      !io->terminated || !io->failure || !io->pos2d)
    return fail(e, BB_ERR_INVALID, "bb_io: required output pointer is NULL");
  if (e->cfg.cameras && (!io->rgbd_0 || !io->rgbd_1)) return fail(e, BB_ERR_INVALID, "bb_io: rgbd_0/rgbd_1 required when cameras are enabled");
  return BB_OK;
}
const int PAIR2TYPE[13] = {0, 1, 2, 3, -1, 4, 5, 6, 7, 8, 9, 10, 11};
}  // namespace

extern "C" {

void bb_default_config(bb_config* c) {
  memset(c, 0, sizeof(*c));
  c->abi_version = BB_ABI_VERSION; c->num_envs = 1; c->precision = 64; c->terrain_type = BB_TERRAIN_PERLIN; c->terrain_seed = -1;
  c->perlin_scale = 25.f; c->perlin_octaves = 4; c->perlin_persistence = 0.2f; c->perlin_lacunarity = 2.f; c->perlin_amplitude = 1.f;
  c->hfield_zscale = 2.f; c->cameras = 1; c->im_h = 64; c->im_w = 64; c->camera_frame_rate = 90.f;
  c->max_ep_steps = 4000; c->max_allowed_tilt = 20.f; c->max_wheel_velocity = 10.f;
  c->reward_type = BB_REWARD_DIRECTIONAL; c->reward_scale = 0.01f; c->action_reg_coef = -0.0001f; c->survival_bonus = 0.02f;
  c->target_direction[1] = 1.f; c->distance_scale = 1.f; c->auto_reset = 1; c->perlin_table = -1;
}
const char* bb_last_error(const bb_engine* e) { return e ? e->err : g_err; }
int bb_num_envs(const bb_engine* e) { return e ? e->N : 0; }
int64_t bb_launch_count(const bb_engine* e) { return e ? e->calls : 0; }
const char* bb_build_info(void) { return "libballbot_cpu_ref abi 2 (fp64 oracle behind the engine ABI; test infrastructure)"; }

int bb_create(const bb_config* cfg, bb_engine** out) {
  if (!cfg || !out) { snprintf(g_err, sizeof(g_err), "bb_create: NULL argument"); return BB_ERR_INVALID; }
  *out = nullptr;
  if (cfg->abi_version != BB_ABI_VERSION || cfg->num_envs < 1 || cfg->precision != 64 || cfg->im_h < 1 || cfg->im_w < 1 || cfg->terrain_type < 0 ||
      cfg->terrain_type > 3 || cfg->reward_type < 0 || cfg->reward_type > 2 || cfg->camera_frame_rate <= 0.f || cfg->seed_stream < 0 || cfg->seed_stream > 1) {
    snprintf(g_err, sizeof(g_err), "bb_create: invalid config for the cpu_ref backend (abi %d, num_envs %d, precision %d)", cfg->abi_version, cfg->num_envs, cfg->precision);
    return BB_ERR_INVALID;
  }
  bb_engine* e = new bb_engine();
  e->cfg = *cfg; e->N = cfg->num_envs; e->calls = 0; e->err[0] = 0;
  const int N = e->N;
  bbo_config oc; bbo_default_config(&oc);
  oc.max_ep_steps = cfg->max_ep_steps; oc.max_allowed_tilt = cfg->max_allowed_tilt; oc.max_wheel_velocity = cfg->max_wheel_velocity;
  oc.camera_frame_rate = cfg->camera_frame_rate; oc.reward_scale = cfg->reward_scale; oc.action_reg_coef = cfg->action_reg_coef;
  oc.survival_bonus = cfg->survival_bonus; oc.target_dir[0] = cfg->target_direction[0]; oc.target_dir[1] = cfg->target_direction[1];
  oc.hfield_zscale = cfg->hfield_zscale; oc.cameras = cfg->cameras; oc.im_h = cfg->im_h; oc.im_w = cfg->im_w;
  oc.reward_type = cfg->reward_type == BB_REWARD_DISTANCE ? 1 : 0; oc.goal[0] = cfg->goal_position[0]; oc.goal[1] = cfg->goal_position[1];
  oc.distance_scale = cfg->distance_scale;
  for (int i = 0; i < N; i++) e->env.push_back(bbo_create(&oc));
  const int nf = cfg->terrain_type == BB_TERRAIN_FLAT ? 0 : (cfg->terrain_type == BB_TERRAIN_SHARED ? 1 : N);
  e->hf.assign(nf, std::vector<float>(HF, 0.f));
  e->episode.assign(N, 0); e->tseed.assign(N, 0); e->ep_len.assign(N, 0); e->ep_ret.assign(N, 0.f); e->rng.assign(5 * (size_t)N, 0ull);
  *out = e;
  return BB_OK;
}
int bb_destroy(bb_engine* e) {
  if (!e) return BB_OK;
  for (bbo_env* o : e->env) bbo_destroy(o);
  delete e;
  return BB_OK;
}

int bb_reset(bb_engine* e, const uint8_t* mask, const int32_t* seeds, const bb_io* io, void*) {
  if (!e) return BB_ERR_INVALID;
  int rc = checkIo(e, io); if (rc) return rc;
  for (int i = 0; i < e->N; i++) {
    if (mask && !mask[i]) continue;
    e->episode[i]++;
    resetEnv(e, io, i, seeds ? seeds[i] : drawSeed(e, i));
  }
  e->calls++;
  return BB_OK;
}
int bb_set_rng_state(bb_engine* e, const uint64_t* state, void*) {
  if (!e || !state) return BB_ERR_INVALID;
  if (e->cfg.seed_stream != 1) return fail(e, BB_ERR_STATE, "bb_set_rng_state: engine was created with seed_stream = 0 (counter-based terrain seeds)");
  for (size_t k = 0; k < 5 * (size_t)e->N; k++) e->rng[k] = state[k];
  return BB_OK;
}

int bb_step(bb_engine* e, const float* actions, const bb_io* io, void*) {
  if (!e || !actions) return e ? fail(e, BB_ERR_INVALID, "bb_step: actions is NULL") : BB_ERR_INVALID;
  int rc = checkIo(e, io); if (rc) return rc;
  for (int i = 0; i < e->N; i++) {
    float obs[16], r, info[4]; uint8_t term, flr;
    bbo_step(e->env[i], actions + 3 * i, obs, &r, &term, &flr, info);
    if (e->cfg.reward_type == BB_REWARD_EXTERNAL) r -= (obs[6] * e->cfg.target_direction[0] + obs[7] * e->cfg.target_direction[1]) * e->cfg.reward_scale;
    writeObs(io, i, obs);
    io->reward[i] = r; io->terminated[i] = term; io->failure[i] = flr; io->pos2d[2 * i] = info[0]; io->pos2d[2 * i + 1] = info[1];
    if (io->status) io->status[i] = 0;
    e->ep_ret[i] += r; e->ep_len[i] += 1;
    if (term) {
      if (io->terminal_obs) memcpy(io->terminal_obs + 16 * i, obs, sizeof(obs));
      if (io->episode_return) io->episode_return[i] = e->ep_ret[i];
      if (io->episode_length) io->episode_length[i] = e->ep_len[i];
    }
    if (term && e->cfg.auto_reset) { e->episode[i]++; resetEnv(e, io, i, drawSeed(e, i)); }
    else if (info[3] != 0.f) writeImages(e, io, i);
  }
  e->calls++;
  return BB_OK;
}
int bb_add_reward(bb_engine* e, const float* term, const bb_io* io, void*) {
  if (!e || !term) return BB_ERR_INVALID;
  int rc = checkIo(e, io); if (rc) return rc;
  for (int i = 0; i < e->N; i++) {
    const float add = term[i] * e->cfg.reward_scale;
    io->reward[i] += add;
    if (io->terminated[i]) { if (io->episode_return) io->episode_return[i] += add; } else e->ep_ret[i] += add;
  }
  return BB_OK;
}

int bb_set_state(bb_engine* e, const double* qpos, const double* qvel, const double* warm, void*) {
  if (!e) return BB_ERR_INVALID;
  for (int i = 0; i < e->N; i++) {
    double t; bbo_get_state(e->env[i], nullptr, nullptr, nullptr, &t);
    bbo_set_state(e->env[i], qpos ? qpos + 17 * i : nullptr, qvel ? qvel + 15 * i : nullptr, warm ? warm + 15 * i : nullptr, t);
  }
  return BB_OK;
}
int bb_get_state(bb_engine* e, double* qpos, double* qvel, double* warm, void*) {
  if (!e) return BB_ERR_INVALID;
  for (int i = 0; i < e->N; i++)
    bbo_get_state(e->env[i], qpos ? qpos + 17 * i : nullptr, qvel ? qvel + 15 * i : nullptr, warm ? warm + 15 * i : nullptr, nullptr);
  return BB_OK;
}
int bb_set_hfield(bb_engine* e, const int32_t* ids, int32_t n, const float* hf, void*) {
  if (!e || !ids || !hf || n < 0) return BB_ERR_INVALID;
  if (e->cfg.terrain_type == BB_TERRAIN_SHARED) {
    if (n != 1) return fail(e, BB_ERR_INVALID, "bb_set_hfield: a shared-terrain engine takes exactly one heightfield");
    memcpy(e->hf[0].data(), hf, sizeof(float) * HF);
    for (int i = 0; i < e->N; i++) bbo_set_hfield(e->env[i], e->hf[0].data());
    return BB_OK;
  }
  if (e->cfg.terrain_type != BB_TERRAIN_EXTERNAL) return fail(e, BB_ERR_INVALID, "bb_set_hfield: engine has no per-env heightfields (flat terrain or Perlin table); create it with BB_TERRAIN_EXTERNAL");
  for (int k = 0; k < n; k++) {
    if (ids[k] < 0 || ids[k] >= e->N) return BB_ERR_INVALID;
    memcpy(e->hf[ids[k]].data(), hf + (size_t)k * HF, sizeof(float) * HF);
    bbo_set_hfield(e->env[ids[k]], e->hf[ids[k]].data());
  }
  return BB_OK;
}
int bb_get_hfield(bb_engine* e, int32_t env, float* out, void*) {
  if (!e || !out || env < 0 || env >= e->N) return BB_ERR_INVALID;
  const float* f = fieldOf(e, env);
  if (f) memcpy(out, f, sizeof(float) * HF); else memset(out, 0, sizeof(float) * HF);
  return BB_OK;
}
int bb_get_terrain_seeds(bb_engine* e, int32_t* seeds, void*) {
  if (!e || !seeds) return BB_ERR_INVALID;
  for (int i = 0; i < e->N; i++) seeds[i] = e->tseed[i];
  return BB_OK;
}
int bb_perlin_terrain(bb_engine* e, const int32_t* seeds, int32_t n, float* out, void*) {
  if (!e || !seeds || !out || n < 1) return BB_ERR_INVALID;
  for (int k = 0; k < n; k++)
    bbo_perlin_terrain(BB_HFIELD_N, e->cfg.perlin_scale, e->cfg.perlin_octaves, e->cfg.perlin_persistence, e->cfg.perlin_lacunarity, e->cfg.perlin_amplitude,
                       seeds[k], out + (size_t)k * HF);
  return BB_OK;
}
int bb_perlin_grid(int32_t, int32_t n, float scale, int32_t octaves, float persistence, float lacunarity, float amplitude, const int32_t* seeds, int32_t nseeds, float* out) {
  if (n < 1 || nseeds < 1 || !seeds || !out || octaves < 1) return BB_ERR_INVALID;
  for (int k = 0; k < nseeds; k++) bbo_perlin_terrain(n, scale, octaves, persistence, lacunarity, amplitude, seeds[k], out + (size_t)k * n * n);
  return BB_OK;
}
int bb_snoise2_grid(int32_t, int32_t n, float scale, int32_t octaves, float persistence, float lacunarity, int32_t base, float* out) {
  if (n < 1 || !out || octaves < 1 || scale == 0.f) return BB_ERR_INVALID;
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) out[i * n + j] = bbo_snoise2((float)(i / (double)scale), (float)(j / (double)scale), octaves, persistence, lacunarity, base);
  return BB_OK;
}
int bb_render_depth(bb_engine* e, float* img0, float* img1, void*) {
  if (!e || !img0 || !img1) return BB_ERR_INVALID;
  const size_t npix = (size_t)e->cfg.im_h * e->cfg.im_w;
  for (int i = 0; i < e->N; i++) { bbo_forward(e->env[i], nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr); bbo_render_depth(e->env[i], 0, img0 + npix * i); bbo_render_depth(e->env[i], 1, img1 + npix * i); }
  return BB_OK;
}
int bb_probe_forward(bb_engine* e, int32_t env, const double* ctrl3, double* out, double* contacts, void*) {
  if (!e || env < 0 || env >= e->N || !ctrl3 || !out || !contacts) return BB_ERR_INVALID;
  double qas[15], qacc[15], M[225]; int ncon = 0, niter = 0;
  bbo_forward(e->env[env], ctrl3, M, nullptr, qas, qacc, &ncon, &niter);
  for (int k = 0; k < 15; k++) {
    out[k] = qacc[k]; out[15 + k] = qas[k];
    double s = 0; for (int j = 0; j < 15; j++) s += M[k * 15 + j] * qas[j];
    out[30 + k] = s;                                                   // qfrc_smooth = M qacc_smooth
  }
  out[45] = ncon; out[46] = niter;
  double dist[BBO_MAXCON], pos[3 * BBO_MAXCON], frame[9 * BBO_MAXCON]; int pair[BBO_MAXCON];
  const int n = bbo_get_contacts(e->env[env], BBO_MAXCON, dist, pos, frame, pair);
  for (int c = 0; c < n && c < BB_PROBE_MAXCON; c++) {
    double* o = contacts + BB_CONTACT_STRIDE * c;
    o[0] = PAIR2TYPE[pair[c]]; o[1] = dist[c];
    for (int j = 0; j < 3; j++) o[2 + j] = pos[3 * c + j];
    for (int j = 0; j < 9; j++) o[5 + j] = frame[9 * c + j];
  }
  return BB_OK;
}
int bb_get_contacts(bb_engine* e, int32_t env, double* contacts, int32_t* ncon, void* s) {
  if (!e || !ncon) return BB_ERR_INVALID;
  const double z[3] = {0, 0, 0}; double out[64];
  const int rc = bb_probe_forward(e, env, z, out, contacts, s); if (rc) return rc;
  *ncon = (int)out[45];
  return BB_OK;
}
int bb_model_constants(double* dA12, double* meaninertia, double* masses3) {
  double mass[8], invw[16], mi;
  bbo_get_model(mass, nullptr, nullptr, invw, &mi);
  const double w[8] = {invw[0], invw[2], invw[4], invw[6], invw[8], invw[10], invw[12], invw[14]};   // translational invweight0 per body
  if (dA12) {
    const double d[12] = {w[7] + w[4], w[7] + w[5], w[7] + w[6], w[7], w[2], w[3], w[4], w[5], w[6], w[7] + w[1], w[7] + w[2], w[7] + w[3]};
    memcpy(dA12, d, sizeof(d));
  }
  if (meaninertia) *meaninertia = mi;
  if (masses3) { masses3[0] = mass[1] + mass[2] + mass[3]; masses3[1] = mass[4]; masses3[2] = mass[7]; }
  return BB_OK;
}

/* GPU-only entry points */
int bb_host_buffers(bb_engine* e, float**, bb_host_io*) { return e ? fail(e, BB_ERR_STATE, "not available in the cpu_ref backend") : BB_ERR_INVALID; }
int bb_step_host(bb_engine* e, const float*, const bb_host_io*) { return e ? fail(e, BB_ERR_STATE, "not available in the cpu_ref backend") : BB_ERR_INVALID; }
int bb_reset_host(bb_engine* e, const uint8_t*, const bb_host_io*) { return e ? fail(e, BB_ERR_STATE, "not available in the cpu_ref backend") : BB_ERR_INVALID; }
int bb_profile_begin(bb_engine* e, int32_t) { return e ? fail(e, BB_ERR_STATE, "not available in the cpu_ref backend") : BB_ERR_INVALID; }
int bb_profile_end(bb_engine* e, double*, int32_t*) { return e ? fail(e, BB_ERR_STATE, "not available in the cpu_ref backend") : BB_ERR_INVALID; }
int bb_gae(const float*, const float*, const uint8_t*, int32_t, int32_t, float, float, float*, float*, void*) { return BB_ERR_STATE; }
int bb_adamw_step(float*, const float*, float*, float*, int32_t, float, float, float, float, float, float, float, double*, double*, void*) { return BB_ERR_STATE; }
int bb_fp64_peak(int32_t, double*) { return BB_ERR_STATE; }

}  // extern "C"
