// bb_engine.cu -- CUDA kernels (sm_100a) and the C ABI (include/ballbot_b200.h) of the batched ballbot engine.
//
// Kernels (all fixed-grid, device-side work lists => no host sync or allocation inside bb_step, CUDA-graph capturable):
//   k_begin_step, k_order   counting sort of the envs by the solver work of their previous step (scheduling only)
//   k_stage<T>(s) x5        split-phase step, one 16-lane group per env: fold RK stage s-1, advance the stage state, smooth
//                           dynamics, contact generation + constraint rows, qacc_smooth; s = 0 loads the state and maps the
//                           action, s = 4 applies the RK4 update and writes proprio obs, reward, termination, episode
//                           statistics and the reset / refresh work lists                  (ballbot_env.py:854-1036)
//   k_newton_p<T, mode>(s) x4  elliptic-cone Newton solve of RK stage s for the envs that have contacts (mj_fwdConstraint): persistent
//                           16-lane groups drawing envs from the stage's work list; mode = solver_mode (0: the reference's exact
//                           line search, 1: strong Wolfe + cone-apex candidates); k_newton<T, mode> is the one-launch-slot-per-env variant
//   k_step_warp<T>          the same arithmetic fused into one launch (step_kernel = 2, cross-check)
//   k_step<T>               one thread per env (step_kernel = 1, cross-check; shares bb_core.cuh with the CPU test harness)
//   k_terrain               simplex-fBm heightfields: the 10,000-field table at bb_create, or per reset            (terrain/perlin.py:8-74)
//   k_reset<T>              spawn-height window max, state reset, reset observation       (ballbot_env.py:528-565,612-634)
//   k_depth<T>              2 x HxW depth ray-cast per refreshing env (hfield DDA + analytic prims) (sensors/rgbd.py:46-82)
// State is one record per env (qpos17 qvel15 qacc_warmstart15, stride 48) that a 16-lane group loads / stores coalesced.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/ballbot_b200.h"
#include "bb_model.h"
#include "bb_group.cuh"

using namespace bb;

namespace {

#ifndef BB_WPB
#define BB_WPB 1          // warps per CTA of the step kernel (2 envs per warp); 1 avoids waiting for the slowest warp of a CTA
#endif
#ifndef BB_WPB_STAGE
#define BB_WPB_STAGE 6    // warps per CTA of k_stage (straight-line phase code: CTA-synchronised phases share instruction fetch).
                          // Measured at 65,536 envs (perlin, step kernels per step): 2 -> 8.84, 3 -> 8.32, 4 -> 8.22, 6 -> 7.82, 12 -> 7.98 ms:
                          // two CTAs of six warps per SM keep one phase's code resident for half of the SM's warps
#endif
#ifndef BB_WARP_MINBLOCKS
#define BB_WARP_MINBLOCKS (12 / BB_WPB)    // fp64: 12 warps per SM (shared memory: 18.3 KB per warp; 168 registers)
#endif
#ifndef BB_WARP_MINBLOCKS32
#define BB_WARP_MINBLOCKS32 (24 / BB_WPB)  // fp32 instantiation: half the shared memory and registers => 24 warps per SM (+14 % measured)
#endif
constexpr int SST = 48;            // per-env stride of the state array  T[N][SST] (one coalesced 384/192-byte record per env)
constexpr int CST = 20;            // per-env stride of the camera configuration array T[N][CST]
constexpr int HF_CELLS = HN * HN;

__constant__ ModelConst<double> c_mc64;
__constant__ ModelConst<float> c_mc32;
template <typename T> __device__ __forceinline__ const ModelConst<T>& cmc();
template <> __device__ __forceinline__ const ModelConst<double>& cmc<double>() { return c_mc64; }
template <> __device__ __forceinline__ const ModelConst<float>& cmc<float>() { return c_mc32; }

struct EnvParams {
  int N; long long env_offset;
  int cameras, im_h, im_w, cam_period;
  int max_ep_steps; float max_tilt, max_wheel_vel;
  int reward_type; float reward_scale, action_reg, survival, tdir[2], goal[2], dist_scale;
  float zscale; int terrain_type, terrain_seed; unsigned long long seed;
  int auto_reset, hf_mode, solver_mode, seed_stream;   // hf_mode: 0 one shared field, 1 one field per env, 2 table of Perlin fields indexed by seed
  float pscale, ppers, plac, pamp; int poct;
};
enum { HF_SHARED = 0, HF_PER_ENV = 1, HF_TABLE = 2 };
struct DevState {
  void* st;        // T[N][SST]
  void* camq;      // T[N][CST]  configuration the cameras see (last RK stage / reset state)
  void* gscr;      // T[N][bbg::GSCR] overflow scratch of the group kernel (contact records beyond shared memory)
  void* rk;        // T[N][bbg::RKN]  split-phase step: RK4 bookkeeping between the stage kernels
  void* ctx;       // T[N][bbg::CTXN] split-phase step: solver input (M, qfrc_smooth, qacc_smooth, contact records)
  int* meta;       // int[N][4]       split-phase step: ncon, nw, Newton iterations, flags
  int* step_count; int* cam_steps; unsigned* episode; int* tseed;
  float* hfield;   // [N][HF_CELLS] (HF_PER_ENV), [HF_CELLS] (HF_SHARED) or [table_n][HF_CELLS] (HF_TABLE)
  float* hmip;     // block maxima (8 x 8 cells, MIP_CELLS per field) of every heightfield in `hfield`, same indexing
  unsigned long long* rng;   // [N][5] numpy PCG64 state per env (seed_stream = 1): state hi/lo, inc hi/lo, has_uint32 | uinteger << 32
  float* ptab;     // [2][293] circle coordinates of the tiled simplex noise (sin, cos) per grid index
  float* ep_ret; int* ep_len;
  int* counters;   // [0] reset-list length, [1] refresh-list length, [2] depth work-unit cursor, [8 + s] / [12 + s] solver work list of stage s: length / cursor
  int* reset_list; int* refresh_list;
  // work-sorted scheduling of the group step kernel: key = Newton iterations of the env's previous step (capped)
  int* work;       // [N] key of the previous step
  int* order;      // [N] env indices sorted by descending key (heavy envs first, similar envs share a warp)
  int* bins;       // [WORK_BINS] histogram of work[] (accumulated by the step kernel) + [WORK_BINS] scatter cursors
  int* alist;      // [N] envs with contacts in the current RK stage, appended by k_stage (counters[8 + stage]), consumed by the
                   // persistent solver kernel through the cursor counters[12 + stage]
};
constexpr int WORK_BINS = 64;

__device__ __forceinline__ unsigned long long splitmix(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
// numpy's PCG64 (XSL-RR 128/64) + Generator.integers(0, 10000): 32-bit halves of the 64-bit outputs are consumed low half first
// (has_uint32 / uinteger buffering of pcg64_next32), bounded by Lemire's multiply-and-reject (buffered_bounded_lemire_uint32)
__device__ __forceinline__ unsigned pcg64Next32(unsigned long long* st) {
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
__device__ __forceinline__ int pcg64Integers10000(unsigned long long* st) {
  const unsigned rng = 9999u, excl = 10000u;
  unsigned long long m = (unsigned long long)pcg64Next32(st) * excl;
  unsigned left = (unsigned)m;
  if (left < excl) {
    const unsigned thr = (0xFFFFFFFFu - rng) % excl;
    while (left < thr) { m = (unsigned long long)pcg64Next32(st) * excl; left = (unsigned)m; }
  }
  return (int)(m >> 32);
}
// r_seed = self._np_random.integers(0, 10000) (ballbot_env.py:506): either the numpy-compatible per-env stream or a counter-based
// hash with the same U{0..9999} law per (env, episode)
__device__ __forceinline__ int drawTerrainSeed(const EnvParams& p, const DevState& d, int env, unsigned episode) {
  if (p.terrain_seed >= 0) return p.terrain_seed;
  if (p.seed_stream == 1) return pcg64Integers10000(d.rng + 5 * (size_t)env);
  unsigned long long h = splitmix(p.seed ^ splitmix((unsigned long long)(p.env_offset + env) * 0x100000001B3ull + episode));
  return (int)(h % 10000ull);
}
// heightfield of env i (field index: the env itself, its terrain seed in the table, or the one shared field)
__device__ __forceinline__ size_t hfIndex(const EnvParams& p, const DevState& d, int i) {
  if (p.hf_mode == HF_PER_ENV) return (size_t)i;
  if (p.hf_mode == HF_TABLE) return (size_t)(p.terrain_seed >= 0 ? 0 : d.tseed[i]);
  return 0;
}
__device__ __forceinline__ const float* hfOf(const EnvParams& p, const DevState& d, int i) { return d.hfield + hfIndex(p, d, i) * HF_CELLS; }
__device__ __forceinline__ const float* mipOf(const EnvParams& p, const DevState& d, int i) { return d.hmip + hfIndex(p, d, i) * MIP_CELLS; }

// --------------------------------------------------------------------------------------------- step
template <typename T>
__global__ void __launch_bounds__(64) k_step(EnvParams p, DevState d, const float* __restrict__ actions, bb_io io) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.N) return;
  T* st = (T*)d.st;
  T qpos[NQ], qvel[NV], warm[NV];
#pragma unroll
  for (int k = 0; k < NQ; k++) qpos[k] = st[(size_t)i * SST + k];
#pragma unroll
  for (int k = 0; k < NV; k++) { qvel[k] = st[(size_t)i * SST + NQ + k]; warm[k] = st[(size_t)i * SST + NQ + NV + k]; }
  const float a0 = actions[3 * i], a1 = actions[3 * i + 1], a2 = actions[3 * i + 2];
  T ctrl[3];
  {  // ballbot_env.py:903-907
    const float av[3] = {a0, a1, a2};
    for (int k = 0; k < 3; k++) { T u = (T)av[k] * (T)p.max_wheel_vel; u = u > (T)p.max_wheel_vel ? (T)p.max_wheel_vel : (u < -(T)p.max_wheel_vel ? -(T)p.max_wheel_vel : u); ctrl[k] = -u; }
  }
  bool bad = false;  // mj_checkPos / mj_checkVel
  for (int k = 0; k < NQ; k++) bad |= !(babs(qpos[k]) < (T)1e10);
  for (int k = 0; k < NV; k++) bad |= !(babs(qvel[k]) < (T)1e10);
  KinOut<T> kin;
  T qlast[NQ];
  int status = 0;
  if (!bad) {
    Scratch<T> s;
    const float* hf = hfOf(p, d, i);
    rk4Step(cmc<T>(), qpos, qvel, warm, ctrl, hf, (T)p.zscale, s, &kin, qlast, p.solver_mode != 0);
    for (int k = 0; k < NQ; k++) bad |= !(babs(qpos[k]) < (T)1e10);
    for (int k = 0; k < NV; k++) bad |= !(babs(qvel[k]) < (T)1e10);
    status = kin.ncon << 8;
  }
  if (bad) {  // numerical failure: flag, report a terminated+failed step with a zero observation; the env is reset
    status |= 1;
    kin.quatB[0] = 1; kin.quatB[1] = kin.quatB[2] = kin.quatB[3] = 0;
    for (int k = 0; k < 3; k++) { kin.cvel_ang[k] = 0; kin.cvel_lin[k] = 0; kin.posB[k] = 0; }
    for (int k = 0; k < NV; k++) qvel[k] = 0;
  }
#pragma unroll
  for (int k = 0; k < NQ; k++) st[(size_t)i * SST + k] = qpos[k];
#pragma unroll
  for (int k = 0; k < NV; k++) { st[(size_t)i * SST + NQ + k] = qvel[k]; st[(size_t)i * SST + NQ + NV + k] = warm[k]; }

  // ---- observation (ballbot_env.py:772-827)
  float ob[16];
  proprioObs(kin, qvel, (T)p.max_wheel_vel, ob, ob + 3, ob + 6, ob + 9);
  ob[12] = a0; ob[13] = a1; ob[14] = a2;
  int cs = d.cam_steps[i] + 1;
  bool refresh = false;
  if (p.cameras && cs >= p.cam_period) { refresh = true; cs = 0; }
  ob[15] = p.cameras ? (float)((double)cs * 0.002) : 0.f;
  // ---- reward (ballbot_env.py:929-937), float32 arithmetic as NumPy >= 2
  float r = 0.f;
  if (p.reward_type == BB_REWARD_DIRECTIONAL) r = (ob[6] * p.tdir[0] + ob[7] * p.tdir[1]) * p.reward_scale;
  else if (p.reward_type == BB_REWARD_DISTANCE) {
    const float dx = p.goal[0] - (float)kin.posB[0], dy = p.goal[1] - (float)kin.posB[1];
    r = (-p.dist_scale * sqrtf(dx * dx + dy * dy)) * p.reward_scale;
  }
  const float nrm = sqrtf(a0 * a0 + a1 * a1 + a2 * a2);
  r += p.action_reg * (nrm * nrm);
  // ---- termination (ballbot_env.py:977-1020)
  const int sc = d.step_count[i] + 1;
  bool term = sc >= p.max_ep_steps, fail = false;
  const double tilt = tiltDegrees(ob);
  if (tilt > (double)p.max_tilt || bad) { fail = true; term = true; } else r += p.survival;
  const float eret = d.ep_ret[i] + r; const int elen = d.ep_len[i] + 1;
  // ---- outputs
  for (int k = 0; k < 3; k++) {
    io.orientation[3 * i + k] = ob[k]; io.angular_vel[3 * i + k] = ob[3 + k]; io.vel[3 * i + k] = ob[6 + k];
    io.motor_state[3 * i + k] = ob[9 + k]; io.actions[3 * i + k] = ob[12 + k];
  }
  io.rel_image_ts[i] = ob[15];
  io.reward[i] = r; io.terminated[i] = term; io.failure[i] = fail;
  io.pos2d[2 * i] = (float)kin.posB[0]; io.pos2d[2 * i + 1] = (float)kin.posB[1];
  if (io.status) io.status[i] = status;
  if (term) {
    if (io.terminal_obs) for (int k = 0; k < 16; k++) io.terminal_obs[16 * i + k] = ob[k];
    if (io.episode_return) io.episode_return[i] = eret;
    if (io.episode_length) io.episode_length[i] = elen;
  }
  d.step_count[i] = sc; d.cam_steps[i] = cs; d.ep_ret[i] = eret; d.ep_len[i] = elen;
  if (term && p.auto_reset) {
    const unsigned ep = d.episode[i] + 1; d.episode[i] = ep;
    d.tseed[i] = drawTerrainSeed(p, d, i, ep);
    d.reset_list[atomicAdd(&d.counters[0], 1)] = i;
  } else if (refresh) {
    T* cq = (T*)d.camq;
    for (int k = 0; k < NQ; k++) cq[(size_t)i * CST + k] = qlast[k];
    d.refresh_list[atomicAdd(&d.counters[1], 1)] = i;
  }
}

// --------------------------------------------------------------------------------------------- step, lane group per env
// One 16-lane group per env (two envs per warp, bb_group.cuh); solver state in dynamic shared memory (bbg::GS<T> per env).
// stepLoad / stepFinish are shared by the fused kernel (k_step_warp) and the split-phase kernels (k_stage / k_newton).

// state record -> S.xq / S.xv / warm, action -> S.ctrl (ballbot_env.py:903-907), structural zeros of M.  Returns "state is not finite".
template <typename T>
__device__ __forceinline__ bool stepLoad(const EnvParams& p, const DevState& d, const float* __restrict__ actions, int i, bbg::GS<T>& S,
                                         const bbg::Ln& L, T& warm) {
  const T* st = (const T*)d.st + (size_t)i * SST;
  bool bad = false;
  for (int k = L.gl; k < NQ + NV; k += bbg::G) {   // one coalesced record per env: qpos[17] qvel[15] warm[15]
    const T v = st[k];
    bad |= !(babs(v) < (T)1e10);
    if (k < NQ) S.xq[k] = v; else S.xv[k - NQ] = v;
  }
  warm = st[NQ + NV + L.gi];
  if (L.gl == NV) S.xv[NV] = 0;
  for (int k = L.gl; k < bbg::MSZ; k += bbg::G) S.M[k] = 0;   // structural zeros of the mass matrix (never written again)
  bad = __any_sync(L.mask, bad);
  if (L.gl < 3) {
    T u = (T)actions[3 * i + L.gl] * (T)p.max_wheel_vel; u = u > (T)p.max_wheel_vel ? (T)p.max_wheel_vel : (u < -(T)p.max_wheel_vel ? -(T)p.max_wheel_vel : u);
    S.ctrl[L.gl] = -u;
  }
  __syncwarp(L.mask);
  return bad;
}
__device__ __forceinline__ bool cameraRefresh(const EnvParams& p, const DevState& d, int i, int& cs) {
  cs = d.cam_steps[i] + 1;
  if (p.cameras && cs >= p.cam_period) { cs = 0; return true; }
  return false;
}
// in: S.xq / S.xv = new state, S.kin = last-stage kinematics, warm.  Writes the state record, observation, reward,
// termination, Monitor accumulators, work key and the reset / refresh work lists.
template <typename T>
__device__ __forceinline__ void stepFinish(const EnvParams& p, const DevState& d, const float* __restrict__ actions, const bb_io& io, int i, bbg::GS<T>& S,
                                           const bbg::Ln& L, T warm, bool bad, int ncmax, int nit) {
  KinOut<T> kin;
  const int nev = nit >> 12; nit &= 0xfff;
  int status = (ncmax << 8) | (nit << 16);   // bit 0: numerical failure, bits 8-15: max contacts of the stages, bits 16+: Newton iterations
  // work of the solver ~ Newton iterations x (Hessian + factorisation + substitutions) + line-search evaluations, about 15 : 1 per
  // ncu's instruction counts; 64 bins cover ~65 iterations (with solver mode 1 the evaluations alone no longer separate the envs)
#ifndef BB_KEY_ITER_WEIGHT
#define BB_KEY_ITER_WEIGHT 15
#endif
  const int kv = nev ? (BB_KEY_ITER_WEIGHT * nit + nev + 15) >> 4 : nit;     // split-phase path: weighted work; fused path: iterations
  const int key = bad ? 0 : (kv < WORK_BINS ? kv : WORK_BINS - 1);
  if (!bad) {
    bool b2 = (L.gl < NV && !(babs(S.xv[L.gl]) < (T)1e10));
    for (int k = L.gl; k < NQ; k += bbg::G) b2 |= !(babs(S.xq[k]) < (T)1e10);
    bad = __any_sync(L.mask, b2);
#pragma unroll
    for (int k = 0; k < 4; k++) kin.quatB[k] = S.kin[k];
#pragma unroll
    for (int k = 0; k < 3; k++) { kin.cvel_ang[k] = S.kin[4 + k]; kin.cvel_lin[k] = S.kin[7 + k]; kin.posB[k] = S.kin[10 + k]; }
  } else status = 0;
  if (bad) {
    status |= 1;
    kin.quatB[0] = 1; kin.quatB[1] = kin.quatB[2] = kin.quatB[3] = 0;
    for (int k = 0; k < 3; k++) { kin.cvel_ang[k] = 0; kin.cvel_lin[k] = 0; kin.posB[k] = 0; }
    if (L.gl < NV) S.xv[L.gl] = 0;
    __syncwarp(L.mask);
  }
  int cs; const bool refresh = cameraRefresh(p, d, i, cs);
  T* st = (T*)d.st + (size_t)i * SST;
  // write the state record back (coalesced)
  for (int k = L.gl; k < NQ + NV; k += bbg::G) st[k] = k < NQ ? S.xq[k] : S.xv[k - NQ];
  if (L.gl < NV) st[NQ + NV + L.gl] = warm;
  // ---- observation / reward / termination: every lane evaluates the few scalars, lane 0 (or a few lanes) store
  const float a0 = actions[3 * i], a1 = actions[3 * i + 1], a2 = actions[3 * i + 2];
  float ob[16];
  proprioObs(kin, (const T*)S.xv, (T)p.max_wheel_vel, ob, ob + 3, ob + 6, ob + 9);
  ob[12] = a0; ob[13] = a1; ob[14] = a2;
  ob[15] = p.cameras ? (float)((double)cs * 0.002) : 0.f;
  float r = 0.f;
  if (p.reward_type == BB_REWARD_DIRECTIONAL) r = (ob[6] * p.tdir[0] + ob[7] * p.tdir[1]) * p.reward_scale;
  else if (p.reward_type == BB_REWARD_DISTANCE) {
    const float dx = p.goal[0] - (float)kin.posB[0], dy = p.goal[1] - (float)kin.posB[1];
    r = (-p.dist_scale * sqrtf(dx * dx + dy * dy)) * p.reward_scale;
  }
  const float nrm = sqrtf(a0 * a0 + a1 * a1 + a2 * a2);
  r += p.action_reg * (nrm * nrm);
  const int sc = d.step_count[i] + 1;
  bool term = sc >= p.max_ep_steps, fail = false;
  const double tilt = tiltDegrees(ob);
  if (tilt > (double)p.max_tilt || bad) { fail = true; term = true; } else r += p.survival;
  const float eret = d.ep_ret[i] + r; const int elen = d.ep_len[i] + 1;
  __syncwarp(L.mask);
  if (L.gl < 15) {   // obs block: orientation, angular_vel, vel, motor_state, actions (3 each)
    float* dst = L.gl < 3 ? io.orientation : (L.gl < 6 ? io.angular_vel : (L.gl < 9 ? io.vel : (L.gl < 12 ? io.motor_state : io.actions)));
    float v = ob[0];
#pragma unroll
    for (int k = 1; k < 15; k++) v = L.gl == k ? ob[k] : v;
    dst[3 * i + L.gl % 3] = v;
  }
  if (term && io.terminal_obs && L.gl < 16) {
    float v = ob[0];
#pragma unroll
    for (int k = 1; k < 16; k++) v = L.gl == k ? ob[k] : v;
    io.terminal_obs[16 * i + L.gl] = v;
  }
  if (L.gl == 0) {
    d.work[i] = key; atomicAdd(&d.bins[key], 1);
    io.rel_image_ts[i] = ob[15];
    io.reward[i] = r; io.terminated[i] = term; io.failure[i] = fail;
    io.pos2d[2 * i] = (float)kin.posB[0]; io.pos2d[2 * i + 1] = (float)kin.posB[1];
    if (io.status) io.status[i] = status;
    if (term) { if (io.episode_return) io.episode_return[i] = eret; if (io.episode_length) io.episode_length[i] = elen; }
    d.step_count[i] = sc; d.cam_steps[i] = cs; d.ep_ret[i] = eret; d.ep_len[i] = elen;
    if (term && p.auto_reset) {
      const unsigned ep = d.episode[i] + 1; d.episode[i] = ep;
      d.tseed[i] = drawTerrainSeed(p, d, i, ep);
      d.reset_list[atomicAdd(&d.counters[0], 1)] = i;
    } else if (refresh) d.refresh_list[atomicAdd(&d.counters[1], 1)] = i;
  }
}

// fused variant: the whole RK4 step of an env inside one launch (step_kernel = 2; cross-check of the split-phase path)
template <typename T>
__global__ void __launch_bounds__(32 * BB_WPB, sizeof(T) == 4 ? BB_WARP_MINBLOCKS32 : BB_WARP_MINBLOCKS) k_step_warp(EnvParams p, DevState d, const float* __restrict__ actions, bb_io io) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const bbg::Ln L = bbg::makeLn();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane / bbg::G;
  const int slot = (blockIdx.x * (blockDim.x >> 5) + warp) * bbg::EPW + grp;
  if (slot >= p.N) return;
  const int i = d.order[slot];
  bbg::GS<T>& S = reinterpret_cast<bbg::GS<T>*>(smem_raw)[warp * bbg::EPW + grp];
  T warm;
  const bool bad = stepLoad(p, d, actions, i, S, L, warm);
  int ncmax = 0, nit = 0;
  if (!bad) {
    int cs; const bool refresh = cameraRefresh(p, d, i, cs);
    const float* hf = hfOf(p, d, i);
    T* gs = (T*)d.gscr + (size_t)i * bbg::GSCR;
    T* cq = refresh ? (T*)d.camq + (size_t)i * CST : nullptr;
    bbg::gRk4(cmc<T>(), S, hf, mipOf(p, d, i), (T)p.zscale, gs, cq, L, warm, p.solver_mode != 0, ncmax, nit);
  }
  stepFinish(p, d, actions, io, i, S, L, warm, bad, ncmax, nit);
}

// split-phase variant (default), see bb_group.cuh "split-phase step".
// k_stage<T>(stage): stage 0 loads the state; stages 1..3 fold the previous stage into the RK4 sums and advance the
// stage state; every stage < 4 then runs mj_forward up to the solver and parks the solver input; stage 4 applies the RK4
// update and finishes the step (observation, reward, termination, work lists).
#ifndef BB_STAGE_MINBLOCKS
#define BB_STAGE_MINBLOCKS (BB_WARP_MINBLOCKS * BB_WPB / BB_WPB_STAGE)
#endif
template <typename T>
__global__ void __launch_bounds__(32 * BB_WPB_STAGE, sizeof(T) == 4 ? BB_WARP_MINBLOCKS32 * BB_WPB / BB_WPB_STAGE : BB_STAGE_MINBLOCKS)
k_stage(EnvParams p, DevState d, const float* __restrict__ actions, bb_io io, int stage) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const bbg::Ln L = bbg::makeLn();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane / bbg::G;
  const int slot = (blockIdx.x * (blockDim.x >> 5) + warp) * bbg::EPW + grp;
  const bool live = slot < p.N;              // threads beyond the last env only take part in the CTA barriers
  // work-sorted order (the same one k_newton uses): the envs of a CTA enter the barrier-synchronised phases together, so
  // free-falling envs share CTAs with free-falling envs and landed ones with landed ones
  const int i = d.order[live ? slot : p.N - 1];
  bool skip = !live;
  bbg::GS<T>& S = reinterpret_cast<bbg::GS<T>*>(smem_raw)[warp * bbg::EPW + grp];
  const ModelConst<T>& mc = cmc<T>();
  T* rk = (T*)d.rk + (size_t)i * bbg::RKN;
  int* meta = d.meta + 4 * (size_t)i;
  const bool dof = L.gl < NV;
  const T h = mc.timestep;
  T v0 = 0, sumv = 0, suma = 0, xv = 0, warm = 0;
  int ncmax = 0;
  if (stage == 0) {
    const bool bad = stepLoad(p, d, actions, i, S, L, warm);
    if (live && L.gl == 0) { meta[bbg::META_NCON] = 0; meta[bbg::META_NW] = 0; meta[bbg::META_NIT] = 0; meta[bbg::META_FLAGS] = bad ? 1 : 0; }
    skip |= bad;
    if (!skip) {
      if (L.gl == 0) bb::normalizeQuats(S.xq);
      __syncwarp(L.mask);
      for (int k = L.gl; k < NQ; k += bbg::G) rk[bbg::RK_Q0 + k] = S.xq[k];
      v0 = S.xv[L.gi]; xv = v0;
      if (dof) { rk[bbg::RK_V0 + L.gl] = v0; rk[bbg::RK_WARM + L.gl] = warm; }
      if (L.gl < 3) rk[bbg::RK_CTRL + L.gl] = S.ctrl[L.gl];
    }
  } else {
    const int flags = meta[bbg::META_FLAGS];
    ncmax = flags >> 8;
    if (flags & 1) {   // non-finite state: nothing was integrated; stage 4 reports the failure
      if (stage == 4 && live) { T w0 = 0; stepLoad(p, d, actions, i, S, L, w0); stepFinish(p, d, actions, io, i, S, L, w0, true, 0, 0); }
      skip = true;
    }
    if (!skip) {
      for (int k = L.gl; k < NQ; k += bbg::G) S.q0[k] = rk[bbg::RK_Q0 + k];
      for (int k = L.gl; k < bbg::MSZ; k += bbg::G) S.M[k] = 0;
      if (L.gl < 3) S.ctrl[L.gl] = rk[bbg::RK_CTRL + L.gl];
      v0 = rk[bbg::RK_V0 + L.gi];
      const T xvp = rk[bbg::RK_XV + L.gi], qacc = rk[bbg::RK_QACC + L.gi];
      if (stage > 1) { sumv = rk[bbg::RK_SUMV + L.gi]; suma = rk[bbg::RK_SUMA + L.gi]; }
      const int sp = stage - 1;                                   // the stage that has just been solved
      const T bw = (sp == 0 || sp == 3) ? (T)(1.0 / 6.0) : (T)(1.0 / 3.0);
      sumv += bw * xvp; suma += bw * qacc;
      const T ha = stage == 4 ? h : (sp == 2 ? h : (T)0.5 * h);
      if (bbg::G == 16 || L.gl < 16) S.vb[0][L.gl] = dof ? (stage == 4 ? sumv : xvp) : (T)0;
      __syncwarp(L.mask);
      if (L.gl == 0) bbg::gIntegrate(S.xq, S.q0, (const T*)S.vb[0], ha);
      xv = v0 + ha * (stage == 4 ? suma : qacc);
      if (dof) S.xv[L.gl] = xv;
      if (L.gl == NV) S.xv[NV] = 0;
      __syncwarp(L.mask);
      if (stage == 4) {
        if (L.gl < 13) S.kin[L.gl] = rk[bbg::RK_KIN + L.gl];
        __syncwarp(L.mask);
        stepFinish(p, d, actions, io, i, S, L, qacc, false, ncmax, meta[bbg::META_NIT]);   // qacc_warmstart := last-stage qacc
      }
    }
  }
  if (stage == 4) return;
  // ---- mj_forward up to the solver (CTA-synchronised phases)
  const float* hf = hfOf(p, d, i);
  if (!skip && p.hf_mode != HF_SHARED && L.gl < 8) {
    // the ~7 x 7 heights under the ball are first touched by the collision phase, ~2 k instructions from here: start the
    // DRAM/L2 fetch of the 8 row segments now (the ball moves less than a cell per stage, so the stage configuration is enough)
    const T gsc = (T)(HN - 1) / ((T)2 * mc.hx);
    int c0 = (int)((S.xq[10] - mc.ball_r + mc.hx) * gsc), r0 = (int)((S.xq[11] - mc.ball_r + mc.hx) * gsc) + L.gl;
    c0 = c0 < 0 ? 0 : (c0 > HN - 8 ? HN - 8 : c0); r0 = r0 < 0 ? 0 : (r0 > HN - 1 ? HN - 1 : r0);
    const float* row = hf + r0 * HN + c0;
    asm volatile("prefetch.global.L1 [%0];" ::"l"(row));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(row + 7));
  }
  T* gs = (T*)d.gscr + (size_t)i * bbg::GSCR;
  int nw, nd; T qfs, qas;
  const int ncon = bbg::gForwardPre(mc, S, hf, mipOf(p, d, i), (T)p.zscale, gs, L, stage == 3, nw, nd, qfs, qas, skip, BB_WPB_STAGE > 1);
  if (skip) return;
  if (dof) {
    rk[bbg::RK_XV + L.gl] = xv;
    if (stage > 0) { rk[bbg::RK_SUMV + L.gl] = sumv; rk[bbg::RK_SUMA + L.gl] = suma; }
    if (ncon == 0) {
      rk[bbg::RK_QACC + L.gl] = qas;
      if (p.solver_mode != 0 && stage < 3) rk[bbg::RK_WARM + L.gl] = qas;
    }
  }
  if (stage == 3) {
    if (L.gl < 13) rk[bbg::RK_KIN + L.gl] = S.kin[L.gl];
    int cs;
    if (cameraRefresh(p, d, i, cs)) { T* cq = (T*)d.camq + (size_t)i * CST; for (int k = L.gl; k < NQ; k += bbg::G) cq[k] = S.xq[k]; }
  }
  if (L.gl == 0) {
    meta[bbg::META_NCON] = ncon; meta[bbg::META_NW] = nw | (nd << 8);
    if (ncon > ncmax) meta[bbg::META_FLAGS] = ncon << 8;
    if (ncon > 0) d.alist[atomicAdd(&d.counters[8 + stage], 1)] = i;     // work list of the solver launch of this stage
  }
  if (ncon > 0) bbg::ctxSave((T*)d.ctx + (size_t)i * bbg::CTXN, S, ncon, nd, qfs, qas, L);
}
// k_newton<T>: constraint solve of one RK stage for the envs that have contacts, in work-sorted order.  Uniform-warp
// solver (GNewton<T, true>): a warp leaves only when neither of its envs has contacts; otherwise both groups run the solver
// loops together and every collective uses the constant full-warp mask.
#ifndef BB_NEWTON_MINBLOCKS
#define BB_NEWTON_MINBLOCKS BB_WARP_MINBLOCKS      // residency (= register cap) of the solver kernel alone, fp64
#endif
template <typename T, int FM>
__global__ void __launch_bounds__(32 * BB_WPB, sizeof(T) == 4 ? BB_WARP_MINBLOCKS32 : BB_NEWTON_MINBLOCKS) k_newton(EnvParams p, DevState d, int stage) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const bbg::Ln L = bbg::makeLn();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane / bbg::G;
  const int slot = (blockIdx.x * (blockDim.x >> 5) + warp) * bbg::EPW + grp;
  const bool live = slot < p.N;
  const int i = d.order[live ? slot : p.N - 1];
  int* meta = d.meta + 4 * (size_t)i;
  int ncon = meta[bbg::META_NCON], nw = meta[bbg::META_NW] & 0xff, nd = meta[bbg::META_NW] >> 8;
  const bool act = live && ncon > 0 && !(meta[bbg::META_FLAGS] & 1);
  if (!__any_sync(0xffffffffu, act)) return;
  if (!act) { ncon = 0; nw = 0; nd = 0; }              // passenger group: empty contact loops, results discarded
  bbg::GS<T>& S = reinterpret_cast<bbg::GS<T>*>(smem_raw)[warp * bbg::EPW + grp];
  T* rk = (T*)d.rk + (size_t)i * bbg::RKN;
  T qfs = 0, qas = 0, warm = 0;
  if (act) {
    bbg::ctxLoad((const T*)d.ctx + (size_t)i * bbg::CTXN, S, ncon, nd, qfs, qas, L);
    warm = L.gl < NV ? rk[bbg::RK_WARM + L.gl] : (T)0;
  }
  __syncwarp();
  T* gs = (T*)d.gscr + (size_t)i * bbg::GSCR;
  bbg::Ln LU = L; LU.mask = 0xffffffffu;               // compile-time constant member mask for everything inlined below
  bbg::GNewton<T, true, FM> nwt(cmc<T>(), S, gs, LU, ncon, nw, nd, FM == 1, qfs, qas);
  int niter;
  const T qacc = nwt.run(warm, niter, act);
  if (!act) return;
  if (L.gl < NV) {
    rk[bbg::RK_QACC + L.gl] = qacc;
    if (p.solver_mode != 0 && stage < 3) rk[bbg::RK_WARM + L.gl] = qacc;
  }
  if (L.gl == 0) meta[bbg::META_NIT] += niter + (nwt.nevals << 12);   // low 12 bits: Newton iterations, above: line-search evaluations
}
// k_newton_p<T, FM>: persistent variant of k_newton (default).  Every 16-lane group draws environments from the stage's work
// list (the envs that have contacts, appended by k_stage in work-sorted CTA order) through an atomic cursor; when its solve
// ends it writes the result back and takes the next environment at the next Newton-iteration boundary, while the other group
// of the warp carries on with its own solve.  In k_newton a group whose solve has finished rides along until its partner is
// done (23 of 32 lanes active on the bench workload); here both halves of a warp stay busy until the list is empty.
// The arithmetic per environment is the same code in the same order (GNewton::begin / iterate), so results are identical.
template <typename T, int FM>
__global__ void __launch_bounds__(32 * BB_WPB, sizeof(T) == 4 ? BB_WARP_MINBLOCKS32 : BB_NEWTON_MINBLOCKS) k_newton_p(EnvParams p, DevState d, int stage) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const bbg::Ln L = bbg::makeLn();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane / bbg::G;
  bbg::GS<T>& S = reinterpret_cast<bbg::GS<T>*>(smem_raw)[warp * bbg::EPW + grp];
  bbg::Ln LU = L; LU.mask = 0xffffffffu;               // compile-time constant member mask for everything inlined below
  bbg::GNewton<T, true, FM> nwt(cmc<T>(), S, (T*)d.gscr, LU, 0, 0, 0, FM == 1, (T)0, (T)0);
  const int total = d.counters[8 + stage];
  int* cursor = d.counters + 12 + stage;
  bool on = false, done = false;
  int i = 0;
#pragma unroll 1
  for (;;) {
    const bool need = !on && !done;
    if (__any_sync(0xffffffffu, need)) {
      int slot = 0;
      if (need && L.gl == 0) slot = atomicAdd(cursor, 1);
      slot = __shfl_sync(0xffffffffu, slot, lane & 16);
      bool fresh = false;
      T warm = 0;
      if (need) {
        if (slot >= total) done = true;
        else {
          fresh = true;
          i = d.alist[slot];
          const int* meta = d.meta + 4 * (size_t)i;
          const int ncon = meta[bbg::META_NCON], nw = meta[bbg::META_NW] & 0xff, nd = meta[bbg::META_NW] >> 8;
          T qfs, qas;
          bbg::ctxLoad((const T*)d.ctx + (size_t)i * bbg::CTXN, S, ncon, nd, qfs, qas, L);
          warm = L.gl < NV ? ((const T*)d.rk + (size_t)i * bbg::RKN)[bbg::RK_WARM + L.gl] : (T)0;
          nwt.attach((T*)d.gscr + (size_t)i * bbg::GSCR, ncon, nw, nd, qfs, qas);
        }
      }
      __syncwarp();
      if (__any_sync(0xffffffffu, fresh)) {
        nwt.begin(warm, fresh);
        if (fresh) on = nwt.iter < cmc<T>().iterations;
      }
    }
    if (!__any_sync(0xffffffffu, on)) { if (__all_sync(0xffffffffu, done)) break; else continue; }
    const bool was = on;
    on = nwt.iterate(on);
    if (was && !on) {                                   // solve finished: write back (group-uniform branch, plain stores)
      T* rk = (T*)d.rk + (size_t)i * bbg::RKN;
      if (L.gl < NV) {
        rk[bbg::RK_QACC + L.gl] = nwt.qacc;
        if (FM == 1 && stage < 3) rk[bbg::RK_WARM + L.gl] = nwt.qacc;
      }
      if (L.gl == 0) d.meta[4 * (size_t)i + bbg::META_NIT] += nwt.iter + (nwt.nevals << 12);
    }
  }
}
// forward-dynamics probe through the group path (same outputs as k_probe)
template <typename T> __global__ void k_probe_warp(EnvParams p, DevState d, int env, const double* ctrl3, double* out, double* dbg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if ((threadIdx.x & 31) >= bbg::G) return;
  const bbg::Ln L = bbg::makeLn();
  bbg::GS<T>& S = reinterpret_cast<bbg::GS<T>*>(smem_raw)[0];
  const T* st = (const T*)d.st + (size_t)env * SST;
  for (int k = L.gl; k < NQ + NV; k += bbg::G) { if (k < NQ) S.xq[k] = st[k]; else S.xv[k - NQ] = st[k]; }
  const T warm = st[NQ + NV + L.gi];
  if (L.gl == NV) S.xv[NV] = 0;
  for (int k = L.gl; k < bbg::MSZ; k += bbg::G) S.M[k] = 0;
  if (L.gl < 3) S.ctrl[L.gl] = (T)ctrl3[L.gl];
  __syncwarp(L.mask);
  const float* hf = hfOf(p, d, env);
  int ncon, niter; T qas, qfs;
  const T qacc = bbg::gForward<T, true>(cmc<T>(), S, hf, mipOf(p, d, env), (T)p.zscale, (T*)d.gscr + (size_t)env * bbg::GSCR, L, warm, p.solver_mode != 0, true, ncon, niter, &qas, &qfs, dbg);
  if (L.gl < NV) { out[L.gl] = (double)qacc; out[15 + L.gl] = (double)qas; out[30 + L.gl] = (double)qfs; }
  if (L.gl == 0) { out[45] = ncon; out[46] = niter; }
}

// explicit reset: mask -> reset list (+ new terrain seed)
__global__ void k_mask_to_list(EnvParams p, DevState d, const uint8_t* __restrict__ mask, const int* __restrict__ seeds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.N) return;
  if (mask && !mask[i]) return;
  const unsigned ep = d.episode[i] + 1; d.episode[i] = ep;
  if (seeds) {   // caller-fixed r_seed (replay); the table only holds seeds 0 .. BB_PERLIN_SEEDS - 1
    const int sd = seeds[i];
    d.tseed[i] = p.hf_mode == HF_TABLE ? (int)((unsigned)sd % (unsigned)BB_PERLIN_SEEDS) : sd;
  } else d.tseed[i] = drawTerrainSeed(p, d, i, ep);
  d.reset_list[atomicAdd(&d.counters[0], 1)] = i;
}
// k_begin_step, one block of WORK_BINS threads: clears the work-list counters and turns the key histogram of the previous step into
// the scatter cursors of k_order (descending keys: the longest solves are scheduled first).
__global__ void k_clear_counters(DevState d) { d.counters[0] = 0; d.counters[1] = 0; d.counters[2] = 0; for (int k = 8; k < 16; k++) d.counters[k] = 0; }
__global__ void k_begin_step(DevState d) {
  __shared__ int h[WORK_BINS];
  const int t = threadIdx.x;
  if (t == 0) { d.counters[0] = 0; d.counters[1] = 0; d.counters[2] = 0; }
  if (t >= 8 && t < 16) d.counters[t] = 0;
  h[t] = d.bins[t];
  __syncthreads();
  int start = 0;
  for (int k = WORK_BINS - 1; k > t; k--) start += h[k];
  d.bins[WORK_BINS + t] = start;
  d.bins[t] = 0;
}
// counting-sort scatter: order[] = env indices grouped by key (the order inside a bin is irrelevant: envs are independent)
__global__ void k_order(int N, DevState d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int key = d.work[i];
  const unsigned peers = __match_any_sync(__activemask(), key);
  const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
  int base = 0;
  if (lane == leader) base = atomicAdd(&d.bins[WORK_BINS + key], __popc(peers));
  base = __shfl_sync(peers, base, leader);
  d.order[base + __popc(peers & ((1u << lane) - 1u))] = i;
}
__global__ void k_init_order(int N, DevState d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) { d.work[i] = 0; d.order[i] = i; }
  if (i < 2 * WORK_BINS) d.bins[i] = i == 0 ? N : 0;
}

// --------------------------------------------------------------------------------------------- simplex fBm terrain
__constant__ unsigned char c_perm[256];
__device__ __forceinline__ int dperm(const int* __restrict__ perm, int i) { return perm[i & 255]; }   // perm: shared-memory copy of c_perm
// All arithmetic below uses unfused round-to-nearest single-precision ops (__fmul_rn/__fadd_rn/__fsub_rn are never
// contracted into FMAs): the tiled-noise coordinates are ~1e4 where one float ulp is 1e-3, so the heights are only
// reproducible if every rounding step matches the plain C evaluation order of noise._simplex.
#define FM(a, b) __fmul_rn((a), (b))
#define FA(a, b) __fadd_rn((a), (b))
#define FS(a, b) __fsub_rn((a), (b))
__device__ __forceinline__ float grad4(int gi, float x, float y, float z, float w) {
  // 32 gradient directions of 4-D simplex noise: one zero component (gi>>3 selects which), signs from the low bits.
  // Branch-free selection of the three participating coordinates (gi differs from lane to lane; a switch serialises).
  const int zc = gi >> 3;
  const float s0 = (gi & 4) ? -1.f : 1.f, s1 = (gi & 2) ? -1.f : 1.f, s2 = (gi & 1) ? -1.f : 1.f;
  const float a = zc == 0 ? y : x, b = zc <= 1 ? z : y, c = zc <= 2 ? w : z;
  return FA(FA(FM(s0, a), FM(s1, b)), FM(s2, c));
}
__device__ float simplex4(const int* __restrict__ perm, float x, float y, float z, float w) {
  const float F4 = 0.309016994f, G4 = 0.138196601f;
  const float sk = FM(FA(FA(FA(x, y), z), w), F4);
  const float fi = floorf(FA(x, sk)), fj = floorf(FA(y, sk)), fk = floorf(FA(z, sk)), fl = floorf(FA(w, sk));
  const float t = FM(FA(FA(FA(fi, fj), fk), fl), G4);
  const float x0 = FS(x, FS(fi, t)), y0 = FS(y, FS(fj, t)), z0 = FS(z, FS(fk, t)), w0 = FS(w, FS(fl, t));
  const int rx = (x0 > y0) + (x0 > z0) + (x0 > w0);
  const int ry = !(x0 > y0) + (y0 > z0) + (y0 > w0);
  const int rz = !(x0 > z0) + !(y0 > z0) + (z0 > w0);
  const int rw = 6 - rx - ry - rz;
  const int I = (int)fi & 255, J = (int)fj & 255, K = (int)fk & 255, L = (int)fl & 255;
  float total = 0.f;
#pragma unroll
  for (int c = 0; c < 5; c++) {
    const int thr = 4 - c;   // corner c steps along the c highest-ranked axes
    const int i1 = c == 0 ? 0 : (rx >= thr), j1 = c == 0 ? 0 : (ry >= thr), k1 = c == 0 ? 0 : (rz >= thr), l1 = c == 0 ? 0 : (rw >= thr);
    const float off = FM((float)c, G4);
    const float xc = FA(FS(x0, (float)i1), off), yc = FA(FS(y0, (float)j1), off), zc = FA(FS(z0, (float)k1), off), wc = FA(FS(w0, (float)l1), off);
    float tt = FS(FS(FS(FS(0.6f, FM(xc, xc)), FM(yc, yc)), FM(zc, zc)), FM(wc, wc));
    if (tt >= 0.f) {
      const int gi = dperm(perm, I + i1 + dperm(perm, J + j1 + dperm(perm, K + k1 + dperm(perm, L + l1)))) & 31;
      tt = FM(tt, tt);
      total = FA(total, FM(FM(tt, tt), grad4(gi, xc, yc, zc, wc)));
    }
  }
  return FM(27.f, total);
}
// snoise2(x, y, octaves, persistence, lacunarity, repeatx=1024, repeaty=1024, base=seed): tiled branch = 4-D fBm on two circles.
// The circle coordinates depend on the grid index only (not on the seed): perlinCircle(idx) = (sin, cos) * R of one axis.
// fast_sin / fast_cos of noise/_noise.h: argument in half-turns, wrapped to [-1, 1] by the 1.5 * 2^24 float trick, parabola + correction
__device__ __forceinline__ float noiseFastSin(float x) {
  const float z = FA(x, 25165824.0f);
  x = FS(x, FS(z, 25165824.0f));
  const float y = FS(x, FM(x, fabsf(x)));
  return FM(y, FA(3.1f, FM(3.6f, fabsf(y))));
}
__device__ __forceinline__ void perlinCircle(int idx, float scale, float& s, float& c) {
  const float x = (float)((double)idx / (double)scale);
  const float rr = (float)(1024.0 * 0.3183098861837907 * 0.5);
  const float xf = (float)((double)x * 2.0 / 1024.0);
  s = FM(noiseFastSin(xf), rr); c = FM(noiseFastSin(FA(xf, 0.5f)), rr);
}
__device__ float perlinFbm(const int* __restrict__ perm, float si, float ci, float sj, float cj, int seed, int oct, float pers, float lac, float amp) {
  const float x = si, y = sj, z = FA((float)seed, ci), w = FA((float)seed, cj);
  float freq = 1.f, a = 1.f, mx = 1.f, total = simplex4(perm, x, y, z, w);
  for (int o = 1; o < oct; o++) {
    freq = FM(freq, lac); a = FM(a, pers); mx = FA(mx, a);
    total = FA(total, FM(simplex4(perm, FM(x, freq), FM(y, freq), FM(z, freq), FM(w, freq)), a));
  }
  const double v = ((double)__fdiv_rn(total, mx) + 1.0) / 2.0 * (double)amp;
  return (float)(v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v));
}
__device__ float perlinHeight(const int* __restrict__ perm, int i, int j, int seed, float scale, int oct, float pers, float lac, float amp) {
  float si, ci, sj, cj;
  perlinCircle(i, scale, si, ci); perlinCircle(j, scale, sj, cj);
  return perlinFbm(perm, si, ci, sj, cj, seed, oct, pers, lac, amp);
}
// 2-D simplex noise of noise._simplex (untiled branch of snoise2: terrain/gradient.py:74-80), same unfused float arithmetic
__device__ float simplex2(const int* __restrict__ perm, float x, float y) {
  const float F2 = 0.3660254037844386f, G2 = 0.21132486540518713f;
  const float s = FM(FA(x, y), F2), i = floorf(FA(x, s)), j = floorf(FA(y, s)), t = FM(FA(i, j), G2);
  float xx[3], yy[3];
  xx[0] = FS(x, FS(i, t)); yy[0] = FS(y, FS(j, t));
  const int i1 = xx[0] > yy[0], j1 = xx[0] <= yy[0];
  xx[2] = FS(FA(xx[0], FM(G2, 2.0f)), 1.0f); yy[2] = FS(FA(yy[0], FM(G2, 2.0f)), 1.0f);
  xx[1] = FA(FS(xx[0], (float)i1), G2); yy[1] = FA(FS(yy[0], (float)j1), G2);
  const int I = (int)i & 255, J = (int)j & 255;
  const int g[3] = {dperm(perm, I + dperm(perm, J)) % 12, dperm(perm, I + i1 + dperm(perm, J + j1)) % 12, dperm(perm, I + 1 + dperm(perm, J + 1)) % 12};
  float n[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const float f = FS(FS(0.5f, FM(xx[c], xx[c])), FM(yy[c], yy[c]));
    if (f > 0.f) {
      // GRAD3 = {1,1,0},{-1,1,0},{1,-1,0},{-1,-1,0},{1,0,1},{-1,0,1},{1,0,-1},{-1,0,-1},{0,1,1},{0,-1,1},{0,1,-1},{0,-1,-1}: only x, y are used
      const int gi = g[c];
      const float gx = gi < 8 ? ((gi & 1) ? -1.f : 1.f) : 0.f;
      const float gy = gi < 4 ? ((gi & 2) ? -1.f : 1.f) : (gi < 8 ? 0.f : ((gi & 1) ? -1.f : 1.f));
      const float f2 = FM(f, f);
      n[c] = FM(FM(f2, f2), FA(FM(gx, xx[c]), FM(gy, yy[c])));
    }
  }
  return FM(FA(FA(n[0], n[1]), n[2]), 70.0f);
}
__global__ void __launch_bounds__(256) k_snoise2_grid(int n, float scale, int oct, float pers, float lac, int base, float* __restrict__ out) {
  __shared__ int perm[256];
  perm[threadIdx.x] = c_perm[threadIdx.x];
  __syncthreads();
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= n * n) return;
  const float x = (float)((double)(cell / n) / (double)scale), y = (float)((double)(cell % n) / (double)scale), z = (float)base;
  float freq = 1.f, amp = 1.f, mx = 1.f, total = simplex2(perm, FA(x, z), FA(y, z));
  for (int o = 1; o < oct; o++) {
    freq = FM(freq, lac); amp = FM(amp, pers); mx = FA(mx, amp);
    total = FA(total, FM(simplex2(perm, FA(FM(x, freq), z), FA(FM(y, freq), z)), amp));
  }
  out[cell] = __fdiv_rn(total, mx);
}
#undef FM
#undef FA
#undef FS
// circle coordinates of the 293 grid indices (engine constant: depends on perlin_scale only)
__global__ void k_iota(int n, int first, int* __restrict__ out) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) out[i] = first + i; }
__global__ void k_perlin_table(float scale, float* __restrict__ tab) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < HN) perlinCircle(i, scale, tab[i], tab[HN + i]);
}
// grid.x covers the cells of one heightfield, grid.y strides over the work list
__global__ void __launch_bounds__(256) k_terrain(EnvParams p, DevState d, const int* __restrict__ list, const int* __restrict__ count,
                                                 int fixed_count, const int* __restrict__ seeds, float* __restrict__ out) {
  __shared__ int perm[256];
  perm[threadIdx.x] = c_perm[threadIdx.x];
  __syncthreads();
  const int n = count ? *count : fixed_count;
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= HF_CELLS) return;
  const int r = cell / HN, c = cell - r * HN;
  const float si = d.ptab[r], ci = d.ptab[HN + r], sj = d.ptab[c], cj = d.ptab[HN + c];
  for (int k = blockIdx.y; k < n; k += gridDim.y) {
    const int env = list ? list[k] : k;
    const int seed = seeds ? seeds[k] : d.tseed[env];
    float* dst = out ? out + (size_t)k * HF_CELLS : d.hfield + (size_t)env * HF_CELLS;
    dst[cell] = perlinFbm(perm, si, ci, sj, cj, seed, p.poct, p.ppers, p.plac, p.pamp);
  }
}

// general n x n grid (registry callable `perlin`, terrain/perlin.py:8-74 at arbitrary n)
__global__ void __launch_bounds__(256) k_perlin_grid(int n, float scale, int oct, float pers, float lac, float amp, const int* __restrict__ seeds,
                                                      int nseeds, float* __restrict__ out) {
  __shared__ int perm[256];
  perm[threadIdx.x] = c_perm[threadIdx.x];
  __syncthreads();
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= n * n) return;
  const int r = cell / n, c = cell - r * n;
  for (int k = blockIdx.y; k < nseeds; k += gridDim.y) out[(size_t)k * n * n + cell] = perlinHeight(perm, r, c, seeds[k], scale, oct, pers, lac, amp);
}

// --------------------------------------------------------------------------------------------- reset
template <typename T>
__global__ void k_reset(EnvParams p, DevState d, bb_io io) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= d.counters[0]) return;
  const int i = d.reset_list[k];
  // spawn height: max of hfield rows/cols 140..151 (ballbot_env.py:546-563, cell_size = 5/293 quirk)
  const float* hf = hfOf(p, d, i);
  float mx = -1e30f;
  for (int r = 140; r < 152; r++) for (int c = 140; c < 152; c++) mx = fmaxf(mx, hf[r * HN + c]);
  const double off = (double)mx * (double)p.zscale + 0.01;
  const double q0[NQ] = {0, 0, 0.24 + off, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.26 + off, 1, 0, 0, 0};
  T* st = (T*)d.st; T* cq = (T*)d.camq;
  for (int j = 0; j < NQ; j++) { st[(size_t)i * SST + j] = (T)q0[j]; cq[(size_t)i * CST + j] = (T)q0[j]; }
  for (int j = NQ; j < SST; j++) st[(size_t)i * SST + j] = (T)0;
  d.step_count[i] = 0; d.cam_steps[i] = 0; d.ep_ret[i] = 0.f; d.ep_len[i] = 0;
  // reset observation: mj_forward at rest => zero rotation vector and velocities (ballbot_env.py:634)
  for (int j = 0; j < 3; j++) {
    io.orientation[3 * i + j] = 0.f; io.angular_vel[3 * i + j] = 0.f; io.vel[3 * i + j] = 0.f; io.motor_state[3 * i + j] = 0.f; io.actions[3 * i + j] = 0.f;
  }
  io.rel_image_ts[i] = 0.f;
  if (p.cameras) d.refresh_list[atomicAdd(&d.counters[1], 1)] = i;
}

// --------------------------------------------------------------------------------------------- depth ray cast
struct F3 { float x, y, z; };
__device__ __forceinline__ F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ F3 operator*(F3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float fdot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
struct Prim { F3 c, u; float r, hl; int type; };  // type 0 sphere, 1 capsule, 2 cylinder
struct Scene { F3 cam_o[2]; F3 cam_x[2], cam_y[2], cam_z[2]; Prim prim[7]; float brad2[7]; float rect[7][4]; };   // rect: screen-space bounds of the work unit's camera

__device__ __forceinline__ float hitSphere(F3 o, F3 dir, F3 c, float rad) {
  const F3 oc = o - c; const float a = fdot(dir, dir), b = fdot(oc, dir), cc = fdot(oc, oc) - rad * rad;
  const float disc = b * b - a * cc; if (disc < 0.f) return -1.f;
  return __fdividef(-b - sqrtf(disc), a);   // entry point only (back faces are culled by the rasteriser); 2-ulp division, depth tolerance 1e-3
}
__device__ __forceinline__ float hitSide(F3 o, F3 dir, const Prim& p) {
  const F3 oc = o - p.c; const float od = fdot(oc, p.u), dd = fdot(dir, p.u);
  const F3 op = oc - p.u * od, dp = dir - p.u * dd;
  const float a = fdot(dp, dp), b = fdot(op, dp), cc = fdot(op, op) - p.r * p.r;
  if (a < 1e-18f) return -1.f;
  const float disc = b * b - a * cc; if (disc < 0.f) return -1.f;
  const float t = __fdividef(-b - sqrtf(disc), a); const float zz = od + t * dd;
  return (zz < -p.hl || zz > p.hl) ? -1.f : t;
}
__device__ float hitPrim(F3 o, F3 dir, const Prim& p) {
  if (p.type == 0) return hitSphere(o, dir, p.c, p.r);
  float best = -1.f; float t = hitSide(o, dir, p); if (t > 0.f) best = t;
  const F3 oc = o - p.c;
  for (int s = -1; s <= 1; s += 2) {
    if (p.type == 1) {
      const F3 e = p.c + p.u * ((float)s * p.hl);
      t = hitSphere(o, dir, e, p.r);
      if (t > 0.f) { const F3 hp = oc + dir * t; if ((float)s * fdot(hp, p.u) >= p.hl && (best < 0.f || t < best)) best = t; }
    } else {
      const float od = fdot(oc, p.u), dd = fdot(dir, p.u);
      if (fabsf(dd) < 1e-18f || (float)s * dd >= 0.f) continue;
      t = __fdividef((float)s * p.hl - od, dd); if (t <= 0.f) continue;
      const F3 hp = oc + dir * t - p.u * ((float)s * p.hl);
      if (fdot(hp, hp) <= p.r * p.r && (best < 0.f || t < best)) best = t;
    }
  }
  return best;
}
// ray vs heightfield: 2-D DDA over the cells in grid coordinates (cell units, raw height units), both triangles of each
// visited cell.  The two triangles of a cell are graphs over the cell (split along the (0,0)-(1,1) diagonal, the same
// diagonal as the collision prisms), so each test is one ray/plane solve plus a range check in cell coordinates.
__device__ float hitHfield(F3 o, F3 dir, const float* __restrict__ hf, float sx, float sz, float tmax) {
  const int n = HN; const float idx = (float)(n - 1) / (2.f * sx), isz = 1.f / sz;
  const float gx0 = (o.x + sx) * idx, gy0 = (o.y + sx) * idx, dgx = dir.x * idx, dgy = dir.y * idx;   // grid coordinates of the ray
  const float oz = o.z * isz, dz = dir.z * isz;                                                       // raw height units
  const bool zx = fabsf(dir.x) < 1e-18f, zy = fabsf(dir.y) < 1e-18f;
  const float ix = zx ? 0.f : 1.f / dgx, iy = zy ? 0.f : 1.f / dgy;
  // clip the ray to the field [0, n-1]^2 (grid units)
  float t0 = 0.f, t1 = tmax;
  if (zx) { if (gx0 < 0.f || gx0 > (float)(n - 1)) return -1.f; }
  else { const float ta = -gx0 * ix, tb = ((float)(n - 1) - gx0) * ix; t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb)); }
  if (zy) { if (gy0 < 0.f || gy0 > (float)(n - 1)) return -1.f; }
  else { const float ta = -gy0 * iy, tb = ((float)(n - 1) - gy0) * iy; t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb)); }
  if (t0 >= t1) return -1.f;
  int cx = (int)floorf(gx0 + (t0 + 1e-6f) * dgx), cy = (int)floorf(gy0 + (t0 + 1e-6f) * dgy);
  cx = min(max(cx, 0), n - 2); cy = min(max(cy, 0), n - 2);
  const int stx = dir.x > 0.f ? 1 : -1, sty = dir.y > 0.f ? 1 : -1;
  const float tdx = zx ? 3e38f : fabsf(ix), tdy = zy ? 3e38f : fabsf(iy);
  float tmx = zx ? 3e38f : ((float)(cx + (stx > 0 ? 1 : 0)) - gx0) * ix;
  float tmy = zy ? 3e38f : ((float)(cy + (sty > 0 ? 1 : 0)) - gy0) * iy;
  float tcur = t0;
  const bool down = dz < 0.f;
  const float e = 1e-5f;
  for (int it = 0; it < 4 * n; it++) {
    if (tcur > t1) return -1.f;                                               // left the field or the depth range
    const float* h = hf + cy * n + cx;
    const float h00 = h[0], h10 = h[1], h01 = h[n], h11 = h[n + 1];
    const bool xs = tmx < tmy;
    const float tn = xs ? tmx : tmy, texit = fminf(tn, t1);
    const float zlow = oz + (down ? texit : tcur) * dz;                       // lowest point of the ray inside this cell
    if (zlow <= fmaxf(fmaxf(h00, h10), fmaxf(h01, h11))) {                    // otherwise the ray passes above both triangles
      const float u0 = gx0 - (float)cx, v0 = gy0 - (float)cy, c0 = h00 - oz;
      float best = -1.f;
      {  // triangle (0,0) (1,1) (1,0): u >= v, z = h00 + (h10 - h00) u + (h11 - h10) v
        const float a = h10 - h00, b = h11 - h10;
        const float den = dz - a * dgx - b * dgy, num = c0 + a * u0 + b * v0;
        if (fabsf(den) > 1e-18f) {
          const float t = __fdividef(num, den), u = u0 + t * dgx, v = v0 + t * dgy;   // 2-ulp division: depth tolerance is 1e-3
          if (t > 0.f && v >= -e && u <= 1.f + e && u >= v - e) best = t;
        }
      }
      {  // triangle (0,1) (0,0) (1,1): v >= u, z = h00 + (h11 - h01) u + (h01 - h00) v
        const float a = h11 - h01, b = h01 - h00;
        const float den = dz - a * dgx - b * dgy, num = c0 + a * u0 + b * v0;
        if (fabsf(den) > 1e-18f) {
          const float t = __fdividef(num, den), u = u0 + t * dgx, v = v0 + t * dgy;   // 2-ulp division: depth tolerance is 1e-3
          if (t > 0.f && u >= -e && v <= 1.f + e && v >= u - e && (best < 0.f || t < best)) best = t;
        }
      }
      if (best > 0.f && best <= tmax) return best;
    }
    // next cell (indices are clamped: beyond the field edge tcur exceeds t1 and the loop ends)
    // (branch-free: neighbouring rays step along different axes)
    cx = min(max(cx + (xs ? stx : 0), 0), n - 2); cy = min(max(cy + (xs ? 0 : sty), 0), n - 2);
    tmx += xs ? tdx : 0.f; tmy += xs ? 0.f : tdy;
    tcur = tn;
  }
  return -1.f;
}
// scene item `item` of env i: 0,1 cameras; 2 ball; 3..5 wheel capsules; 6 tower cylinder; 7,8 camera sticks (one thread each)
template <typename T>
__device__ void buildScene(const ModelConst<float>& mc, const T* __restrict__ cq, int stride, int i, Scene& sc, int item) {
  float q[NQ]; for (int k = 0; k < NQ; k++) q[k] = (float)cq[(size_t)i * stride + k];
  const Rot<float> RB = quat2rot(q[3], q[4], q[5], q[6]);
  const V3<float> pB = mk(q[0], q[1], q[2]);
  auto toF = [](const V3<float>& v) { return f3(v.x, v.y, v.z); };
  if (item < 2) {
    const int c = item;
    sc.cam_o[c] = toF(pB + rot(RB, ld3(mc.cam_pos[c])));
    const float* m = mc.cam_rot[c];   // row-major camera->base rotation
    sc.cam_x[c] = toF(rot(RB, mk(m[0], m[3], m[6]))); sc.cam_y[c] = toF(rot(RB, mk(m[1], m[4], m[7]))); sc.cam_z[c] = toF(rot(RB, mk(m[2], m[5], m[8])));
    return;
  }
  const int g = item - 2;
  Prim& pr = sc.prim[g];
  if (g == 0) {
    const Rot<float> RL = quat2rot(q[13], q[14], q[15], q[16]);
    pr.type = 0; pr.c = toF(mk(q[10], q[11], q[12]) + rot(RL, mk(0.f, 0.f, mc.dz))); pr.r = mc.ball_r; pr.hl = 0.f; pr.u = f3(0, 0, 1);
  } else if (g < 4) {
    const int w = g - 1;
    float sq, cq2; sincosf(q[7 + w], &sq, &cq2);
    const V3<float> a = ld3(mc.ax[w]);
    const V3<float> si = rodrigues(a, ld3(mc.s0[w]), sq, cq2), ui = rodrigues(a, ld3(mc.u0[w]), sq, cq2);
    pr.type = 1; pr.r = mc.wheel_r; pr.hl = mc.wheel_hl;
    pr.c = toF(pB + rot(RB, ld3(mc.anc[w]) + si)); pr.u = toF(rot(RB, ui));
  } else if (g == 4) {
    pr.type = 2; pr.r = mc.tower_r; pr.hl = mc.tower_hl; pr.c = toF(pB + rot(RB, ld3(mc.tower_c))); pr.u = toF(RB.c2);
  } else {
    const int k = g - 5;
    pr.type = 1; pr.r = mc.stick_r; pr.hl = mc.stick_hl;
    pr.c = toF(pB + rot(RB, ld3(mc.stick_c[k]))); pr.u = toF(rot(RB, ld3(mc.stick_u[k])));
  }
  // squared bounding-sphere radius (sphere r, capsule hl + r, cylinder sqrt(r^2 + hl^2)), 1 mm slack
  const float br = pr.type == 0 ? pr.r : (pr.type == 1 ? pr.hl + pr.r : sqrtf(pr.r * pr.r + pr.hl * pr.hl));
  sc.brad2[g] = (br + 1e-3f) * (br + 1e-3f);
}
// Persistent warps, no CTA barriers: a unit of work is (env from the work list, camera, quarter of the image); every warp
// draws units from a global counter (d.counters[2], cleared at the start of each bb_step / bb_reset / bb_render_depth),
// builds the scene in its own shared-memory slot (9 lanes, one item each) and renders the unit's 8 x 4-pixel tiles
// (neighbouring rays traverse similar cells => less divergence than row segments).  Consecutive units belong to the same
// env, so the warps of a CTA share heightfield lines in L1/L2.
#ifndef BB_DEPTH_PARTS
#define BB_DEPTH_PARTS 8
#endif
constexpr int DEPTH_PARTS = BB_DEPTH_PARTS;
template <typename T>
__global__ void __launch_bounds__(256) k_depth(EnvParams p, DevState d, const int* __restrict__ list, const int* __restrict__ count,
                                               int fixed_count, const T* __restrict__ cfgq, int cfg_stride, float* __restrict__ img0, float* __restrict__ img1) {
  __shared__ Scene scs[8];
  const int lane = threadIdx.x & 31;
  Scene& sc = scs[threadIdx.x >> 5];
  const int n = count ? *count : fixed_count;
  const int units = n * 2 * DEPTH_PARTS;
  const int npix = p.im_h * p.im_w;
  const int tiles_x = (p.im_w + 7) >> 3, tiles = tiles_x * ((p.im_h + 3) >> 2), chunk = (tiles + DEPTH_PARTS - 1) / DEPTH_PARTS;
  const int lx = lane & 7, ly = lane >> 3;
  for (;;) {
    int u = 0;
    if (lane == 0) u = atomicAdd(&d.counters[2], 1);
    u = __shfl_sync(0xffffffffu, u, 0);
    if (u >= units) break;
    const int k = u / (2 * DEPTH_PARTS), cam = (u / DEPTH_PARTS) & 1, part = u % DEPTH_PARTS;
    const int env = list ? list[k] : k;
    __syncwarp();
    if (lane < 9) buildScene(c_mc32, cfgq, cfg_stride, env, sc, lane);
    __syncwarp();
    const float* hf = hfOf(p, d, env);
    float* out = (cam ? img1 : img0) + (size_t)env * npix;
    const F3 o = sc.cam_o[cam];
    const float aspect = (float)p.im_w / p.im_h;
    // ---- screen-space bounds of the primitives' bounding spheres for this camera (lanes 0..6), then one 7-bit mask per 8 x 4
    // tile of this unit (lanes 0..chunk-1): a ray only visits the primitives whose bounds overlap its tile (conservative: the
    // per-ray bounding-sphere test and the exact test are unchanged, so the image is identical).  xn = x / depth of a sphere at
    // camera coordinates (xc, yc, zc) spans tan(theta0 -+ phi), theta0 = atan(xc / zc), sin(phi) = R / |(xc, zc)|.
    if (lane < 7) {
      const F3 oc = sc.prim[lane].c - o;
      const float xc = fdot(oc, sc.cam_x[cam]), yc = fdot(oc, sc.cam_y[cam]), zc = -fdot(oc, sc.cam_z[cam]);
      const float R2 = sc.brad2[lane], R = sqrtf(R2) * 1.001f + 1e-6f;
      float x0 = -3e38f, x1 = 3e38f, y0 = -3e38f, y1 = 3e38f;
      if (zc + R < 0.f) { x0 = 3e38f; x1 = -3e38f; }                       // entirely behind the camera plane: never visible
      else if (zc > R) {
        const float iz = 1.f / zc;
        { const float t0 = xc * iz, tp = R * rsqrtf(fmaxf(xc * xc + zc * zc - R * R, 1e-30f)); x0 = (t0 - tp) / (1.f + t0 * tp); x1 = (t0 + tp) / (1.f - t0 * tp);
          if (!(1.f - t0 * tp > 1e-6f)) x1 = 3e38f; if (!(1.f + t0 * tp > 1e-6f)) x0 = -3e38f; }
        { const float t0 = yc * iz, tp = R * rsqrtf(fmaxf(yc * yc + zc * zc - R * R, 1e-30f)); y0 = (t0 - tp) / (1.f + t0 * tp); y1 = (t0 + tp) / (1.f - t0 * tp);
          if (!(1.f - t0 * tp > 1e-6f)) y1 = 3e38f; if (!(1.f + t0 * tp > 1e-6f)) y0 = -3e38f; }
      }
      sc.rect[lane][0] = x0 - 1e-4f; sc.rect[lane][1] = x1 + 1e-4f; sc.rect[lane][2] = y0 - 1e-4f; sc.rect[lane][3] = y1 + 1e-4f;
    }
    __syncwarp();
    const int tend = min(tiles, (part + 1) * chunk);
    unsigned mymask = 0;
    {
      const int tile = part * chunk + lane;
      if (chunk <= 32 && lane < chunk && tile < tend) {
        const int r0 = (tile / tiles_x) * 4, c0 = (tile % tiles_x) * 8;
        const float xa = (2.f * (c0 + 0.5f) / p.im_w - 1.f) * aspect, xb = (2.f * (c0 + 7.5f) / p.im_w - 1.f) * aspect;
        const float yb = 1.f - 2.f * (r0 + 0.5f) / p.im_h, ya = 1.f - 2.f * (r0 + 3.5f) / p.im_h;
#pragma unroll
        for (int g = 0; g < 7; g++)
          if (sc.rect[g][0] <= xb && sc.rect[g][1] >= xa && sc.rect[g][2] <= yb && sc.rect[g][3] >= ya) mymask |= 1u << g;
      }
    }
    for (int tile = part * chunk; tile < tend; tile++) {
      const unsigned tmask = chunk <= 32 ? __shfl_sync(0xffffffffu, mymask, (tile - part * chunk) & 31) : 0x7fu;   // (larger images: no tile culling)
      const int r = (tile / tiles_x) * 4 + ly, c = (tile % tiles_x) * 8 + lx;
      if (r >= p.im_h || c >= p.im_w) continue;
      const int px = r * p.im_w + c;
      const float xn = (2.f * (c + 0.5f) / p.im_w - 1.f) * aspect, yn = 1.f - 2.f * (r + 0.5f) / p.im_h;  // fovy 90
      const F3 dir = sc.cam_x[cam] * xn + sc.cam_y[cam] * yn - sc.cam_z[cam];
      const float inv_dd = 1.f / fdot(dir, dir);
      float best = 1.0f;   // depth >= 1 is clipped to 1 (sensors/rgbd.py:74)
#pragma unroll 1
      for (unsigned mm = tmask; mm; mm &= mm - 1) {
        const int g = __ffs(mm) - 1;
        const Prim& pr = sc.prim[g];
        const F3 oc = pr.c - o; const float along = fdot(oc, dir) * inv_dd;   // bounding-sphere reject before the exact test
        const F3 perp = oc - dir * along;
        if (fdot(perp, perp) > sc.brad2[g]) continue;
        const float t = hitPrim(o, dir, pr); if (t > 1e-4f && t < best) best = t;
      }
      const float t = hitHfield(o, dir, hf, c_mc32.hx, p.zscale, best);
      if (t > 1e-4f && t < best) best = t;
      out[px] = best;
    }
  }
}


// ---- rasterising variant of the depth observation (default for images of up to 64 x 64 pixels).
// The reference renders with OpenGL, i.e. it rasterises; k_depth above casts one ray per pixel and walks the heightfield cell
// by cell (~14 cells x 2 plane tests per ray).  Here one CTA owns one (env, camera) image with a z-buffer in shared memory:
//   1. every pixel gets the nearest robot primitive (ball, wheels, tower, sticks) exactly like the ray-caster;
//   2. the threads sweep the heightfield cells under the view pyramid (clipped at depth 1): a cell's four corners are taken to
//      camera space, cells outside the pyramid / beyond the clip / behind the camera are culled, the others are scanned over
//      the bounding box of their projection, each pixel running the SAME ray / plane solve and in-triangle test as hitHfield
//      for the cell's two triangles; hits go into the z-buffer with atomicMin on the float bits (t > 0 => ordered as uint).
// A pixel is now tested against the ~2-4 triangles whose projection covers it instead of every cell its ray crosses, and the
// depth values are bit-identical to the ray-caster's for the winning triangle (same expressions, same epsilons).
constexpr int RASTER_MAX_PIX = 64 * 64, RASTER_BATCH = 2048;
template <typename T>
__global__ void __launch_bounds__(128) k_depth_raster(EnvParams p, DevState d, const int* __restrict__ list, const int* __restrict__ count,
                                                      int fixed_count, const T* __restrict__ cfgq, int cfg_stride, float* __restrict__ img0, float* __restrict__ img1) {
  __shared__ Scene sc;
  __shared__ unsigned zb[RASTER_MAX_PIX];
  __shared__ int2 s_work[RASTER_BATCH];
  __shared__ int s_unit, s_nwork;
  const int tid = threadIdx.x;
  const int n = count ? *count : fixed_count;
  const int units = n * 2, npix = p.im_h * p.im_w, W = p.im_w, H = p.im_h;
  const float aspect = (float)W / (float)H;
  const float sx = c_mc32.hx, idx = (float)(HN - 1) / (2.f * sx), isz = 1.f / p.zscale;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_unit = atomicAdd(&d.counters[2], 1);
    __syncthreads();
    const int u = s_unit;
    if (u >= units) break;
    const int k = u >> 1, cam = u & 1;
    const int env = list ? list[k] : k;
    if (tid < 9) buildScene(c_mc32, cfgq, cfg_stride, env, sc, tid);
    __syncthreads();
    const float* hf = hfOf(p, d, env);
    float* out = (cam ? img1 : img0) + (size_t)env * npix;
    const F3 o = sc.cam_o[cam], ax = sc.cam_x[cam], ay = sc.cam_y[cam], az = sc.cam_z[cam];
    // ---- 1. primitives, one pixel per thread and iteration
    for (int px = tid; px < npix; px += blockDim.x) {
      const int r = px / W, c = px - r * W;
      const float xn = (2.f * (c + 0.5f) / W - 1.f) * aspect, yn = 1.f - 2.f * (r + 0.5f) / H;
      const F3 dir = ax * xn + ay * yn - az;
      const float inv_dd = 1.f / fdot(dir, dir);
      float best = 1.0f;
#pragma unroll 1
      for (int g = 0; g < 7; g++) {
        const Prim& pr = sc.prim[g];
        const F3 oc = pr.c - o; const float along = fdot(oc, dir) * inv_dd;
        const F3 perp = oc - dir * along;
        if (fdot(perp, perp) > sc.brad2[g]) continue;
        const float t = hitPrim(o, dir, pr); if (t > 1e-4f && t < best) best = t;
      }
      zb[px] = __float_as_uint(best);
    }
    // ---- 2. heightfield cells under the view pyramid (apex + the four corner rays at depth 1), in grid coordinates
    float gxlo = (o.x + sx) * idx, gxhi = gxlo, gylo = (o.y + sx) * idx, gyhi = gylo;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const F3 cdir = ax * ((q & 1) ? aspect : -aspect) + ay * ((q & 2) ? 1.f : -1.f) - az;
      const float gx = (o.x + cdir.x + sx) * idx, gy = (o.y + cdir.y + sx) * idx;
      gxlo = fminf(gxlo, gx); gxhi = fmaxf(gxhi, gx); gylo = fminf(gylo, gy); gyhi = fmaxf(gyhi, gy);
    }
    const int cx0 = max(0, (int)floorf(gxlo)), cx1 = min(HN - 2, (int)floorf(gxhi)), cy0 = max(0, (int)floorf(gylo)), cy1 = min(HN - 2, (int)floorf(gyhi));
    const int ncx = cx1 - cx0 + 1, ncy = cy1 - cy0 + 1, ncell = ncx > 0 && ncy > 0 ? ncx * ncy : 0;
    __syncthreads();
    const float gx0 = (o.x + sx) * idx, gy0 = (o.y + sx) * idx, oz = o.z * isz;
    const float dxw = 1.f / idx, e = 1e-5f;
    // Two phases per batch of cells, so that a few large projections do not serialise a warp: (A) one thread per cell culls and
    // computes the pixel bounding box, survivors go to a shared work list; (B) 8-lane groups take cells from the list and sweep
    // the box's pixels in parallel.
    for (int base = 0; base < ncell; base += RASTER_BATCH) {
      if (tid == 0) s_nwork = 0;
      __syncthreads();
      const int lim = min(ncell, base + RASTER_BATCH);
      for (int cell = base + tid; cell < lim; cell += blockDim.x) {
        const int cy = cy0 + cell / ncx, cx = cx0 + cell % ncx;
        const float* h = hf + cy * HN + cx;
        const float h00 = h[0], h10 = h[1], h01 = h[HN], h11 = h[HN + 1];
        float xc[4], yc[4], dp[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const F3 v = f3((float)(cx + (q & 1)) * dxw - sx - o.x, (float)(cy + (q >> 1)) * dxw - sx - o.y, (q == 0 ? h00 : (q == 1 ? h10 : (q == 2 ? h01 : h11))) * p.zscale - o.z);
          xc[q] = fdot(v, ax); yc[q] = fdot(v, ay); dp[q] = -fdot(v, az);
        }
        const float dmin = fminf(fminf(dp[0], dp[1]), fminf(dp[2], dp[3])), dmax = fmaxf(fmaxf(dp[0], dp[1]), fmaxf(dp[2], dp[3]));
        if (dmax <= 1e-3f || dmin >= 1.0f) continue;                     // behind the near plane or beyond the clip (depth is linear on a triangle)
        const float NEAR = 1e-3f;
        float xlo = 3e38f, xhi = -3e38f, ylo = 3e38f, yhi = -3e38f;
#pragma unroll
        for (int q = 0; q < 4; q++) {                                    // quad order 0-1-3-2; edges crossing the near plane add their crossing point
          const int qa = q == 0 ? 0 : (q == 1 ? 1 : (q == 2 ? 3 : 2)), qb = q == 0 ? 1 : (q == 1 ? 3 : (q == 2 ? 2 : 0));
          if (dp[qa] >= NEAR) {
            const float id = __fdividef(1.f, dp[qa]), xq = xc[qa] * id, yq = yc[qa] * id;
            xlo = fminf(xlo, xq); xhi = fmaxf(xhi, xq); ylo = fminf(ylo, yq); yhi = fmaxf(yhi, yq);
          }
          if ((dp[qa] >= NEAR) != (dp[qb] >= NEAR)) {
            const float w = (NEAR - dp[qa]) / (dp[qb] - dp[qa]);
            const float xq = (xc[qa] + w * (xc[qb] - xc[qa])) * (1.f / NEAR), yq = (yc[qa] + w * (yc[qb] - yc[qa])) * (1.f / NEAR);
            xlo = fminf(xlo, xq); xhi = fmaxf(xhi, xq); ylo = fminf(ylo, yq); yhi = fmaxf(yhi, yq);
          }
        }
        if (xlo > aspect || xhi < -aspect || ylo > 1.f || yhi < -1.f) continue;     // outside the view pyramid
        xlo = fmaxf(xlo, -aspect); xhi = fminf(xhi, aspect); ylo = fmaxf(ylo, -1.f); yhi = fminf(yhi, 1.f);
        // pixel centre (c + 0.5) <-> xn = (2 (c + 0.5) / W - 1) aspect; the in-triangle epsilon (1e-5 cells) is far below a pixel
        const int c0 = max(0, (int)ceilf((xlo / aspect + 1.f) * 0.5f * W - 0.5f - 0.01f)), c1 = min(W - 1, (int)floorf((xhi / aspect + 1.f) * 0.5f * W - 0.5f + 0.01f));
        const int r0 = max(0, (int)ceilf((1.f - yhi) * 0.5f * H - 0.5f - 0.01f)), r1 = min(H - 1, (int)floorf((1.f - ylo) * 0.5f * H - 0.5f + 0.01f));
        if (c1 < c0 || r1 < r0) continue;                                // no pixel centre inside the box
        const int slot = atomicAdd(&s_nwork, 1);
        s_work[slot] = make_int2(cell, c0 | (c1 << 8) | (r0 << 16) | (r1 << 24));
      }
      __syncthreads();
      const int nwork = s_nwork;
      for (int wk = tid >> 3; wk < nwork; wk += blockDim.x >> 3) {
        const int2 it = s_work[wk];
        const int cy = cy0 + it.x / ncx, cx = cx0 + it.x % ncx;
        const int c0 = it.y & 255, c1 = (it.y >> 8) & 255, r0 = (it.y >> 16) & 255, r1 = (it.y >> 24) & 255;
        const float* h = hf + cy * HN + cx;
        const float h00 = h[0], h10 = h[1], h01 = h[HN], h11 = h[HN + 1];
        const float a1 = h10 - h00, b1 = h11 - h10, a2 = h11 - h01, b2 = h01 - h00;
        const float u0 = gx0 - (float)cx, v0 = gy0 - (float)cy, cc0 = h00 - oz;
        const float n1 = cc0 + a1 * u0 + b1 * v0, n2 = cc0 + a2 * u0 + b2 * v0;
        const int bw = c1 - c0 + 1, npx = bw * (r1 - r0 + 1);
        for (int q = tid & 7; q < npx; q += 8) {
          const int r = r0 + q / bw, c = c0 + q % bw;
          const float yn = 1.f - 2.f * (r + 0.5f) / H, xn = (2.f * (c + 0.5f) / W - 1.f) * aspect;
          const F3 dir = ax * xn + ay * yn - az;
          const float dgx = dir.x * idx, dgy = dir.y * idx, dz = dir.z * isz;
          float best = -1.f;
          {
            const float den = dz - a1 * dgx - b1 * dgy;
            if (fabsf(den) > 1e-18f) {
              const float t = __fdividef(n1, den), uu = u0 + t * dgx, vv = v0 + t * dgy;
              if (t > 0.f && vv >= -e && uu <= 1.f + e && uu >= vv - e) best = t;
            }
          }
          {
            const float den = dz - a2 * dgx - b2 * dgy;
            if (fabsf(den) > 1e-18f) {
              const float t = __fdividef(n2, den), uu = u0 + t * dgx, vv = v0 + t * dgy;
              if (t > 0.f && uu >= -e && vv <= 1.f + e && vv >= uu - e && (best < 0.f || t < best)) best = t;
            }
          }
          if (best > 1e-4f && best < 1.0f) atomicMin(&zb[r * W + c], __float_as_uint(best));
        }
      }
      __syncthreads();
    }
    __syncthreads();
    for (int px = tid; px < npix; px += blockDim.x) out[px] = __uint_as_float(zb[px]);
  }
}

// --------------------------------------------------------------------------------------------- misc kernels
template <typename T> __global__ void k_set_state(int N, T* st, const double* qpos, const double* qvel, const double* warm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= N) return;
  if (qpos) for (int k = 0; k < NQ; k++) st[(size_t)i * SST + k] = (T)qpos[(size_t)i * NQ + k];
  if (qvel) for (int k = 0; k < NV; k++) st[(size_t)i * SST + NQ + k] = (T)qvel[(size_t)i * NV + k];
  if (warm) for (int k = 0; k < NV; k++) st[(size_t)i * SST + NQ + NV + k] = (T)warm[(size_t)i * NV + k];
}
template <typename T> __global__ void k_get_state(int N, const T* st, double* qpos, double* qvel, double* warm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= N) return;
  if (qpos) for (int k = 0; k < NQ; k++) qpos[(size_t)i * NQ + k] = (double)st[(size_t)i * SST + k];
  if (qvel) for (int k = 0; k < NV; k++) qvel[(size_t)i * NV + k] = (double)st[(size_t)i * SST + NQ + k];
  if (warm) for (int k = 0; k < NV; k++) warm[(size_t)i * NV + k] = (double)st[(size_t)i * SST + NQ + NV + k];
}
template <typename T> __global__ void k_copy_camq(int N, const T* st, T* cq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= N) return;
  for (int k = 0; k < NQ; k++) cq[(size_t)i * CST + k] = st[(size_t)i * SST + k];
}
__global__ void k_scatter_hfield(const int* __restrict__ ids, int n, const float* __restrict__ src, float* __restrict__ dst) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n * HF_CELLS) return;
  const int k = (int)(t / HF_CELLS); const int cell = (int)(t - (size_t)k * HF_CELLS);
  dst[(size_t)ids[k] * HF_CELLS + cell] = src[t];
}
// block maxima of a heightfield: mip cell (br, bc) = max over the vertices of cells [8 br, 8 br + 8) x [8 bc, 8 bc + 8), i.e. vertex
// rows / cols 8 b .. 8 b + 8.  Field f = list[k] (reset list, id list) or k; grid.y strides over the fields.
__global__ void k_hf_mip(const float* __restrict__ hf, float* __restrict__ mip, const int* __restrict__ list, const int* __restrict__ count, int fixed_count) {
  const int n = count ? *count : fixed_count;
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= MIP_CELLS) return;
  const int br = cell / MIPN, bc = cell - br * MIPN;
  const int r1 = min(8 * br + 8, HN - 1), c1 = min(8 * bc + 8, HN - 1);
  for (int k = blockIdx.y; k < n; k += gridDim.y) {
    const size_t f = list ? (size_t)list[k] : (size_t)k;
    const float* h = hf + f * HF_CELLS;
    float m = -1e30f;
    for (int r = 8 * br; r <= r1; r++) for (int c = 8 * bc; c <= c1; c++) m = fmaxf(m, h[r * HN + c]);
    mip[f * MIP_CELLS + cell] = m;
  }
}
__global__ void k_get_hfield(EnvParams p, DevState d, int env, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < HF_CELLS) out[c] = hfOf(p, d, env)[c];
}
__global__ void k_add_reward(int N, float scale, const float* __restrict__ term, bb_io io, float* ep_ret) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= N) return;
  const float add = term[i] * scale;
  io.reward[i] += add;
  if (io.terminated[i]) { if (io.episode_return) io.episode_return[i] += add; } else ep_ret[i] += add;
}
__global__ void k_pack_obs16(int N, bb_io io, float* __restrict__ obs16) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= N) return;
  for (int k = 0; k < 3; k++) {
    obs16[16 * i + k] = io.orientation[3 * i + k]; obs16[16 * i + 3 + k] = io.angular_vel[3 * i + k]; obs16[16 * i + 6 + k] = io.vel[3 * i + k];
    obs16[16 * i + 9 + k] = io.motor_state[3 * i + k]; obs16[16 * i + 12 + k] = io.actions[3 * i + k];
  }
  obs16[16 * i + 15] = io.rel_image_ts[i];
}

// single forward-dynamics evaluation at the current state of one env: solver / contact probe for parity tests
template <typename T> __global__ void k_probe(EnvParams p, DevState d, int env, const double* ctrl3, double* out, double* crows) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const T* st = (const T*)d.st;
  T qpos[NQ], qvel[NV], warm[NV], ctrl[3], qacc[NV];
  for (int k = 0; k < NQ; k++) qpos[k] = st[(size_t)env * SST + k];
  for (int k = 0; k < NV; k++) { qvel[k] = st[(size_t)env * SST + NQ + k]; warm[k] = st[(size_t)env * SST + NQ + NV + k]; }
  for (int k = 0; k < 3; k++) ctrl[k] = (T)ctrl3[k];
  Scratch<T> s; KinOut<T> kin;
  const float* hf = hfOf(p, d, env);
  forwardDynamics(cmc<T>(), qpos, qvel, ctrl, warm, hf, (T)p.zscale, s, qacc, &kin, p.solver_mode != 0);
  for (int k = 0; k < NV; k++) { out[k] = (double)qacc[k]; out[15 + k] = (double)s.qas[k]; out[30 + k] = (double)s.qfs[k]; }
  out[45] = kin.ncon; out[46] = kin.niter; out[47] = (double)cmc<T>().timestep; out[48] = cmc<T>().iterations; out[49] = cmc<T>().ls_iterations;
  out[50] = (double)cmc<T>().meaninertia; out[51] = (double)cmc<T>().K; out[52] = (double)cmc<T>().B; out[53] = (double)cmc<T>().dA[3];
  for (int c = 0; c < s.nc; c++) {
    double* o = crows + BB_CONTACT_STRIDE * c;
    o[0] = s.ctype[c]; o[1] = (double)s.cDist[c];
    for (int j = 0; j < 3; j++) o[2 + j] = (double)s.cP[c][j];
    for (int j = 0; j < 9; j++) o[5 + j] = (double)s.cF[c][j];
  }
}
__global__ void k_probe_count(const double* out, int* ncon) { *ncon = (int)out[45]; }

const unsigned char h_perm[256] = {
    151, 160, 137, 91, 90, 15, 131, 13, 201, 95, 96, 53, 194, 233, 7, 225, 140, 36, 103, 30, 69, 142, 8, 99, 37, 240, 21, 10, 23,
    190, 6, 148, 247, 120, 234, 75, 0, 26, 197, 62, 94, 252, 219, 203, 117, 35, 11, 32, 57, 177, 33, 88, 237, 149, 56, 87, 174, 20,
    125, 136, 171, 168, 68, 175, 74, 165, 71, 134, 139, 48, 27, 166, 77, 146, 158, 231, 83, 111, 229, 122, 60, 211, 133, 230, 220,
    105, 92, 41, 55, 46, 245, 40, 244, 102, 143, 54, 65, 25, 63, 161, 1, 216, 80, 73, 209, 76, 132, 187, 208, 89, 18, 169, 200, 196,
    135, 130, 116, 188, 159, 86, 164, 100, 109, 198, 173, 186, 3, 64, 52, 217, 226, 250, 124, 123, 5, 202, 38, 147, 118, 126, 255,
    82, 85, 212, 207, 206, 59, 227, 47, 16, 58, 17, 182, 189, 28, 42, 223, 183, 170, 213, 119, 248, 152, 2, 44, 154, 163, 70, 221,
    153, 101, 155, 167, 43, 172, 9, 129, 22, 39, 253, 19, 98, 108, 110, 79, 113, 224, 232, 178, 185, 112, 104, 218, 246, 97, 228,
    251, 34, 242, 193, 238, 210, 144, 12, 191, 179, 162, 241, 81, 51, 145, 235, 249, 14, 239, 107, 49, 192, 214, 31, 181, 199, 106,
    157, 184, 84, 204, 176, 115, 121, 50, 45, 127, 4, 150, 254, 138, 236, 205, 93, 222, 114, 67, 29, 24, 72, 243, 141, 128, 195, 78,
    66, 215, 61, 156, 180};

char g_create_error[256] = "";

}  // namespace

// ================================================================================================= engine object
struct bb_engine {
  bb_config cfg;
  EnvParams p;
  DevState d;
  int N;
  int table_n;                     // fields in the Perlin table (HF_TABLE)
  int sm_count;                    // multiprocessors of the device (grid of the persistent solver kernel)
  int raster;                      // depth observation: 1 = k_depth_raster (images of <= 64 x 64 pixels), 0 = k_depth ray-caster
  double* probe_out;               // scratch of bb_get_contacts
  size_t tsize;
  int64_t launches;
  char err[256];
  // optional per-kernel CUDA-event timing (bb_profile_begin/end): 5 events per bb_step
  cudaEvent_t* prof_ev; int prof_cap, prof_n;
  // host-path staging (allocated lazily)
  bool host_ready;
  cudaStream_t hstream;
  float *h_act, *d_act;            // pinned / device actions
  bb_io dio;                       // device output buffers of the host path
  float *d_obs16, *h_obs16, *h_reward, *h_pos2d, *h_term_obs, *h_epret;
  uint8_t *h_term, *h_fail, *d_mask, *h_mask;
  int32_t* h_eplen;
};

#define BB_CUDA(call)                                                                                         \
  do {                                                                                                        \
    cudaError_t _e = (call);                                                                                  \
    if (_e != cudaSuccess) {                                                                                  \
      snprintf(e->err, sizeof(e->err), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return BB_ERR_CUDA;                                                                                     \
    }                                                                                                         \
  } while (0)

static int fail(bb_engine* e, int code, const char* msg) { snprintf(e->err, sizeof(e->err), "%s", msg); return code; }
// Every entry point runs on the engine's own device whatever the caller's current device is, and leaves the caller's
// current device untouched (two engines on different GPUs in one process; torch's current device != cfg.device).
struct DevGuard {
  int prev; bool sw;
  explicit DevGuard(int dev) : prev(-1), sw(false) { if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) { cudaSetDevice(dev); sw = true; } }
  ~DevGuard() { if (sw) cudaSetDevice(prev); }
};
static inline int blocksFor(int n, int b) { return (n + b - 1) / b; }

static int checkIo(bb_engine* e, const bb_io* io) {
  if (!io || !io->orientation || !io->angular_vel || !io->vel || !io->motor_state || !io->actions || !io->rel_image_ts || !io->reward ||
      !io->terminated || !io->failure || !io->pos2d)
    return fail(e, BB_ERR_INVALID, "bb_io: required output pointer is NULL");
  if (e->cfg.cameras && (!io->rgbd_0 || !io->rgbd_1)) return fail(e, BB_ERR_INVALID, "bb_io: rgbd_0/rgbd_1 required when cameras are enabled");
  return BB_OK;
}

// persistent grid of k_depth: 4 CTAs of 8 warps per SM, never more warps than work units
static inline int depthGrid(int N) { const int need = (N * 2 * DEPTH_PARTS + 7) / 8; return need < 148 * 4 ? need : 148 * 4; }
// terrain regeneration + state reset + depth refresh for whatever is in the work lists
static int launchResetAndRender(bb_engine* e, const bb_io* io, cudaStream_t s, bool do_reset, cudaEvent_t* ev = nullptr) {
  const int N = e->N;
  if (do_reset) {
    if (e->cfg.terrain_type == BB_TERRAIN_PERLIN && e->p.hf_mode == HF_PER_ENV) {
      dim3 grid(blocksFor(HF_CELLS, 256), N < 128 ? N : 128);
      k_terrain<<<grid, 256, 0, s>>>(e->p, e->d, e->d.reset_list, e->d.counters, 0, nullptr, nullptr);
      k_hf_mip<<<dim3(blocksFor(MIP_CELLS, 128), N < 128 ? N : 128), 128, 0, s>>>(e->d.hfield, e->d.hmip, e->d.reset_list, e->d.counters, 0);
      e->launches += 2;
    }
    if (ev) cudaEventRecord(ev[2], s);
    if (e->cfg.precision == 64) k_reset<double><<<blocksFor(N, 128), 128, 0, s>>>(e->p, e->d, *io);
    else k_reset<float><<<blocksFor(N, 128), 128, 0, s>>>(e->p, e->d, *io);
    e->launches++;
  } else if (ev) cudaEventRecord(ev[2], s);
  if (ev) cudaEventRecord(ev[3], s);
  if (e->cfg.cameras) {
    const int grid = depthGrid(N);
    if (e->raster) {
      const int rgrid = N * 2 < 148 * 8 ? N * 2 : 148 * 8;
      if (e->cfg.precision == 64) k_depth_raster<double><<<rgrid, 128, 0, s>>>(e->p, e->d, e->d.refresh_list, e->d.counters + 1, 0, (const double*)e->d.camq, CST, io->rgbd_0, io->rgbd_1);
      else k_depth_raster<float><<<rgrid, 128, 0, s>>>(e->p, e->d, e->d.refresh_list, e->d.counters + 1, 0, (const float*)e->d.camq, CST, io->rgbd_0, io->rgbd_1);
    } else if (e->cfg.precision == 64) k_depth<double><<<grid, 256, 0, s>>>(e->p, e->d, e->d.refresh_list, e->d.counters + 1, 0, (const double*)e->d.camq, CST, io->rgbd_0, io->rgbd_1);
    else k_depth<float><<<grid, 256, 0, s>>>(e->p, e->d, e->d.refresh_list, e->d.counters + 1, 0, (const float*)e->d.camq, CST, io->rgbd_0, io->rgbd_1);
    e->launches++;
  }
  if (ev) cudaEventRecord(ev[4], s);
  return BB_OK;
}

extern "C" {

void bb_default_config(bb_config* c) {
  memset(c, 0, sizeof(*c));
  c->abi_version = BB_ABI_VERSION; c->num_envs = 1; c->env_offset = 0; c->device = 0; c->precision = 64;
  c->terrain_type = BB_TERRAIN_PERLIN; c->terrain_seed = -1;
  c->perlin_scale = 25.f; c->perlin_octaves = 4; c->perlin_persistence = 0.2f; c->perlin_lacunarity = 2.f; c->perlin_amplitude = 1.f;
  c->hfield_zscale = 2.f; c->cameras = 1; c->im_h = 64; c->im_w = 64; c->camera_frame_rate = 90.f;
  c->max_ep_steps = 4000; c->max_allowed_tilt = 20.f; c->max_wheel_velocity = 10.f;
  c->reward_type = BB_REWARD_DIRECTIONAL; c->reward_scale = 0.01f; c->action_reg_coef = -0.0001f; c->survival_bonus = 0.02f;
  c->target_direction[0] = 0.f; c->target_direction[1] = 1.f; c->goal_position[0] = 0.f; c->goal_position[1] = 0.f; c->distance_scale = 1.f;
  c->seed = 0; c->auto_reset = 1; c->step_kernel = 0; c->solver_mode = 0; c->perlin_table = -1; c->seed_stream = 0; c->depth_kernel = 1;
}

const char* bb_last_error(const bb_engine* e) { return e ? e->err : g_create_error; }
int bb_num_envs(const bb_engine* e) { return e ? e->N : 0; }
int64_t bb_launch_count(const bb_engine* e) { return e ? e->launches : 0; }

const char* bb_build_info(void) {
#ifndef BB_SRC_HASH
#define BB_SRC_HASH "unknown"
#endif
  return "libballbot_b200 abi 2 built " __DATE__ " " __TIME__ " src " BB_SRC_HASH;
}
int bb_model_constants(double* dA12, double* meaninertia, double* masses3) {
  ModelConst<double> m; buildModelConst(m);
  if (dA12) for (int i = 0; i < NCT; i++) dA12[i] = m.dA[i];
  if (meaninertia) *meaninertia = m.meaninertia;
  if (masses3) { masses3[0] = m.m0; masses3[1] = m.mw; masses3[2] = m.mL; }
  return BB_OK;
}

int bb_create(const bb_config* cfg, bb_engine** out) {
  if (!cfg || !out) { snprintf(g_create_error, sizeof(g_create_error), "bb_create: NULL argument"); return BB_ERR_INVALID; }
  *out = nullptr;
  if (cfg->abi_version != BB_ABI_VERSION || cfg->num_envs < 1 || (cfg->precision != 32 && cfg->precision != 64) || cfg->im_h < 1 || cfg->im_w < 1 ||
      cfg->terrain_type < 0 || cfg->terrain_type > 4 || cfg->reward_type < 0 || cfg->reward_type > 2 || cfg->camera_frame_rate <= 0.f ||
      cfg->perlin_table < -1 || cfg->perlin_table > 1 || cfg->seed_stream < 0 || cfg->seed_stream > 1 || cfg->perlin_octaves < 1) {
    snprintf(g_create_error, sizeof(g_create_error), "bb_create: invalid config (abi %d, num_envs %d, precision %d)", cfg->abi_version, cfg->num_envs, cfg->precision);
    return BB_ERR_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= cfg->device) {
    snprintf(g_create_error, sizeof(g_create_error), "bb_create: CUDA device %d not available (%d devices); this engine has no CPU fallback", cfg->device, ndev);
    return BB_ERR_NO_DEVICE;
  }
  bb_engine* e = new (std::nothrow) bb_engine();
  if (!e) return BB_ERR_INVALID;
  memset(e, 0, sizeof(*e));
  e->cfg = *cfg; e->N = cfg->num_envs;
  const int N = e->N;
  auto bail = [&](int code) { snprintf(g_create_error, sizeof(g_create_error), "%s", e->err); bb_destroy(e); return code; };
#define BB_CUDA_C(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { snprintf(e->err, sizeof(e->err), "%s failed: %s", #call, cudaGetErrorString(_e)); return bail(BB_ERR_CUDA); } } while (0)
  DevGuard guard(cfg->device);
  {
    ModelConst<double> m64; buildModelConst(m64);
    ModelConst<float> m32; narrowModel(m64, m32);
    BB_CUDA_C(cudaMemcpyToSymbol(c_mc64, &m64, sizeof(m64)));
    BB_CUDA_C(cudaMemcpyToSymbol(c_mc32, &m32, sizeof(m32)));
    BB_CUDA_C(cudaMemcpyToSymbol(c_perm, h_perm, sizeof(h_perm)));

  }
  EnvParams& p = e->p;
  p.N = N; p.env_offset = cfg->env_offset; p.cameras = cfg->cameras; p.im_h = cfg->im_h; p.im_w = cfg->im_w;
  {  // camera cadence: smallest k with k*dt >= 1/frame_rate, timestamps accumulated like mjData.time (ballbot_env.py:745-750)
    double t = 0; int k = 0; const double want = 1.0 / (double)cfg->camera_frame_rate;
    do { t += 0.002; k++; } while (t < want && k < 100000);
    p.cam_period = k;
  }
  p.max_ep_steps = cfg->max_ep_steps; p.max_tilt = cfg->max_allowed_tilt; p.max_wheel_vel = cfg->max_wheel_velocity;
  p.reward_type = cfg->reward_type; p.reward_scale = cfg->reward_scale; p.action_reg = cfg->action_reg_coef; p.survival = cfg->survival_bonus;
  p.tdir[0] = cfg->target_direction[0]; p.tdir[1] = cfg->target_direction[1]; p.goal[0] = cfg->goal_position[0]; p.goal[1] = cfg->goal_position[1];
  p.dist_scale = cfg->distance_scale; p.zscale = cfg->hfield_zscale; p.terrain_type = cfg->terrain_type; p.terrain_seed = cfg->terrain_seed;
  p.seed = cfg->seed; p.auto_reset = cfg->auto_reset; p.solver_mode = cfg->solver_mode; p.seed_stream = cfg->seed_stream;
  p.hf_mode = cfg->terrain_type == BB_TERRAIN_EXTERNAL ? HF_PER_ENV : HF_SHARED;
  if (cfg->terrain_type == BB_TERRAIN_TABLE) {   // caller-provided fields for every seed 0 .. BB_PERLIN_SEEDS - 1 (seed-dependent plugin terrains)
    p.hf_mode = HF_TABLE; e->table_n = cfg->terrain_seed >= 0 ? 1 : BB_PERLIN_SEEDS;
  }
  if (cfg->terrain_type == BB_TERRAIN_PERLIN) {
    // a fixed terrain seed means ONE field for every env and every episode; random seeds mean at most BB_PERLIN_SEEDS fields
    const bool table = cfg->terrain_seed >= 0 || cfg->perlin_table == 1 || (cfg->perlin_table == -1 && N >= 2048);
    p.hf_mode = table ? HF_TABLE : HF_PER_ENV;
    e->table_n = table ? (cfg->terrain_seed >= 0 ? 1 : BB_PERLIN_SEEDS) : 0;
  }
  p.pscale = cfg->perlin_scale; p.ppers = cfg->perlin_persistence; p.plac = cfg->perlin_lacunarity; p.pamp = cfg->perlin_amplitude; p.poct = cfg->perlin_octaves;
  e->tsize = cfg->precision == 64 ? 8 : 4;
  e->raster = cfg->depth_kernel != 1 && cfg->im_h * cfg->im_w <= RASTER_MAX_PIX;
  e->sm_count = 148;
  { int smc = 0; if (cudaDeviceGetAttribute(&smc, cudaDevAttrMultiProcessorCount, cfg->device) == cudaSuccess && smc > 0) e->sm_count = smc; }
  DevState& d = e->d;
  BB_CUDA_C(cudaMalloc(&d.st, e->tsize * SST * N));
  BB_CUDA_C(cudaMalloc(&d.camq, e->tsize * CST * N));
  BB_CUDA_C(cudaMalloc(&d.gscr, e->tsize * (size_t)bbg::GSCR * N));
  BB_CUDA_C(cudaMalloc(&d.rk, e->tsize * (size_t)bbg::RKN * N)); BB_CUDA_C(cudaMalloc(&d.ctx, e->tsize * (size_t)bbg::CTXN * N));
  BB_CUDA_C(cudaMalloc(&d.meta, sizeof(int) * 4 * (size_t)N));
  BB_CUDA_C(cudaMemset(d.rk, 0, e->tsize * (size_t)bbg::RKN * N)); BB_CUDA_C(cudaMemset(d.meta, 0, sizeof(int) * 4 * (size_t)N));
  BB_CUDA_C(cudaMalloc(&d.step_count, sizeof(int) * N)); BB_CUDA_C(cudaMalloc(&d.cam_steps, sizeof(int) * N));
  BB_CUDA_C(cudaMalloc(&d.episode, sizeof(unsigned) * N)); BB_CUDA_C(cudaMalloc(&d.tseed, sizeof(int) * N));
  BB_CUDA_C(cudaMalloc(&d.ep_ret, sizeof(float) * N)); BB_CUDA_C(cudaMalloc(&d.ep_len, sizeof(int) * N));
  BB_CUDA_C(cudaMalloc(&d.counters, sizeof(int) * 16));
  BB_CUDA_C(cudaMalloc(&d.work, sizeof(int) * N)); BB_CUDA_C(cudaMalloc(&d.order, sizeof(int) * N)); BB_CUDA_C(cudaMalloc(&d.alist, sizeof(int) * N)); BB_CUDA_C(cudaMalloc(&d.bins, sizeof(int) * 2 * WORK_BINS));
  {  // CTAs of more than one warp may need more than the default 48 KB of dynamic shared memory
    const int epb = BB_WPB * bbg::EPW;
    const int sm64 = (int)(epb * sizeof(bbg::GS<double>)), sm32 = (int)(epb * sizeof(bbg::GS<float>));
    BB_CUDA_C(cudaFuncSetAttribute(k_step_warp<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm64));
    BB_CUDA_C(cudaFuncSetAttribute(k_step_warp<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm32));
    const int sepb = BB_WPB_STAGE * bbg::EPW;
    BB_CUDA_C(cudaFuncSetAttribute(k_stage<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sepb * sizeof(bbg::GS<double>))));
    BB_CUDA_C(cudaFuncSetAttribute(k_stage<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sepb * sizeof(bbg::GS<float>))));
    BB_CUDA_C(cudaFuncSetAttribute(k_newton<double, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm64));
    BB_CUDA_C(cudaFuncSetAttribute(k_newton<float, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm32));
    BB_CUDA_C(cudaFuncSetAttribute(k_newton<double, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm64));
    BB_CUDA_C(cudaFuncSetAttribute(k_newton<float, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm32));
    BB_CUDA_C(cudaFuncSetAttribute(k_newton_p<double, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm64));
    BB_CUDA_C(cudaFuncSetAttribute(k_newton_p<float, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm32));
    BB_CUDA_C(cudaFuncSetAttribute(k_newton_p<double, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm64));
    BB_CUDA_C(cudaFuncSetAttribute(k_newton_p<float, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm32));
  }
  k_init_order<<<blocksFor(N > 2 * WORK_BINS ? N : 2 * WORK_BINS, 256), 256>>>(N, d);
  BB_CUDA_C(cudaMalloc(&d.reset_list, sizeof(int) * N)); BB_CUDA_C(cudaMalloc(&d.refresh_list, sizeof(int) * N));
  const size_t hfbytes = sizeof(float) * HF_CELLS * (p.hf_mode == HF_PER_ENV ? (size_t)N : (p.hf_mode == HF_TABLE ? (size_t)e->table_n : 1));
  BB_CUDA_C(cudaMalloc(&d.hfield, hfbytes));
  if (p.hf_mode != HF_TABLE || cfg->terrain_type == BB_TERRAIN_TABLE) BB_CUDA_C(cudaMemset(d.hfield, 0, hfbytes));
  BB_CUDA_C(cudaMalloc(&d.ptab, sizeof(float) * 2 * HN));
  k_perlin_table<<<blocksFor(HN, 128), 128>>>(cfg->perlin_scale, d.ptab);
  const size_t nfields = p.hf_mode == HF_PER_ENV ? (size_t)N : (p.hf_mode == HF_TABLE ? (size_t)e->table_n : 1);
  BB_CUDA_C(cudaMalloc(&d.hmip, sizeof(float) * MIP_CELLS * nfields));
  BB_CUDA_C(cudaMemset(d.hmip, 0, sizeof(float) * MIP_CELLS * nfields));
  BB_CUDA_C(cudaMalloc(&d.rng, sizeof(unsigned long long) * 5 * (size_t)N));
  BB_CUDA_C(cudaMemset(d.rng, 0, sizeof(unsigned long long) * 5 * (size_t)N));
  BB_CUDA_C(cudaMalloc(&e->probe_out, sizeof(double) * 64));
  if (p.hf_mode == HF_TABLE && cfg->terrain_type == BB_TERRAIN_PERLIN) {
    // every Perlin field the reference can ever draw (r_seed in 0 .. 9999, ballbot_env.py:506), generated once by the same
    // kernel that regenerates per-env fields: resets then only select a field, and HBM use no longer grows with num_envs
    int* seeds = nullptr;
    BB_CUDA_C(cudaMalloc(&seeds, sizeof(int) * e->table_n));
    k_iota<<<blocksFor(e->table_n, 256), 256>>>(e->table_n, cfg->terrain_seed >= 0 ? cfg->terrain_seed : 0, seeds);
    dim3 grid(blocksFor(HF_CELLS, 256), e->table_n < 592 ? e->table_n : 592);
    k_terrain<<<grid, 256>>>(p, d, nullptr, nullptr, e->table_n, seeds, d.hfield);
    k_hf_mip<<<dim3(blocksFor(MIP_CELLS, 128), e->table_n < 592 ? e->table_n : 592), 128>>>(d.hfield, d.hmip, nullptr, nullptr, e->table_n);
    BB_CUDA_C(cudaDeviceSynchronize());
    cudaFree(seeds);
  }
  BB_CUDA_C(cudaMemset(d.st, 0, e->tsize * SST * N)); BB_CUDA_C(cudaMemset(d.camq, 0, e->tsize * CST * N));
  BB_CUDA_C(cudaMemset(d.step_count, 0, sizeof(int) * N)); BB_CUDA_C(cudaMemset(d.cam_steps, 0, sizeof(int) * N));
  BB_CUDA_C(cudaMemset(d.episode, 0, sizeof(unsigned) * N)); BB_CUDA_C(cudaMemset(d.tseed, 0, sizeof(int) * N));
  BB_CUDA_C(cudaMemset(d.ep_ret, 0, sizeof(float) * N)); BB_CUDA_C(cudaMemset(d.ep_len, 0, sizeof(int) * N));
  BB_CUDA_C(cudaMemset(d.counters, 0, sizeof(int) * 16));
  BB_CUDA_C(cudaDeviceSynchronize());
#undef BB_CUDA_C
  *out = e;
  return BB_OK;
}

int bb_destroy(bb_engine* e) {
  if (!e) return BB_OK;
  DevGuard guard(e->cfg.device);
  DevState& d = e->d;
  cudaFree(d.st); cudaFree(d.camq); cudaFree(d.gscr); cudaFree(d.rk); cudaFree(d.ctx); cudaFree(d.meta); cudaFree(d.step_count); cudaFree(d.cam_steps); cudaFree(d.episode); cudaFree(d.tseed);
  cudaFree(d.work); cudaFree(d.order); cudaFree(d.bins); cudaFree(d.alist);
  cudaFree(d.ep_ret); cudaFree(d.ep_len); cudaFree(d.counters); cudaFree(d.reset_list); cudaFree(d.refresh_list); cudaFree(d.hfield); cudaFree(d.hmip); cudaFree(d.ptab); cudaFree(d.rng); cudaFree(e->probe_out);
  if (e->prof_ev) { for (int i = 0; i < 5 * e->prof_cap; i++) cudaEventDestroy(e->prof_ev[i]); free(e->prof_ev); }
  if (e->host_ready) {
    cudaFreeHost(e->h_act); cudaFree(e->d_act); cudaFree(e->d_obs16); cudaFreeHost(e->h_obs16); cudaFreeHost(e->h_reward); cudaFreeHost(e->h_pos2d);
    cudaFreeHost(e->h_term_obs); cudaFreeHost(e->h_epret); cudaFreeHost(e->h_term); cudaFreeHost(e->h_fail); cudaFreeHost(e->h_eplen);
    cudaFree(e->d_mask); cudaFreeHost(e->h_mask);
    bb_io& o = e->dio;
    cudaFree(o.orientation); cudaFree(o.angular_vel); cudaFree(o.vel); cudaFree(o.motor_state); cudaFree(o.actions); cudaFree(o.rel_image_ts);
    cudaFree(o.rgbd_0); cudaFree(o.rgbd_1); cudaFree(o.reward); cudaFree(o.terminated); cudaFree(o.failure); cudaFree(o.pos2d);
    cudaFree(o.terminal_obs); cudaFree(o.episode_return); cudaFree(o.episode_length); cudaFree(o.status);
    cudaStreamDestroy(e->hstream);
  }
  delete e;
  return BB_OK;
}

int bb_reset(bb_engine* e, const uint8_t* mask_dev, const int32_t* seeds_dev, const bb_io* io, void* stream) {
  if (!e) return BB_ERR_INVALID;
  int rc = checkIo(e, io); if (rc) return rc;
  DevGuard guard(e->cfg.device);
  cudaStream_t s = (cudaStream_t)stream;
  const int N = e->N;
  k_clear_counters<<<1, 1, 0, s>>>(e->d);
  k_mask_to_list<<<blocksFor(N, 256), 256, 0, s>>>(e->p, e->d, mask_dev, seeds_dev);
  e->launches += 2;
  rc = launchResetAndRender(e, io, s, true); if (rc) return rc;
  BB_CUDA(cudaGetLastError());
  return BB_OK;
}

int bb_set_rng_state(bb_engine* e, const uint64_t* state_dev, void* stream) {
  if (!e || !state_dev) return BB_ERR_INVALID;
  if (e->cfg.seed_stream != 1) return fail(e, BB_ERR_STATE, "bb_set_rng_state: engine was created with seed_stream = 0 (counter-based terrain seeds)");
  DevGuard guard(e->cfg.device);
  BB_CUDA(cudaMemcpyAsync(e->d.rng, state_dev, sizeof(unsigned long long) * 5 * (size_t)e->N, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return BB_OK;
}

int bb_step(bb_engine* e, const float* actions_dev, const bb_io* io, void* stream) {
  if (!e || !actions_dev) return e ? fail(e, BB_ERR_INVALID, "bb_step: actions is NULL") : BB_ERR_INVALID;
  int rc = checkIo(e, io); if (rc) return rc;
  DevGuard guard(e->cfg.device);
  cudaStream_t s = (cudaStream_t)stream;
  const int N = e->N;
  const int bs = N >= 148 * 64 * 2 ? 64 : 32;
  cudaEvent_t* ev = nullptr;
  if (e->prof_ev && e->prof_n < e->prof_cap) { ev = e->prof_ev + 5 * (size_t)e->prof_n; e->prof_n++; }
  if (e->cfg.step_kernel == 1) k_clear_counters<<<1, 1, 0, s>>>(e->d);
  else { k_begin_step<<<1, WORK_BINS, 0, s>>>(e->d); k_order<<<blocksFor(N, 256), 256, 0, s>>>(N, e->d); e->launches++; }
  if (ev) cudaEventRecord(ev[0], s);
  if (e->cfg.step_kernel == 1) {   // thread-per-env reference mapping
    if (e->cfg.precision == 64) k_step<double><<<blocksFor(N, bs), bs, 0, s>>>(e->p, e->d, actions_dev, *io);
    else k_step<float><<<blocksFor(N, bs), bs, 0, s>>>(e->p, e->d, actions_dev, *io);
  } else {                         // lane group per env: split-phase (default) or fused
    const int wpb = BB_WPB;
    const int epb = wpb * bbg::EPW;   // envs per CTA
    const int grid = blocksFor(N, epb), bt = wpb * 32;
    const size_t sm64 = epb * sizeof(bbg::GS<double>), sm32 = epb * sizeof(bbg::GS<float>);
    if (e->cfg.step_kernel == 2) {
      if (e->cfg.precision == 64) k_step_warp<double><<<grid, bt, sm64, s>>>(e->p, e->d, actions_dev, *io);
      else k_step_warp<float><<<grid, bt, sm32, s>>>(e->p, e->d, actions_dev, *io);
    } else {
      const int sepb = BB_WPB_STAGE * bbg::EPW, sgrid = blocksFor(N, sepb);
      for (int stage = 0; stage <= 4; stage++) {
        if (e->cfg.precision == 64) k_stage<double><<<sgrid, BB_WPB_STAGE * 32, sepb * sizeof(bbg::GS<double>), s>>>(e->p, e->d, actions_dev, *io, stage);
        else k_stage<float><<<sgrid, BB_WPB_STAGE * 32, sepb * sizeof(bbg::GS<float>), s>>>(e->p, e->d, actions_dev, *io, stage);
        if (stage == 4) break;
        const bool fs = e->cfg.solver_mode != 0;
#ifndef BB_NEWTON_PERSISTENT
#define BB_NEWTON_PERSISTENT 1
#endif
        if (BB_NEWTON_PERSISTENT) {   // persistent groups drawing from the stage's work list: one wave of resident CTAs
          const int res = e->sm_count * (e->cfg.precision == 64 ? BB_NEWTON_MINBLOCKS : BB_WARP_MINBLOCKS32);
          const int pg = grid < res ? grid : res;
          if (e->cfg.precision == 64) { if (fs) k_newton_p<double, 1><<<pg, bt, sm64, s>>>(e->p, e->d, stage); else k_newton_p<double, 0><<<pg, bt, sm64, s>>>(e->p, e->d, stage); }
          else { if (fs) k_newton_p<float, 1><<<pg, bt, sm32, s>>>(e->p, e->d, stage); else k_newton_p<float, 0><<<pg, bt, sm32, s>>>(e->p, e->d, stage); }
        } else if (e->cfg.precision == 64) { if (fs) k_newton<double, 1><<<grid, bt, sm64, s>>>(e->p, e->d, stage); else k_newton<double, 0><<<grid, bt, sm64, s>>>(e->p, e->d, stage); }
        else { if (fs) k_newton<float, 1><<<grid, bt, sm32, s>>>(e->p, e->d, stage); else k_newton<float, 0><<<grid, bt, sm32, s>>>(e->p, e->d, stage); }
      }
      e->launches += 8;
    }
  }
  if (ev) cudaEventRecord(ev[1], s);
  e->launches += 2;
  rc = launchResetAndRender(e, io, s, e->cfg.auto_reset != 0, ev); if (rc) return rc;
  BB_CUDA(cudaGetLastError());
  return BB_OK;
}

int bb_add_reward(bb_engine* e, const float* term_dev, const bb_io* io, void* stream) {
  if (!e || !term_dev) return BB_ERR_INVALID;
  DevGuard guard(e->cfg.device);
  int rc = checkIo(e, io); if (rc) return rc;
  k_add_reward<<<blocksFor(e->N, 256), 256, 0, (cudaStream_t)stream>>>(e->N, e->cfg.reward_scale, term_dev, *io, e->d.ep_ret);
  e->launches++;
  BB_CUDA(cudaGetLastError());
  return BB_OK;
}

int bb_set_state(bb_engine* e, const double* qpos, const double* qvel, const double* warm, void* stream) {
  if (!e) return BB_ERR_INVALID;
  DevGuard guard(e->cfg.device);
  cudaStream_t s = (cudaStream_t)stream; const int N = e->N;
  if (e->cfg.precision == 64) { k_set_state<double><<<blocksFor(N, 128), 128, 0, s>>>(N, (double*)e->d.st, qpos, qvel, warm); k_copy_camq<double><<<blocksFor(N, 128), 128, 0, s>>>(N, (const double*)e->d.st, (double*)e->d.camq); }
  else { k_set_state<float><<<blocksFor(N, 128), 128, 0, s>>>(N, (float*)e->d.st, qpos, qvel, warm); k_copy_camq<float><<<blocksFor(N, 128), 128, 0, s>>>(N, (const float*)e->d.st, (float*)e->d.camq); }
  e->launches += 2;
  BB_CUDA(cudaGetLastError());
  return BB_OK;
}
int bb_get_state(bb_engine* e, double* qpos, double* qvel, double* warm, void* stream) {
  if (!e) return BB_ERR_INVALID;
  DevGuard guard(e->cfg.device);
  cudaStream_t s = (cudaStream_t)stream; const int N = e->N;
  if (e->cfg.precision == 64) k_get_state<double><<<blocksFor(N, 128), 128, 0, s>>>(N, (const double*)e->d.st, qpos, qvel, warm);
  else k_get_state<float><<<blocksFor(N, 128), 128, 0, s>>>(N, (const float*)e->d.st, qpos, qvel, warm);
  e->launches++;
  BB_CUDA(cudaGetLastError());
  return BB_OK;
}
int bb_set_hfield(bb_engine* e, const int32_t* ids, int32_t n, const float* hf, void* stream) {
  if (!e || !ids || !hf || n < 0) return BB_ERR_INVALID;
  DevGuard guard(e->cfg.device);
  if (e->cfg.terrain_type == BB_TERRAIN_SHARED) {   // the one field shared by every env
    if (n != 1) return fail(e, BB_ERR_INVALID, "bb_set_hfield: a shared-terrain engine takes exactly one heightfield");
    BB_CUDA(cudaMemcpyAsync(e->d.hfield, hf, sizeof(float) * HF_CELLS, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    k_hf_mip<<<dim3(blocksFor(MIP_CELLS, 128), 1), 128, 0, (cudaStream_t)stream>>>(e->d.hfield, e->d.hmip, nullptr, nullptr, 1);
    e->launches += 2;
    BB_CUDA(cudaGetLastError());
    return BB_OK;
  }
  if (e->cfg.terrain_type == BB_TERRAIN_TABLE) {   // ids = table slots (terrain seeds); host-side bound check is the caller's (ids live on the device)
    if (n > e->table_n) return fail(e, BB_ERR_INVALID, "bb_set_hfield: more fields than table slots");
  } else if (e->p.hf_mode != HF_PER_ENV) return fail(e, BB_ERR_INVALID, "bb_set_hfield: engine has no per-env heightfields (flat terrain or Perlin table); create it with BB_TERRAIN_EXTERNAL");
  if (n == 0) return BB_OK;
  const size_t tot = (size_t)n * HF_CELLS;
  k_scatter_hfield<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ids, n, hf, e->d.hfield);
  k_hf_mip<<<dim3(blocksFor(MIP_CELLS, 128), n < 128 ? n : 128), 128, 0, (cudaStream_t)stream>>>(e->d.hfield, e->d.hmip, ids, nullptr, n);
  e->launches += 2;
  BB_CUDA(cudaGetLastError());
  return BB_OK;
}
int bb_get_hfield(bb_engine* e, int32_t env, float* out, void* stream) {
  if (!e || !out || env < 0 || env >= e->N) return BB_ERR_INVALID;
  DevGuard guard(e->cfg.device);
  k_get_hfield<<<blocksFor(HF_CELLS, 256), 256, 0, (cudaStream_t)stream>>>(e->p, e->d, env, out);   // the field index of table mode lives on the device
  e->launches++;
  BB_CUDA(cudaGetLastError());
  return BB_OK;
}
int bb_get_terrain_seeds(bb_engine* e, int32_t* seeds, void* stream) {
  if (!e || !seeds) return BB_ERR_INVALID;
  DevGuard guard(e->cfg.device);
  BB_CUDA(cudaMemcpyAsync(seeds, e->d.tseed, sizeof(int) * e->N, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return BB_OK;
}
int bb_perlin_terrain(bb_engine* e, const int32_t* seeds, int32_t n, float* out, void* stream) {
  if (!e || !seeds || !out || n < 1) return BB_ERR_INVALID;
  DevGuard guard(e->cfg.device);
  dim3 grid(blocksFor(HF_CELLS, 256), n < 128 ? n : 128);
  k_terrain<<<grid, 256, 0, (cudaStream_t)stream>>>(e->p, e->d, nullptr, nullptr, n, seeds, out);
  e->launches++;
  BB_CUDA(cudaGetLastError());
  return BB_OK;
}
int bb_render_depth(bb_engine* e, float* img0, float* img1, void* stream) {
  if (!e || !img0 || !img1) return BB_ERR_INVALID;
  DevGuard guard(e->cfg.device);
  const int N = e->N; cudaStream_t s = (cudaStream_t)stream;
  const int grid = depthGrid(N);
  BB_CUDA(cudaMemsetAsync(e->d.counters + 2, 0, sizeof(int), s));   // work-unit cursor of k_depth
  if (e->raster) {
    const int rgrid = N * 2 < 148 * 8 ? N * 2 : 148 * 8;
    if (e->cfg.precision == 64) k_depth_raster<double><<<rgrid, 128, 0, s>>>(e->p, e->d, nullptr, nullptr, N, (const double*)e->d.st, SST, img0, img1);
    else k_depth_raster<float><<<rgrid, 128, 0, s>>>(e->p, e->d, nullptr, nullptr, N, (const float*)e->d.st, SST, img0, img1);
  } else if (e->cfg.precision == 64) k_depth<double><<<grid, 256, 0, s>>>(e->p, e->d, nullptr, nullptr, N, (const double*)e->d.st, SST, img0, img1);
  else k_depth<float><<<grid, 256, 0, s>>>(e->p, e->d, nullptr, nullptr, N, (const float*)e->d.st, SST, img0, img1);
  e->launches++;
  BB_CUDA(cudaGetLastError());
  return BB_OK;
}

// replaces reading mjData.contact / qacc / solver_niter after mj_forward in a debugger (parity tests): one forward-dynamics
// evaluation for env `env` at its current state; out_dev double[64], contact arrays double[53 | 159 | 477] (device)
int bb_probe_forward(bb_engine* e, int32_t env, const double* ctrl3_dev, double* out_dev, double* contacts_dev, void* stream) {
  if (!e || env < 0 || env >= e->N || !ctrl3_dev || !out_dev || !contacts_dev) return BB_ERR_INVALID;
  DevGuard guard(e->cfg.device);
  if (e->cfg.step_kernel == 1) {
    if (e->cfg.precision == 64) k_probe<double><<<1, 32, 0, (cudaStream_t)stream>>>(e->p, e->d, env, ctrl3_dev, out_dev, contacts_dev);
    else k_probe<float><<<1, 32, 0, (cudaStream_t)stream>>>(e->p, e->d, env, ctrl3_dev, out_dev, contacts_dev);
  } else {   // lane-group path: the production collision / solver code (gForward), contacts written by the lanes that found them
    if (e->cfg.precision == 64) k_probe_warp<double><<<1, 32, sizeof(bbg::GS<double>), (cudaStream_t)stream>>>(e->p, e->d, env, ctrl3_dev, out_dev, contacts_dev);
    else k_probe_warp<float><<<1, 32, sizeof(bbg::GS<float>), (cudaStream_t)stream>>>(e->p, e->d, env, ctrl3_dev, out_dev, contacts_dev);
  }
  e->launches++;
  BB_CUDA(cudaGetLastError());
  return BB_OK;
}
__global__ void k_zero3(double* p) { if (threadIdx.x < 3) p[threadIdx.x] = 0.0; }
int bb_get_contacts(bb_engine* e, int32_t env, double* contacts_dev, int32_t* ncon_dev, void* stream) {
  if (!e || env < 0 || env >= e->N || !contacts_dev || !ncon_dev) return BB_ERR_INVALID;
  DevGuard guard(e->cfg.device);
  double* ctrl = e->probe_out + 56;    // three zeros behind the probe outputs
  k_zero3<<<1, 32, 0, (cudaStream_t)stream>>>(ctrl);
  int rc = bb_probe_forward(e, env, ctrl, e->probe_out, contacts_dev, stream); if (rc) return rc;
  k_probe_count<<<1, 1, 0, (cudaStream_t)stream>>>(e->probe_out, ncon_dev);
  e->launches += 2;
  BB_CUDA(cudaGetLastError());
  return BB_OK;
}

// host-in / host-out Perlin heightfields for arbitrary n (plugin-API callable): replaces generate_perlin_terrain
int bb_perlin_grid(int32_t device, int32_t n, float scale, int32_t octaves, float persistence, float lacunarity, float amplitude,
                   const int32_t* seeds_host, int32_t nseeds, float* out_host) {
  if (n < 1 || nseeds < 1 || !seeds_host || !out_host || octaves < 1) return BB_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device) { snprintf(g_create_error, sizeof(g_create_error), "bb_perlin_grid: CUDA device %d not available; no CPU fallback", device); return BB_ERR_NO_DEVICE; }
  DevGuard guard(device);
  if (cudaMemcpyToSymbol(c_perm, h_perm, sizeof(h_perm)) != cudaSuccess) return BB_ERR_CUDA;
  int* dseeds = nullptr; float* dout = nullptr; const size_t cells = (size_t)n * n;
  int rc = BB_OK;
  if (cudaMalloc(&dseeds, sizeof(int) * nseeds) != cudaSuccess || cudaMalloc(&dout, sizeof(float) * cells * nseeds) != cudaSuccess) rc = BB_ERR_CUDA;
  if (rc == BB_OK && cudaMemcpy(dseeds, seeds_host, sizeof(int) * nseeds, cudaMemcpyHostToDevice) != cudaSuccess) rc = BB_ERR_CUDA;
  if (rc == BB_OK) {
    dim3 grid((unsigned)((cells + 255) / 256), nseeds < 128 ? nseeds : 128);
    k_perlin_grid<<<grid, 256>>>(n, scale, octaves, persistence, lacunarity, amplitude, dseeds, nseeds, dout);
    if (cudaMemcpy(out_host, dout, sizeof(float) * cells * nseeds, cudaMemcpyDeviceToHost) != cudaSuccess) rc = BB_ERR_CUDA;
  }
  cudaFree(dseeds); cudaFree(dout);
  if (rc != BB_OK) snprintf(g_create_error, sizeof(g_create_error), "bb_perlin_grid: CUDA error %s", cudaGetErrorString(cudaGetLastError()));
  return rc;
}

// host-in / host-out raw 2-D simplex fBm (untiled noise.snoise2): out[i * n + j] = snoise2(i / scale, j / scale, octaves, persistence, lacunarity, base)
int bb_snoise2_grid(int32_t device, int32_t n, float scale, int32_t octaves, float persistence, float lacunarity, int32_t base, float* out_host) {
  if (n < 1 || !out_host || octaves < 1 || scale == 0.f) return BB_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device) { snprintf(g_create_error, sizeof(g_create_error), "bb_snoise2_grid: CUDA device %d not available; no CPU fallback", device); return BB_ERR_NO_DEVICE; }
  DevGuard guard(device);
  if (cudaMemcpyToSymbol(c_perm, h_perm, sizeof(h_perm)) != cudaSuccess) return BB_ERR_CUDA;
  float* dout = nullptr; const size_t cells = (size_t)n * n;
  if (cudaMalloc(&dout, sizeof(float) * cells) != cudaSuccess) return BB_ERR_CUDA;
  k_snoise2_grid<<<(unsigned)((cells + 255) / 256), 256>>>(n, scale, octaves, persistence, lacunarity, base, dout);
  const int rc = cudaMemcpy(out_host, dout, sizeof(float) * cells, cudaMemcpyDeviceToHost) == cudaSuccess ? BB_OK : BB_ERR_CUDA;
  cudaFree(dout);
  return rc;
}

// per-kernel device timing of the next `max_steps` bb_step calls (CUDA events on the caller's stream)
int bb_profile_begin(bb_engine* e, int32_t max_steps) {
  if (!e || max_steps < 1) return BB_ERR_INVALID;
  DevGuard guard(e->cfg.device);
  if (e->prof_ev) { for (int i = 0; i < 5 * e->prof_cap; i++) cudaEventDestroy(e->prof_ev[i]); free(e->prof_ev); e->prof_ev = nullptr; }
  e->prof_ev = (cudaEvent_t*)malloc(sizeof(cudaEvent_t) * 5 * (size_t)max_steps);
  if (!e->prof_ev) return fail(e, BB_ERR_INVALID, "bb_profile_begin: out of host memory");
  for (int i = 0; i < 5 * max_steps; i++) BB_CUDA(cudaEventCreate(&e->prof_ev[i]));
  e->prof_cap = max_steps; e->prof_n = 0;
  return BB_OK;
}
// ms4 = total milliseconds spent in {k_step, k_terrain, k_reset, k_depth} over the profiled steps; synchronises
int bb_profile_end(bb_engine* e, double* ms4, int32_t* nsteps) {
  if (!e || !e->prof_ev || !ms4) return BB_ERR_INVALID;
  DevGuard guard(e->cfg.device);
  double acc[4] = {0, 0, 0, 0};
  if (e->prof_n > 0) BB_CUDA(cudaEventSynchronize(e->prof_ev[5 * (size_t)(e->prof_n - 1) + 4]));
  for (int k = 0; k < e->prof_n; k++)
    for (int j = 0; j < 4; j++) { float ms = 0.f; BB_CUDA(cudaEventElapsedTime(&ms, e->prof_ev[5 * (size_t)k + j], e->prof_ev[5 * (size_t)k + j + 1])); acc[j] += ms; }
  for (int j = 0; j < 4; j++) ms4[j] = acc[j];
  if (nsteps) *nsteps = e->prof_n;
  for (int i = 0; i < 5 * e->prof_cap; i++) cudaEventDestroy(e->prof_ev[i]);
  free(e->prof_ev); e->prof_ev = nullptr; e->prof_cap = e->prof_n = 0;
  return BB_OK;
}

// ------------------------------------------------------------------------------------------------- host-buffer path
static int hostInit(bb_engine* e) {
  if (e->host_ready) return BB_OK;
  const int N = e->N; const size_t npix = (size_t)e->cfg.im_h * e->cfg.im_w;
  BB_CUDA(cudaStreamCreateWithFlags(&e->hstream, cudaStreamNonBlocking));
  BB_CUDA(cudaMallocHost(&e->h_act, sizeof(float) * 3 * N)); BB_CUDA(cudaMalloc(&e->d_act, sizeof(float) * 3 * N));
  bb_io& o = e->dio; memset(&o, 0, sizeof(o));
  BB_CUDA(cudaMalloc(&o.orientation, sizeof(float) * 3 * N)); BB_CUDA(cudaMalloc(&o.angular_vel, sizeof(float) * 3 * N));
  BB_CUDA(cudaMalloc(&o.vel, sizeof(float) * 3 * N)); BB_CUDA(cudaMalloc(&o.motor_state, sizeof(float) * 3 * N));
  BB_CUDA(cudaMalloc(&o.actions, sizeof(float) * 3 * N)); BB_CUDA(cudaMalloc(&o.rel_image_ts, sizeof(float) * N));
  if (e->cfg.cameras) { BB_CUDA(cudaMalloc(&o.rgbd_0, sizeof(float) * npix * N)); BB_CUDA(cudaMalloc(&o.rgbd_1, sizeof(float) * npix * N)); }
  BB_CUDA(cudaMalloc(&o.reward, sizeof(float) * N)); BB_CUDA(cudaMalloc(&o.terminated, N)); BB_CUDA(cudaMalloc(&o.failure, N));
  BB_CUDA(cudaMalloc(&o.pos2d, sizeof(float) * 2 * N)); BB_CUDA(cudaMalloc(&o.terminal_obs, sizeof(float) * 16 * N));
  BB_CUDA(cudaMalloc(&o.episode_return, sizeof(float) * N)); BB_CUDA(cudaMalloc(&o.episode_length, sizeof(int) * N));
  BB_CUDA(cudaMalloc(&o.status, sizeof(int) * N));
  BB_CUDA(cudaMemset(o.terminal_obs, 0, sizeof(float) * 16 * N)); BB_CUDA(cudaMemset(o.episode_return, 0, sizeof(float) * N));
  BB_CUDA(cudaMemset(o.episode_length, 0, sizeof(int) * N));
  BB_CUDA(cudaMalloc(&e->d_obs16, sizeof(float) * 16 * N)); BB_CUDA(cudaMallocHost(&e->h_obs16, sizeof(float) * 16 * N));
  BB_CUDA(cudaMallocHost(&e->h_reward, sizeof(float) * N)); BB_CUDA(cudaMallocHost(&e->h_pos2d, sizeof(float) * 2 * N));
  BB_CUDA(cudaMallocHost(&e->h_term_obs, sizeof(float) * 16 * N)); BB_CUDA(cudaMallocHost(&e->h_epret, sizeof(float) * N));
  BB_CUDA(cudaMallocHost(&e->h_term, N)); BB_CUDA(cudaMallocHost(&e->h_fail, N)); BB_CUDA(cudaMallocHost(&e->h_eplen, sizeof(int) * N));
  BB_CUDA(cudaMalloc(&e->d_mask, N)); BB_CUDA(cudaMallocHost(&e->h_mask, N));
  e->host_ready = true;
  return BB_OK;
}
static int hostReadback(bb_engine* e, const bb_host_io* out) {
  const int N = e->N; cudaStream_t s = e->hstream; const bb_io& o = e->dio;
  k_pack_obs16<<<blocksFor(N, 256), 256, 0, s>>>(N, o, e->d_obs16); e->launches++;
  BB_CUDA(cudaMemcpyAsync(e->h_obs16, e->d_obs16, sizeof(float) * 16 * N, cudaMemcpyDeviceToHost, s));
  BB_CUDA(cudaMemcpyAsync(e->h_reward, o.reward, sizeof(float) * N, cudaMemcpyDeviceToHost, s));
  BB_CUDA(cudaMemcpyAsync(e->h_term, o.terminated, N, cudaMemcpyDeviceToHost, s));
  BB_CUDA(cudaMemcpyAsync(e->h_fail, o.failure, N, cudaMemcpyDeviceToHost, s));
  BB_CUDA(cudaMemcpyAsync(e->h_pos2d, o.pos2d, sizeof(float) * 2 * N, cudaMemcpyDeviceToHost, s));
  if (out->terminal_obs) BB_CUDA(cudaMemcpyAsync(e->h_term_obs, o.terminal_obs, sizeof(float) * 16 * N, cudaMemcpyDeviceToHost, s));
  if (out->episode_return) BB_CUDA(cudaMemcpyAsync(e->h_epret, o.episode_return, sizeof(float) * N, cudaMemcpyDeviceToHost, s));
  if (out->episode_length) BB_CUDA(cudaMemcpyAsync(e->h_eplen, o.episode_length, sizeof(int) * N, cudaMemcpyDeviceToHost, s));
  const size_t npix = (size_t)e->cfg.im_h * e->cfg.im_w;
  if (e->cfg.cameras && out->img_0) BB_CUDA(cudaMemcpyAsync(out->img_0, o.rgbd_0, sizeof(float) * npix * N, cudaMemcpyDeviceToHost, s));
  if (e->cfg.cameras && out->img_1) BB_CUDA(cudaMemcpyAsync(out->img_1, o.rgbd_1, sizeof(float) * npix * N, cudaMemcpyDeviceToHost, s));
  BB_CUDA(cudaStreamSynchronize(s));
  if (out->obs16 && out->obs16 != e->h_obs16) memcpy(out->obs16, e->h_obs16, sizeof(float) * 16 * N);
  if (out->reward && out->reward != e->h_reward) memcpy(out->reward, e->h_reward, sizeof(float) * N);
  if (out->terminated && out->terminated != e->h_term) memcpy(out->terminated, e->h_term, N);
  if (out->failure && out->failure != e->h_fail) memcpy(out->failure, e->h_fail, N);
  if (out->pos2d && out->pos2d != e->h_pos2d) memcpy(out->pos2d, e->h_pos2d, sizeof(float) * 2 * N);
  if (out->terminal_obs && out->terminal_obs != e->h_term_obs) memcpy(out->terminal_obs, e->h_term_obs, sizeof(float) * 16 * N);
  if (out->episode_return && out->episode_return != e->h_epret) memcpy(out->episode_return, e->h_epret, sizeof(float) * N);
  if (out->episode_length && out->episode_length != e->h_eplen) memcpy(out->episode_length, e->h_eplen, sizeof(int) * N);
  return BB_OK;
}
int bb_host_buffers(bb_engine* e, float** actions_host, bb_host_io* out) {
  if (!e || !out) return BB_ERR_INVALID;
  DevGuard guard(e->cfg.device);
  int rc = hostInit(e); if (rc) return rc;
  if (actions_host) *actions_host = e->h_act;
  memset(out, 0, sizeof(*out));
  out->obs16 = e->h_obs16; out->reward = e->h_reward; out->terminated = e->h_term; out->failure = e->h_fail; out->pos2d = e->h_pos2d;
  out->terminal_obs = e->h_term_obs; out->episode_return = e->h_epret; out->episode_length = e->h_eplen;
  return BB_OK;
}
int bb_step_host(bb_engine* e, const float* actions_host, const bb_host_io* out) {
  if (!e || !actions_host || !out) return BB_ERR_INVALID;
  DevGuard guard(e->cfg.device);
  int rc = hostInit(e); if (rc) return rc;
  if (actions_host != e->h_act) memcpy(e->h_act, actions_host, sizeof(float) * 3 * e->N);
  BB_CUDA(cudaMemcpyAsync(e->d_act, e->h_act, sizeof(float) * 3 * e->N, cudaMemcpyHostToDevice, e->hstream));
  rc = bb_step(e, e->d_act, &e->dio, e->hstream); if (rc) return rc;
  return hostReadback(e, out);
}
int bb_reset_host(bb_engine* e, const uint8_t* mask_host, const bb_host_io* out) {
  if (!e || !out) return BB_ERR_INVALID;
  DevGuard guard(e->cfg.device);
  int rc = hostInit(e); if (rc) return rc;
  const uint8_t* dm = nullptr;
  if (mask_host) { memcpy(e->h_mask, mask_host, e->N); BB_CUDA(cudaMemcpyAsync(e->d_mask, e->h_mask, e->N, cudaMemcpyHostToDevice, e->hstream)); dm = e->d_mask; }
  BB_CUDA(cudaMemsetAsync(e->dio.reward, 0, sizeof(float) * e->N, e->hstream));
  BB_CUDA(cudaMemsetAsync(e->dio.terminated, 0, e->N, e->hstream)); BB_CUDA(cudaMemsetAsync(e->dio.failure, 0, e->N, e->hstream));
  BB_CUDA(cudaMemsetAsync(e->dio.pos2d, 0, sizeof(float) * 2 * e->N, e->hstream));
  rc = bb_reset(e, dm, nullptr, &e->dio, e->hstream); if (rc) return rc;
  return hostReadback(e, out);
}

}  // extern "C"
