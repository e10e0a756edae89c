// bb_group.cuh -- lane-group-per-env forward dynamics + RK4 for the B200 ballbot engine (device only).
//
// A group of G = 16 lanes integrates one environment, so one warp carries two environments through the same
// instruction stream (the 15 dofs of the model fit one 16-lane group; with a full warp per env more than half of the
// lanes idle in every solver phase).  All group collectives (__shfl_sync / __syncwarp / __ballot_sync) are issued with
// the group's own member mask, so the two halves of a warp are logically independent and only pay for divergence
// where their trip counts differ (Newton iterations, line-search evaluations, contact batches).
//
// Layout of the solver (mj_fwdConstraint / mj_solNewton semantics, reference call ballbot_env.py:912):
//   * lane i < 15 of a group owns dof i: qfrc_smooth, qacc_smooth, qacc, M qacc, gradient, search direction and the
//     RK4 accumulators are ONE register each per lane;
//   * the block-diagonal mass matrix (9 x 9 base/wheels + 6 x 6 ball) sits in shared memory, both triangles stored;
//   * the Hessian H = M + sum_c J_c' W_c J_c is assembled with column i in the registers of lane i and factorised in
//     place by a rolled right-looking Cholesky: per pivot one broadcast, one rsqrt, the scaled column goes to the
//     shared factor array (row stride 17), and every lane shifts its column up by one inside the update FMA so that the
//     pivot row is always register 0; forward substitution is fused, backward substitution reads the lane's own row;
//   * contact c is owned by lane c % G for the per-contact scalar work (cone zones, line-search coefficients);
//     its record (3 Jacobian rows + 24 scalars) lives in shared memory (3 dense + 8 terrain records) and spills to a
//     global scratch beyond that (deep impacts, robot geoms touching the terrain);
//   * contact order: [nw ball x wheel pairs (anisotropic friction)] [nd - nw other dense pairs: ball x stick / tower, stick /
//     wheel capsule x heightfield] [ncon - nd ball x heightfield prisms]; "dense" = 16-column Jacobian rows (any dof), the
//     ball x heightfield rows only hold the ball dofs;
//   * the exact line search is a state machine around ONE evaluation call site, so two environments that are in
//     different phases of their searches still share every evaluation instruction; evaluated points live in
//     shared-memory slots and the brackets are slot indices;
//   * GNewton<T, true> (k_newton) is the uniform-warp variant: both groups run every loop together, a finished group
//     rides along with its state frozen, and all collectives use the compile-time full-warp mask.
// The arithmetic restates the same algorithm as bb_core.cuh (thread-per-env cross-check and CPU test harness); only
// summation orders and the rsqrt-based square roots differ.
#pragma once
#include "bb_core.cuh"

#ifndef BB_GROUP
#define BB_GROUP 16
#endif

namespace bbg {
using namespace bb;

constexpr int G = BB_GROUP;                 // lanes per environment (16: two envs per warp, 32: one env per warp)
constexpr int EPW = 32 / G;                 // environments per warp
constexpr int MSZ = 118;                    // mass matrix: 9 x 9 base/wheel block + 6 x 6 ball block (both triangles), padded
constexpr int LS_ = 17;                     // row stride of the Cholesky factor (column j of L in row j; odd => conflict-free transposed reads)
// contact records (in T).  Wheel pair: rows of 16 (dofs 0..14 + zero pad); terrain pair: rows of 8 = dofs 8..15 (only the
// ball dofs 9..14 are non-zero).  Scalar block at OS: W(6) frc(3) D0 jar(3) jv(3) LS(8).
constexpr int CRW = 74, CRH = 50;           // strides == 10 (mod 16): one-record-per-lane access is bank-conflict-free
constexpr int OSW = 48, OSH = 24;
constexpr int O_W = 0, O_FRC = 6, O_D0 = 9, O_JAR = 10, O_JV = 13, O_LS = 16;
#ifndef BB_NHS
#define BB_NHS 8
#endif
constexpr int NHS = BB_NHS;                 // terrain records resident in shared memory
constexpr int NDS = 3;                      // dense records resident in shared memory
constexpr int NDMAX = 27;                   // dense contact capacity (3 wheel pairs + 24 others)
constexpr int GSD = (NDMAX - NDS) * CRW;    // dense overflow records come first in the per-env global scratch
constexpr int GSCR = GSD + (MAXH - NHS) * CRH;    // per-env global overflow scratch (in T)
constexpr int NCMAX = NDMAX + MAXH;         // contact capacity of one forward pass
// geometry block published by the smooth-dynamics pass
constexpr int LSP_SLOTS = 8, LSP_W = 4, LSP_ALPHA = 0, LSP_COST = 1, LSP_D1 = 2, LSP_NXT = 3;   // line-search point slots
constexpr int GE_PB = 0, GE_PL = 3, GE_RB = 6, GE_RL = 15, GE_AW = 24, GE_HW = 33, GE_CC = 42, GE_CU = 51, GE_N = 60;

template <typename T> struct GS {
  T xq[20], xv[16], q0[20], ctrl[4];
  T M[MSZ];
  T vb[2][16];          // vectors published by the dof lanes (broadcast reads)
  union {
    T Lf[NV * LS_ + 3]; // Cholesky factor, Lf[j * 17 + i] = L(i, j) for i > j and 0 for i <= j (Newton solve); the last
                        // pivot's (unused) trailing update reads up to index 256
    T geo[GE_N];        // geometry block of the smooth-dynamics pass (consumed by gCollide before the first factorisation)
  };
  T kin[16];            // quatB(4) cvel_ang(3) cvel_lin(3) posB(3) of the last evaluated stage
  T lsp[LSP_SLOTS * LSP_W];   // line-search points (alpha, cost, d1, next Newton alpha), see lsEval
  T wrec[3 * CRW];
  T hrec[NHS * CRH];
  unsigned char cst[80];   // per contact: bits 0-1 zone (0 satisfied, 1 quadratic, 2 cone)
};

struct Ln { int gl; int gi; unsigned mask; };   // lane in group, clamped dof index, member mask of the group
__device__ __forceinline__ Ln makeLn() {
  Ln L; const int lane = threadIdx.x & 31;
  L.gl = lane & (G - 1); L.gi = L.gl < NV ? L.gl : NV - 1;
  L.mask = G == 32 ? 0xffffffffu : (0xffffu << (lane & 16));
  return L;
}
template <typename T> __device__ __forceinline__ T gsum(T v, unsigned mask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}
template <typename T> __device__ __forceinline__ T gget(T v, int src, unsigned mask) { return __shfl_sync(mask, v, src, G); }

template <typename T> struct V2T;
template <> struct V2T<double> { typedef double2 t; };
template <> struct V2T<float> { typedef float2 t; };
template <typename T> __device__ __forceinline__ typename V2T<T>::t ld2(const T* p) { return *reinterpret_cast<const typename V2T<T>::t*>(p); }

__device__ __forceinline__ float brsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ double brsqrt(double x) { return rsqrt(x); }
// s = sqrt(x), is = 1/sqrt(x) for x > 0 from one reciprocal square root (+ one correction step for s)
__device__ __forceinline__ void sqrtInv(float x, float& s, float& is) { is = rsqrtf(x); s = x * is; }
__device__ __forceinline__ void sqrtInv(double x, double& s, double& is) {
  is = rsqrt(x);
  const double s0 = x * is;
  s = fma(fma(-s0, s0, x), 0.5 * is, s0);
}

// a / b for b >= 1e-15 (the clamped second derivative of the line search): reciprocal seed + two Newton steps + one
// residual correction instead of the IEEE division sequence (shorter dependent chain; the result is within 1 ulp)
__device__ __forceinline__ float fdivPos(float a, float b) { return a / b; }
__device__ __forceinline__ double fdivPos(double a, double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  r = fma(fma(-b, r, 1.0), r, r);
  r = fma(fma(-b, r, 1.0), r, r);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
}

// record of contact c; nd = number of dense contacts (they come first)
template <typename T> __device__ __forceinline__ T* crecD(GS<T>& S, T* gs, int c) { return c < NDS ? S.wrec + c * CRW : gs + (c - NDS) * CRW; }
template <typename T> __device__ __forceinline__ T* crecH(GS<T>& S, T* gs, int hc) { return hc < NHS ? S.hrec + hc * CRH : gs + GSD + (hc - NHS) * CRH; }
template <typename T> __device__ __forceinline__ T* crec(GS<T>& S, T* gs, int c, int nd) { return c < nd ? crecD(S, gs, c) : crecH(S, gs, c - nd); }

// ---------------------------------------------------------------------------------------------- smooth dynamics
// Every lane evaluates the (small, serial) smooth dynamics of its environment redundantly; the results go to shared
// memory (mass matrix, qfrc_smooth -> vb[0], contact geometry, observation kinematics).
template <typename T> __device__ __noinline__ void gSmooth(const ModelConst<T>& mc, GS<T>& S, bool wantKin) {
  Geo<T> g; V3<T> capC[3], capU[3]; KinOut<T> kin;
  smoothDynamics<T, false, 1>(mc, S.xq, S.xv, S.ctrl, S.M, S.vb[0], g, capC, capU, wantKin ? &kin : (KinOut<T>*)nullptr);
  T* ge = S.geo;
  st3(ge + GE_PB, g.pB); st3(ge + GE_PL, g.pL);
  st3(ge + GE_RB, g.RB.c0); st3(ge + GE_RB + 3, g.RB.c1); st3(ge + GE_RB + 6, g.RB.c2);
  st3(ge + GE_RL, g.RL.c0); st3(ge + GE_RL + 3, g.RL.c1); st3(ge + GE_RL + 6, g.RL.c2);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    st3(ge + GE_AW + 3 * i, g.aw[i]); st3(ge + GE_HW + 3 * i, g.hw[i]);
    st3(ge + GE_CC + 3 * i, capC[i]); st3(ge + GE_CU + 3 * i, capU[i]);
  }
  if (wantKin) {
#pragma unroll
    for (int k = 0; k < 4; k++) S.kin[k] = kin.quatB[k];
#pragma unroll
    for (int k = 0; k < 3; k++) { S.kin[4 + k] = kin.cvel_ang[k]; S.kin[7 + k] = kin.cvel_lin[k]; S.kin[10 + k] = kin.posB[k]; }
  }
}

// ---------------------------------------------------------------------------------------------- Hessian + Cholesky + solve
// Assembles column gi of H = M + sum_{c < ncon, zone != 0} J_c' W_c J_c in registers, factorises H = L L' across the
// group and returns x[gi] of H x = b (b: one value per dof lane).  ncon = 0 gives M^-1 b (qacc_smooth).
template <typename T>
__device__ __forceinline__ T gHessSolve(GS<T>& S, const T* gs, int ncon, int nd, T b, const Ln L) {
  const int gi = L.gi;
  T h[NV];
  {
    const bool top = gi < 9;
    const T* ma = S.M + (top ? gi : 0);
    const T* mb = S.M + 81 + (top ? 0 : gi - 9);
#pragma unroll
    for (int k = 0; k < 9; k++) { const T v = ma[k * 9]; h[k] = top ? v : (T)0; }
#pragma unroll
    for (int k = 9; k < NV; k++) { const T v = mb[(k - 9) * 6]; h[k] = top ? (T)0 : v; }
  }
  // dense pairs: 15-column rows
#pragma unroll 1
  for (int c = 0; c < nd; c++) {
    if ((S.cst[c] & 3) == 0) continue;
    const T* rec = c < NDS ? (const T*)(S.wrec + c * CRW) : gs + (c - NDS) * CRW;
    const T a0 = rec[gi], a1 = rec[16 + gi], a2 = rec[32 + gi];
    const typename V2T<T>::t w01 = ld2(rec + OSW), w23 = ld2(rec + OSW + 2), w45 = ld2(rec + OSW + 4);
    const T t0 = w01.x * a0 + w23.y * a1 + w45.x * a2;   // W = [w0 w3 w4; w3 w1 w5; w4 w5 w2]
    const T t1 = w23.y * a0 + w01.y * a1 + w45.y * a2;
    const T t2 = w45.x * a0 + w45.y * a1 + w23.x * a2;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
      const typename V2T<T>::t r0 = ld2(rec + k), r1 = ld2(rec + 16 + k), r2 = ld2(rec + 32 + k);
      h[k] += r0.x * t0 + r1.x * t1 + r2.x * t2;
      if (k + 1 < NV) h[k + 1] += r0.y * t0 + r1.y * t1 + r2.y * t2;
    }
  }
  // terrain pair: only the ball dofs 9..14 (rows hold dofs 8..15)
  const bool bl = gi >= 9;
#pragma unroll 1
  for (int c = nd; c < ncon; c++) {
    if ((S.cst[c] & 3) == 0) continue;
    const int hc = c - nd;
    const T* rec = hc < NHS ? (const T*)(S.hrec + hc * CRH) : gs + GSD + (hc - NHS) * CRH;
    const int o = bl ? gi - 8 : 0;
    T a0 = rec[o], a1 = rec[8 + o], a2 = rec[16 + o];
    if (!bl) { a0 = 0; a1 = 0; a2 = 0; }
    const typename V2T<T>::t w01 = ld2(rec + OSH), w23 = ld2(rec + OSH + 2), w45 = ld2(rec + OSH + 4);
    const T t0 = w01.x * a0 + w23.y * a1 + w45.x * a2;
    const T t1 = w23.y * a0 + w01.y * a1 + w45.y * a2;
    const T t2 = w45.x * a0 + w45.y * a1 + w23.x * a2;
#pragma unroll
    for (int k = 8; k < 16; k += 2) {
      const typename V2T<T>::t r0 = ld2(rec + k - 8), r1 = ld2(rec + k), r2 = ld2(rec + 8 + k);
      if (k >= 9) h[k] += r0.x * t0 + r1.x * t1 + r2.x * t2;
      if (k + 1 < NV) h[k + 1] += r0.y * t0 + r1.y * t1 + r2.y * t2;
    }
  }
  // ---- right-looking Cholesky, column per lane, forward substitution fused.  The loop over the pivots is rolled: after
  // step j every lane shifts its column up by one (folded into the update FMA), so the pivot row is always h[0] and
  // the code is one short body (the fully unrolled triangle was 2.7 k instructions and thrashed the instruction cache).
  // The trailing update touches the 14 - j live rows only up to the granularity of three loop bodies (14 / 9 / 4 rows).
  T myinv = 0, y = b;
  T* Lf = S.Lf;
  // Pivot floor.  fp64: MuJoCo's absolute mjMINVAL.  fp32: the Hessian spans ~10 orders of magnitude (contact stiffness vs
  // wheel inertia), so cancellation can leave a pivot tiny or negative; it is floored relative to the dof's own assembled
  // diagonal entry (travels with the pivot in a second broadcast) instead of letting 1 / sqrt(pivot) explode.
  T dfloor = (T)1e-15;
  if (sizeof(T) == 4) {
    T d0 = h[0];
#pragma unroll
    for (int k = 1; k < NV; k++) d0 = gi == k ? h[k] : d0;
    dfloor = (T)1e-6 * d0;
  }
#define BB_PIVOT_STEP(NUPD)                                                                                               \
  {                                                                                                                       \
    T piv = gget(h[0], j, L.mask);                                                                                        \
    const T pfl = sizeof(T) == 4 ? gget(dfloor, j, L.mask) : (T)1e-15;                                                    \
    piv = piv < pfl ? pfl : piv;                                                                                          \
    const T inv = brsqrt(piv);                                                                                            \
    const T l = h[0] * inv;                       /* lanes i >= j: L(i,j)  (lane j: sqrt(pivot)) */                      \
    if (gi == j) myinv = inv;                                                                                             \
    const T yj = gget(y, j, L.mask) * inv;        /* y_j = (b_j - sum_{k<j} L(j,k) y_k) / L(j,j) */                      \
    if (gi == j) y = yj;                                                                                                  \
    const T lz = gi > j ? l : (T)0;                                                                                       \
    y -= lz * yj;                                                                                                         \
    Lf[j * LS_ + L.gl] = lz;                                                                                              \
    __syncwarp(L.mask);                                                                                                   \
    /* lanes i > j: A(r,i) -= L(i,j) L(r,j) for the rows r = j+1 .., stored shifted: h[k] <- A(j+1+k, i) */               \
    const T* cj = Lf + j * LS_ + j + 1;                                                                                   \
    _Pragma("unroll") for (int k = 0; k < (NUPD); k++) h[k] = h[k + 1] - lz * cj[k];                                      \
  }
  int j = 0;
#pragma unroll 1
  for (; j < 5; j++) BB_PIVOT_STEP(14)
#pragma unroll 1
  for (; j < 10; j++) BB_PIVOT_STEP(9)
#pragma unroll 1
  for (; j < NV; j++) BB_PIVOT_STEP(4)
#undef BB_PIVOT_STEP
  // ---- backward substitution L' x = y (lane k reads its own row of Lf: L(i,k), i > k)
  T x = 0, s = 0;
  const T* lrow = Lf + gi * LS_;
#pragma unroll 1
  for (int i = NV - 1; i >= 0; i--) {
    const T cand = (y - s) * myinv;               // valid on lane i
    const T xi = gget(cand, i, L.mask);
    if (gi == i) x = xi;
    s += lrow[i] * xi;                            // lanes k < i: L(i,k); zero for k >= i
  }
  return x;
}

// x = M^-1 b for the block-diagonal mass matrix (qacc_smooth): the 9 x 9 base/wheel block on lanes 0..8 and the 6 x 6 ball
// block on lanes 9..14 are factorised side by side with the same shifting-column scheme as gHessSolve, so the sequential
// chain is 9 pivots instead of 15.  The arithmetic inside each block is the one the dense 15 x 15 factorisation performs
// (its cross-block updates are exact zeros), so the result is bit-identical to gHessSolve(ncon = 0).
template <typename T>
__device__ __forceinline__ T gMassSolve(GS<T>& S, T b, const Ln L) {
  const int gi = L.gi;
  const bool top = gi < 9;
  const int blk = top ? 0 : 9, nb = top ? 9 : 6, li = gi - blk;
  T h[9];
  {
    const T* ma = S.M + (top ? gi : 0);
    const T* mb = S.M + 81 + (top ? 0 : gi - 9);
#pragma unroll
    for (int k = 0; k < 6; k++) { const T va = ma[k * 9], vb = mb[k * 6]; h[k] = top ? va : vb; }
#pragma unroll
    for (int k = 6; k < 9; k++) { const T va = ma[k * 9]; h[k] = top ? va : (T)0; }
  }
  T myinv = 0, y = b;
  T* Lf = S.Lf;
#pragma unroll 1
  for (int j = 0; j < 9; j++) {
    const bool actv = j < nb;
    const int src = actv ? blk + j : L.gl;
    T piv = gget(h[0], src, L.mask);
    piv = piv < (T)1e-15 ? (T)1e-15 : piv;
    const T inv = brsqrt(piv);
    const T l = h[0] * inv;
    const bool mine = actv && li == j;
    if (mine) myinv = inv;
    const T yj = gget(y, src, L.mask) * inv;
    if (mine) y = yj;
    const bool below = actv && li > j;
    const T lz = below ? l : (T)0;
    if (below) y -= l * yj;                       // (predicated: a finished block's lanes see non-finite garbage pivots)
    Lf[j * LS_ + L.gl] = lz;
    __syncwarp(L.mask);
    const T* cj = Lf + j * LS_ + blk + j + 1;
#pragma unroll
    for (int k = 0; k < 8; k++) h[k] = h[k + 1] - lz * cj[k];
  }
  T x = 0, s = 0;
  const T* lrow = Lf + li * LS_ + blk;
#pragma unroll 1
  for (int i = 8; i >= 0; i--) {
    const bool actv = i < nb;
    const T cand = (y - s) * myinv;               // valid on the lane with local index i
    const T xi = gget(cand, actv ? blk + i : L.gl, L.mask);
    if (actv && li == i) x = xi;
    if (actv) s += lrow[i] * xi;                  // L(blk + i, gi) for li < i, zero otherwise
  }
  return x;
}

// r[gi] = sum_k M(gi,k) v[k] for a published vector v (shared, 16 entries, v[15] finite)
template <typename T> __device__ __forceinline__ T gSymv(const GS<T>& S, const T* v, int gi) {
  const bool top = gi < 9;
  const T* m = top ? S.M + gi : S.M + 81 + (gi - 9);
  const T* vv = top ? v : v + 9;
  const int st = top ? 9 : 6;
  T acc = 0;
#pragma unroll
  for (int k = 0; k < 6; k++) acc += m[k * st] * vv[k];
  if (top) {
#pragma unroll
    for (int k = 6; k < 9; k++) acc += m[k * 9] * vv[k];
  }
  return acc;
}

// out[k] = J_c[k] . v for the contact record `rec` (wheel: 15 dofs, terrain: dofs 8..15); v: published vector
template <typename T> __device__ __forceinline__ void rowsDot(const T* rec, bool wheel, const T* v, T* out) {
  const int n2 = wheel ? 8 : 4, rs = wheel ? 16 : 8;
  const T* vv = wheel ? v : v + 8;
  T a0 = 0, a1 = 0, a2 = 0;
#pragma unroll 2
  for (int p = 0; p < n2; p++) {
    const typename V2T<T>::t x = ld2(vv + 2 * p), r0 = ld2(rec + 2 * p), r1 = ld2(rec + rs + 2 * p), r2 = ld2(rec + 2 * rs + 2 * p);
    a0 += r0.x * x.x + r0.y * x.y; a1 += r1.x * x.x + r1.y * x.y; a2 += r2.x * x.x + r2.y * x.y;
  }
  out[0] = a0; out[1] = a1; out[2] = a2;
}

// ---------------------------------------------------------------------------------------------- per-contact cone math
// zone logic of mj_constraintUpdate for one elliptic contact; returns cost; state: 0 satisfied, 1 quadratic, 2 cone
template <typename T>
__device__ __forceinline__ T coneLane(const ModelConst<T>& mc, int k, T D0, const T* jar, T* frc, T* h, int& state, bool wantH) {
  const T mu = mc.mu[k], f1 = mc.f1[k], f2 = mc.f2[k];
  const T U0 = jar[0] * mu, U1 = jar[1] * f1, U2 = jar[2] * f2;
  const T N = U0, Tsq = U1 * U1 + U2 * U2;
  T Tn = 0, iT = 0;
  if (Tsq > 0) sqrtInv(Tsq, Tn, iT);
  if (N >= mu * Tn || (Tn <= 0 && N >= 0)) { frc[0] = frc[1] = frc[2] = 0; state = 0; return 0; }
  if (mu * N + Tn <= 0 || (Tn <= 0 && N < 0)) {
    const T D1 = D0 * mc.d1r[k], D2 = D0 * mc.d2r[k];
    frc[0] = -D0 * jar[0]; frc[1] = -D1 * jar[1]; frc[2] = -D2 * jar[2]; state = 1;
    if (wantH) { h[0] = D0; h[1] = D1; h[2] = D2; h[3] = 0; h[4] = 0; h[5] = 0; }
    return (T)0.5 * (D0 * jar[0] * jar[0] + D1 * jar[1] * jar[1] + D2 * jar[2] * jar[2]);
  }
  const T Dm = D0 * mc.dmr[k];
  const T NT = N - mu * Tn;
  frc[0] = -Dm * NT * mu;
  const T sc = -frc[0] * iT;
  frc[1] = sc * f1 * U1; frc[2] = sc * f2 * U2;
  state = 2;
  if (wantH) {
    const T muN_T3 = mu * N * iT * iT * iT, dg = mu * mu - mu * N * iT;
    h[0] = Dm * mu * mu;
    h[1] = Dm * f1 * f1 * (muN_T3 * U1 * U1 + dg);
    h[2] = Dm * f2 * f2 * (muN_T3 * U2 * U2 + dg);
    h[3] = Dm * mu * f1 * (-mu * U1 * iT);
    h[4] = Dm * mu * f2 * (-mu * U2 * iT);
    h[5] = Dm * f1 * f2 * (muN_T3 * U1 * U2);
  }
  return (T)0.5 * Dm * NT * NT;
}

// One evaluated point of the 1-D line-search objective: alpha, cost, first / second derivative and the alpha of the next
// 1-D Newton step (alpha - d1 / d2, computed once here so that the search logic needs no further divisions).
template <typename T> struct LsPt { T alpha, cost, d1, d2, nxt; };
template <typename T> struct LsCtx { T qG0, qG1, qG2; };
template <typename T> struct Sum3 { T a, b, c; };
template <typename T> __device__ __forceinline__ Sum3<T> gsum3(T a, T b, T c, unsigned mask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) { a += __shfl_xor_sync(mask, a, o); b += __shfl_xor_sync(mask, b, o); c += __shfl_xor_sync(mask, c, o); }
  Sum3<T> r; r.a = a; r.b = b; r.c = c; return r;
}

// cost and derivatives of the 1-D line-search objective at alpha (PrimalEval); uniform over the group.  The point is
// also parked in slot `slot` of S.lsp so that the search logic can refer to older points by index.
template <typename T>
__device__ __forceinline__ LsPt<T> lsEval(const ModelConst<T>& mc, GS<T>& S, const T* gs, int ncon, int nw, int nd, const LsCtx<T> q, T alpha, int slot, const Ln L,
                                          const T* rec0) {
  T cost = 0, d1 = 0, d2 = 0;
#pragma unroll 1
  for (int c = L.gl; c < ncon; c += G) {
    const T* sc = (c < G ? rec0 : (const T*)crec(S, (T*)gs, c, nd)) + (c < nd ? OSW : OSH);
    const int k = c < nw ? 0 : 1;
    const T mu = mc.mu[k];
    const typename V2T<T>::t l01 = ld2(sc + O_LS), l23 = ld2(sc + O_LS + 2), l4q = ld2(sc + O_LS + 4), lq = ld2(sc + O_LS + 6);
    const T U0 = l01.x, V0 = l01.y, UU = l23.x, UV = l23.y, VV = l4q.x;
    const T N = U0 + alpha * V0, Tsq = UU + alpha * ((T)2 * UV + alpha * VV);
    bool bottom = false;
    if (Tsq <= 0) bottom = N < 0;
    else {
      T Tn, iT; sqrtInv(Tsq, Tn, iT);
      if (N >= mu * Tn) {}
      else if (mu * N + Tn <= 0) bottom = true;
      else {
        const T Dm = sc[O_D0] * mc.dmr[k];
        const T b = UV + alpha * VV;
        const T T1 = b * iT, T2 = VV * iT - b * T1 * iT * iT;
        const T NT = N - mu * Tn, dNT = V0 - mu * T1;
        cost += (T)0.5 * Dm * NT * NT; d1 += Dm * NT * dNT; d2 += Dm * (dNT * dNT - NT * mu * T2);
      }
    }
    if (bottom) { cost += l4q.y + alpha * (lq.x + alpha * lq.y); d1 += lq.x + (T)2 * alpha * lq.y; d2 += (T)2 * lq.y; }
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) { cost += __shfl_xor_sync(L.mask, cost, o); d1 += __shfl_xor_sync(L.mask, d1, o); d2 += __shfl_xor_sync(L.mask, d2, o); }
  LsPt<T> p; p.alpha = alpha;
  p.cost = q.qG0 + alpha * (q.qG1 + alpha * q.qG2) + cost; p.d1 = q.qG1 + (T)2 * alpha * q.qG2 + d1; p.d2 = (T)2 * q.qG2 + d2;
  if (p.d2 < (T)1e-15) p.d2 = (T)1e-15;
  p.nxt = alpha - fdivPos(p.d1, p.d2);
  if (L.gl < LSP_W) S.lsp[slot * LSP_W + L.gl] = L.gl == LSP_ALPHA ? alpha : (L.gl == LSP_COST ? p.cost : (L.gl == LSP_D1 ? p.d1 : p.nxt));
  __syncwarp(L.mask);
  return p;
}

// ---------------------------------------------------------------------------------------------- group Newton solver
// U (uniform warp): both groups of the warp run every loop of the solver together -- a group whose solve has finished (or
// that has no contacts) keeps executing as a passenger with its state frozen -- so every collective is reached by the whole
// warp and L.mask can be the compile-time constant 0xffffffff (set by the caller; everything below is inlined).  That
// removes the convergence pre-check (MATCH / REDUX / VOTE / branch) ptxas emits in front of every shuffle group with a
// run-time member mask: ~5 % of the solver's instructions and ~12 % of its stall samples.
template <typename T, bool U = false, int FM = -1> struct GNewton {
  const ModelConst<T>& mc; GS<T>& S; T* gs; const Ln L; int ncon, nw, nd;   // nw wheel pairs <= nd dense contacts <= ncon
  const bool fast;   // solver_mode 1: strong-Wolfe line search with cone-apex candidates (searchFast), same minimiser of the outer problem
                     // FM: -1 = `fast` decides at run time (fused / probe kernels), 0 / 1 = compile-time choice (k_newton)
  __device__ __forceinline__ bool isFast() const { return FM < 0 ? fast : FM == 1; }
  T qfs, qas;        // dof-lane registers
  T qacc, Ma, grad, search, Mv;
  T cost, gauss, gnorm2;
  int nevals = 0;    // line-search evaluations of this solve (work key of the scheduler)
  int iter = 0;      // Newton iterations of this solve
  T* rec0;           // record of the contact this lane owns in every per-contact loop (c = gl; loop-invariant for the whole solve)
  __device__ GNewton(const ModelConst<T>& m, GS<T>& s, T* g, const Ln l, int n, int w, int dn, bool f, T qf, T qa)
      : mc(m), S(s), gs(g), L(l), ncon(n), nw(w), nd(dn), fast(f), qfs(qf), qas(qa) { rec0 = crec(S, gs, L.gl, nd); }
  __device__ __forceinline__ T* recOf(int c) const { return c < G ? rec0 : crec(S, gs, c, nd); }
  // persistent solver kernel: the group takes its next environment (its solver input has been loaded into S)
  __device__ __forceinline__ void attach(T* g, int n, int w, int dn, T qf, T qa) {
    gs = g; ncon = n; nw = w; nd = dn; qfs = qf; qas = qa; rec0 = crec(S, gs, L.gl, nd); nevals = 0; iter = 0;
  }

  // forces / zones / cone Hessian blocks at the current jar (contact lanes), cost, gradient (dof lanes)
  __device__ __forceinline__ void costGrad() {
    T cpart = 0;
#pragma unroll 1
    for (int c = L.gl; c < ncon; c += G) {
      T* sc = recOf(c) + (c < nd ? OSW : OSH);
      T h[6], f[3]; int st;
      const T jr[3] = {sc[O_JAR], sc[O_JAR + 1], sc[O_JAR + 2]};
      cpart += coneLane(mc, c < nw ? 0 : 1, sc[O_D0], jr, f, h, st, true);
      S.cst[c] = (unsigned char)st;
      sc[O_FRC] = f[0]; sc[O_FRC + 1] = f[1]; sc[O_FRC + 2] = f[2];
      if (st) {
#pragma unroll
        for (int m = 0; m < 6; m++) sc[O_W + m] = h[m];
      }
    }
    const bool dof = L.gl < NV;
    __syncwarp(L.mask);
    T g = Ma - qfs;
    const int gi = L.gi;
#pragma unroll 1
    for (int c = 0; c < nd; c++) {
      if ((S.cst[c] & 3) == 0) continue;
      const T* rec = crecD(S, gs, c);
      g -= rec[gi] * rec[OSW + O_FRC] + rec[16 + gi] * rec[OSW + O_FRC + 1] + rec[32 + gi] * rec[OSW + O_FRC + 2];
    }
    if (gi >= 9) {
#pragma unroll 1
      for (int c = nd; c < ncon; c++) {
        if ((S.cst[c] & 3) == 0) continue;
        const int hc = c - nd;
        const T* rec = crecH(S, gs, hc);
        g -= rec[gi - 8] * rec[OSH + O_FRC] + rec[gi] * rec[OSH + O_FRC + 1] + rec[8 + gi] * rec[OSH + O_FRC + 2];
      }
    }
    grad = dof ? g : (T)0;
    const Sum3<T> r3 = gsum3(dof ? (T)0.5 * (Ma - qfs) * (qacc - qas) : (T)0, cpart, grad * grad, L.mask);
    gauss = r3.a; cost = r3.a + r3.b; gnorm2 = r3.c;
  }

  // solver_mode 1 (restated for one thread in bb_core.cuh, Newton::lineSearchFast).  The search direction is the exact
  // Newton direction, so phi(0) = cost, phi'(0) = grad . search, phi''(0) = -phi'(0) need no evaluation and the first 1-D
  // Newton point is alpha = 1.  Safeguarded 1-D Newton iteration on phi' inside a bracket [lo, hi], stopped at the strong
  // Wolfe conditions (c1 = 1e-4, c2 = 0.1).  phi' jumps at the apex of a friction cone; for the omniwheel pairs (friction
  // 1 : 0.001) every search line passes within ~1e-6 of an apex and the minimiser very often sits on it: the apex of wheel
  // pair c is at alpha_c = -UV_c / VV_c, and a step that would jump across it lands on it first (the reference's search
  // gets there by bisection, ~20 evaluations).  All state is group-uniform; one evaluation call site.
  // exits of searchFast other than the Wolfe test (bracket collapsed to rounding, evaluation cap): the best Armijo point,
  // unless (fp32) its improvement is inside the rounding noise of the cost -- then the solve has converged (alpha = 0)
  __device__ __forceinline__ static T fallbackAlpha(bool have, T bestA, T bestC, T c0) {
    if (!have) return 0;
    if (sizeof(T) == 4 && c0 - bestC <= (T)1e-5 * babs(c0)) return 0;
    return bestA;
  }
  __device__ __forceinline__ T searchFast(const LsCtx<T> q, const T gtol, bool on) {
    const bool dof = L.gl < NV;
    const T d0 = gsum(dof ? grad * search : (T)0, L.mask);
    const T c0 = cost, wtol = bmax((T)0.1 * babs(d0), gtol);
    // rounding floor of the working precision: cost differences below it are noise, a bracket cannot shrink below ~1 ulp
    const T cslack = sizeof(T) == 4 ? (T)4e-7 * babs(c0) : (T)0, wrel = sizeof(T) == 4 ? (T)1e-6 : (T)1e-12;
    T lo = 0, hi = -1, a = 1, wprev = (T)1e30, bestA = 0, bestC = c0, result = 0;
    T kk0 = -1, kk1 = -1, kk2 = -1;
    bool have = false, kready = false;
    int k = 0; unsigned kused = 0;
#pragma unroll 1
    for (;;) {
      if (U && !__any_sync(0xffffffffu, on)) break;
      const LsPt<T> p = lsEval(mc, S, gs, ncon, nw, nd, q, a, 1, L, rec0);
      if (U && !on) continue;                             // passenger: state frozen
      nevals++;
      const bool armijo = p.cost <= c0 + (T)1e-4 * a * d0 + cslack;
      if (armijo && (!have || p.cost < bestC)) { bestA = a; bestC = p.cost; have = true; }
      bool done = false;
      if (armijo && babs(p.d1) <= wtol) { result = a; done = true; }
      else {
        if (!armijo || p.d1 > 0) hi = a; else lo = a;
        T an = p.nxt;
        if (hi < 0) { if (!(an > a * (T)1.1)) an = a * (T)1.1; if (an > a * 4) an = a * 4; }
        else {
          const T w = hi - lo, mid = (T)0.5 * (lo + hi);
          if (w < wrel * hi) { result = fallbackAlpha(have, bestA, bestC, c0); done = true; }
          if ((!armijo && p.d1 < 0) || !(an > lo && an < hi) || (k >= 2 && (k & 1) == 0 && w > (T)0.5 * wprev)) an = mid;
          if ((k & 1) == 0) wprev = w;
        }
        if (!kready) {                                    // apex positions of the wheel pairs (first rejected point only)
          kready = true;
#pragma unroll
          for (int c = 0; c < 3; c++) {
            T kv = -1;
            if (c < nw) {
              const T* ls = S.wrec + c * CRW + OSW + O_LS;
              const T UV = ls[3], VV = ls[4];
              if (VV > (T)1e-30) kv = -fdivPos(UV, VV);
            }
            if (c == 0) kk0 = kv; else if (c == 1) kk1 = kv; else kk2 = kv;
          }
        }
        int kb = -1; T kbest = 0;
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const T kv = c == 0 ? kk0 : (c == 1 ? kk1 : kk2);
          const bool between = an > a ? (kv > a && kv < an) : (kv < a && kv > an);
          if (!((kused >> c) & 1u) && kv > 0 && between && (kb < 0 || babs(kv - a) < babs(kbest - a))) { kb = c; kbest = kv; }
        }
        if (kb >= 0) { an = kbest; kused |= 1u << kb; }
        a = an;
        if (++k >= 30 && !done) { result = fallbackAlpha(have, bestA, bestC, c0); done = true; }
      }
      if (done) { if (U) { on = false; continue; } else break; }
    }
    return result;
  }

  // exact line search of mj_solNewton (PrimalSearch) as a state machine around a single evaluation site.  Evaluated
  // points live in S.lsp (slots of 4 words); the brackets p1 / p2 and their pending Newton points are slot indices, so
  // "p1 = candidate" is an integer move and the whole search needs a handful of registers.
  __device__ __forceinline__ T lineSearch(T scale, bool on = true) {
    const bool dof = L.gl < NV;
    if (G == 16 || L.gl < 16) S.vb[0][L.gl] = dof ? search : (T)0;
    __syncwarp(L.mask);
    Mv = gSymv(S, S.vb[0], L.gi);
    const Sum3<T> r3 = gsum3(dof ? search * search : (T)0, dof ? search * (Ma - qfs) : (T)0, dof ? (T)0.5 * search * Mv : (T)0, L.mask);
    const T sn2 = r3.a;
    if (U) on = on && !(sn2 < (T)1e-30);
    else if (sn2 < (T)1e-30) return 0;
    T snorm, isn; sqrtInv(sn2, snorm, isn);
    T gtol = mc.tolerance * mc.ls_tolerance * snorm / scale;
    // jv = J search and the per-contact coefficients of the search (PrimalPrepare)
#pragma unroll 1
    for (int c = L.gl; c < ncon; c += G) {
      const bool wheel = c < nd;
      T* rec = recOf(c);
      T* sc = rec + (wheel ? OSW : OSH);
      T w[3]; rowsDot(rec, wheel, S.vb[0], w);
      sc[O_JV] = w[0]; sc[O_JV + 1] = w[1]; sc[O_JV + 2] = w[2];
      const int k = c < nw ? 0 : 1;
      const T D0 = sc[O_D0], D1 = D0 * mc.d1r[k], D2 = D0 * mc.d2r[k];
      const T mu = mc.mu[k], f1 = mc.f1[k], f2 = mc.f2[k];
      const T j0 = sc[O_JAR], j1 = sc[O_JAR + 1], j2 = sc[O_JAR + 2];
      const T u1 = j1 * f1, u2 = j2 * f2, v1 = w[1] * f1, v2 = w[2] * f2;
      T* ls = sc + O_LS;
      ls[0] = j0 * mu; ls[1] = w[0] * mu;
      ls[2] = u1 * u1 + u2 * u2; ls[3] = u1 * v1 + u2 * v2; ls[4] = v1 * v1 + v2 * v2;
      ls[5] = (T)0.5 * (D0 * j0 * j0 + D1 * j1 * j1 + D2 * j2 * j2);
      ls[6] = D0 * j0 * w[0] + D1 * j1 * w[1] + D2 * j2 * w[2];
      ls[7] = (T)0.5 * (D0 * w[0] * w[0] + D1 * w[1] * w[1] + D2 * w[2] * w[2]);
    }
    LsCtx<T> q;
    q.qG0 = gauss; q.qG1 = r3.b; q.qG2 = r3.c;
    __syncwarp(L.mask);
    if (isFast()) return searchFast(q, gtol, on);
    // states: 0 p0, 1 first Newton point, 2 one-sided Newton iteration, 3 p1next, 4 midpoint, 5 p1 re-bracket, 6 p2 re-bracket
    const T* P = S.lsp;
    int i1 = 0, i2 = 0, i1n = 0, i2n = 0, imid = 0, ic0 = 0;   // slots of p1, p2, p1next, p2next, pmid, candidate 0
    T p0cost = 0, dir = 1, result = 0, a = 0;
    int st = 0, it = 0, b1 = 0, dst = 0;
    const int maxit = mc.ls_iterations;
#pragma unroll 1
    for (;;) {
      if (U && !__any_sync(0xffffffffu, on)) break;      // both searches of the warp have finished
      const LsPt<T> p = lsEval(mc, S, gs, ncon, nw, nd, q, a, dst, L, rec0);
      if (U && !on) continue;                             // passenger: state frozen
      nevals++;
      bool done = false;
      int src = -1;          // slot whose Newton step is evaluated next (-1: `a` has been set explicitly)
      if (st == 0) {
        p0cost = p.cost;
        a = p.nxt; st = 1;
      } else if (st <= 2) {
        T d = p.d1;
        if (st == 1) {       // p1 = better of p0, first Newton point
          i1 = p0cost < p.cost ? 0 : dst;
          d = P[i1 * LSP_W + LSP_D1];
          dir = d < 0 ? (T)1 : (T)-1;
        } else { i1 = dst; it++; }
        if (babs(d) < gtol || it >= maxit) { result = P[i1 * LSP_W + LSP_ALPHA]; done = true; }
        else if (d * dir <= -gtol) { i2 = i1; src = i1; st = 2; }          // keep iterating on the same side
        else if (st == 1) { result = P[i1 * LSP_W + LSP_ALPHA]; done = true; }   // (non-finite derivative)
        else { i2n = i1; src = i1; st = 3; }                              // crossed the minimum: p2 .. p1 is a bracket
      } else if (st == 3) {
        i1n = dst; st = 4;
        a = (T)0.5 * (P[i1 * LSP_W + LSP_ALPHA] + P[i2 * LSP_W + LSP_ALPHA]);
      } else {
        // candidates of this trip = {p1next, p2next, pmid} as they were when the midpoint was evaluated (the reference
        // copies them before the first bracket update, so the second update still sees the old p1next: slot ic0)
        if (st == 4) { imid = dst; it++; b1 = -1; ic0 = i1n; }
        else if (st == 5) i1n = dst;
        else i2n = dst;
        const T dc0 = P[ic0 * LSP_W + LSP_D1], dc1 = P[i2n * LSP_W + LSP_D1], dc2 = P[imid * LSP_W + LSP_D1];
        if (st == 4) {
          if (babs(dc0) < gtol) { result = P[ic0 * LSP_W + LSP_ALPHA]; done = true; }
          else if (babs(dc1) < gtol) { result = P[i2n * LSP_W + LSP_ALPHA]; done = true; }
          else if (babs(dc2) < gtol) { result = p.alpha; done = true; }
        }
        if (!done) {
          bool need = false;
#pragma unroll 1
          for (int side = st == 4 ? 0 : 1; side < 2 && !need && st != 6; side++) {   // bracket(p1, ..), then bracket(p2, ..)
            int ip = side ? i2 : i1, flag = 0;
            T dp = P[ip * LSP_W + LSP_D1];
#pragma unroll
            for (int i = 0; i < 3; i++) {
              const T dc = i == 0 ? dc0 : (i == 1 ? dc1 : dc2);
              const int ic = i == 0 ? ic0 : (i == 1 ? i2n : imid);
              if (dp < 0 && dc < 0 && dp < dc) { ip = ic; dp = dc; flag = 1; }
              else if (dp > 0 && dc > 0 && dp > dc) { ip = ic; dp = dc; flag = 2; }
            }
            if (side == 0) { i1 = ip; b1 = flag; } else i2 = ip;
            if (flag) { src = ip; st = 5 + side; need = true; }
            else if (side == 1 && !b1) { result = P[imid * LSP_W + LSP_COST] < p0cost ? P[imid * LSP_W + LSP_ALPHA] : (T)0; done = true; }
          }
          if (!need && !done) {   // next trip of the bracketing loop
            if (it < maxit) { a = (T)0.5 * (P[i1 * LSP_W + LSP_ALPHA] + P[i2 * LSP_W + LSP_ALPHA]); st = 4; }
            else {
              const T c1 = P[i1 * LSP_W + LSP_COST], c2 = P[i2 * LSP_W + LSP_COST];
              if (c1 <= c2 && c1 < p0cost) result = P[i1 * LSP_W + LSP_ALPHA];
              else if (c2 <= c1 && c2 < p0cost) result = P[i2 * LSP_W + LSP_ALPHA];
              else result = 0;
              done = true;
            }
          }
        }
      }
      if (done) { if (U) { on = false; continue; } else break; }
      if (src >= 0) a = P[src * LSP_W + LSP_NXT];
      // next evaluation goes to a slot that none of the live points occupies (slot 0 = p0 stays)
      const unsigned live = 1u | (1u << i1) | (1u << i2) | (1u << i1n) | (1u << i2n) | (1u << imid) | (1u << ic0);
      dst = __ffs(~live) - 1;
    }
    return result;
  }

  // fp32 only: the reference's stopping rule compares the cost improvement with tolerance / scale ~ 1e-8, far below the
  // rounding noise of a single-precision cost (~1e-7 |cost|); an improvement inside that noise ends the solve (without this
  // a few envs per 10^5 ran to the iteration cap and held their whole launch back)
  __device__ __forceinline__ static bool belowNoise(T old, T now) { return sizeof(T) == 4 && old - now <= (T)5e-7 * babs(old); }
  // ---- uniform-warp pieces (U): begin() = warm-start choice + first cost / gradient for the groups that are `fresh`; the other
  // group of the warp rides along with its state untouched (its per-contact loops are empty, register updates predicated,
  // costGrad() is idempotent for a group that did not move).  iterate() = one Newton iteration for the groups that are `on`.
  __device__ __forceinline__ void begin(T warm, bool fresh) {
    const bool dof = L.gl < NV;
    const int nc = fresh ? ncon : 0;
    // warm-start choice: total cost at qacc_warmstart (-> vb[0]) and at qacc_smooth (-> vb[1])
    if (fresh && (G == 16 || L.gl < 16)) { S.vb[0][L.gl] = dof ? warm : (T)0; S.vb[1][L.gl] = dof ? qas : (T)0; }
    __syncwarp(L.mask);
    const T mw = gSymv(S, S.vb[0], L.gi);
    T cw = dof ? (T)0.5 * (mw - qfs) * (warm - qas) : (T)0, cs = 0;
#pragma unroll 1
    for (int c = L.gl; c < nc; c += G) {
      const bool wheel = c < nd;
      T* rec = recOf(c);
      T* sc = rec + (wheel ? OSW : OSH);
      T jw[3], js[3], f[3], h[6]; int st;
      rowsDot(rec, wheel, S.vb[0], jw); rowsDot(rec, wheel, S.vb[1], js);
#pragma unroll
      for (int m = 0; m < 3; m++) { const T ar = sc[O_JV + m]; jw[m] -= ar; js[m] -= ar; }
      cw += coneLane(mc, c < nw ? 0 : 1, sc[O_D0], jw, f, h, st, false);
      cs += coneLane(mc, c < nw ? 0 : 1, sc[O_D0], js, f, h, st, false);
      // park both candidates: jar <- warm-start residual, LS[0..2] <- smooth residual
      sc[O_JAR] = jw[0]; sc[O_JAR + 1] = jw[1]; sc[O_JAR + 2] = jw[2];
      sc[O_LS] = js[0]; sc[O_LS + 1] = js[1]; sc[O_LS + 2] = js[2];
    }
    const Sum3<T> r3 = gsum3(cw, cs, (T)0, L.mask);
    const bool useSmooth = r3.a > r3.b;
    if (fresh) {
      if (useSmooth) {
        qacc = qas; Ma = qfs;
#pragma unroll 1
        for (int c = L.gl; c < nc; c += G) {
          T* sc = recOf(c) + (c < nd ? OSW : OSH);
          sc[O_JAR] = sc[O_LS]; sc[O_JAR + 1] = sc[O_LS + 1]; sc[O_JAR + 2] = sc[O_LS + 2];
        }
      } else { qacc = warm; Ma = mw; }
    }
    __syncwarp(L.mask);
    costGrad();
  }
  __device__ __forceinline__ bool iterate(bool on) {
    const T scale = (T)1 / (mc.meaninertia * (T)NV);
    search = -gHessSolve(S, gs, ncon, nd, grad, L);
    const T alpha = lineSearch(scale, on);
    const bool step = on && alpha != 0;               // alpha == 0 ends the solve without touching the state
    if (step) {
      qacc += alpha * search; Ma += alpha * Mv;
#pragma unroll 1
      for (int c = L.gl; c < ncon; c += G) {
        T* sc = recOf(c) + (c < nd ? OSW : OSH);
        sc[O_JAR] += alpha * sc[O_JV]; sc[O_JAR + 1] += alpha * sc[O_JV + 1]; sc[O_JAR + 2] += alpha * sc[O_JV + 2];
      }
    }
    const T old = cost;
    costGrad();                                       // idempotent for a group that did not step
    if (step) iter++;
    return step && iter < mc.iterations && !(scale * (old - cost) < mc.tolerance || scale * bsqrt(gnorm2) < mc.tolerance || belowNoise(old, cost));
  }

  // in: aref parked in the jv slot of every record, `warm` = qacc_warmstart of this dof lane.  Returns qacc; niter by reference.
  __device__ __forceinline__ T run(T warm, int& niter, bool act = true) {
    if (U) {
      iter = 0;
      begin(warm, true);
      bool on = act && iter < mc.iterations;
#pragma unroll 1
      while (__any_sync(0xffffffffu, on)) on = iterate(on);
      niter = iter;
      return qacc;
    }
    const bool dof = L.gl < NV;
    // warm-start choice: total cost at qacc_warmstart (-> vb[0]) and at qacc_smooth (-> vb[1])
    if (G == 16 || L.gl < 16) { S.vb[0][L.gl] = dof ? warm : (T)0; S.vb[1][L.gl] = dof ? qas : (T)0; }
    __syncwarp(L.mask);
    const T mw = gSymv(S, S.vb[0], L.gi);
    T cw = dof ? (T)0.5 * (mw - qfs) * (warm - qas) : (T)0, cs = 0;
#pragma unroll 1
    for (int c = L.gl; c < ncon; c += G) {
      const bool wheel = c < nd;
      T* rec = recOf(c);
      T* sc = rec + (wheel ? OSW : OSH);
      T jw[3], js[3], f[3], h[6]; int st;
      rowsDot(rec, wheel, S.vb[0], jw); rowsDot(rec, wheel, S.vb[1], js);
#pragma unroll
      for (int m = 0; m < 3; m++) { const T ar = sc[O_JV + m]; jw[m] -= ar; js[m] -= ar; }
      cw += coneLane(mc, c < nw ? 0 : 1, sc[O_D0], jw, f, h, st, false);
      cs += coneLane(mc, c < nw ? 0 : 1, sc[O_D0], js, f, h, st, false);
      // park both candidates: jar <- warm-start residual, LS[0..2] <- smooth residual
      sc[O_JAR] = jw[0]; sc[O_JAR + 1] = jw[1]; sc[O_JAR + 2] = jw[2];
      sc[O_LS] = js[0]; sc[O_LS + 1] = js[1]; sc[O_LS + 2] = js[2];
    }
    const Sum3<T> r3 = gsum3(cw, cs, (T)0, L.mask);
    const bool useSmooth = r3.a > r3.b;
    if (useSmooth) {
      qacc = qas; Ma = qfs;
#pragma unroll 1
      for (int c = L.gl; c < ncon; c += G) {
        T* sc = recOf(c) + (c < nd ? OSW : OSH);
        sc[O_JAR] = sc[O_LS]; sc[O_JAR + 1] = sc[O_LS + 1]; sc[O_JAR + 2] = sc[O_LS + 2];
      }
    } else { qacc = warm; Ma = mw; }
    __syncwarp(L.mask);
    const T scale = (T)1 / (mc.meaninertia * (T)NV);
    iter = 0;
    costGrad();
    {
#pragma unroll 1
    while (iter < mc.iterations) {
      search = -gHessSolve(S, gs, ncon, nd, grad, L);
      const T alpha = lineSearch(scale);
      if (alpha == 0) break;
      qacc += alpha * search; Ma += alpha * Mv;
#pragma unroll 1
      for (int c = L.gl; c < ncon; c += G) {
        T* sc = recOf(c) + (c < nd ? OSW : OSH);
        sc[O_JAR] += alpha * sc[O_JV]; sc[O_JAR + 1] += alpha * sc[O_JV + 1]; sc[O_JAR + 2] += alpha * sc[O_JV + 2];
      }
      const T old = cost;
      costGrad();
      iter++;
      if (scale * (old - cost) < mc.tolerance || scale * bsqrt(gnorm2) < mc.tolerance || belowNoise(old, cost)) break;
    }
    }
    niter = iter;
    return qacc;
  }
};

// ---------------------------------------------------------------------------------------------- contact records
// impedance / regulariser / reference acceleration (mj_makeImpedance, mj_referenceConstraint) -> scalar block
template <typename T>
__device__ __forceinline__ void finishRecord(const ModelConst<T>& mc, T* sc, int ty, T dist, const T* vel) {
  // (reciprocals of the model constants and ONE division for D0 = 1 / max(1e-15, R0), R0 = (1 - imp) dA / imp: the four fp64
  // divisions of the literal formula were 10 % of k_stage's instructions; the results differ by <= 2 ulp)
  const T x = babs(dist) * mc.inv_width;
  T imp;
  if (x >= 1) imp = mc.solimp[1];
  else if (x <= 0) imp = mc.solimp[0];
  else {
    const T om = (T)1 - x;
    const T y = x <= mc.solimp[3] ? x * x * mc.inv_mid : (T)1 - om * om * mc.inv_1mmid;
    imp = mc.solimp[0] + y * (mc.solimp[1] - mc.solimp[0]);
  }
  sc[O_D0] = imp / bmax((T)1e-15 * imp, ((T)1 - imp) * mc.dA[ty]);
  // aref is parked in the jv slot until the solver has formed jar = J qacc - aref
  sc[O_JV] = -mc.B * vel[0] - mc.K * imp * dist; sc[O_JV + 1] = -mc.B * vel[1]; sc[O_JV + 2] = -mc.B * vel[2];
}

// capsule against one heightfield prism; a single out-of-line copy keeps the stage kernel's code small
template <typename T>
__device__ __noinline__ bool gCapsulePrism(const V3<T>& p0, const V3<T>& p1, T r, const V3<T>& ta, const V3<T>& tb, const V3<T>& tc, T& dist, V3<T>& n, V3<T>& pos) {
  return capsulePrism(p0, p1, r, ta, tb, tc, dist, n, pos);
}
// dense constraint rows of a contact at `pos` with frame F: base columns (b1 = +1 when the base tree is body 2), hinge column of
// wheel wi (or -1), ball columns (the ball is body 1 of every robot pair) or zeros; returns the contact-frame velocity
template <typename T>
__device__ __forceinline__ void gDenseRows(const GS<T>& S, const T* ge, T* rec, const T* F, const V3<T>& pos, int wi, bool ball, T* vel) {
  const V3<T> pB = ld3(ge + GE_PB), pL = ld3(ge + GE_PL);
  const Rot<T> RB = {ld3(ge + GE_RB), ld3(ge + GE_RB + 3), ld3(ge + GE_RB + 6)};
  const Rot<T> RL = {ld3(ge + GE_RL), ld3(ge + GE_RL + 3), ld3(ge + GE_RL + 6)};
  const V3<T> rB = pos - pB, rL = pos - pL;
  const V3<T> cb0 = cross(RB.c0, rB), cb1 = cross(RB.c1, rB), cb2 = cross(RB.c2, rB);
  const V3<T> cl0 = cross(rL, RL.c0), cl1 = cross(rL, RL.c1), cl2 = cross(rL, RL.c2);
  const int w = wi < 0 ? 0 : wi;
  V3<T> ah = cross(ld3(ge + GE_AW + 3 * w), pos - ld3(ge + GE_HW + 3 * w));
  if (wi < 0) ah = mk((T)0, (T)0, (T)0);
  const T bs = ball ? (T)1 : (T)0;
#pragma unroll 1
  for (int k = 0; k < 3; k++) {
    T* jr = rec + k * 16;
    const V3<T> fk = mk(F[3 * k], F[3 * k + 1], F[3 * k + 2]);
    jr[0] = fk.x; jr[1] = fk.y; jr[2] = fk.z;
    jr[3] = dot(fk, cb0); jr[4] = dot(fk, cb1); jr[5] = dot(fk, cb2);
    const T hq = dot(fk, ah);
    jr[6] = wi == 0 ? hq : (T)0; jr[7] = wi == 1 ? hq : (T)0; jr[8] = wi == 2 ? hq : (T)0;
    jr[9] = -bs * fk.x; jr[10] = -bs * fk.y; jr[11] = -bs * fk.z;
    jr[12] = bs * dot(fk, cl0); jr[13] = bs * dot(fk, cl1); jr[14] = bs * dot(fk, cl2); jr[15] = 0;
    T v = 0;
    for (int q = 0; q < NV; q++) v += jr[q] * S.xv[q];
    vel[k] = v;
  }
}
template <bool DBG, typename T> __device__ __forceinline__ void gDbgContact(double* dbg, int c, int ty, T dist, const V3<T>& pos, const T* F) {
  if (!DBG) return;
  double* o = dbg + 14 * c;
  o[0] = ty; o[1] = (double)dist; o[2] = (double)pos.x; o[3] = (double)pos.y; o[4] = (double)pos.z;
  for (int k = 0; k < 9; k++) o[5 + k] = (double)F[k];
}

// Contact generation with the constraint rows built in place by the lane that found the contact:
//   phase 1  lanes 0..5: ball x wheel capsules (patched sphere-capsule, tools/mujoco_fix.patch:9-18), ball x camera sticks
//            (same patched routine), ball x tower cylinder
//   phase 2  camera-stick and wheel capsules x heightfield prisms of each capsule's sub-grid (dense rows, appended)
//   phase 3  ball x heightfield prisms in the reference scan order
// Returns the contact count; nw = ball x wheel contacts (first), nd = all dense contacts (phases 1 + 2).
// DBG: every contact is also written to dbg[c] = {type, dist, pos3, frame9} (bb_probe_forward; never in the step kernels).
template <typename T, bool DBG = false>
__device__ __noinline__ int gCollide(const ModelConst<T>& mc, GS<T>& S, const float* __restrict__ hf, const float* __restrict__ mip, T zscale, T* gs, const Ln L,
                                     int& nwOut, int& ndOut, double* dbg = nullptr) {
  const int gl = L.gl;
  const unsigned lt = (1u << gl) - 1u;
  const int gsh = (threadIdx.x & 31) & ~(G - 1);     // bit position of the group's lane 0 in a ballot
  const unsigned gmask = (G == 32) ? 0xffffffffu : 0xffffu;
  const T* ge = S.geo;
  const V3<T> pB = ld3(ge + GE_PB), pL = ld3(ge + GE_PL);
  const Rot<T> RB = {ld3(ge + GE_RB), ld3(ge + GE_RB + 3), ld3(ge + GE_RB + 6)};
  const Rot<T> RL = {ld3(ge + GE_RL), ld3(ge + GE_RL + 3), ld3(ge + GE_RL + 6)};
  const V3<T> bc = pL + RL.c2 * mc.dz;
  const T br = mc.ball_r;
  // ---- phase 1: the ball against the robot's own geoms
  bool hit = false; T dist = 0; V3<T> n = mk((T)0, (T)0, (T)1), pos = n;
  V3<T> cu = n;
  if (gl < 3) {
    cu = ld3(ge + GE_CU + 3 * gl);
    hit = sphereCapsule(bc, br, ld3(ge + GE_CC + 3 * gl), cu, mc.wheel_r, mc.wheel_hl, dist, n, pos);
  } else if (gl < 5) {
    cu = rot(RB, ld3(mc.stick_u[gl - 3]));
    hit = sphereCapsule(bc, br, pB + rot(RB, ld3(mc.stick_c[gl - 3])), cu, mc.stick_r, mc.stick_hl, dist, n, pos);
  } else if (gl == 5) {
    hit = sphereCylinder(bc, br, pB + rot(RB, ld3(mc.tower_c)), RB.c2, mc.tower_r, mc.tower_hl, dist, n, pos);
  }
  unsigned m = (__ballot_sync(L.mask, hit) >> gsh) & gmask;
  const int nw = __popc(m & 7u);
  int nd = __popc(m);
  if (hit) {
    const int c = __popc(m & lt), ty = gl < 3 ? gl : (gl < 5 ? 10 + gl - 3 : 9);
    T* rec = crecD(S, gs, c);
    T F[9] = {n.x, n.y, n.z, cu.x, cu.y, cu.z, 0, 0, 0};
    makeFrame(F, gl < 5);
    T vel[3];
    gDenseRows(S, ge, rec, F, pos, gl < 3 ? gl : -1, true, vel);
    finishRecord(mc, rec + OSW, ty, dist, vel);
    S.cst[c] = 0;
    gDbgContact<DBG>(dbg, c, ty, dist, pos, F);
  }
  // ---- coarse heightfield reject (exact): lane g < 6 takes geom g (sticks, wheels, ball), reads the block maxima of the
  // 8 x 8-cell mip under the geom's sub-grid and compares with the geom's lowest point; a geom whose sub-grid lies entirely
  // below it would have every prism z-rejected, so its scan (phases 2 / 3) is skipped without touching the heightfield
  const T sx = mc.hx;
  bool need = false;
  if (gl < 6) {
    V3<T> lo, hi;
    if (gl < 5) {
      V3<T> cc; T rad, hl;
      if (gl < 2) { cc = pB + rot(RB, ld3(mc.stick_c[gl])); cu = rot(RB, ld3(mc.stick_u[gl])); rad = mc.stick_r; hl = mc.stick_hl; }
      else { cc = ld3(ge + GE_CC + 3 * (gl - 2)); cu = ld3(ge + GE_CU + 3 * (gl - 2)); rad = mc.wheel_r; hl = mc.wheel_hl; }
      const V3<T> p0 = cc - cu * hl, p1 = cc + cu * hl;
      lo = mk(bmin(p0.x, p1.x) - rad, bmin(p0.y, p1.y) - rad, bmin(p0.z, p1.z) - rad);
      hi = mk(bmax(p0.x, p1.x) + rad, bmax(p0.y, p1.y) + rad, bmax(p0.z, p1.z) + rad);
    } else { lo = mk(bc.x - br, bc.y - br, bc.z - br); hi = mk(bc.x + br, bc.y + br, bc.z + br); }
    HfGrid<T> gr;
    if (hfieldSubgrid(sx, mc.hbase, zscale, lo, hi, gr) && gr.cmax > gr.cmin && gr.rmax > gr.rmin) {
      float zm = -1e30f;
      for (int br_ = gr.rmin >> 3; br_ <= (gr.rmax - 1) >> 3; br_++)
        for (int bc_ = gr.cmin >> 3; bc_ <= (gr.cmax - 1) >> 3; bc_++) zm = fmaxf(zm, mip[br_ * MIPN + bc_]);
      need = !((T)zm * zscale < gr.zmin);
    }
  }
  const unsigned needm = (__ballot_sync(L.mask, need) >> gsh) & gmask;
  // ---- phase 2: camera sticks (gi 0, 1) and wheel capsules (gi 2..4) against the heightfield
#pragma unroll 1
  for (int gi = 0; gi < 5; gi++) {
    if (!((needm >> gi) & 1u)) continue;
    V3<T> cc; T rad, hl;
    if (gi < 2) { cc = pB + rot(RB, ld3(mc.stick_c[gi])); cu = rot(RB, ld3(mc.stick_u[gi])); rad = mc.stick_r; hl = mc.stick_hl; }
    else { cc = ld3(ge + GE_CC + 3 * (gi - 2)); cu = ld3(ge + GE_CU + 3 * (gi - 2)); rad = mc.wheel_r; hl = mc.wheel_hl; }
    const int ty = 4 + gi, wi = gi < 2 ? -1 : gi - 2;   // contact types 4, 5 (sticks) and 6..8 (wheels)
    const V3<T> p0 = cc - cu * hl, p1 = cc + cu * hl;
    const V3<T> lo = mk(bmin(p0.x, p1.x) - rad, bmin(p0.y, p1.y) - rad, bmin(p0.z, p1.z) - rad);
    const V3<T> hi = mk(bmax(p0.x, p1.x) + rad, bmax(p0.y, p1.y) + rad, bmax(p0.z, p1.z) + rad);
    HfGrid<T> gr;
    if (!hfieldSubgrid(sx, mc.hbase, zscale, lo, hi, gr)) continue;
    const int ncols = gr.cmax - gr.cmin, nrows = gr.rmax - gr.rmin, nprism = ncols > 0 && nrows > 0 ? 2 * ncols * nrows : 0;
    int cnt = 0;
#pragma unroll 1
    for (int base = 0; base < nprism && cnt < MAXH; base += G) {
      const int p = base + gl;
      hit = false;
      if (p < nprism) {
        const int cell = p >> 1, r = gr.rmin + cell / ncols, c = gr.cmin + cell % ncols;
        V3<T> ta, tb, tc; hfieldTriangle(hf, zscale, sx, gr.dx, r, c, p & 1, ta, tb, tc);
        if (!(ta.z < gr.zmin && tb.z < gr.zmin && tc.z < gr.zmin) && !capsuleAbovePlane(p0, p1, rad, ta, tb, tc))
          hit = gCapsulePrism(p0, p1, rad, ta, tb, tc, dist, n, pos);
      }
      m = (__ballot_sync(L.mask, hit) >> gsh) & gmask;
      if (m == 0) continue;
      const int rank = cnt + __popc(m & lt), c = nd + rank;
      if (hit && rank < MAXH && c < NDMAX) {
        T* rec = crecD(S, gs, c);
        T F[9] = {n.x, n.y, n.z, 0, 0, 0, 0, 0, 0};
        makeFrame(F, false);
        T vel[3];
        gDenseRows(S, ge, rec, F, pos, wi, false, vel);
        finishRecord(mc, rec + OSW, ty, dist, vel);
        S.cst[c] = 0;
        gDbgContact<DBG>(dbg, c, ty, dist, pos, F);
      }
      cnt += __popc(m);
    }
    nd += cnt < MAXH ? cnt : MAXH;
    nd = nd < NDMAX ? nd : NDMAX;
  }
  // ---- phase 3: ball vs heightfield prisms
  int cnt = 0;
  const bool skip = !((needm >> 5) & 1u);   // box test and mip test of lane 5
  if (!skip) {
    const T gsc = (T)(HN - 1) / ((T)2 * sx);
    int cmin = (int)bfloor((bc.x - br + sx) * gsc), cmax = (int)bceil((bc.x + br + sx) * gsc);
    int rmin = (int)bfloor((bc.y - br + sx) * gsc), rmax = (int)bceil((bc.y + br + sx) * gsc);
    cmin = cmin < 0 ? 0 : cmin; rmin = rmin < 0 ? 0 : rmin; cmax = cmax > HN - 1 ? HN - 1 : cmax; rmax = rmax > HN - 1 ? HN - 1 : rmax;
    const int ncols = cmax - cmin, nrows = rmax - rmin, nprism = ncols > 0 && nrows > 0 ? 2 * ncols * nrows : 0;
    const T dx = (T)2 * sx / (T)(HN - 1), zmin = bc.z - br;
#pragma unroll 1
    for (int base = 0; base < nprism && cnt < MAXH; base += G) {
      const int p = base + gl;
      hit = false;
      if (p < nprism) {
        const int cell = p >> 1, k = p & 1, r = rmin + cell / ncols, c = cmin + cell % ncols;
        const T x0 = dx * (T)c - sx, x1 = dx * (T)(c + 1) - sx, y0 = dx * (T)r - sx, y1 = dx * (T)(r + 1) - sx;
        const T ex = bc.x < x0 ? x0 - bc.x : (bc.x > x1 ? bc.x - x1 : (T)0), ey = bc.y < y0 ? y0 - bc.y : (bc.y > y1 ? bc.y - y1 : (T)0);
        if (ex * ex + ey * ey < br * br) {
          const T h00 = (T)hf[r * HN + c] * zscale, h10 = (T)hf[r * HN + c + 1] * zscale;
          const T h01 = (T)hf[(r + 1) * HN + c] * zscale, h11 = (T)hf[(r + 1) * HN + c + 1] * zscale;
          const V3<T> v01 = mk(x0, y1, h01), v00 = mk(x0, y0, h00), v11 = mk(x1, y1, h11), v10 = mk(x1, y0, h10);
          const V3<T> ta = k ? v00 : v01, tb = k ? v11 : v00, tc = k ? v10 : v11;
          if (!(ta.z < zmin && tb.z < zmin && tc.z < zmin)) {
            const V3<T> q = closestOnTriangle(bc, ta, tb, tc);
            const V3<T> dv = bc - q;
            V3<T> nn = cross(tb - ta, tc - ta); if (nn.z < 0) nn = -nn;
            if (dot(bc - ta, nn) < 0) {   // centre under the top plane: contact only inside this prism's column
              const V3<T> e1 = tb - ta, e2 = tc - ta, ap = bc - ta;
              const T u = e1.x * e2.y - e1.y * e2.x;
              const T sa = (ap.x * e2.y - ap.y * e2.x) / u, tt = (e1.x * ap.y - e1.y * ap.x) / u;
              if (!(sa < 0 || tt < 0 || sa + tt > 1)) {
                hit = true; n = nn * ((T)1 / bsqrt(dot(nn, nn))); dist = dot(ap, n) - br; pos = bc - n * (br + (T)0.5 * dist);
              }
            } else {
              const T dl = bsqrt(dot(dv, dv));
              if (dl < br && dl >= (T)1e-15) { hit = true; dist = dl - br; n = dv * ((T)1 / dl); pos = q + n * ((T)0.5 * dist); }
            }
          }
        }
      }
      m = (__ballot_sync(L.mask, hit) >> gsh) & gmask;
      if (m == 0) continue;
      const int rank = cnt + __popc(m & lt);
      if (hit && rank < MAXH) {
        const int c = nd + rank;
        T* rec = crecH(S, gs, rank);
        T F[9] = {n.x, n.y, n.z, 0, 0, 0, 0, 0, 0};
        makeFrame(F, false);
        const V3<T> rL = pos - pL;
        // relative velocity = ball - world: linear +f, angular f . (c_m x ... ) with the sign of body 2
        const V3<T> cl0 = cross(RL.c0, rL), cl1 = cross(RL.c1, rL), cl2 = cross(RL.c2, rL);
        T vel[3];
#pragma unroll 1
        for (int k = 0; k < 3; k++) {
          T* jr = rec + k * 8;
          const V3<T> fk = mk(F[3 * k], F[3 * k + 1], F[3 * k + 2]);
          jr[0] = 0; jr[1] = fk.x; jr[2] = fk.y; jr[3] = fk.z;
          jr[4] = dot(fk, cl0); jr[5] = dot(fk, cl1); jr[6] = dot(fk, cl2); jr[7] = 0;
          T v = 0;
          for (int q = 1; q < 7; q++) v += jr[q] * S.xv[8 + q];
          vel[k] = v;
        }
        finishRecord(mc, rec + OSH, 3, dist, vel);
        S.cst[c] = 0;
        gDbgContact<DBG>(dbg, c, 3, dist, pos, F);
      }
      cnt += __popc(m);
    }
    cnt = cnt < MAXH ? cnt : MAXH;
  }
  __syncwarp(L.mask);
  nwOut = nw; ndOut = nd;
  return nd + cnt;
}

// ---------------------------------------------------------------------------------------------- one mj_forward (group)
// Everything of mj_forward before the constraint solver: kinematics / mass matrix / bias (gSmooth), contact generation and
// constraint rows (gCollide), qacc_smooth.  in: S.xq, S.xv, S.ctrl   out: contact count (nw wheel contacts first),
// qfrc_smooth / qacc_smooth of this dof lane, S.M, the contact records (and S.xq normalised, S.kin when wantKin).
// cta_sync: the warps of the CTA enter the three phases together (every thread of the CTA must make the call; `skip`
// marks threads that only take part in the barriers), so the large straight-line phase code is fetched once per CTA.
template <typename T, bool DBG = false>
__device__ __forceinline__ int gForwardPre(const ModelConst<T>& mc, GS<T>& S, const float* __restrict__ hf, const float* __restrict__ mip, T zscale, T* gs, const Ln L, bool wantKin,
                                           int& nw, int& nd, T& qfs, T& qas, bool skip = false, bool cta_sync = false, double* dbg = nullptr) {
  int ncon = 0;
  nw = 0; nd = 0; qfs = 0; qas = 0;
  if (cta_sync) __syncthreads();
  if (!skip) {
    if (L.gl == 0) normalizeQuats(S.xq);
    __syncwarp(L.mask);
    gSmooth(mc, S, wantKin);
    __syncwarp(L.mask);
    qfs = L.gl < NV ? S.vb[0][L.gi] : (T)0;
    __syncwarp(L.mask);
  }
  if (cta_sync) __syncthreads();
  if (!skip) ncon = gCollide<T, DBG>(mc, S, hf, mip, zscale, gs, L, nw, nd, dbg);   // consumes S.geo, which shares storage with the Cholesky factor
  if (cta_sync) __syncthreads();
  if (!skip) qas = gMassSolve(S, qfs, L);                     // qacc_smooth = M^-1 qfrc_smooth
  return ncon;
}
// in: S.xq, S.xv, S.ctrl, warm (dof-lane register)   out: returns qacc of this dof lane (and S.xq normalised).
template <typename T, bool DBG = false>
__device__ __noinline__ T gForward(const ModelConst<T>& mc, GS<T>& S, const float* __restrict__ hf, const float* __restrict__ mip, T zscale, T* gs, const Ln L, T warm, bool fast,
                                   bool wantKin, int& nconOut, int& niterOut, T* qasOut = nullptr, T* qfsOut = nullptr, double* dbg = nullptr) {
  int nw, nd; T qfs, qas;
  const int ncon = gForwardPre<T, DBG>(mc, S, hf, mip, zscale, gs, L, wantKin, nw, nd, qfs, qas, false, false, dbg);
  if (qasOut) { *qasOut = qas; *qfsOut = qfs; }
  nconOut = ncon; niterOut = 0;
  if (ncon == 0) return qas;
  GNewton<T> nwt(mc, S, gs, L, ncon, nw, nd, fast, qfs, qas);
  int niter;
  const T qacc = nwt.run(warm, niter);
  niterOut = niter;
  return qacc;
}

// quaternion/position integration is done by lane 0 through one shared (non-inlined) copy of the code
template <typename T> __device__ __noinline__ void gIntegrate(T* dst, const T* src, const T* vel, T h) {
  for (int i = 0; i < NQ; i++) dst[i] = src[i];
  integratePos(dst, vel, h);
}

// ---------------------------------------------------------------------------------------------- RK4 (group)
// in: S.xq/S.xv = state, S.ctrl, warm ; out: S.xq/S.xv = new state, warm = last-stage qacc, S.kin = last-stage kinematics.
// qlast (global, NQ) receives the last-stage configuration when non-null.
template <typename T>
__device__ void gRk4(const ModelConst<T>& mc, GS<T>& S, const float* __restrict__ hf, const float* __restrict__ mip, T zscale, T* gs, T* qlast, const Ln L, T& warm, bool chain_warm,
                     int& ncmax, int& nitsum) {
  const T h = mc.timestep;
  const bool dof = L.gl < NV;
  if (L.gl == 0) normalizeQuats(S.xq);
  __syncwarp(L.mask);
  for (int k = L.gl; k < NQ; k += G) S.q0[k] = S.xq[k];
  const T v0 = S.xv[L.gi];
  T xv = v0, sumv = 0, suma = 0, qacc = 0;
  __syncwarp(L.mask);
  ncmax = 0; nitsum = 0;
#pragma unroll 1
  for (int st = 0; st < 5; st++) {
    if (st < 4) {
      int nc, ni;
      qacc = gForward(mc, S, hf, mip, zscale, gs, L, warm, chain_warm, st == 3, nc, ni);
      ncmax = nc > ncmax ? nc : ncmax; nitsum += ni;
      const T bw = (st == 0 || st == 3) ? (T)(1.0 / 6.0) : (T)(1.0 / 3.0);
      sumv += bw * xv; suma += bw * qacc;
      if (st == 3 && qlast) { for (int k = L.gl; k < NQ; k += G) qlast[k] = S.xq[k]; }
      // fast mode: stages 2..4 start their Newton solve from the previous stage's solution instead of the previous
      // step's qacc_warmstart (same unique minimiser within the solver tolerance, fewer iterations)
      if (chain_warm && st < 3) warm = qacc;
      __syncwarp(L.mask);
    }
    // stage advance (st < 3: X0 + a_st h (v_st, acc_st)) or final update (st == 4: X0 + h sum_j B_j (v_j, acc_j))
    if (st != 3) {
      const T ha = st == 4 ? h : ((st == 2) ? h : (T)0.5 * h);
      if (st == 4) { if (G == 16 || L.gl < 16) S.vb[0][L.gl] = dof ? sumv : (T)0; __syncwarp(L.mask); }
      if (L.gl == 0) gIntegrate(S.xq, S.q0, st == 4 ? (const T*)S.vb[0] : (const T*)S.xv, ha);
      __syncwarp(L.mask);
      xv = v0 + ha * (st == 4 ? suma : qacc);
      if (st == 4) warm = qacc;
      if (dof) S.xv[L.gl] = xv;
      __syncwarp(L.mask);
    }
  }
}


// ---------------------------------------------------------------------------------------------- split-phase step
// The fused kernel (gRk4) keeps a whole RK4 step of one env inside one warp.  Its instruction stream is ~150 KB and the
// single-warp CTAs of an SM sit at unrelated program counters, so the kernel is bound by instruction fetch (ncu:
// no_instruction is the top stall, fp32 and higher occupancy buy nothing).  The split-phase path runs every RK stage as
// two launches -- k_stage (state advance, smooth dynamics, collision, constraint rows: straight-line code that all warps
// stream through together) and k_newton (only the solver loop, ~30 KB of hot code) -- and parks the per-env context in
// HBM between them: RK bookkeeping (RKN words) and the solver input (CTXN words: M, qfrc_smooth, qacc_smooth, contact
// records).  ~6 KB per env and stage, i.e. < 0.3 ms per step of HBM time at 65,536 envs.
constexpr int RK_Q0 = 0, RK_V0 = 20, RK_SUMV = 36, RK_SUMA = 52, RK_XV = 68, RK_QACC = 84, RK_WARM = 100, RK_CTRL = 116, RK_KIN = 120, RKN = 136;
constexpr int CTX_M = 0, CTX_QFS = 120, CTX_QAS = 136, CTX_W = 152, CTX_H = CTX_W + 3 * CRW + 2, CTXN = CTX_H + NHS * CRH;
constexpr int META_NCON = 0, META_NW = 1, META_NIT = 2, META_FLAGS = 3;   // int[N][4]; META_NW = nw | nd << 8; flags: bit 0 bad, bits 8.. max ncon

template <typename T> __device__ __forceinline__ void gcopy(T* __restrict__ dst, const T* __restrict__ src, int n, const Ln L) {
  // n even, both 16-byte aligned
  for (int k = 2 * L.gl; k < n; k += 2 * G) *reinterpret_cast<typename V2T<T>::t*>(dst + k) = ld2(src + k);
}
// solver input of one env -> HBM (after gForwardPre) and back (before GNewton::run)
template <typename T> __device__ __forceinline__ void ctxSave(T* __restrict__ cx, const GS<T>& S, int ncon, int nd, T qfs, T qas, const Ln L) {
  gcopy(cx + CTX_M, S.M, MSZ, L);
  if (G == 16 || L.gl < 16) { cx[CTX_QFS + L.gl] = qfs; cx[CTX_QAS + L.gl] = qas; }
  gcopy(cx + CTX_W, S.wrec, (nd < NDS ? nd : NDS) * CRW, L);    // dense records beyond NDS already live in the global scratch
  const int nh = ncon - nd < NHS ? ncon - nd : NHS;
  gcopy(cx + CTX_H, S.hrec, nh * CRH, L);
}
template <typename T> __device__ __forceinline__ void ctxLoad(const T* __restrict__ cx, GS<T>& S, int ncon, int nd, T& qfs, T& qas, const Ln L) {
  gcopy(S.M, cx + CTX_M, MSZ, L);
  qfs = cx[CTX_QFS + L.gi]; qas = cx[CTX_QAS + L.gi];
  if (L.gl >= NV) { qfs = 0; qas = 0; }
  gcopy(S.wrec, cx + CTX_W, (nd < NDS ? nd : NDS) * CRW, L);
  const int nh = ncon - nd < NHS ? ncon - nd : NHS;
  gcopy(S.hrec, cx + CTX_H, nh * CRH, L);
  __syncwarp(L.mask);
}

}  // namespace bbg
