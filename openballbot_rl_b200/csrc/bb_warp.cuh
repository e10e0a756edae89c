// bb_warp.cuh -- warp-per-env forward dynamics + RK4 for the B200 ballbot engine (device only).
//
// One warp integrates one environment.  The solver state lives in shared memory (mass matrix, Hessian, constraint
// Jacobian rows, 15-vectors), per-contact solver quantities live in registers of the lane that owns the contact
// (contact c -> lane c % 32, slot c / 32), and the dense linear algebra is cooperative:
//   * 15x15 Cholesky with one matrix row per lane held in registers, columns exchanged by shuffles;
//   * Hessian assembly H = M + J' W J with the 120 packed entries spread over the lanes;
//   * J v per contact lane, J' f per dof lane with forces broadcast by shuffle;
//   * exact line search with warp-reduced cost/derivatives (identical sums on every lane => uniform control flow).
// The arithmetic follows the same restatement of mj_forward / mj_solNewton as bb_core.cuh (the thread-per-env
// reference implementation used by the probe and the CPU test harness); only summation orders differ.
#pragma once
#include "bb_core.cuh"

namespace bbw {
using namespace bb;

constexpr unsigned FULL = 0xffffffffu;
// Per-contact record (in T): 3 Jacobian rows (stride JST), weight block W(6), jar(3), jv(3), frc(3), D0,
// line-search coefficients LS(8) = U0 V0 UU UV VV q0 q1 q2 (valid during one line search).
constexpr int JST = 17;
constexpr int O_W = 3 * JST, O_JAR = O_W + 6, O_JV = O_JAR + 3, O_FRC = O_JV + 3, O_D0 = O_FRC + 3, O_LS = O_D0 + 1;
constexpr int CR = O_LS + 8;                // 75 words: odd stride => conflict-free one-contact-per-lane access
constexpr int NCS = 12;                     // contact records resident in shared memory; the rest spills to a global scratch
constexpr int GSCR = (NC - NCS) * CR;       // per-env global overflow scratch (in T)
constexpr int STG = 11;                     // staging entry: pos3 n3 dist type hint3

__constant__ unsigned char c_tri_i[NTRI], c_tri_j[NTRI];

template <typename T> struct WS {
  T xq[20], xv[16], q0[20], v0[16], sumv[16], suma[16], warm[16], ctrl[4];
  T M[NTRI], H[NTRI];
  T qfs[16], qas[16], qacc[16], Ma[16], grad[16], search[16], Mv[16], Mgrad[16], col[16];
  T rec[NCS * CR];    // contact records; also the contact staging area during collision (NC * STG <= NCS * CR)
  unsigned char cst[64];   // per contact: bits 0-1 zone (0 satisfied, 1 quadratic, 2 cone), bit 2 heightfield pair
};

template <typename T> __device__ __forceinline__ T wsum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
template <typename T> __device__ __forceinline__ T* crec(WS<T>& S, T* gs, int c) { return c < NCS ? S.rec + c * CR : gs + (c - NCS) * CR; }
__device__ __forceinline__ float qdiv(float a, float b) { return a / b; }
__device__ __noinline__ double qdiv(double a, double b) { return a / b; }
__device__ __forceinline__ float qsqrt(float a) { return sqrtf(a); }
__device__ __noinline__ double qsqrt(double a) { return sqrt(a); }
__device__ __forceinline__ float brsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ double brsqrt(double x) { return rsqrt(x); }

// ---------------------------------------------------------------------------------------------- Cholesky + solve
// Factor the packed SPD matrix A (shared, NTRI) in place and solve A x = b (b, x shared 15-vectors, may alias).
// Right-looking: every lane keeps its (up to 4) packed entries in registers; per column the owners publish the raw
// column through `col` (shared, 16), everybody scales by rsqrt(pivot) and updates its trailing entries.  Rolled loops
// keep the code small (instruction-cache footprint matters more here than shared-memory traffic).
template <typename T> __device__ __noinline__ void wCholSolve(T* A, T* col, const T* b, T* x, int lane) {
  int ei[4], ej[4]; T a[4];
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const int e = lane + 32 * m, ee = e < NTRI ? e : 0;
    ei[m] = c_tri_i[ee]; ej[m] = e < NTRI ? c_tri_j[ee] : 99; a[m] = A[ee];
  }
  // entries of surplus slots (e >= NTRI) get ej = 99 / ei = 0: they never publish and their updates are discarded
#pragma unroll 1
  for (int j = 0; j < NV; j++) {
#pragma unroll
    for (int m = 0; m < 4; m++) if (ej[m] == j) col[ei[m]] = a[m];
    __syncwarp();
    T piv = col[j];
    piv = piv < (T)1e-15 ? (T)1e-15 : piv;
    const T inv = brsqrt(piv);
    const T inv2 = inv * inv;
#pragma unroll
    for (int m = 0; m < 4; m++) {   // branch-free: selects instead of divergent blocks
      const int cj = ej[m] < NV ? ej[m] : 0;
      const T ci = col[ei[m]], cjv = col[cj];
      const T upd = a[m] - ci * cjv * inv2;                                  // trailing update  A(i,k) -= L(i,j) L(k,j)
      const T fin = (ei[m] == j) ? piv * inv : a[m] * inv;                   // final L(i,j)
      a[m] = ej[m] == j ? fin : (ej[m] > j ? upd : a[m]);
    }
    __syncwarp();
  }
#pragma unroll
  for (int m = 0; m < 4; m++) { const int e = lane + 32 * m; if (e < NTRI) A[e] = a[m]; }
  __syncwarp();
  // inverse diagonal once, then forward (L y = b) and backward (L' x = y) substitution with one value per dof lane
  const T myinv = lane < NV ? qdiv((T)1, A[lane * (lane + 1) / 2 + lane]) : (T)0;
  const int rowoff = lane < NV ? lane * (lane + 1) / 2 : 0;
  T y = lane < NV ? b[lane] : (T)0;
#pragma unroll 1
  for (int k = 0; k < NV; k++) {
    const T yk = __shfl_sync(FULL, y * myinv, k);
    if (lane == k) y = yk; else if (lane > k && lane < NV) y -= A[rowoff + k] * yk;
  }
#pragma unroll 1
  for (int k = NV - 1; k >= 0; k--) {
    const T xk = __shfl_sync(FULL, y * myinv, k);
    if (lane == k) y = xk; else if (lane < k) y -= A[k * (k + 1) / 2 + lane] * xk;
  }
  __syncwarp();
  if (lane < NV) x[lane] = y;
  __syncwarp();
}
// r = A v for packed symmetric A (shared); lane i < 15 owns r[i]
template <typename T> __device__ __forceinline__ T wSymvRow(const T* A, const T* v, int lane) {
  T acc = 0;
  if (lane < NV) {
#pragma unroll
    for (int j = 0; j < NV; j++) acc += A[tidx(lane, j)] * v[j];
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------- collision
// Fills the staging area (aliasing S.J) with the penetrating contacts in the reference scan order; returns their count.
template <typename T>
__device__ int wCollide(const ModelConst<T>& mc, const Geo<T>& g, const V3<T>* capC, const V3<T>* capU, const float* __restrict__ hf,
                        T zscale, T* stage, int lane) {
  const unsigned lt = (1u << lane) - 1u;
  const V3<T> bc = g.pL + rot(g.RL, mk((T)0, (T)0, mc.dz));
  const T br = mc.ball_r;
  bool hit = false; T dist = 0; V3<T> n = mk((T)0, (T)0, (T)1), pos = n, hint = mk((T)0, (T)0, (T)0);
  if (lane < 3) {   // patched sphere-capsule pairs (tools/mujoco_fix.patch:9-18)
    T x = dot(capU[lane], bc - capC[lane]);
    x = x > mc.wheel_hl ? mc.wheel_hl : (x < -mc.wheel_hl ? -mc.wheel_hl : x);
    const V3<T> dif = capC[lane] + capU[lane] * x - bc;
    const T cd = bsqrt(dot(dif, dif)), mind = br + mc.wheel_r;
    if (cd < mind) { hit = true; n = dif * ((T)1 / cd); dist = cd - mind; pos = bc + n * (br + (T)0.5 * dist); hint = capU[lane]; }
  }
  unsigned m = __ballot_sync(FULL, hit);
  if (hit) {
    T* e = stage + __popc(m & lt) * STG;
    st3(e, pos); st3(e + 3, n); e[6] = dist; e[7] = (T)lane; st3(e + 8, hint);
  }
  int nc = __popc(m);
  const T sx = mc.hx;
  const bool skip = (sx < bc.x - br) || (-sx > bc.x + br) || (sx < bc.y - br) || (-sx > bc.y + br) || (zscale < bc.z - br) || (-mc.hbase > bc.z + br);
  if (!skip) {
    const T gsc = (T)(HN - 1) / ((T)2 * sx);
    int cmin = (int)bfloor((bc.x - br + sx) * gsc), cmax = (int)bceil((bc.x + br + sx) * gsc);
    int rmin = (int)bfloor((bc.y - br + sx) * gsc), rmax = (int)bceil((bc.y + br + sx) * gsc);
    cmin = cmin < 0 ? 0 : cmin; rmin = rmin < 0 ? 0 : rmin; cmax = cmax > HN - 1 ? HN - 1 : cmax; rmax = rmax > HN - 1 ? HN - 1 : rmax;
    const int ncols = cmax - cmin, nrows = rmax - rmin, nprism = ncols > 0 && nrows > 0 ? 2 * ncols * nrows : 0;
    const T dx = (T)2 * sx / (T)(HN - 1), zmin = bc.z - br;
    int cnt = 0;
    for (int base = 0; base < nprism && cnt < MAXH; base += 32) {
      const int p = base + lane;
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
      m = __ballot_sync(FULL, hit);
      const int rank = cnt + __popc(m & lt);
      if (hit && rank < MAXH) {
        T* e = stage + (nc + rank) * STG;
        st3(e, pos); st3(e + 3, n); e[6] = dist; e[7] = (T)3; e[8] = 0; e[9] = 0; e[10] = 0;
      }
      cnt += __popc(m);
    }
    nc += cnt < MAXH ? cnt : MAXH;
  }
  __syncwarp();
  return nc;
}

// ---------------------------------------------------------------------------------------------- per-contact cone math
// zone logic of mj_constraintUpdate for one elliptic contact; returns cost; state: 0 satisfied, 1 quadratic, 2 cone
template <typename T>
__device__ __forceinline__ T coneLane(const ModelConst<T>& mc, int k, T D0, const T* jar, T* frc, T* h, int& state) {
  const T mu = mc.mu[k], f1 = mc.f1[k], f2 = mc.f2[k];
  const T D1 = D0 * mc.d1r[k], D2 = D0 * mc.d2r[k];
  const T U0 = jar[0] * mu, U1 = jar[1] * f1, U2 = jar[2] * f2;
  const T N = U0, Tn = qsqrt(U1 * U1 + U2 * U2);
  if (N >= mu * Tn || (Tn <= 0 && N >= 0)) { frc[0] = frc[1] = frc[2] = 0; state = 0; return 0; }
  if (mu * N + Tn <= 0 || (Tn <= 0 && N < 0)) {
    frc[0] = -D0 * jar[0]; frc[1] = -D1 * jar[1]; frc[2] = -D2 * jar[2]; state = 1;
    h[0] = D0; h[1] = D1; h[2] = D2; h[3] = 0; h[4] = 0; h[5] = 0;
    return (T)0.5 * (D0 * jar[0] * jar[0] + D1 * jar[1] * jar[1] + D2 * jar[2] * jar[2]);
  }
  const T Dm = D0 * mc.dmr[k];
  const T NT = N - mu * Tn;
  frc[0] = -Dm * NT * mu;
  const T iT = qdiv((T)1, Tn);
  const T sc = -frc[0] * iT;
  frc[1] = sc * f1 * U1; frc[2] = sc * f2 * U2;
  state = 2;
  const T muN_T3 = mu * N * iT * iT * iT, dg = mu * mu - mu * N * iT;
  h[0] = Dm * mu * mu;
  h[1] = Dm * f1 * f1 * (muN_T3 * U1 * U1 + dg);
  h[2] = Dm * f2 * f2 * (muN_T3 * U2 * U2 + dg);
  h[3] = Dm * mu * f1 * (-mu * U1 * iT);
  h[4] = Dm * mu * f2 * (-mu * U2 * iT);
  h[5] = Dm * f1 * f2 * (muN_T3 * U1 * U2);
  return (T)0.5 * Dm * NT * NT;
}
// r[k] = J_c[k] . v for the contact record `rec` (v: shared 15-vector)
template <typename T> __device__ __forceinline__ void rowsDot(const T* rec, const T* v, T* out) {
  T a0 = 0, a1 = 0, a2 = 0;
#pragma unroll 5
  for (int m = 0; m < NV; m++) { const T vm = v[m]; a0 += rec[m] * vm; a1 += rec[JST + m] * vm; a2 += rec[2 * JST + m] * vm; }
  out[0] = a0; out[1] = a1; out[2] = a2;
}

template <typename T> struct LsCtx { T qG0, qG1, qG2; };
// per-contact line-search coefficients (PrimalPrepare), computed once per line search by the owning lane
template <typename T> __device__ __forceinline__ void lsPrepare(const ModelConst<T>& mc, WS<T>& S, T* gs, int lane, int ncon) {
  for (int c = lane; c < ncon; c += 32) {
    T* rec = crec(S, gs, c);
    const int k = (S.cst[c] >> 2) & 1;
    const T D0 = rec[O_D0], D1 = D0 * mc.d1r[k], D2 = D0 * mc.d2r[k];
    const T mu = mc.mu[k], f1 = mc.f1[k], f2 = mc.f2[k];
    const T j0 = rec[O_JAR], j1 = rec[O_JAR + 1], j2 = rec[O_JAR + 2], w0 = rec[O_JV], w1 = rec[O_JV + 1], w2 = rec[O_JV + 2];
    const T u1 = j1 * f1, u2 = j2 * f2, v1 = w1 * f1, v2 = w2 * f2;
    T* ls = rec + O_LS;
    ls[0] = j0 * mu; ls[1] = w0 * mu;
    ls[2] = u1 * u1 + u2 * u2; ls[3] = u1 * v1 + u2 * v2; ls[4] = v1 * v1 + v2 * v2;
    ls[5] = (T)0.5 * (D0 * j0 * j0 + D1 * j1 * j1 + D2 * j2 * j2);
    ls[6] = D0 * j0 * w0 + D1 * j1 * w1 + D2 * j2 * w2;
    ls[7] = (T)0.5 * (D0 * w0 * w0 + D1 * w1 * w1 + D2 * w2 * w2);
  }
}
// cost and derivatives of the 1-D line-search objective at alpha (PrimalEval); warp-uniform result
template <typename T>
__device__ __noinline__ LsPt<T> lsEval(const ModelConst<T>& mc, WS<T>& S, T* gs, int lane, int ncon, const LsCtx<T>& q, T alpha) {
  T cost = 0, d1 = 0, d2 = 0;
  for (int c = lane; c < ncon; c += 32) {
    const T* rec = crec(S, gs, c);
    const int k = (S.cst[c] >> 2) & 1;
    const T* ls = rec + O_LS;
    const T mu = mc.mu[k];
    const T U0 = ls[0], V0 = ls[1], UU = ls[2], UV = ls[3], VV = ls[4];
    const T N = U0 + alpha * V0, Tsq = UU + alpha * ((T)2 * UV + alpha * VV);
    bool bottom = false;
    if (Tsq <= 0) bottom = N < 0;
    else {
      const T Tn = qsqrt(Tsq);
      if (N >= mu * Tn) {}
      else if (mu * N + Tn <= 0) bottom = true;
      else {
        const T Dm = rec[O_D0] * mc.dmr[k];
        const T iT = qdiv((T)1, Tn);
        const T N1 = V0, T1 = (UV + alpha * VV) * iT, T2 = VV * iT - (UV + alpha * VV) * T1 * iT * iT;
        const T NT = N - mu * Tn, dNT = N1 - mu * T1;
        cost += (T)0.5 * Dm * NT * NT; d1 += Dm * NT * dNT; d2 += Dm * (dNT * dNT - NT * mu * T2);
      }
    }
    if (bottom) { cost += ls[5] + alpha * (ls[6] + alpha * ls[7]); d1 += ls[6] + (T)2 * alpha * ls[7]; d2 += (T)2 * ls[7]; }
  }
  // contacts live on lanes 0..ncon-1: reduce only over the populated half-warps
  if (ncon <= 8) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) { cost += __shfl_xor_sync(FULL, cost, o); d1 += __shfl_xor_sync(FULL, d1, o); d2 += __shfl_xor_sync(FULL, d2, o); }
    cost = __shfl_sync(FULL, cost, 0); d1 = __shfl_sync(FULL, d1, 0); d2 = __shfl_sync(FULL, d2, 0);
  } else if (ncon <= 16) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) { cost += __shfl_xor_sync(FULL, cost, o); d1 += __shfl_xor_sync(FULL, d1, o); d2 += __shfl_xor_sync(FULL, d2, o); }
    cost = __shfl_sync(FULL, cost, 0); d1 = __shfl_sync(FULL, d1, 0); d2 = __shfl_sync(FULL, d2, 0);
  } else { cost = wsum(cost); d1 = wsum(d1); d2 = wsum(d2); }
  LsPt<T> p; p.alpha = alpha;
  p.cost = q.qG0 + alpha * (q.qG1 + alpha * q.qG2) + cost; p.d1 = q.qG1 + (T)2 * alpha * q.qG2 + d1; p.d2 = (T)2 * q.qG2 + d2;
  if (p.d2 < (T)1e-15) p.d2 = (T)1e-15;
  return p;
}

// ---------------------------------------------------------------------------------------------- warp Newton solver
template <typename T> struct WNewton {
  const ModelConst<T>& mc; WS<T>& S; T* gs; int lane, ncon;
  LsCtx<T> q; T cost, gauss;
  bool fast;   // fast solver mode: inexact line search (stop when |phi'| <= 1e-3 |phi'(0)|), same minimiser of the outer problem
  __device__ WNewton(const ModelConst<T>& m, WS<T>& s, T* g, int l, int n, bool f) : mc(m), S(s), gs(g), lane(l), ncon(n), fast(f) {}

  // forces/zones/cost at the current jar, Hessian, factorisation, gradient, Newton direction (S.Mgrad)
  __device__ __forceinline__ void update() {
    T cpart = 0;
    for (int c = lane; c < ncon; c += 32) {
      T* rec = crec(S, gs, c);
      const int k = (S.cst[c] >> 2) & 1;
      T h[6], f[3]; int st;
      const T jr[3] = {rec[O_JAR], rec[O_JAR + 1], rec[O_JAR + 2]};
      cpart += coneLane(mc, k, rec[O_D0], jr, f, h, st);
      S.cst[c] = (unsigned char)(st | (k << 2));
      rec[O_FRC] = f[0]; rec[O_FRC + 1] = f[1]; rec[O_FRC + 2] = f[2];
      if (st) { for (int m = 0; m < 6; m++) rec[O_W + m] = h[m]; }
    }
    const T gpart = lane < NV ? (T)0.5 * (S.Ma[lane] - S.qfs[lane]) * (S.qacc[lane] - S.qas[lane]) : (T)0;
    gauss = wsum(gpart);
    cost = gauss + wsum(cpart);
    __syncwarp();
    // ---- H = M + sum_c J_c' W_c J_c (packed entries spread over the lanes) and grad = Ma - qfs - J' f (dof lanes)
    T acc[4]; int ei[4], ej[4];
#pragma unroll
    for (int m = 0; m < 4; m++) {
      const int e = lane + 32 * m, ee = e < NTRI ? e : 0;
      ei[m] = c_tri_i[ee]; ej[m] = c_tri_j[ee]; acc[m] = S.M[ee];
    }
    const int col = lane < NV ? lane : 0;
    T gacc = lane < NV ? S.Ma[lane] - S.qfs[lane] : (T)0;
    const int nres = ncon < NCS ? ncon : NCS;
#pragma unroll 1
    for (int c = 0; c < ncon; c++) {
      const int st = S.cst[c];
      if ((st & 3) == 0) continue;
      const T* rec = c < nres ? (const T*)(S.rec + c * CR) : (const T*)(gs + (c - NCS) * CR);
      const T* r0 = rec; const T* r1 = rec + JST; const T* r2 = rec + 2 * JST;
      gacc -= r0[col] * rec[O_FRC] + r1[col] * rec[O_FRC + 1] + r2[col] * rec[O_FRC + 2];
      const int jmin = (st & 4) ? 9 : 0;   // heightfield contacts only touch the ball dofs
      const T w0 = rec[O_W], w1 = rec[O_W + 1], w2 = rec[O_W + 2], w3 = rec[O_W + 3], w4 = rec[O_W + 4], w5 = rec[O_W + 5];
#pragma unroll
      for (int m = 0; m < 4; m++) if (ej[m] >= jmin) {
        const T a0 = r0[ei[m]], a1 = r1[ei[m]], a2 = r2[ei[m]], b0 = r0[ej[m]], b1 = r1[ej[m]], b2 = r2[ej[m]];
        acc[m] += a0 * (w0 * b0 + w3 * b1 + w4 * b2) + a1 * (w3 * b0 + w1 * b1 + w5 * b2) + a2 * (w4 * b0 + w5 * b1 + w2 * b2);
      }
    }
#pragma unroll
    for (int m = 0; m < 4; m++) { const int e = lane + 32 * m; if (e < NTRI) S.H[e] = acc[m]; }
    if (lane < NV) S.grad[lane] = gacc;
    __syncwarp();
    wCholSolve(S.H, S.col, S.grad, S.Mgrad, lane);
  }
  __device__ __forceinline__ int bracket(LsPt<T>& p, const LsPt<T>* cand, LsPt<T>& pnext) {
    int flag = 0;
    for (int i = 0; i < 3; i++) {
      if (p.d1 < 0 && cand[i].d1 < 0 && p.d1 < cand[i].d1) { p = cand[i]; flag = 1; }
      else if (p.d1 > 0 && cand[i].d1 > 0 && p.d1 > cand[i].d1) { p = cand[i]; flag = 2; }
    }
    if (flag) pnext = lsEval(mc, S, gs, lane, ncon, q, p.alpha - qdiv(p.d1, p.d2));
    return flag;
  }
  __device__ __forceinline__ T lineSearch(T scale) {
    const T sv = lane < NV ? S.search[lane] : (T)0;
    const T sn = qsqrt(wsum(sv * sv));
    if (sn < (T)1e-15) return 0;
    T gtol = mc.tolerance * mc.ls_tolerance * sn / scale;
    const T mv = wSymvRow(S.M, S.search, lane);
    if (lane < NV) S.Mv[lane] = mv;
    for (int c = lane; c < ncon; c += 32) { T* rec = crec(S, gs, c); rowsDot(rec, S.search, rec + O_JV); }
    lsPrepare(mc, S, gs, lane, ncon);
    q.qG0 = gauss;
    q.qG1 = wsum(lane < NV ? sv * (S.Ma[lane] - S.qfs[lane]) : (T)0);
    q.qG2 = wsum(lane < NV ? (T)0.5 * sv * mv : (T)0);
    __syncwarp();
    int it = 0;
    const LsPt<T> p0 = lsEval(mc, S, gs, lane, ncon, q, (T)0);
    if (fast) gtol = bmax(gtol, (T)1e-3 * babs(p0.d1));
    LsPt<T> p1 = lsEval(mc, S, gs, lane, ncon, q, p0.alpha - qdiv(p0.d1, p0.d2)), p2 = p0, pmid = p0, p1n = p0, p2n = p0;
    if (p0.cost < p1.cost) p1 = p0;
    if (babs(p1.d1) < gtol) return p1.alpha;
    const T dir = p1.d1 < 0 ? (T)1 : (T)-1;
    bool upd = false;
    while (p1.d1 * dir <= -gtol && it < mc.ls_iterations) {
      p2 = p1; upd = true;
      p1 = lsEval(mc, S, gs, lane, ncon, q, p1.alpha - qdiv(p1.d1, p1.d2)); it++;
      if (babs(p1.d1) < gtol) return p1.alpha;
    }
    if (it >= mc.ls_iterations || !upd) return p1.alpha;
    p2n = p1; p1n = lsEval(mc, S, gs, lane, ncon, q, p1.alpha - qdiv(p1.d1, p1.d2));
    while (it < mc.ls_iterations) {
      pmid = lsEval(mc, S, gs, lane, ncon, q, (T)0.5 * (p1.alpha + p2.alpha)); it++;
      const LsPt<T> cand[3] = {p1n, p2n, pmid};
      for (int i = 0; i < 3; i++) if (babs(cand[i].d1) < gtol) return cand[i].alpha;
      const int b1 = bracket(p1, cand, p1n), b2 = bracket(p2, cand, p2n);
      if (!b1 && !b2) return pmid.cost < p0.cost ? pmid.alpha : (T)0;
    }
    if (p1.cost <= p2.cost && p1.cost < p0.cost) return p1.alpha;
    if (p2.cost <= p1.cost && p2.cost < p0.cost) return p2.alpha;
    return 0;
  }
  // result in S.qacc; returns the iteration count. On entry rec[O_JV..] holds aref of every contact.
  __device__ int run() {
    // warm-start choice: total cost at qacc_warmstart (w=0) and at qacc_smooth (w=1)
    T cw[2];
#pragma unroll 1
    for (int w = 0; w < 2; w++) {
      const T* v = w ? S.qas : S.warm;
      const T mv = wSymvRow(S.M, v, lane);
      T part = lane < NV ? (T)0.5 * (mv - S.qfs[lane]) * (v[lane] - S.qas[lane]) : (T)0;
      for (int c = lane; c < ncon; c += 32) {
        const T* rec = crec(S, gs, c);
        T jr[3], f[3], h[6]; int st;
        rowsDot(rec, v, jr);
        jr[0] -= rec[O_JV]; jr[1] -= rec[O_JV + 1]; jr[2] -= rec[O_JV + 2];
        part += coneLane(mc, (S.cst[c] >> 2) & 1, rec[O_D0], jr, f, h, st);
      }
      cw[w] = wsum(part);
    }
    if (lane < NV) S.qacc[lane] = cw[0] > cw[1] ? S.qas[lane] : S.warm[lane];
    __syncwarp();
    const T ma = wSymvRow(S.M, S.qacc, lane);
    if (lane < NV) S.Ma[lane] = ma;
    for (int c = lane; c < ncon; c += 32) {
      T* rec = crec(S, gs, c);
      T jr[3]; rowsDot(rec, S.qacc, jr);
      rec[O_JAR] = jr[0] - rec[O_JV]; rec[O_JAR + 1] = jr[1] - rec[O_JV + 1]; rec[O_JAR + 2] = jr[2] - rec[O_JV + 2];
    }
    __syncwarp();
    const T scale = (T)1 / (mc.meaninertia * (T)NV);
    int iter = 0; bool first = true; T old = 0;
#pragma unroll 1
    for (;;) {
      update();
      if (!first) {
        const T gv = lane < NV ? S.grad[lane] : (T)0;
        const T gn = wsum(gv * gv);
        iter++;
        if (scale * (old - cost) < mc.tolerance || scale * qsqrt(gn) < mc.tolerance) break;
      }
      first = false;
      if (iter >= mc.iterations) break;
      if (lane < NV) S.search[lane] = -S.Mgrad[lane];
      __syncwarp();
      const T alpha = lineSearch(scale);
      if (alpha == 0) break;
      if (lane < NV) { S.qacc[lane] += alpha * S.search[lane]; S.Ma[lane] += alpha * S.Mv[lane]; }
      for (int c = lane; c < ncon; c += 32) {
        T* rec = crec(S, gs, c);
        rec[O_JAR] += alpha * rec[O_JV]; rec[O_JAR + 1] += alpha * rec[O_JV + 1]; rec[O_JAR + 2] += alpha * rec[O_JV + 2];
      }
      __syncwarp();
      old = cost;
    }
    return iter;
  }
};

// ---------------------------------------------------------------------------------------------- one mj_forward (warp)
// in: S.xq, S.xv, S.ctrl, S.warm    out: S.qacc (and S.xq normalised).  kin: observation kinematics of this stage.
template <typename T>
__device__ __noinline__ void wForward(const ModelConst<T>& mc, WS<T>& S, const float* __restrict__ hf, T zscale, T* gs, KinOut<T>& kin, int lane,
                                      bool fast) {
  if (lane == 0) normalizeQuats(S.xq);
  for (int e = lane; e < NTRI; e += 32) S.M[e] = 0;
  __syncwarp();
  int ncon;
  {
    Geo<T> g; V3<T> capC[3], capU[3];
    // every lane evaluates the (small, serial) smooth dynamics redundantly; the stores to S.M / S.qfs carry identical values
    smoothDynamics<T, false>(mc, S.xq, S.xv, S.ctrl, S.M, S.qfs, g, capC, capU, &kin);
    __syncwarp();
    for (int e = lane; e < NTRI; e += 32) S.H[e] = S.M[e];
    __syncwarp();
    wCholSolve(S.H, S.col, S.qfs, S.qas, lane);   // qacc_smooth = M^-1 qfrc_smooth
    ncon = wCollide(mc, g, capC, capU, hf, zscale, S.rec, lane);
    if (ncon == 0) {
      if (lane < NV) S.qacc[lane] = S.qas[lane];
      __syncwarp();
      kin.ncon = 0; kin.niter = 0;
      return;
    }
    // ---- constraint assembly: every lane first reads its staged contacts (<= 2), then writes the records (they alias)
    T stg[2][STG];
#pragma unroll
    for (int sl = 0; sl < 2; sl++) {
      const int c = lane + 32 * sl;
      if (c < ncon) { for (int k = 0; k < STG; k++) stg[sl][k] = S.rec[c * STG + k]; }
    }
    __syncwarp();
#pragma unroll 1
    for (int sl = 0; sl < 2; sl++) {
      const int c = lane + 32 * sl;
      if (c >= ncon) continue;
      const T* sg = sl ? stg[1] : stg[0];
      const int ty = (int)sg[7];
      T F[9] = {sg[3], sg[4], sg[5], sg[8], sg[9], sg[10], 0, 0, 0};
      makeFrame(F, ty != 3);
      const V3<T> P = mk(sg[0], sg[1], sg[2]);
      const T dist = sg[6];
      const V3<T> rL = P - g.pL;
      T* rec = crec(S, gs, c);
      // world-frame relative-velocity (body2 - body1) Jacobian: [lin | base-angular 3x3 | hinge | -lin | ball-angular 3x3]
      const T sgn = ty != 3 ? (T)1 : (T)-1;      // ball is body1 for the wheel pairs, body2 for the terrain pair
      V3<T> cb[3], cl[3], ah = mk((T)0, (T)0, (T)0);
      cl[0] = cross(rL, g.RL.c0) * sgn; cl[1] = cross(rL, g.RL.c1) * sgn; cl[2] = cross(rL, g.RL.c2) * sgn;
      if (ty != 3) {
        const V3<T> rB = P - g.pB;
        cb[0] = cross(g.RB.c0, rB); cb[1] = cross(g.RB.c1, rB); cb[2] = cross(g.RB.c2, rB);
        const V3<T> aw = ty == 0 ? g.aw[0] : (ty == 1 ? g.aw[1] : g.aw[2]);
        const V3<T> hw = ty == 0 ? g.hw[0] : (ty == 1 ? g.hw[1] : g.hw[2]);
        ah = cross(aw, P - hw);
      } else { cb[0] = ah; cb[1] = ah; cb[2] = ah; }
      T vel[3];
#pragma unroll 1
      for (int k = 0; k < 3; k++) {
        T* jr = rec + k * JST;
        const V3<T> fk = mk(F[3 * k], F[3 * k + 1], F[3 * k + 2]);
        const T wl = ty != 3 ? (T)1 : (T)0;
        jr[0] = fk.x * wl; jr[1] = fk.y * wl; jr[2] = fk.z * wl;
        jr[3] = dot(fk, cb[0]); jr[4] = dot(fk, cb[1]); jr[5] = dot(fk, cb[2]);
        const T hq = dot(fk, ah);
        jr[6] = ty == 0 ? hq : (T)0; jr[7] = ty == 1 ? hq : (T)0; jr[8] = ty == 2 ? hq : (T)0;
        jr[9] = -fk.x * sgn; jr[10] = -fk.y * sgn; jr[11] = -fk.z * sgn;
        jr[12] = dot(fk, cl[0]); jr[13] = dot(fk, cl[1]); jr[14] = dot(fk, cl[2]);
        T v = 0;
        for (int m = 0; m < NV; m++) v += jr[m] * S.xv[m];
        vel[k] = v;
      }
      // impedance / regulariser / reference acceleration (mj_makeImpedance, mj_referenceConstraint)
      const T x = babs(dist) / mc.solimp[2];
      T imp;
      if (x >= 1) imp = mc.solimp[1];
      else if (x <= 0) imp = mc.solimp[0];
      else {
        const T mid = mc.solimp[3];
        const T y = x <= mid ? x * x / mid : (T)1 - ((T)1 - x) * ((T)1 - x) / ((T)1 - mid);
        imp = mc.solimp[0] + y * (mc.solimp[1] - mc.solimp[0]);
      }
      const T R0 = bmax((T)1e-15, ((T)1 - imp) * mc.dA[ty] / imp);
      rec[O_D0] = (T)1 / R0;
      S.cst[c] = (unsigned char)((ty == 3 ? 1 : 0) << 2);
      // aref is parked in the jv slot until the solver has formed jar = J qacc - aref
      rec[O_JV] = -mc.B * vel[0] - mc.K * imp * dist; rec[O_JV + 1] = -mc.B * vel[1]; rec[O_JV + 2] = -mc.B * vel[2];
    }
    __syncwarp();
  }
  WNewton<T> nw(mc, S, gs, lane, ncon, fast);
  const int niter = nw.run();
  kin.ncon = ncon; kin.niter = niter;
}

// quaternion/position integration is done by lane 0 through one shared (non-inlined) copy of the code
template <typename T> __device__ __noinline__ void wIntegrate(T* dst, const T* src, const T* vel, T h) {
  for (int i = 0; i < NQ; i++) dst[i] = src[i];
  integratePos(dst, vel, h);
}

// ---------------------------------------------------------------------------------------------- RK4 (warp)
// in: S.xq/S.xv = state, S.warm, S.ctrl ; out: S.xq/S.xv = new state, S.warm = last-stage qacc, kin = last-stage kinematics.
// qlast (global, NQ) receives the last-stage configuration when non-null.
template <typename T>
__device__ void wRk4(const ModelConst<T>& mc, WS<T>& S, const float* __restrict__ hf, T zscale, T* gs, KinOut<T>& kin, T* qlast, int lane,
                     bool chain_warm) {
  const T h = mc.timestep;
  if (lane == 0) normalizeQuats(S.xq);
  __syncwarp();
  if (lane < NQ) S.q0[lane] = S.xq[lane];
  if (lane < NV) { S.v0[lane] = S.xv[lane]; S.sumv[lane] = 0; S.suma[lane] = 0; }
  __syncwarp();
  int ncmax = 0, nit = 0;
#pragma unroll 1
  for (int st = 0; st < 5; st++) {
    if (st < 4) {
      wForward(mc, S, hf, zscale, gs, kin, lane, chain_warm);
      ncmax = kin.ncon > ncmax ? kin.ncon : ncmax; nit += kin.niter;
      const T bw = (st == 0 || st == 3) ? (T)(1.0 / 6.0) : (T)(1.0 / 3.0);
      if (lane < NV) { S.sumv[lane] += bw * S.xv[lane]; S.suma[lane] += bw * S.qacc[lane]; }
      if (st == 3 && qlast && lane < NQ) qlast[lane] = S.xq[lane];
      // fast mode: stages 2..4 start their Newton solve from the previous stage's solution instead of the previous
      // step's qacc_warmstart (same unique minimiser within the solver tolerance, far fewer iterations)
      if (chain_warm && st < 3 && lane < NV) S.warm[lane] = S.qacc[lane];
      __syncwarp();
    }
    // stage advance (st < 3: X0 + a_st h (v_st, acc_st)) or final update (st == 4: X0 + h sum_j B_j (v_j, acc_j))
    if (st != 3) {
      const T ha = st == 4 ? h : ((st == 2) ? h : (T)0.5 * h);
      const T* vel = st == 4 ? S.sumv : S.xv;
      const T* ac = st == 4 ? S.suma : S.qacc;
      if (lane == 0) wIntegrate(S.xq, S.q0, vel, ha);
      __syncwarp();
      if (lane < NV) { const T nv = S.v0[lane] + ha * ac[lane]; if (st == 4) S.warm[lane] = S.qacc[lane]; S.xv[lane] = nv; }
      __syncwarp();
    }
  }
  kin.ncon = ncmax; kin.niter = nit;
}

}  // namespace bbw
