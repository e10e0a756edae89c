// bb_core.cuh -- per-env physics core of the B200 ballbot engine (sm_100a), templated on the real type.
//
// One CUDA thread integrates one environment.  Everything here is written for THIS model
// (reference: ballbot_gym/models/ballbot.xml:35-93) instead of a generic kinematic tree:
//   * subsystem A = base + 2 camera bodies (welded) + 3 hinge wheels, formulated in the BASE-LOCAL frame with
//     classical Newton-Euler (no spatial c-frame), 9 dofs;
//   * subsystem L = the ball, 6 dofs, closed-form mass matrix;
//   * contacts are matrix-free: rows of efc_J are never stored, J*v / J'*f / J'WJ are evaluated from the
//     contact point and frame;
//   * the elliptic-cone Newton solver (MuJoCo mj_solNewton semantics, reference call ballbot_env.py:912)
//     works on a packed 15x15 lower-triangular Hessian in per-thread local memory.
// The functions are __host__ __device__ so that tests/hostcore can run the very same arithmetic on the CPU
// against the oracle without a GPU; the product path only ever calls them from the kernels in bb_engine.cu.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define BB_HD __host__ __device__ __forceinline__
#define BB_HDN __host__ __device__
#define BB_NOINL __host__ __device__ __noinline__
#else
#define BB_HD inline
#define BB_HDN
#define BB_NOINL inline
#endif

#if defined(BB_STATS) && !defined(__CUDA_ARCH__)
static long bb_stats_ls_evals = 0, bb_stats_newton = 0;   // host-only instrumentation for scripts/exp (never defined in the product build)
#endif

namespace bb {

constexpr int NQ = 17, NV = 15, HN = 293;
constexpr int MIPB = 8, MIPN = (HN - 1 + MIPB - 1) / MIPB, MIP_CELLS = MIPN * MIPN;   // block maxima of 8 x 8 cells (37 x 37 per field)
constexpr int MAXH = 50;          // mjMAXCONPAIR: max ball-hfield contacts per forward pass
constexpr int NC = 64;            // contact capacity of one forward pass (3 wheel pairs + <= 50 ball-terrain prisms + the other pairs)
// contact types (index of ModelConst::dA):
//   0..2  ball x wheel_i (explicit pairs, anisotropic friction, ballbot.xml:89-93)     3      heightfield x ball
//   4,5   heightfield x camera stick i (cam bodies are welded to the base)              6..8   heightfield x wheel_i capsule
//   9     ball x tower cylinder                                                         10,11  ball x camera stick i
constexpr int NCT = 12;
BB_HD bool ctHasBase(int ty) { return ty != 3; }
BB_HD bool ctHasBall(int ty) { return ty <= 3 || ty >= 9; }
BB_HD int ctWheel(int ty) { return ty <= 2 ? ty : ((ty >= 6 && ty <= 8) ? ty - 6 : -1); }
BB_HD int ctFric(int ty) { return ty <= 2 ? 0 : 1; }
constexpr int NTRI = NV * (NV + 1) / 2;

// ---------------------------------------------------------------------------------------------- math
BB_HD float bsqrt(float x) { return sqrtf(x); }
BB_HD double bsqrt(double x) { return sqrt(x); }
BB_HD float babs(float x) { return fabsf(x); }
BB_HD double babs(double x) { return fabs(x); }
BB_HD void bsincos(float a, float* s, float* c) { *s = sinf(a); *c = cosf(a); }
BB_HD void bsincos(double a, double* s, double* c) { *s = sin(a); *c = cos(a); }
BB_HD float batan2(float a, float b) { return atan2f(a, b); }
BB_HD double batan2(double a, double b) { return atan2(a, b); }
BB_HD float bfloor(float a) { return floorf(a); }
BB_HD double bfloor(double a) { return floor(a); }
BB_HD float bceil(float a) { return ceilf(a); }
BB_HD double bceil(double a) { return ceil(a); }
template <typename T> BB_HD T bmax(T a, T b) { return a > b ? a : b; }
template <typename T> BB_HD T bmin(T a, T b) { return a < b ? a : b; }

template <typename T> struct V3 { T x, y, z; };
template <typename T> BB_HD V3<T> mk(T x, T y, T z) { V3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <typename T> BB_HD V3<T> ld3(const T* p) { return mk(p[0], p[1], p[2]); }
template <typename T> BB_HD void st3(T* p, const V3<T>& v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
template <typename T> BB_HD V3<T> operator+(const V3<T>& a, const V3<T>& b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> BB_HD V3<T> operator-(const V3<T>& a, const V3<T>& b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> BB_HD V3<T> operator*(const V3<T>& a, T s) { return mk(a.x * s, a.y * s, a.z * s); }
template <typename T> BB_HD V3<T> operator*(T s, const V3<T>& a) { return mk(a.x * s, a.y * s, a.z * s); }
template <typename T> BB_HD V3<T> operator-(const V3<T>& a) { return mk(-a.x, -a.y, -a.z); }
template <typename T> BB_HD T dot(const V3<T>& a, const V3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> BB_HD V3<T> cross(const V3<T>& a, const V3<T>& b) {
  return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// 3x3 rotation stored by columns (c0,c1,c2): R v = c0 v.x + c1 v.y + c2 v.z ; R' v = (c0.v, c1.v, c2.v)
template <typename T> struct Rot { V3<T> c0, c1, c2; };
template <typename T> BB_HD V3<T> rot(const Rot<T>& R, const V3<T>& v) { return R.c0 * v.x + R.c1 * v.y + R.c2 * v.z; }
template <typename T> BB_HD V3<T> rotT(const Rot<T>& R, const V3<T>& v) { return mk(dot(R.c0, v), dot(R.c1, v), dot(R.c2, v)); }
template <typename T> BB_HD Rot<T> quat2rot(T w, T x, T y, T z) {
  Rot<T> R;
  R.c0 = mk(w * w + x * x - y * y - z * z, 2 * (x * y + w * z), 2 * (x * z - w * y));
  R.c1 = mk(2 * (x * y - w * z), w * w - x * x + y * y - z * z, 2 * (y * z + w * x));
  R.c2 = mk(2 * (x * z + w * y), 2 * (y * z - w * x), w * w - x * x - y * y + z * z);
  return R;
}
// symmetric 3x3 as (xx,yy,zz,xy,xz,yz)
template <typename T> struct S3 { T xx, yy, zz, xy, xz, yz; };
template <typename T> BB_HD V3<T> smul(const S3<T>& s, const V3<T>& v) {
  return mk(s.xx * v.x + s.xy * v.y + s.xz * v.z, s.xy * v.x + s.yy * v.y + s.yz * v.z, s.xz * v.x + s.yz * v.y + s.zz * v.z);
}
// point-mass inertia m(|r|^2 1 - r r')
template <typename T> BB_HD void addPointInertia(S3<T>& s, T m, const V3<T>& r) {
  T rr = dot(r, r);
  s.xx += m * (rr - r.x * r.x); s.yy += m * (rr - r.y * r.y); s.zz += m * (rr - r.z * r.z);
  s.xy -= m * r.x * r.y; s.xz -= m * r.x * r.z; s.yz -= m * r.y * r.z;
}
BB_HD int tidx(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

// ---------------------------------------------------------------------------------------------- model constants
// Derived on the host in double (bb_model.h) from the numbers of ballbot.xml, then narrowed to T.
template <typename T> struct ModelConst {
  // base group (base + cam bodies) in the base frame
  T m0, mA;        // mass of group, of the whole A tree
  T c0[3];         // group COM
  T I0c[6];        // group inertia about its COM
  T I0o[6];        // group inertia about the base origin
  // wheels, base frame, q = 0
  T ax[3][3];      // hinge axis
  T anc[3][3];     // hinge anchor
  T s0[3][3];      // wheel COM - anchor
  T u0[3][3];      // capsule axis
  T mw, It, Ia;    // wheel mass, transverse / axial inertia (capsule about its axis u)
  T armature, damping;
  T wheel_r, wheel_hl;
  // ball
  T mL, IL, dz, ball_r;
  // contact model
  T dA[NCT];       // diagApprox (sum of the two bodies' invweight0) per contact type
  T K, B;          // solref -> stiffness / damping of aref
  T solimp[5];
  T inv_width, inv_mid, inv_1mmid;   // 1 / solimp[2], 1 / solimp[3], 1 / (1 - solimp[3]) (the last two are exact powers of two for the default solimp)
  T mu[2], f1[2], f2[2], d1r[2], d2r[2];   // [0] wheel pairs, [1] hfield pair
  T dmr[2];        // 1 / (mu^2 (1 + mu^2)): cone-zone stiffness ratio Dm / D0
  T meaninertia, timestep, grav;           // gravity = (0,0,-grav)
  T tolerance, ls_tolerance;
  // heightfield
  T hx, hbase;     // half extent (5), base (0.1); z scale is per engine (ramp/gradient mutate it)
  // cameras in base frame (pos, rotation columns)
  T cam_pos[2][3], cam_rot[2][9];
  // render geoms in base frame: tower cylinder, sticks
  T tower_c[3], tower_r, tower_hl;
  T stick_c[2][3], stick_u[2][3], stick_r, stick_hl;
  // ---- integer members last: everything above is a homogeneous array of T (see narrowModel)
  int iterations, ls_iterations;
};

// per-stage geometry shared by all contact operators
template <typename T> struct Geo {
  V3<T> pB, pL;
  Rot<T> RB, RL;
  V3<T> aw[3], hw[3];   // world hinge axes / anchors
};

template <typename T> struct Scratch {
  T M[NTRI], H[NTRI];
  T qfs[NV], qas[NV], Ma[NV], grad[NV], search[NV], Mv[NV], Mgrad[NV];
  int nc;
  unsigned char ctype[NC], cstate[NC];
  T cP[NC][3], cF[NC][9], cD[NC], cAref[NC][3], cJar[NC][3], cJv[NC][3], cFrc[NC][3], cHc[NC][6];
  T cDist[NC];
};

// kinematic quantities the observation needs (ballbot_env.py:772-800), taken at the evaluated stage state
template <typename T> struct KinOut { T quatB[4]; T cvel_ang[3], cvel_lin[3]; T posB[3]; int ncon, niter; };

// ---------------------------------------------------------------------------------------------- packed Cholesky
template <typename T> BB_HD void cholPacked(T* A) {
  for (int j = 0; j < NV; j++) {
    T s = A[tidx(j, j)];
    for (int k = 0; k < j; k++) { T l = A[j * (j + 1) / 2 + k]; s -= l * l; }
    if (s < (T)1e-15) s = (T)1e-15;
    s = bsqrt(s);
    A[j * (j + 1) / 2 + j] = s;
    T inv = (T)1 / s;
    for (int i = j + 1; i < NV; i++) {
      T t = A[i * (i + 1) / 2 + j];
      for (int k = 0; k < j; k++) t -= A[i * (i + 1) / 2 + k] * A[j * (j + 1) / 2 + k];
      A[i * (i + 1) / 2 + j] = t * inv;
    }
  }
}
template <typename T> BB_HD void cholSolvePacked(const T* L, const T* b, T* x) {
  for (int i = 0; i < NV; i++) {
    T s = b[i];
    for (int k = 0; k < i; k++) s -= L[i * (i + 1) / 2 + k] * x[k];
    x[i] = s / L[i * (i + 1) / 2 + i];
  }
  for (int i = NV - 1; i >= 0; i--) {
    T s = x[i];
    for (int k = i + 1; k < NV; k++) s -= L[k * (k + 1) / 2 + i] * x[k];
    x[i] = s / L[i * (i + 1) / 2 + i];
  }
}
template <typename T> BB_HD void symvPacked(const T* A, const T* v, T* r) {
  for (int i = 0; i < NV; i++) r[i] = 0;
  for (int i = 0; i < NV; i++) {
    T acc = 0, vi = v[i];
    for (int j = 0; j < i; j++) { T a = A[i * (i + 1) / 2 + j]; acc += a * v[j]; r[j] += a * vi; }
    r[i] += acc + A[i * (i + 1) / 2 + i] * vi;
  }
}

// ---------------------------------------------------------------------------------------------- smooth dynamics
// Rodrigues rotation of v about unit axis a
template <typename T> BB_HD V3<T> rodrigues(const V3<T>& a, const V3<T>& v, T s, T c) {
  return v * c + cross(a, v) * s + a * (dot(a, v) * ((T)1 - c));
}

template <typename T> BB_HD void normalizeQuat4(T* q) {
  const T n = bsqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < (T)1e-15) { q[0] = (T)1; q[1] = (T)0; q[2] = (T)0; q[3] = (T)0; }
  else { const T inv = (T)1 / n; q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv; }
}
template <typename T> BB_HD void normalizeQuats(T* qpos) { normalizeQuat4(qpos + 3); normalizeQuat4(qpos + 13); }

// Fills s.M (packed), s.qfs (= passive - bias + actuator), geometry g, wheel capsule world frames, obs kinematics.
// M (packed lower triangle, NTRI) and qfs (NV) are raw pointers so that the same code serves the thread-per-env scratch
// and the warp-per-env shared-memory layout (ZERO_M=false: the caller has zeroed M and normalised the quaternions).
// ML selects the layout of M: 0 = packed lower triangle (tidx), 1 = the two diagonal blocks of the block-diagonal mass
// matrix, both triangles written: base+wheels 9 x 9 at M[i * 9 + j], ball 6 x 6 at M[81 + (i - 9) * 6 + (j - 9)] (the
// group kernel zeroes the never-written entries once per launch).
template <int ML, typename T> BB_HD void setM(T* M, int i, int j, T v) {
  if (ML == 0) M[tidx(i, j)] = v;
  else if (i < 9) { M[i * 9 + j] = v; M[j * 9 + i] = v; }
  else { M[81 + (i - 9) * 6 + (j - 9)] = v; M[81 + (j - 9) * 6 + (i - 9)] = v; }
}
template <typename T, bool ZERO_M = true, int ML = 0>
BB_HD void smoothDynamics(const ModelConst<T>& mc, T* qpos, const T* qvel, const T* ctrl, T* M, T* qfs, Geo<T>& g,
                          V3<T>* capC, V3<T>* capU, KinOut<T>* kin) {
  if (ZERO_M) normalizeQuats(qpos);   // mj_kinematics normalises free-joint quaternions in place
  g.pB = ld3(qpos); g.pL = ld3(qpos + 10);
  g.RB = quat2rot(qpos[3], qpos[4], qpos[5], qpos[6]);
  g.RL = quat2rot(qpos[13], qpos[14], qpos[15], qpos[16]);
  const V3<T> w = ld3(qvel + 3);                    // base angular velocity, base frame
  const V3<T> gl = rotT(g.RB, mk((T)0, (T)0, -mc.grav));  // gravity in base frame

  // ---- wheels (base frame)
  V3<T> a[3], si[3], ri[3], ui[3];
  S3<T> IA; IA.xx = mc.I0o[0]; IA.yy = mc.I0o[1]; IA.zz = mc.I0o[2]; IA.xy = mc.I0o[3]; IA.xz = mc.I0o[4]; IA.yz = mc.I0o[5];
  V3<T> mcA = ld3(mc.c0) * mc.m0;                    // first moment of the whole tree about the base origin
  const V3<T> c0 = ld3(mc.c0);
  S3<T> I0c; I0c.xx = mc.I0c[0]; I0c.yy = mc.I0c[1]; I0c.zz = mc.I0c[2]; I0c.xy = mc.I0c[3]; I0c.xz = mc.I0c[4]; I0c.yz = mc.I0c[5];
  // group bias wrench
  V3<T> F = (cross(w, cross(w, c0)) - gl) * mc.m0;
  V3<T> Fsum = F;
  V3<T> Tsum = cross(c0, F) + cross(w, smul(I0c, w));
  if (ZERO_M) { for (int i = 0; i < NTRI; i++) M[i] = 0; }
  T biasq[3];
  V3<T> pw[3], Lw[3];
  for (int i = 0; i < 3; i++) {
    T sq, cq; bsincos(qpos[7 + i], &sq, &cq);
    a[i] = ld3(mc.ax[i]);
    const V3<T> h = ld3(mc.anc[i]);
    si[i] = rodrigues(a[i], ld3(mc.s0[i]), sq, cq);
    ui[i] = rodrigues(a[i], ld3(mc.u0[i]), sq, cq);
    ri[i] = h + si[i];
    // wheel inertia about its COM: It*1 + (Ia-It) u u'
    const T dI = mc.Ia - mc.It;
    S3<T> Iw; Iw.xx = mc.It + dI * ui[i].x * ui[i].x; Iw.yy = mc.It + dI * ui[i].y * ui[i].y; Iw.zz = mc.It + dI * ui[i].z * ui[i].z;
    Iw.xy = dI * ui[i].x * ui[i].y; Iw.xz = dI * ui[i].x * ui[i].z; Iw.yz = dI * ui[i].y * ui[i].z;
    IA.xx += Iw.xx; IA.yy += Iw.yy; IA.zz += Iw.zz; IA.xy += Iw.xy; IA.xz += Iw.xz; IA.yz += Iw.yz;
    addPointInertia(IA, mc.mw, ri[i]);
    mcA = mcA + ri[i] * mc.mw;
    // mass-matrix column of the hinge
    const V3<T> axs = cross(a[i], si[i]);            // COM velocity per unit hinge rate
    pw[i] = axs * mc.mw;
    const V3<T> Iwa = smul(Iw, a[i]);
    Lw[i] = Iwa + cross(ri[i], pw[i]);
    setM<ML>(M, 6 + i, 6 + i, dot(a[i], Iwa) + mc.mw * dot(axs, axs) + mc.armature);
    // bias (qdd = 0): classical accelerations in the rotating base frame
    const T qd = qvel[6 + i];
    const V3<T> wi = w + a[i] * qd;
    const V3<T> al = cross(w, a[i]) * qd;
    const V3<T> acc = cross(w, cross(w, h)) + cross(al, si[i]) + cross(wi, cross(wi, si[i]));
    const V3<T> Fi = (acc - gl) * mc.mw;
    const V3<T> Ni = smul(Iw, al) + cross(wi, smul(Iw, wi));
    Fsum = Fsum + Fi;
    Tsum = Tsum + cross(ri[i], Fi) + Ni;
    biasq[i] = dot(a[i], cross(si[i], Fi) + Ni);
  }
  // ---- M_A
  setM<ML>(M, 0, 0, mc.mA); setM<ML>(M, 1, 1, mc.mA); setM<ML>(M, 2, 2, mc.mA);
  {  // M[v, w_k] = d(momentum)/d(w_k) = R_B (e_k x mcA)
    const V3<T> k0 = rot(g.RB, cross(mk((T)1, (T)0, (T)0), mcA));
    const V3<T> k1 = rot(g.RB, cross(mk((T)0, (T)1, (T)0), mcA));
    const V3<T> k2 = rot(g.RB, cross(mk((T)0, (T)0, (T)1), mcA));
    setM<ML>(M, 3, 0, k0.x); setM<ML>(M, 3, 1, k0.y); setM<ML>(M, 3, 2, k0.z);
    setM<ML>(M, 4, 0, k1.x); setM<ML>(M, 4, 1, k1.y); setM<ML>(M, 4, 2, k1.z);
    setM<ML>(M, 5, 0, k2.x); setM<ML>(M, 5, 1, k2.y); setM<ML>(M, 5, 2, k2.z);
  }
  setM<ML>(M, 3, 3, IA.xx); setM<ML>(M, 4, 4, IA.yy); setM<ML>(M, 5, 5, IA.zz);
  setM<ML>(M, 4, 3, IA.xy); setM<ML>(M, 5, 3, IA.xz); setM<ML>(M, 5, 4, IA.yz);
  for (int i = 0; i < 3; i++) {
    const V3<T> pwW = rot(g.RB, pw[i]);
    setM<ML>(M, 6 + i, 0, pwW.x); setM<ML>(M, 6 + i, 1, pwW.y); setM<ML>(M, 6 + i, 2, pwW.z);
    setM<ML>(M, 6 + i, 3, Lw[i].x); setM<ML>(M, 6 + i, 4, Lw[i].y); setM<ML>(M, 6 + i, 5, Lw[i].z);
  }
  // ---- ball
  const V3<T> wl = ld3(qvel + 12);
  const V3<T> d = mk((T)0, (T)0, mc.dz);
  const V3<T> glL = rotT(g.RL, mk((T)0, (T)0, -mc.grav));
  const V3<T> FL = (cross(wl, cross(wl, d)) - glL) * mc.mL;
  const V3<T> FLw = rot(g.RL, FL);
  const V3<T> TL = cross(d, FL);
  setM<ML>(M, 9, 9, mc.mL); setM<ML>(M, 10, 10, mc.mL); setM<ML>(M, 11, 11, mc.mL);
  {
    const V3<T> md = d * mc.mL;
    const V3<T> k0 = rot(g.RL, cross(mk((T)1, (T)0, (T)0), md));
    const V3<T> k1 = rot(g.RL, cross(mk((T)0, (T)1, (T)0), md));
    const V3<T> k2 = rot(g.RL, cross(mk((T)0, (T)0, (T)1), md));
    setM<ML>(M, 12, 9, k0.x); setM<ML>(M, 12, 10, k0.y); setM<ML>(M, 12, 11, k0.z);
    setM<ML>(M, 13, 9, k1.x); setM<ML>(M, 13, 10, k1.y); setM<ML>(M, 13, 11, k1.z);
    setM<ML>(M, 14, 9, k2.x); setM<ML>(M, 14, 10, k2.y); setM<ML>(M, 14, 11, k2.z);
  }
  setM<ML>(M, 12, 12, mc.IL + mc.mL * mc.dz * mc.dz); setM<ML>(M, 13, 13, mc.IL + mc.mL * mc.dz * mc.dz); setM<ML>(M, 14, 14, mc.IL);
  // ---- qfrc_smooth = passive - bias + actuator
  const V3<T> FsW = rot(g.RB, Fsum);
  qfs[0] = -FsW.x; qfs[1] = -FsW.y; qfs[2] = -FsW.z;
  qfs[3] = -Tsum.x; qfs[4] = -Tsum.y; qfs[5] = -Tsum.z;
  for (int i = 0; i < 3; i++) {
    T u = ctrl[i]; u = u > (T)10 ? (T)10 : (u < (T)-10 ? (T)-10 : u);   // ctrlrange, ballbot.xml:84-86
    qfs[6 + i] = -mc.damping * qvel[6 + i] - biasq[i] + u;
  }
  qfs[9] = -FLw.x; qfs[10] = -FLw.y; qfs[11] = -FLw.z;
  qfs[12] = -TL.x; qfs[13] = -TL.y; qfs[14] = -TL.z;
  // ---- world geometry for contacts
  for (int i = 0; i < 3; i++) {
    g.aw[i] = rot(g.RB, a[i]);
    g.hw[i] = g.pB + rot(g.RB, ld3(mc.anc[i]));
    capC[i] = g.pB + rot(g.RB, ri[i]);
    capU[i] = rot(g.RB, ui[i]);
  }
  if (kin) {
    kin->quatB[0] = qpos[3]; kin->quatB[1] = qpos[4]; kin->quatB[2] = qpos[5]; kin->quatB[3] = qpos[6];
    const V3<T> wW = rot(g.RB, w);
    const V3<T> comOff = rot(g.RB, mcA * ((T)1 / mc.mA));
    const V3<T> vl = ld3(qvel) + cross(wW, comOff);   // MuJoCo cvel linear part: velocity at the subtree COM
    st3(kin->cvel_ang, wW); st3(kin->cvel_lin, vl); st3(kin->posB, g.pB);
  }
}

// ---------------------------------------------------------------------------------------------- collision
template <typename T> BB_HD void makeFrame(T* f, bool haveT1) {
  // f[0..2] normal (unit), f[3..5] optional tangent hint   (mju_makeFrame)
  V3<T> n = ld3(f), t = ld3(f + 3);
  const T ny = n.y;
  const bool midY = (ny < (T)0.5) && (ny > (T)-0.5);
  if (!haveT1 || dot(t, t) < (T)0.25) { t = midY ? mk((T)0, (T)1, (T)0) : mk((T)0, (T)0, (T)1); }
  t = t - n * dot(n, t);
  t = t * ((T)1 / bsqrt(dot(t, t)));
  st3(f + 3, t); st3(f + 6, cross(n, t));
}
template <typename T> BB_HD V3<T> closestOnTriangle(const V3<T>& p, const V3<T>& a, const V3<T>& b, const V3<T>& c) {
  const V3<T> ab = b - a, ac = c - a, ap = p - a;
  const T d1 = dot(ab, ap), d2 = dot(ac, ap);
  if (d1 <= 0 && d2 <= 0) return a;
  const V3<T> bp = p - b; const T d3 = dot(ab, bp), d4 = dot(ac, bp);
  if (d3 >= 0 && d4 <= d3) return b;
  const T vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) return a + ab * (d1 / (d1 - d3));
  const V3<T> cp = p - c; const T d5 = dot(ab, cp), d6 = dot(ac, cp);
  if (d6 >= 0 && d5 <= d6) return c;
  const T vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) return a + ac * (d2 / (d2 - d6));
  const T va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) return b + (c - b) * ((d4 - d3) / ((d4 - d3) + (d5 - d6)));
  const T den = (T)1 / (va + vb + vc);
  return a + ab * (vb * den) + ac * (vc * den);
}

// closest points of the segments p1-q1 and p2-q2 (returns the squared distance)
template <typename T> BB_HD T closestSegSeg(const V3<T>& p1, const V3<T>& q1, const V3<T>& p2, const V3<T>& q2, V3<T>& c1, V3<T>& c2) {
  const V3<T> d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
  const T a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r), tiny = (T)1e-15;
  T s, t;
  if (a <= tiny && e <= tiny) { s = 0; t = 0; }
  else if (a <= tiny) { s = 0; t = f / e; t = t < 0 ? (T)0 : (t > 1 ? (T)1 : t); }
  else {
    const T c = dot(d1, r);
    if (e <= tiny) { t = 0; s = -c / a; s = s < 0 ? (T)0 : (s > 1 ? (T)1 : s); }
    else {
      const T b = dot(d1, d2), den = a * e - b * b;
      s = den > tiny ? (b * f - c * e) / den : (T)0; s = s < 0 ? (T)0 : (s > 1 ? (T)1 : s);
      t = (b * s + f) / e;
      if (t < 0) { t = 0; s = -c / a; s = s < 0 ? (T)0 : (s > 1 ? (T)1 : s); }
      else if (t > 1) { t = 1; s = (b - c) / a; s = s < 0 ? (T)0 : (s > 1 ? (T)1 : s); }
    }
  }
  c1 = p1 + d1 * s; c2 = p2 + d2 * t;
  const V3<T> dv = c1 - c2;
  return dot(dv, dv);
}
// closest points between the segment p0-p1 (ps) and the triangle abc (pt): segment end points against the triangle, then the
// segment against the three edges; the first minimum in that order wins
template <typename T>
BB_HD void closestSegTriangle(const V3<T>& p0, const V3<T>& p1, const V3<T>& a, const V3<T>& b, const V3<T>& c, V3<T>& ps, V3<T>& pt) {
  T best = (T)1e30;
#pragma unroll 1
  for (int i = 0; i < 2; i++) {
    const V3<T> e = i ? p1 : p0;
    const V3<T> q = closestOnTriangle(e, a, b, c);
    const V3<T> dv = e - q; const T d2 = dot(dv, dv);
    if (d2 < best) { best = d2; ps = e; pt = q; }
  }
#pragma unroll 1
  for (int i = 0; i < 3; i++) {
    const V3<T> e0 = i == 0 ? a : (i == 1 ? b : c), e1 = i == 0 ? b : (i == 1 ? c : a);
    V3<T> s1, s2;
    const T d2 = closestSegSeg(p0, p1, e0, e1, s1, s2);
    if (d2 < best) { best = d2; ps = s1; pt = s2; }
  }
}
// capsule (segment p0-p1, radius r) against the prism under the top triangle (ta, tb, tc): closest feature of the top surface
// patch; a segment point under the top plane only counts inside the prism's column (same conventions as the ball pair)
// exact early-out of capsulePrism: both segment ends are more than r above the top plane => every segment point is, and the
// distance to the triangle is at least the distance to its plane (no contact; the deep branch needs a point under the plane)
template <typename T>
BB_HD bool capsuleAbovePlane(const V3<T>& p0, const V3<T>& p1, T r, const V3<T>& ta, const V3<T>& tb, const V3<T>& tc) {
  const V3<T> nc = cross(tb - ta, tc - ta);
  const T sg = nc.z < 0 ? (T)-1 : (T)1;
  const T h0 = sg * dot(p0 - ta, nc), h1 = sg * dot(p1 - ta, nc), rn = r * r * dot(nc, nc);     // heights scaled by |nc|
  return h0 > 0 && h1 > 0 && h0 * h0 > rn && h1 * h1 > rn;
}
template <typename T>
BB_HD bool capsulePrism(const V3<T>& p0, const V3<T>& p1, T r, const V3<T>& ta, const V3<T>& tb, const V3<T>& tc, T& dist, V3<T>& n, V3<T>& pos) {
  const V3<T> e1 = tb - ta, e2 = tc - ta;
  const V3<T> nc = cross(e1, e2);
  const V3<T> nn = nc * ((nc.z < 0 ? (T)-1 : (T)1) / bsqrt(dot(nc, nc)));   // upward unit normal of the top plane
  V3<T> ps = p0, pt = ta;
  closestSegTriangle(p0, p1, ta, tb, tc, ps, pt);
  const T h0 = dot(p0 - ta, nn), h1 = dot(p1 - ta, nn);
  const V3<T> pl = h0 <= h1 ? p0 : p1; const T hl = h0 <= h1 ? h0 : h1;
  if (hl < 0) {
    const V3<T> ap = pl - ta;
    const T u = e1.x * e2.y - e1.y * e2.x;
    const T sa = (ap.x * e2.y - ap.y * e2.x) / u, tt = (e1.x * ap.y - e1.y * ap.x) / u;
    if (!(sa < 0 || tt < 0 || sa + tt > 1)) { dist = hl - r; n = nn; pos = pl - nn * (r + (T)0.5 * dist); return true; }
    if (dot(ps - ta, nn) < 0) return false;
  }
  const V3<T> dv = ps - pt; const T dl = bsqrt(dot(dv, dv));
  if (dl >= r || dl < (T)1e-15) return false;
  dist = dl - r; n = dv * ((T)1 / dl); pos = pt + n * ((T)0.5 * dist);
  return true;
}
template <typename T>
BB_NOINL bool capsulePrismNoInline(const V3<T>& p0, const V3<T>& p1, T r, const V3<T>& ta, const V3<T>& tb, const V3<T>& tc, T& dist, V3<T>& n, V3<T>& pos) {
  return capsulePrism(p0, p1, r, ta, tb, tc, dist, n, pos);
}
// heightfield sub-grid of an axis-aligned box [lo, hi] (mjc_ConvexHField): false = the box misses the field
template <typename T> struct HfGrid { int cmin, cmax, rmin, rmax; T dx, zmin; };
template <typename T> BB_HD bool hfieldSubgrid(T sx, T hbase, T zscale, const V3<T>& lo, const V3<T>& hi, HfGrid<T>& gr) {
  if ((sx < lo.x) || (-sx > hi.x) || (sx < lo.y) || (-sx > hi.y) || (zscale < lo.z) || (-hbase > hi.z)) return false;
  const T gs = (T)(HN - 1) / ((T)2 * sx);
  int cmin = (int)bfloor((lo.x + sx) * gs), cmax = (int)bceil((hi.x + sx) * gs);
  int rmin = (int)bfloor((lo.y + sx) * gs), rmax = (int)bceil((hi.y + sx) * gs);
  gr.cmin = cmin < 0 ? 0 : cmin; gr.rmin = rmin < 0 ? 0 : rmin; gr.cmax = cmax > HN - 1 ? HN - 1 : cmax; gr.rmax = rmax > HN - 1 ? HN - 1 : rmax;
  gr.dx = (T)2 * sx / (T)(HN - 1); gr.zmin = lo.z;
  return true;
}
// top triangle k (0 / 1) of heightfield cell (r, c) in the triangle-strip order of mjc_ConvexHField
template <typename T> BB_HD void hfieldTriangle(const float* hf, T zscale, T sx, T dx, int r, int c, int k, V3<T>& ta, V3<T>& tb, V3<T>& tc) {
  const T x0 = dx * (T)c - sx, x1 = dx * (T)(c + 1) - sx, y0 = dx * (T)r - sx, y1 = dx * (T)(r + 1) - sx;
  const V3<T> v00 = mk(x0, y0, (T)hf[r * HN + c] * zscale), v11 = mk(x1, y1, (T)hf[(r + 1) * HN + c + 1] * zscale);
  if (k == 0) { ta = mk(x0, y1, (T)hf[(r + 1) * HN + c] * zscale); tb = v00; tc = v11; }
  else { ta = v00; tb = v11; tc = mk(x1, y0, (T)hf[r * HN + c + 1] * zscale); }
}
// sphere (centre bc, radius br) against a capsule (centre cc, axis cu, radius r, half length hl): nearest point on the segment,
// then sphere-sphere (mjraw_SphereCapsule); normal points from the sphere to the capsule
template <typename T> BB_HD bool sphereCapsule(const V3<T>& bc, T br, const V3<T>& cc, const V3<T>& cu, T r, T hl, T& dist, V3<T>& n, V3<T>& pos) {
  T x = dot(cu, bc - cc);
  x = x > hl ? hl : (x < -hl ? -hl : x);
  const V3<T> dif = cc + cu * x - bc;
  const T cd = bsqrt(dot(dif, dif)), mind = br + r;
  if (cd >= mind) return false;
  n = dif * ((T)1 / cd); dist = cd - mind; pos = bc + n * (br + (T)0.5 * dist);
  return true;
}
// sphere against a cylinder (mjraw_SphereCylinder: side / cap / corner); normal points from the sphere to the cylinder
template <typename T> BB_HD bool sphereCylinder(const V3<T>& bc, T br, const V3<T>& cc, const V3<T>& ax, T rad, T hgt, T& dist, V3<T>& n, V3<T>& pos) {
  const V3<T> vec = bc - cc;
  const T x = dot(vec, ax);
  const V3<T> pp = vec - ax * x; const T pp2 = dot(pp, pp);
  bool side = babs(x) < hgt, cap = pp2 < rad * rad;
  if (side && cap) { if (hgt - babs(x) < rad - bsqrt(pp2)) side = false; else cap = false; }
  const T sg = x > 0 ? (T)1 : (T)-1;
  V3<T> tgt; T trad;
  if (side) { tgt = cc + ax * x; trad = rad; }
  else if (cap) {
    const V3<T> pc = cc + ax * (sg * hgt);
    const T dd = sg * dot(bc - pc, ax) - br;
    if (!(dd < 0)) return false;
    dist = dd; n = ax * (-sg); pos = bc + n * (br + (T)0.5 * dist);
    return true;
  } else { tgt = cc + pp * (rad / bsqrt(pp2)) + ax * (sg * hgt); trad = 0; }
  const V3<T> dif = tgt - bc; const T cd = bsqrt(dot(dif, dif));
  if (!(cd < br + trad) || !(cd > (T)1e-15)) return false;
  dist = cd - br - trad; n = dif * ((T)1 / cd); pos = bc + n * (br + (T)0.5 * dist);
  return true;
}

// Contact generation: 3 patched sphere-capsule pairs (tools/mujoco_fix.patch:9-18) + ball vs heightfield prisms.
// Only penetrating contacts (dist < 0) are recorded since margin = gap = 0.
template <typename T>
BB_HD void collide(const ModelConst<T>& mc, const Geo<T>& g, const V3<T>* capC, const V3<T>* capU, const float* hf, T zscale,
                   Scratch<T>& s) {
  int nc = 0;
  const V3<T> bc = g.pL + rot(g.RL, mk((T)0, (T)0, mc.dz));
  const T br = mc.ball_r;
  for (int i = 0; i < 3; i++) {
    T x = dot(capU[i], bc - capC[i]);
    x = x > mc.wheel_hl ? mc.wheel_hl : (x < -mc.wheel_hl ? -mc.wheel_hl : x);
    const V3<T> np = capC[i] + capU[i] * x;
    const V3<T> dif = np - bc;
    const T cd = bsqrt(dot(dif, dif)), mind = br + mc.wheel_r;
    if (cd >= mind) continue;
    const V3<T> n = dif * ((T)1 / cd);
    const T dist = cd - mind;
    st3(s.cF[nc], n); st3(s.cF[nc] + 3, capU[i]);
    makeFrame(s.cF[nc], true);
    st3(s.cP[nc], bc + n * (br + (T)0.5 * dist));
    s.cDist[nc] = dist; s.ctype[nc] = (unsigned char)i;
    nc++;
  }
  // heightfield: sub-grid of the ball's AABB, two prisms per cell, scan order of mjc_ConvexHField
  {
    const T sx = mc.hx;
    bool skip = (sx < bc.x - br) || (-sx > bc.x + br) || (sx < bc.y - br) || (-sx > bc.y + br) || (zscale < bc.z - br) ||
                (-mc.hbase > bc.z + br);
    if (!skip) {
      const T gs = (T)(HN - 1) / ((T)2 * sx);
      int cmin = (int)bfloor((bc.x - br + sx) * gs), cmax = (int)bceil((bc.x + br + sx) * gs);
      int rmin = (int)bfloor((bc.y - br + sx) * gs), rmax = (int)bceil((bc.y + br + sx) * gs);
      cmin = cmin < 0 ? 0 : cmin; rmin = rmin < 0 ? 0 : rmin; cmax = cmax > HN - 1 ? HN - 1 : cmax; rmax = rmax > HN - 1 ? HN - 1 : rmax;
      const T dx = (T)2 * sx / (T)(HN - 1);
      const T zmin = bc.z - br;
      int cnt = 0;
      for (int r = rmin; r < rmax && cnt < MAXH; r++) {
        const T y0 = dx * (T)r - sx, y1 = dx * (T)(r + 1) - sx;
        // row-level reject: closest y of the strip to the centre
        const T ey = bc.y < y0 ? y0 - bc.y : (bc.y > y1 ? bc.y - y1 : (T)0);
        if (ey >= br) continue;
        T hlo = (T)hf[r * HN + cmin] * zscale, hhi = (T)hf[(r + 1) * HN + cmin] * zscale;
        for (int c = cmin; c < cmax && cnt < MAXH; c++) {
          const T x0 = dx * (T)c - sx, x1 = dx * (T)(c + 1) - sx;
          const T hlo1 = (T)hf[r * HN + c + 1] * zscale, hhi1 = (T)hf[(r + 1) * HN + c + 1] * zscale;
          const T ex = bc.x < x0 ? x0 - bc.x : (bc.x > x1 ? bc.x - x1 : (T)0);
          if (ex * ex + ey * ey < br * br) {
            const V3<T> v01 = mk(x0, y1, hhi), v00 = mk(x0, y0, hlo), v11 = mk(x1, y1, hhi1), v10 = mk(x1, y0, hlo1);
            for (int k = 0; k < 2 && cnt < MAXH; k++) {
              // triangle strip order: (c,r+1),(c,r),(c+1,r+1) then (c,r),(c+1,r+1),(c+1,r)
              const V3<T> ta = k ? v00 : v01, tb = k ? v11 : v00, tc = k ? v10 : v11;
              if (ta.z < zmin && tb.z < zmin && tc.z < zmin) continue;
              const V3<T> q = closestOnTriangle(bc, ta, tb, tc);
              const V3<T> dv = bc - q;
              V3<T> nn = cross(tb - ta, tc - ta); if (nn.z < 0) nn = -nn;
              const T h = dot(bc - ta, nn);
              T dist; V3<T> n, pos;
              if (h < 0) {  // centre under the top plane: contact only inside this prism's column
                const V3<T> e1 = tb - ta, e2 = tc - ta, ap = bc - ta;
                const T u = e1.x * e2.y - e1.y * e2.x;
                const T sa = (ap.x * e2.y - ap.y * e2.x) / u, tt = (e1.x * ap.y - e1.y * ap.x) / u;
                if (sa < 0 || tt < 0 || sa + tt > 1) continue;
                n = nn * ((T)1 / bsqrt(dot(nn, nn)));
                dist = dot(ap, n) - br;
                pos = bc - n * (br + (T)0.5 * dist);
              } else {
                const T dl = bsqrt(dot(dv, dv));
                if (dl >= br || dl < (T)1e-15) continue;
                dist = dl - br; n = dv * ((T)1 / dl); pos = q + n * ((T)0.5 * dist);
              }
              if (nc < NC) {
                st3(s.cF[nc], n); makeFrame(s.cF[nc], false);
                st3(s.cP[nc], pos); s.cDist[nc] = dist; s.ctype[nc] = 3;
                nc++;
              }
              cnt++;
            }
          }
          hlo = hlo1; hhi = hhi1;
        }
      }
    }
  }
  // ---- the other colliding geoms (ballbot.xml:41-69): camera sticks and wheel capsules against the heightfield, then the ball
  // against the sticks (patched sphere-capsule) and the tower cylinder.  Order = the oracle's.
#pragma unroll
  for (int gi = 0; gi < 5; gi++) {   // unrolled: capC / capU stay statically indexed
    V3<T> cc, cu; T rad, hl; int ty;
    if (gi < 2) { cc = g.pB + rot(g.RB, ld3(mc.stick_c[gi])); cu = rot(g.RB, ld3(mc.stick_u[gi])); rad = mc.stick_r; hl = mc.stick_hl; ty = 4 + gi; }
    else { cc = capC[gi - 2]; cu = capU[gi - 2]; rad = mc.wheel_r; hl = mc.wheel_hl; ty = 6 + gi - 2; }
    const V3<T> p0 = cc - cu * hl, p1 = cc + cu * hl;
    const V3<T> lo = mk(bmin(p0.x, p1.x) - rad, bmin(p0.y, p1.y) - rad, bmin(p0.z, p1.z) - rad);
    const V3<T> hi = mk(bmax(p0.x, p1.x) + rad, bmax(p0.y, p1.y) + rad, bmax(p0.z, p1.z) + rad);
    HfGrid<T> gr;
    if (!hfieldSubgrid(mc.hx, mc.hbase, zscale, lo, hi, gr)) continue;
    int cnt = 0;
    for (int r = gr.rmin; r < gr.rmax && cnt < MAXH; r++)
      for (int c = gr.cmin; c < gr.cmax && cnt < MAXH; c++)
        for (int k = 0; k < 2 && cnt < MAXH; k++) {
          V3<T> ta, tb, tc; hfieldTriangle(hf, zscale, mc.hx, gr.dx, r, c, k, ta, tb, tc);
          if (ta.z < gr.zmin && tb.z < gr.zmin && tc.z < gr.zmin) continue;
          if (capsuleAbovePlane(p0, p1, rad, ta, tb, tc)) continue;
          T dist; V3<T> n, pos;
          if (!capsulePrismNoInline(p0, p1, rad, ta, tb, tc, dist, n, pos)) continue;
          if (nc < NC) { st3(s.cF[nc], n); makeFrame(s.cF[nc], false); st3(s.cP[nc], pos); s.cDist[nc] = dist; s.ctype[nc] = (unsigned char)ty; nc++; }
          cnt++;
        }
  }
  for (int i = 0; i < 2; i++) {
    const V3<T> cc = g.pB + rot(g.RB, ld3(mc.stick_c[i])), cu = rot(g.RB, ld3(mc.stick_u[i]));
    T dist; V3<T> n, pos;
    if (sphereCapsule(bc, br, cc, cu, mc.stick_r, mc.stick_hl, dist, n, pos) && nc < NC) {
      st3(s.cF[nc], n); st3(s.cF[nc] + 3, cu); makeFrame(s.cF[nc], true);
      st3(s.cP[nc], pos); s.cDist[nc] = dist; s.ctype[nc] = (unsigned char)(10 + i); nc++;
    }
  }
  {
    T dist; V3<T> n, pos;
    if (sphereCylinder(bc, br, g.pB + rot(g.RB, ld3(mc.tower_c)), g.RB.c2, mc.tower_r, mc.tower_hl, dist, n, pos) && nc < NC) {
      st3(s.cF[nc], n); makeFrame(s.cF[nc], false); st3(s.cP[nc], pos); s.cDist[nc] = dist; s.ctype[nc] = 9; nc++;
    }
  }
  s.nc = nc;
}

// ---------------------------------------------------------------------------------------------- matrix-free contact operators
// relative velocity (body2 - body1) of contact c for generalized vector v, expressed in the contact frame
template <typename T> struct VelCtx { V3<T> vB, wB, vL, wL; const T* v; };
template <typename T> BB_HD VelCtx<T> velCtx(const Geo<T>& g, const T* v) {
  VelCtx<T> c; c.vB = ld3(v); c.wB = rot(g.RB, ld3(v + 3)); c.vL = ld3(v + 9); c.wL = rot(g.RL, ld3(v + 12)); c.v = v; return c;
}
template <typename T> BB_HD void contactVel(const Geo<T>& g, const Scratch<T>& s, int c, const VelCtx<T>& vc, T* out) {
  const V3<T> P = ld3(s.cP[c]);
  // relative velocity body2 - body1: the ball is body 2 against the heightfield and body 1 against every robot geom
  const int ty = s.ctype[c], wi = ctWheel(ty);
  V3<T> rel = mk((T)0, (T)0, (T)0);
  if (ctHasBase(ty)) {
    rel = vc.vB + cross(vc.wB, P - g.pB);
    if (wi >= 0) rel = rel + cross(g.aw[wi], P - g.hw[wi]) * vc.v[6 + wi];
  }
  if (ctHasBall(ty)) {
    const V3<T> vball = vc.vL + cross(vc.wL, P - g.pL);
    rel = ty == 3 ? vball : rel - vball;
  }
  const T* f = s.cF[c];
  out[0] = f[0] * rel.x + f[1] * rel.y + f[2] * rel.z;
  out[1] = f[3] * rel.x + f[4] * rel.y + f[5] * rel.z;
  out[2] = f[6] * rel.x + f[7] * rel.y + f[8] * rel.z;
}
// r -= J' f  (accumulated over all contacts)
template <typename T> BB_HD void subJtF(const Geo<T>& g, const Scratch<T>& s, T* r) {
  V3<T> FB = mk((T)0, (T)0, (T)0), TB = FB, FLs = FB, TLs = FB; T tq[3] = {0, 0, 0};
  for (int c = 0; c < s.nc; c++) {
    const T* f = s.cF[c]; const T* fc = s.cFrc[c];
    if (fc[0] == 0 && fc[1] == 0 && fc[2] == 0) continue;
    const V3<T> Fw = mk(f[0] * fc[0] + f[3] * fc[1] + f[6] * fc[2], f[1] * fc[0] + f[4] * fc[1] + f[7] * fc[2],
                        f[2] * fc[0] + f[5] * fc[1] + f[8] * fc[2]);
    const V3<T> P = ld3(s.cP[c]);
    const int ty = s.ctype[c], wi = ctWheel(ty);
    if (ctHasBase(ty)) {
      FB = FB + Fw; TB = TB + cross(P - g.pB, Fw);
      if (wi >= 0) tq[wi] += dot(g.aw[wi], cross(P - g.hw[wi], Fw));
    }
    if (ctHasBall(ty)) {
      if (ty == 3) { FLs = FLs + Fw; TLs = TLs + cross(P - g.pL, Fw); }
      else { FLs = FLs - Fw; TLs = TLs - cross(P - g.pL, Fw); }
    }
  }
  const V3<T> tb = rotT(g.RB, TB), tl = rotT(g.RL, TLs);
  r[0] -= FB.x; r[1] -= FB.y; r[2] -= FB.z; r[3] -= tb.x; r[4] -= tb.y; r[5] -= tb.z;
  r[6] -= tq[0]; r[7] -= tq[1]; r[8] -= tq[2];
  r[9] -= FLs.x; r[10] -= FLs.y; r[11] -= FLs.z; r[12] -= tl.x; r[13] -= tl.y; r[14] -= tl.z;
}

// elliptic-cone zone logic of mj_constraintUpdate for one contact; returns cost, fills force/state/(cone Hessian)
template <typename T>
BB_HD T coneUpdate(const ModelConst<T>& mc, Scratch<T>& s, int c, const T* jar, bool wantH) {
  const int k = ctFric(s.ctype[c]);
  const T mu = mc.mu[k], f1 = mc.f1[k], f2 = mc.f2[k];
  const T D0 = s.cD[c], D1 = D0 * mc.d1r[k], D2 = D0 * mc.d2r[k];
  const T U0 = jar[0] * mu, U1 = jar[1] * f1, U2 = jar[2] * f2;
  const T N = U0, T2 = U1 * U1 + U2 * U2, Tn = bsqrt(T2);
  T* fc = s.cFrc[c];
  if (N >= mu * Tn || (Tn <= 0 && N >= 0)) { fc[0] = fc[1] = fc[2] = 0; s.cstate[c] = 0; return 0; }
  if (mu * N + Tn <= 0 || (Tn <= 0 && N < 0)) {
    fc[0] = -D0 * jar[0]; fc[1] = -D1 * jar[1]; fc[2] = -D2 * jar[2]; s.cstate[c] = 1;
    return (T)0.5 * (D0 * jar[0] * jar[0] + D1 * jar[1] * jar[1] + D2 * jar[2] * jar[2]);
  }
  const T Dm = D0 / (mu * mu * ((T)1 + mu * mu));
  const T NT = N - mu * Tn;
  fc[0] = -Dm * NT * mu;
  const T sc = -fc[0] / Tn;
  fc[1] = sc * f1 * U1; fc[2] = sc * f2 * U2;
  s.cstate[c] = 2;
  if (wantH) {
    // cone Hessian in contact-row space (xx,yy,zz,xy,xz,yz) with rows (n,t1,t2)
    const T iT = (T)1 / Tn, muN_T3 = mu * N * iT * iT * iT, dg = mu * mu - mu * N * iT;
    T* h = s.cHc[c];
    h[0] = Dm * mu * mu;
    h[1] = Dm * f1 * f1 * (muN_T3 * U1 * U1 + dg);
    h[2] = Dm * f2 * f2 * (muN_T3 * U2 * U2 + dg);
    h[3] = Dm * mu * f1 * (-mu * U1 * iT);
    h[4] = Dm * mu * f2 * (-mu * U2 * iT);
    h[5] = Dm * f1 * f2 * (muN_T3 * U1 * U2);
  }
  return (T)0.5 * Dm * NT * NT;
}

// H += J_c' W J_c for contact c with W (3x3 symmetric, contact-row space)
template <typename T> BB_HD void addContactHessian(const Geo<T>& g, Scratch<T>& s, int c, const T* Wc) {
  // world-space weight Ww = F' W F
  const T* f = s.cF[c];
  T Ww[9];
  {
    // rows of F: n=f[0..2], t1=f[3..5], t2=f[6..8];  WF = W * F
    T WF[9];
    for (int j = 0; j < 3; j++) {
      WF[j] = Wc[0] * f[j] + Wc[3] * f[3 + j] + Wc[4] * f[6 + j];
      WF[3 + j] = Wc[3] * f[j] + Wc[1] * f[3 + j] + Wc[5] * f[6 + j];
      WF[6 + j] = Wc[4] * f[j] + Wc[5] * f[3 + j] + Wc[2] * f[6 + j];
    }
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Ww[3 * i + j] = f[i] * WF[j] + f[3 + i] * WF[3 + j] + f[6 + i] * WF[6 + j];
  }
  const V3<T> P = ld3(s.cP[c]);
  const int ty = s.ctype[c], wi = ctWheel(ty);
  V3<T> gcol[13]; int idx[13]; int n = 0;
  const V3<T> rL = P - g.pL;
  if (ctHasBase(ty)) {
    const V3<T> rB = P - g.pB;
    gcol[n] = mk((T)1, (T)0, (T)0); idx[n++] = 0; gcol[n] = mk((T)0, (T)1, (T)0); idx[n++] = 1; gcol[n] = mk((T)0, (T)0, (T)1); idx[n++] = 2;
    gcol[n] = cross(g.RB.c0, rB); idx[n++] = 3; gcol[n] = cross(g.RB.c1, rB); idx[n++] = 4; gcol[n] = cross(g.RB.c2, rB); idx[n++] = 5;
    if (wi >= 0) { gcol[n] = cross(g.aw[wi], P - g.hw[wi]); idx[n++] = 6 + wi; }
    if (ctHasBall(ty)) {
      gcol[n] = mk((T)-1, (T)0, (T)0); idx[n++] = 9; gcol[n] = mk((T)0, (T)-1, (T)0); idx[n++] = 10; gcol[n] = mk((T)0, (T)0, (T)-1); idx[n++] = 11;
      gcol[n] = cross(rL, g.RL.c0); idx[n++] = 12; gcol[n] = cross(rL, g.RL.c1); idx[n++] = 13; gcol[n] = cross(rL, g.RL.c2); idx[n++] = 14;
    }
  } else {
    gcol[n] = mk((T)1, (T)0, (T)0); idx[n++] = 9; gcol[n] = mk((T)0, (T)1, (T)0); idx[n++] = 10; gcol[n] = mk((T)0, (T)0, (T)1); idx[n++] = 11;
    gcol[n] = cross(g.RL.c0, rL); idx[n++] = 12; gcol[n] = cross(g.RL.c1, rL); idx[n++] = 13; gcol[n] = cross(g.RL.c2, rL); idx[n++] = 14;
  }
  for (int a = 0; a < n; a++) {
    const V3<T> ya = mk(Ww[0] * gcol[a].x + Ww[1] * gcol[a].y + Ww[2] * gcol[a].z, Ww[3] * gcol[a].x + Ww[4] * gcol[a].y + Ww[5] * gcol[a].z,
                        Ww[6] * gcol[a].x + Ww[7] * gcol[a].y + Ww[8] * gcol[a].z);
    const int ia = idx[a];
    for (int b = 0; b <= a; b++) s.H[ia * (ia + 1) / 2 + idx[b]] += dot(gcol[b], ya);   // idx is increasing
  }
}

// ---------------------------------------------------------------------------------------------- Newton solver
template <typename T> struct LsPt { T alpha, cost, d1, d2; };

template <typename T> struct Newton {
  const ModelConst<T>& mc; const Geo<T>& g; Scratch<T>& s;
  T qG0, qG1, qG2;   // Gauss quadratic along the search direction
  T cost, gauss;
  bool fast;         // solver_mode 1: lineSearchFast instead of the reference's exact search (same minimiser, fewer evaluations)
  BB_HDN Newton(const ModelConst<T>& m, const Geo<T>& gg, Scratch<T>& ss, bool f = false) : mc(m), g(gg), s(ss), fast(f) {}

  // total cost at arbitrary qacc (warm-start test); uses s.Mv / s.cJv as temporaries
  BB_HD T costAt(const T* qa) {
    symvPacked(s.M, qa, s.Mv);
    T gs = 0; for (int i = 0; i < NV; i++) gs += (T)0.5 * (s.Mv[i] - s.qfs[i]) * (qa[i] - s.qas[i]);
    const VelCtx<T> vc = velCtx(g, qa);
    T cs = 0;
    for (int c = 0; c < s.nc; c++) {
      T jr[3]; contactVel(g, s, c, vc, jr);
      jr[0] -= s.cAref[c][0]; jr[1] -= s.cAref[c][1]; jr[2] -= s.cAref[c][2];
      cs += coneUpdate(mc, s, c, jr, false);
    }
    return gs + cs;
  }
  // forces/states/cost at the current jar, Hessian + factorisation, gradient and Newton direction
  BB_HD void update(const T* qacc) {
    T cs = 0;
    for (int i = 0; i < NTRI; i++) s.H[i] = s.M[i];
    for (int c = 0; c < s.nc; c++) {
      cs += coneUpdate(mc, s, c, s.cJar[c], true);
      if (s.cstate[c] == 1) {
        const int k = ctFric(s.ctype[c]); const T D0 = s.cD[c];
        const T W[6] = {D0, D0 * mc.d1r[k], D0 * mc.d2r[k], 0, 0, 0};
        addContactHessian(g, s, c, W);
      } else if (s.cstate[c] == 2) addContactHessian(g, s, c, s.cHc[c]);
    }
    gauss = 0; for (int i = 0; i < NV; i++) gauss += (T)0.5 * (s.Ma[i] - s.qfs[i]) * (qacc[i] - s.qas[i]);
    cost = gauss + cs;
    cholPacked(s.H);
    for (int i = 0; i < NV; i++) s.grad[i] = s.Ma[i] - s.qfs[i];
    subJtF(g, s, s.grad);
    cholSolvePacked(s.H, s.grad, s.Mgrad);
  }
  BB_HD LsPt<T> eval(T alpha) {
#if defined(BB_STATS) && !defined(__CUDA_ARCH__)
    bb_stats_ls_evals++;
#endif
    LsPt<T> p; p.alpha = alpha;
    p.cost = qG0 + alpha * (qG1 + alpha * qG2); p.d1 = qG1 + (T)2 * alpha * qG2; p.d2 = (T)2 * qG2;
    for (int c = 0; c < s.nc; c++) {
      const int k = ctFric(s.ctype[c]);
      const T mu = mc.mu[k], f1 = mc.f1[k], f2 = mc.f2[k];
      const T* jr = s.cJar[c]; const T* jv = s.cJv[c];
      const T D0 = s.cD[c], D1 = D0 * mc.d1r[k], D2 = D0 * mc.d2r[k];
      const T U0 = jr[0] * mu, V0 = jv[0] * mu;
      const T u1 = jr[1] * f1, u2 = jr[2] * f2, v1 = jv[1] * f1, v2 = jv[2] * f2;
      const T UU = u1 * u1 + u2 * u2, UV = u1 * v1 + u2 * v2, VV = v1 * v1 + v2 * v2;
      const T N = U0 + alpha * V0, Tsq = UU + alpha * ((T)2 * UV + alpha * VV);
      bool bottom = false;
      if (Tsq <= 0) { bottom = N < 0; }
      else {
        const T Tn = bsqrt(Tsq);
        if (N >= mu * Tn) {}
        else if (mu * N + Tn <= 0) bottom = true;
        else {
          const T Dm = D0 / (mu * mu * ((T)1 + mu * mu));
          const T N1 = V0, T1 = (UV + alpha * VV) / Tn, T2 = VV / Tn - (UV + alpha * VV) * T1 / Tsq;
          const T NT = N - mu * Tn, dNT = N1 - mu * T1;
          p.cost += (T)0.5 * Dm * NT * NT; p.d1 += Dm * NT * dNT; p.d2 += Dm * (dNT * dNT - NT * mu * T2);
        }
      }
      if (bottom) {
        const T q0 = (T)0.5 * (D0 * jr[0] * jr[0] + D1 * jr[1] * jr[1] + D2 * jr[2] * jr[2]);
        const T q1 = D0 * jr[0] * jv[0] + D1 * jr[1] * jv[1] + D2 * jr[2] * jv[2];
        const T q2 = (T)0.5 * (D0 * jv[0] * jv[0] + D1 * jv[1] * jv[1] + D2 * jv[2] * jv[2]);
        p.cost += q0 + alpha * (q1 + alpha * q2); p.d1 += q1 + (T)2 * alpha * q2; p.d2 += (T)2 * q2;
      }
    }
    if (p.d2 < (T)1e-15) p.d2 = (T)1e-15;
    return p;
  }
  BB_HD int bracket(LsPt<T>& p, const LsPt<T>* cand, LsPt<T>& pnext) {
    int flag = 0;
    for (int i = 0; i < 3; i++) {
      if (p.d1 < 0 && cand[i].d1 < 0 && p.d1 < cand[i].d1) { p = cand[i]; flag = 1; }
      else if (p.d1 > 0 && cand[i].d1 > 0 && p.d1 > cand[i].d1) { p = cand[i]; flag = 2; }
    }
    if (flag) pnext = eval(p.alpha - p.d1 / p.d2);
    return flag;
  }
  // solver_mode 1.  The search direction is the exact Newton direction, so phi(0) = cost, phi'(0) = grad . search and
  // phi''(0) = -phi'(0) are known without an evaluation and the first 1-D Newton point is alpha = 1.  The search is a
  // safeguarded 1-D Newton iteration on phi' inside a bracket [lo, hi] that stops at the strong Wolfe conditions
  // (c1 = 1e-4, c2 = 0.1) instead of the reference's |phi'| < tolerance * ls_tolerance * |search| / scale.  phi' jumps at
  // the apex of a friction cone (T = 0): for the omniwheel pairs (friction 1 : 0.001) the tangential residual is almost
  // one-dimensional, every search line passes within ~1e-6 of an apex and the minimiser very often sits on it.  The apex
  // of wheel pair c is at alpha_c = -UV_c / VV_c (minimum of T^2 along the line), so a Newton step that would jump across
  // it lands on it first; from there the iteration converges inside the apex's narrow smooth valley.  (The reference's
  // search reaches the same point by bisection: ~20 evaluations.)
  BB_HD T lineSearchFast(T gtol) {
    T d0 = 0; for (int i = 0; i < NV; i++) d0 += s.grad[i] * s.search[i];
    const T c0 = cost, wtol = bmax((T)0.1 * babs(d0), gtol);
    // rounding floor of the working precision: cost differences below it are noise, a bracket cannot shrink below ~1 ulp
    const T cslack = sizeof(T) == 4 ? (T)4e-7 * babs(c0) : (T)0, wrel = sizeof(T) == 4 ? (T)1e-6 : (T)1e-12;
    T kink[3]; int nk = -1; unsigned kused = 0;
    T lo = 0, hi = -1, a = 1, wprev = (T)1e30, bestA = 0, bestC = c0; bool have = false;
    for (int k = 0; k < 30; k++) {
      const LsPt<T> p = eval(a);
      const bool armijo = p.cost <= c0 + (T)1e-4 * a * d0 + cslack;
      if (armijo && (!have || p.cost < bestC)) { bestA = a; bestC = p.cost; have = true; }
      if (armijo && babs(p.d1) <= wtol) return a;
      if (!armijo || p.d1 > 0) hi = a; else lo = a;
      T an = a - p.d1 / p.d2;
      if (hi < 0) { if (!(an > a * (T)1.1)) an = a * (T)1.1; if (an > a * 4) an = a * 4; }
      else {
        const T w = hi - lo, mid = (T)0.5 * (lo + hi);
        if (w < wrel * hi) break;
        // bisect when the cost rose although phi' < 0 (a cone switched), when the step leaves the bracket, or when two
        // evaluations did not halve the bracket
        if ((!armijo && p.d1 < 0) || !(an > lo && an < hi) || (k >= 2 && (k & 1) == 0 && w > (T)0.5 * wprev)) an = mid;
        if ((k & 1) == 0) wprev = w;
      }
      if (nk < 0) {   // apex positions of the anisotropic (wheel) pairs, computed when the unit step was not accepted
        nk = 0;
        for (int c = 0; c < s.nc && nk < 3; c++) {
          const int kf = ctFric(s.ctype[c]); if (kf != 0) continue;
          const T f1 = mc.f1[kf], f2 = mc.f2[kf];
          const T u1 = s.cJar[c][1] * f1, u2 = s.cJar[c][2] * f2, v1 = s.cJv[c][1] * f1, v2 = s.cJv[c][2] * f2;
          const T UV = u1 * v1 + u2 * v2, VV = v1 * v1 + v2 * v2;
          kink[nk++] = VV > (T)1e-30 ? -UV / VV : (T)-1;
        }
      }
      int kb = -1;
      for (int i = 0; i < nk; i++) {
        const T kk = kink[i];
        if ((kused >> i) & 1u || !(kk > 0)) continue;
        const bool between = an > a ? (kk > a && kk < an) : (kk < a && kk > an);
        if (between && (kb < 0 || babs(kk - a) < babs(kink[kb] - a))) kb = i;
      }
      if (kb >= 0) { an = kink[kb]; kused |= 1u << kb; }
      a = an;
    }
    if (sizeof(T) == 4 && have && c0 - bestC <= (T)1e-5 * babs(c0)) return 0;   // fp32: only noise-level improvement was found: converged
    return have ? bestA : (T)0;
  }
  BB_HD T lineSearch(T scale) {
    T sn = 0; for (int i = 0; i < NV; i++) sn += s.search[i] * s.search[i];
    sn = bsqrt(sn);
    if (sn < (T)1e-15) return 0;
    const T gtol = mc.tolerance * mc.ls_tolerance * sn / scale;
    symvPacked(s.M, s.search, s.Mv);
    const VelCtx<T> vc = velCtx(g, (const T*)s.search);
    for (int c = 0; c < s.nc; c++) contactVel(g, s, c, vc, s.cJv[c]);
    qG0 = gauss; qG1 = 0; qG2 = 0;
    for (int i = 0; i < NV; i++) { qG1 += s.search[i] * (s.Ma[i] - s.qfs[i]); qG2 += (T)0.5 * s.search[i] * s.Mv[i]; }
    if (fast) return lineSearchFast(gtol);
    int it = 0;
    const LsPt<T> p0 = eval((T)0);
    LsPt<T> p1 = eval(p0.alpha - p0.d1 / p0.d2), p2 = p0, pmid, p1n, p2n;
    if (p0.cost < p1.cost) p1 = p0;
    if (babs(p1.d1) < gtol) return p1.alpha;
    const T dir = p1.d1 < 0 ? (T)1 : (T)-1;
    bool upd = false;
    while (p1.d1 * dir <= -gtol && it < mc.ls_iterations) {
      p2 = p1; upd = true;
      p1 = eval(p1.alpha - p1.d1 / p1.d2); it++;
      if (babs(p1.d1) < gtol) return p1.alpha;
    }
    if (it >= mc.ls_iterations || !upd) return p1.alpha;
    p2n = p1; p1n = eval(p1.alpha - p1.d1 / p1.d2);
    while (it < mc.ls_iterations) {
      pmid = eval((T)0.5 * (p1.alpha + p2.alpha)); it++;
      const LsPt<T> cand[3] = {p1n, p2n, pmid};
      for (int i = 0; i < 3; i++) if (babs(cand[i].d1) < gtol) return cand[i].alpha;
      const int b1 = bracket(p1, cand, p1n), b2 = bracket(p2, cand, p2n);
      if (!b1 && !b2) return pmid.cost < p0.cost ? pmid.alpha : (T)0;
    }
    if (p1.cost <= p2.cost && p1.cost < p0.cost) return p1.alpha;
    if (p2.cost <= p1.cost && p2.cost < p0.cost) return p2.alpha;
    return 0;
  }
  // returns iteration count; qacc in/out (on entry: unused), warm = qacc_warmstart
  BB_HD int run(const T* warm, T* qacc) {
    const T cw = costAt(warm), cs0 = costAt(s.qas);
    const T* start = cw > cs0 ? s.qas : warm;
    for (int i = 0; i < NV; i++) qacc[i] = start[i];
    symvPacked(s.M, qacc, s.Ma);
    {
      const VelCtx<T> vc = velCtx(g, (const T*)qacc);
      for (int c = 0; c < s.nc; c++) {
        contactVel(g, s, c, vc, s.cJar[c]);
        s.cJar[c][0] -= s.cAref[c][0]; s.cJar[c][1] -= s.cAref[c][1]; s.cJar[c][2] -= s.cAref[c][2];
      }
    }
    update(qacc);
    for (int i = 0; i < NV; i++) s.search[i] = -s.Mgrad[i];
    const T scale = (T)1 / (mc.meaninertia * (T)NV);
    int iter = 0;
    while (iter < mc.iterations) {
      const T alpha = lineSearch(scale);
      if (alpha == 0) break;
      for (int i = 0; i < NV; i++) { qacc[i] += alpha * s.search[i]; s.Ma[i] += alpha * s.Mv[i]; }
      for (int c = 0; c < s.nc; c++) { s.cJar[c][0] += alpha * s.cJv[c][0]; s.cJar[c][1] += alpha * s.cJv[c][1]; s.cJar[c][2] += alpha * s.cJv[c][2]; }
      const T old = cost;
      update(qacc);
#if defined(BB_STATS) && !defined(__CUDA_ARCH__)
      bb_stats_newton++;
#endif
      T gn = 0; for (int i = 0; i < NV; i++) gn += s.grad[i] * s.grad[i];
      iter++;
      if (scale * (old - cost) < mc.tolerance || scale * bsqrt(gn) < mc.tolerance) break;
      if (sizeof(T) == 4 && old - cost <= (T)5e-7 * babs(old)) break;   // fp32: improvement inside the rounding noise of the cost
      for (int i = 0; i < NV; i++) s.search[i] = -s.Mgrad[i];
    }
    return iter;
  }
};

// ---------------------------------------------------------------------------------------------- one mj_forward
template <typename T>
BB_NOINL void forwardDynamics(const ModelConst<T>& mc, T* qpos, const T* qvel, const T* ctrl, const T* warm, const float* hf, T zscale,
                           Scratch<T>& s, T* qacc, KinOut<T>* kin, bool fast = false) {
  Geo<T> g; V3<T> capC[3], capU[3];
  smoothDynamics(mc, qpos, qvel, ctrl, s.M, s.qfs, g, capC, capU, kin);
  // qacc_smooth = M^-1 qfrc_smooth
  for (int i = 0; i < NTRI; i++) s.H[i] = s.M[i];
  cholPacked(s.H);
  cholSolvePacked(s.H, s.qfs, s.qas);
  collide(mc, g, capC, capU, hf, zscale, s);
  int niter = 0;
  if (s.nc == 0) { for (int i = 0; i < NV; i++) qacc[i] = s.qas[i]; }
  else {
    // impedance, regulariser, reference acceleration (mj_makeImpedance / mj_referenceConstraint)
    const VelCtx<T> vc = velCtx(g, qvel);
    for (int c = 0; c < s.nc; c++) {
      const T dist = s.cDist[c];
      const T x = babs(dist) / mc.solimp[2];
      T imp;
      if (x >= 1) imp = mc.solimp[1];
      else if (x <= 0) imp = mc.solimp[0];
      else {
        const T mid = mc.solimp[3];   // power == 2 (MuJoCo default solimp), ballbot.xml has no override
        const T y = x <= mid ? x * x / mid : (T)1 - ((T)1 - x) * ((T)1 - x) / ((T)1 - mid);
        imp = mc.solimp[0] + y * (mc.solimp[1] - mc.solimp[0]);
      }
      const T R0 = bmax((T)1e-15, ((T)1 - imp) * mc.dA[s.ctype[c]] / imp);
      s.cD[c] = (T)1 / R0;
      T vel[3]; contactVel(g, s, c, vc, vel);
      s.cAref[c][0] = -mc.B * vel[0] - mc.K * imp * dist;
      s.cAref[c][1] = -mc.B * vel[1];
      s.cAref[c][2] = -mc.B * vel[2];
    }
    Newton<T> nw(mc, g, s, fast);
    niter = nw.run(warm, qacc);
  }
  if (kin) { kin->ncon = s.nc; kin->niter = niter; }
}

// free-joint position integration (mj_integratePos): pos += h v ; quat <- quat * exp(h w_local)
template <typename T> BB_HD void integratePos(T* qpos, const T* v, T h) {
  for (int b = 0; b < 2; b++) {
    T* q = qpos + (b ? 10 : 0); const T* u = v + (b ? 9 : 0);
    q[0] += h * u[0]; q[1] += h * u[1]; q[2] += h * u[2];
    const T wn = bsqrt(u[3] * u[3] + u[4] * u[4] + u[5] * u[5]);
    T n = bsqrt(q[3] * q[3] + q[4] * q[4] + q[5] * q[5] + q[6] * q[6]);
    n = n < (T)1e-15 ? (T)1 : (T)1 / n;
    const T w0 = q[3] * n, x0 = q[4] * n, y0 = q[5] * n, z0 = q[6] * n;
    if (wn < (T)1e-15) { q[3] = w0; q[4] = x0; q[5] = y0; q[6] = z0; continue; }
    T sa, ca; bsincos((T)0.5 * h * wn, &sa, &ca);
    const T ax = u[3] / wn * sa, ay = u[4] / wn * sa, az = u[5] / wn * sa;
    q[3] = w0 * ca - x0 * ax - y0 * ay - z0 * az;
    q[4] = w0 * ax + x0 * ca + y0 * az - z0 * ay;
    q[5] = w0 * ay - x0 * az + y0 * ca + z0 * ax;
    q[6] = w0 * az + x0 * ay - y0 * ax + z0 * ca;
  }
  qpos[7] += h * v[6]; qpos[8] += h * v[7]; qpos[9] += h * v[8];
}

// one mj_step with integrator RK4 (ballbot.xml:5). qpos/qvel/warm are updated in place; kin receives the
// kinematics of the LAST stage evaluation (what the reference env reads stale after mj_step).
template <typename T>
BB_HD void rk4Step(const ModelConst<T>& mc, T* qpos, T* qvel, T* warm, const T* ctrl, const float* hf, T zscale, Scratch<T>& s,
                   KinOut<T>* kin, T* qlast = nullptr, bool fast = false) {
  const T h = mc.timestep;
  // stage state (xq, xv), saved initial state (q0, v0), RK4-weighted sums of stage velocities / accelerations
  T q0[NQ], v0[NV], xq[NQ], xv[NV], sumv[NV], suma[NV], acc[NV], wchain[NV];
  normalizeQuats(qpos);   // mj_kinematics normalises qpos in place during the first forward pass
  for (int i = 0; i < NQ; i++) { q0[i] = qpos[i]; xq[i] = qpos[i]; }
  for (int i = 0; i < NV; i++) { v0[i] = qvel[i]; xv[i] = qvel[i]; sumv[i] = (T)0; suma[i] = (T)0; acc[i] = (T)0; }
  int ncmax = 0, nit = 0;
#pragma unroll 1
  for (int st = 0; st < 4; st++) {
    // solver_mode 1 chains the warm start through the stages (acc = the previous stage's solution); the reference restarts
    // every stage from qacc_warmstart of the previous step
    forwardDynamics(mc, xq, xv, ctrl, (fast && st > 0) ? (const T*)wchain : (const T*)warm, hf, zscale, s, acc, kin, fast);
    if (fast) { for (int i = 0; i < NV; i++) wchain[i] = acc[i]; }
    if (kin) { ncmax = kin->ncon > ncmax ? kin->ncon : ncmax; nit += kin->niter; }
    const T bw = (st == 0 || st == 3) ? (T)(1.0 / 6.0) : (T)(1.0 / 3.0);       // RK4_B
    for (int i = 0; i < NV; i++) { sumv[i] += bw * xv[i]; suma[i] += bw * acc[i]; }
    if (st == 3) {
      if (qlast) { for (int i = 0; i < NQ; i++) qlast[i] = xq[i]; }   // configuration the renderer sees (stale, App. C #2)
    } else {
      const T ha = (st == 2) ? h : (T)0.5 * h;                                   // RK4_A: 1/2, 1/2, 1
      for (int i = 0; i < NQ; i++) xq[i] = q0[i];
      integratePos(xq, xv, ha);
      for (int i = 0; i < NV; i++) xv[i] = v0[i] + ha * acc[i];
    }
  }
  for (int i = 0; i < NQ; i++) qpos[i] = q0[i];
  integratePos(qpos, sumv, h);
  for (int i = 0; i < NV; i++) { qvel[i] = v0[i] + h * suma[i]; warm[i] = acc[i]; }
  if (kin) { kin->ncon = ncmax; kin->niter = nit; }
}

// ---------------------------------------------------------------------------------------------- env-level helpers
// proprioceptive observation (ballbot_env.py:772-811), float32 like the reference (_default_dtype)
template <typename T>
BB_HD void proprioObs(const KinOut<T>& k, const T* qvel, T max_wheel_vel, float* orient, float* angvel, float* vel, float* motor) {
  // numpy-quaternion as_rotation_vector = 2 log(q)
  const T w = k.quatB[0], x = k.quatB[1], y = k.quatB[2], z = k.quatB[3];
  const T b = bsqrt(x * x + y * y + z * z);
  T rx = 0, ry = 0, rz = 0;
  if (b <= (T)1e-14 * babs(w)) { if (w < 0) rx = (T)6.283185307179586; }
  else { const T f = (T)2 * batan2(b, w) / b; rx = f * x; ry = f * y; rz = f * z; }
  orient[0] = (float)rx; orient[1] = (float)ry; orient[2] = (float)rz;
  for (int i = 0; i < 3; i++) {
    float a = (float)k.cvel_lin[i], v = (float)k.cvel_ang[i], m = (float)qvel[1 + i] / (float)max_wheel_vel;
    angvel[i] = a < -2.f ? -2.f : (a > 2.f ? 2.f : a);     // "angular_vel" <- cvel[3:6] (sic)
    vel[i] = v < -2.f ? -2.f : (v > 2.f ? 2.f : v);        // "vel" <- cvel[0:3]
    motor[i] = m < -2.f ? -2.f : (m > 2.f ? 2.f : m);      // qvel[1:4]/10 (sic, App. C #3)
  }
}
// tilt angle in degrees from the float32 rotation vector (ballbot_env.py:989-1006), double math like numpy
BB_HD double tiltDegrees(const float* orient) {
  const double rx = orient[0], ry = orient[1], rz = orient[2];
  const double th = sqrt(rx * rx + ry * ry + rz * rz);
  double qx = 0, qy = 0;
  if (th > 1e-300) { const double sc = sin(0.5 * th) / th; qx = rx * sc; qy = ry * sc; }
  return acos(1.0 - 2.0 * (qx * qx + qy * qy)) * 57.29577951308232;
}

}  // namespace bb
