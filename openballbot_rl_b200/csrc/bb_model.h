// bb_model.h -- host-side derivation of the engine's model constants (double precision) from the numbers in
// the reference MJCF (ballbot_gym/models/ballbot.xml:3-5,23,35-93) and MuJoCo's compile rules
// (inertia from geoms, euler "xyz" intrinsic in degrees, fromto capsules, invweight0 at qpos0).
// Runs once in bb_create(); the hot path only reads the resulting ModelConst<T> from constant memory.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>

#include "bb_core.cuh"

namespace bb {

namespace mdl {
struct M33 { double m[9]; };
inline M33 mmul(const M33& a, const M33& b) {
  M33 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.m[3 * i + j] = a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j] + a.m[3 * i + 2] * b.m[6 + j];
  return r;
}
inline void mv(const M33& a, const double* v, double* r) {
  double x = a.m[0] * v[0] + a.m[1] * v[1] + a.m[2] * v[2], y = a.m[3] * v[0] + a.m[4] * v[1] + a.m[5] * v[2],
         z = a.m[6] * v[0] + a.m[7] * v[1] + a.m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
inline M33 rotAxis(int axis, double deg) {
  const double a = deg * M_PI / 180.0, c = std::cos(a), s = std::sin(a);
  M33 r = {{1, 0, 0, 0, 1, 0, 0, 0, 1}};
  if (axis == 0) { r.m[4] = c; r.m[5] = -s; r.m[7] = s; r.m[8] = c; }
  if (axis == 1) { r.m[0] = c; r.m[2] = s; r.m[6] = -s; r.m[8] = c; }
  if (axis == 2) { r.m[0] = c; r.m[1] = -s; r.m[3] = s; r.m[4] = c; }
  return r;
}
inline M33 euler(double ex, double ey, double ez) { return mmul(mmul(rotAxis(0, ex), rotAxis(1, ey)), rotAxis(2, ez)); }  // R = Rx Ry Rz
// solid primitives: mass and (transverse, axial) inertia about the centre, axis = local z
inline void capsule(double r, double hl, double rho, double* m, double* It, double* Ia) {
  const double h = 2 * hl, vc = M_PI * r * r * h, vs = 4.0 / 3.0 * M_PI * r * r * r, mc = rho * vc, ms = rho * vs;
  *m = mc + ms;
  *It = mc * (3 * r * r + h * h) / 12.0 + 0.4 * ms * r * r + ms * h * (3 * r + 2 * h) / 8.0;
  *Ia = mc * r * r / 2.0 + 0.4 * ms * r * r;
}
inline void cylinder(double r, double hl, double rho, double* m, double* It, double* Ia) {
  const double h = 2 * hl; *m = rho * M_PI * r * r * h; *It = *m * (3 * r * r + h * h) / 12.0; *Ia = *m * r * r / 2.0;
}
// accumulate an axisymmetric piece (axis u) at c into (mass, first moment, inertia about the origin)
struct Acc { double m, mc[3], I[6]; };
inline void addAxisym(Acc& a, double m, const double* c, const double* u, double It, double Ia) {
  a.m += m; for (int k = 0; k < 3; k++) a.mc[k] += m * c[k];
  const double d = Ia - It, cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
  a.I[0] += It + d * u[0] * u[0] + m * (cc - c[0] * c[0]); a.I[1] += It + d * u[1] * u[1] + m * (cc - c[1] * c[1]);
  a.I[2] += It + d * u[2] * u[2] + m * (cc - c[2] * c[2]);
  a.I[3] += d * u[0] * u[1] - m * c[0] * c[1]; a.I[4] += d * u[0] * u[2] - m * c[0] * c[2]; a.I[5] += d * u[1] * u[2] - m * c[1] * c[2];
}
}  // namespace mdl

inline void buildModelConst(ModelConst<double>& mc) {
  using namespace mdl;
  std::memset(&mc, 0, sizeof(mc));
  // ---- options (ballbot.xml:3-5 + MuJoCo defaults)
  mc.timestep = 0.002; mc.grav = 9.81; mc.tolerance = 1e-8; mc.ls_tolerance = 0.01; mc.iterations = 100; mc.ls_iterations = 50;
  const double impratio = 1.0, solref[2] = {0.02, 1.0}, solimp[5] = {0.9, 0.95, 0.001, 0.5, 2.0};
  for (int k = 0; k < 5; k++) mc.solimp[k] = solimp[k];
  mc.inv_width = 1.0 / solimp[2]; mc.inv_mid = 1.0 / solimp[3]; mc.inv_1mmid = 1.0 / (1.0 - solimp[3]);
  const double tc = std::fmax(solref[0], 2 * mc.timestep);
  mc.K = 1.0 / (solimp[1] * solimp[1] * tc * tc * solref[1] * solref[1]);
  mc.B = 2.0 / (solimp[1] * tc);
  // explicit pairs friction="0.001 1.0" (ballbot.xml:90-92); ball-terrain: geom defaults (1, 1)
  const double fw[2] = {0.001, 1.0}, fh[2] = {1.0, 1.0};
  mc.f1[0] = fw[0]; mc.f2[0] = fw[1]; mc.f1[1] = fh[0]; mc.f2[1] = fh[1];
  for (int k = 0; k < 2; k++) {
    mc.d1r[k] = impratio;                                                   // R_t1 = R_n / impratio
    mc.d2r[k] = mc.d1r[k] * (mc.f2[k] * mc.f2[k]) / (mc.f1[k] * mc.f1[k]);  // R_t2 = R_t1 mu1^2 / mu2^2
    mc.mu[k] = mc.f1[k] * std::sqrt(1.0 / impratio);                        // regularised cone
    mc.dmr[k] = 1.0 / (mc.mu[k] * mc.mu[k] * (1.0 + mc.mu[k] * mc.mu[k]));
  }
  mc.hx = 5.0; mc.hbase = 0.1;
  // ---- base group: tower cylinder + ballast box + two camera sticks (cone meshes: file missing, density 1 -> dropped)
  Acc g; std::memset(&g, 0, sizeof(g));
  double m, It, Ia;
  { cylinder(0.11, 0.14, 23.6, &m, &It, &Ia); const double c[3] = {0, 0, 0.2}, u[3] = {0, 0, 1}; addAxisym(g, m, c, u, It, Ia);
    mc.tower_c[0] = 0; mc.tower_c[1] = 0; mc.tower_c[2] = 0.2; mc.tower_r = 0.11; mc.tower_hl = 0.14; }
  { const double mb = 400.0 * 0.008, ib = mb / 3.0 * 0.02, c[3] = {0, 0, 0.002}, u[3] = {0, 0, 1}; addAxisym(g, mb, c, u, ib, ib); }
  for (int i = 0; i < 2; i++) {
    const double sg = i ? 1.0 : -1.0;
    const double bpos[3] = {i ? -0.17 : 0.17, -0.01, -0.06};   // cam_0: +0.17, cam_1: -0.17
    const M33 Rb = euler(180, i ? 30 : -30, 0);
    const double cl[3] = {0.1 * sg, 0, 0}, ul[3] = {sg, 0, 0};
    double c[3], u[3]; mv(Rb, cl, c); mv(Rb, ul, u);
    for (int k = 0; k < 3; k++) c[k] += bpos[k];
    capsule(0.01, 0.1, 1000.0, &m, &It, &Ia); addAxisym(g, m, c, u, It, Ia);
    for (int k = 0; k < 3; k++) { mc.stick_c[i][k] = c[k]; mc.stick_u[i][k] = u[k]; mc.cam_pos[i][k] = bpos[k]; }
    const M33 Rc = mmul(Rb, euler(180, 0, 0));
    for (int k = 0; k < 9; k++) mc.cam_rot[i][k] = Rc.m[k];   // row-major: world_dir = Rc * cam_dir
  }
  mc.stick_r = 0.01; mc.stick_hl = 0.1;
  mc.m0 = g.m;
  for (int k = 0; k < 3; k++) mc.c0[k] = g.mc[k] / g.m;
  for (int k = 0; k < 6; k++) mc.I0o[k] = g.I[k];
  {
    const double* c = mc.c0; const double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    mc.I0c[0] = g.I[0] - g.m * (cc - c[0] * c[0]); mc.I0c[1] = g.I[1] - g.m * (cc - c[1] * c[1]); mc.I0c[2] = g.I[2] - g.m * (cc - c[2] * c[2]);
    mc.I0c[3] = g.I[3] + g.m * c[0] * c[1]; mc.I0c[4] = g.I[4] + g.m * c[0] * c[2]; mc.I0c[5] = g.I[5] + g.m * c[1] * c[2];
  }
  // ---- wheels
  capsule(0.025, 0.02, 620.0, &mc.mw, &mc.It, &mc.Ia);
  mc.armature = 0.005; mc.damping = 0.8; mc.wheel_r = 0.025; mc.wheel_hl = 0.02;
  {
    double a[3] = {-0.15316554764123935, -0.6903189805903613, -0.7071067953657663};
    const double an = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    for (int k = 0; k < 3; k++) a[k] /= an;
    const double anchor[3] = {0, 0, 0.0293}, gpos[3] = {-0.018, -0.08, -0.053}, ez[3] = {0, 0, 1};
    const M33 Rg = euler(-45, 9, 0);
    double ug[3]; mv(Rg, ez, ug);
    for (int i = 0; i < 3; i++) {
      const M33 Rb = euler(0, 0, 120.0 * i);
      double t[3];
      mv(Rb, a, t); for (int k = 0; k < 3; k++) mc.ax[i][k] = t[k];
      mv(Rb, anchor, t); for (int k = 0; k < 3; k++) mc.anc[i][k] = t[k] + (k == 2 ? -0.001 : 0.0);
      const double sl[3] = {gpos[0] - anchor[0], gpos[1] - anchor[1], gpos[2] - anchor[2]};
      mv(Rb, sl, t); for (int k = 0; k < 3; k++) mc.s0[i][k] = t[k];
      mv(Rb, ug, t); for (int k = 0; k < 3; k++) mc.u0[i][k] = t[k];
    }
  }
  mc.mA = mc.m0 + 3 * mc.mw;
  // ---- ball
  mc.ball_r = 0.09; mc.dz = -0.14;
  mc.mL = 55.0 * 4.0 / 3.0 * M_PI * 0.09 * 0.09 * 0.09; mc.IL = 0.4 * mc.mL * 0.09 * 0.09;
  // ---- invweight0 / meaninertia at qpos0 (engine_setconst set0 semantics): M from the engine's own smooth dynamics
  {
    static Scratch<double> s;   // large; setup only
    Geo<double> ge; V3<double> cC[3], cU[3];
    double qpos[NQ] = {0, 0, 0.24, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.26, 1, 0, 0, 0}, qvel[NV] = {0}, ctrl[3] = {0, 0, 0};
    smoothDynamics(mc, qpos, qvel, ctrl, s.M, s.qfs, ge, cC, cU, (KinOut<double>*)nullptr);
    double tr = 0; for (int i = 0; i < NV; i++) tr += s.M[tidx(i, i)];
    mc.meaninertia = tr / NV;
    for (int i = 0; i < NTRI; i++) s.H[i] = s.M[i];
    cholPacked(s.H);
    auto invw = [&](const double J[3][NV]) {
      double acc = 0;
      for (int r = 0; r < 3; r++) { double x[NV]; cholSolvePacked(s.H, J[r], x); for (int k = 0; k < NV; k++) acc += J[r][k] * x[k]; }
      return std::fmax(1e-15, acc / 3.0);
    };
    // ball COM (sphere centre): v = vL + wL x d
    double Jb[3][NV]; std::memset(Jb, 0, sizeof(Jb));
    for (int k = 0; k < 3; k++) Jb[k][9 + k] = 1;
    { const double d[3] = {0, 0, mc.dz};
      for (int k = 0; k < 3; k++) { double e[3] = {0, 0, 0}; e[k] = 1; const double cx[3] = {e[1] * d[2] - e[2] * d[1], e[2] * d[0] - e[0] * d[2], e[0] * d[1] - e[1] * d[0]};
        for (int r = 0; r < 3; r++) Jb[r][12 + k] = cx[r]; } }
    const double wb = invw(Jb);
    // invweight0 of a base-tree body at its own COM r (base frame; R_B = 1 at qpos0): v = v_B + w x r (+ hinge column)
    auto invwBase = [&](const double* r, int wheel) {
      double Jw[3][NV]; std::memset(Jw, 0, sizeof(Jw));
      for (int k = 0; k < 3; k++) Jw[k][k] = 1;
      for (int k = 0; k < 3; k++) { double e[3] = {0, 0, 0}; e[k] = 1; const double cx[3] = {e[1] * r[2] - e[2] * r[1], e[2] * r[0] - e[0] * r[2], e[0] * r[1] - e[1] * r[0]};
        for (int q = 0; q < 3; q++) Jw[q][3 + k] = cx[q]; }
      if (wheel >= 0) { const double* a = mc.ax[wheel]; const double* sv = mc.s0[wheel]; const double cx[3] = {a[1] * sv[2] - a[2] * sv[1], a[2] * sv[0] - a[0] * sv[2], a[0] * sv[1] - a[1] * sv[0]};
        for (int q = 0; q < 3; q++) Jw[q][6 + wheel] = cx[q]; }
      return invw(Jw);
    };
    for (int i = 0; i < 3; i++) {
      const double r[3] = {mc.anc[i][0] + mc.s0[i][0], mc.anc[i][1] + mc.s0[i][1], mc.anc[i][2] + mc.s0[i][2]};
      const double ww = invwBase(r, i);
      mc.dA[i] = wb + ww;          // ball x wheel_i
      mc.dA[6 + i] = ww;           // heightfield x wheel_i (the world body has no inverse weight)
    }
    mc.dA[3] = wb;                 // heightfield x ball
    for (int i = 0; i < 2; i++) {  // camera bodies: their only massive geom is the stick, COM = stick centre
      const double wc = invwBase(mc.stick_c[i], -1);
      mc.dA[4 + i] = wc; mc.dA[10 + i] = wb + wc;
    }
    {  // base body proper: tower cylinder + ballast box
      double mt, it, ia; cylinder(0.11, 0.14, 23.6, &mt, &it, &ia);
      const double mb = 400.0 * 0.008, r[3] = {0, 0, (mt * 0.2 + mb * 0.002) / (mt + mb)};
      mc.dA[9] = wb + invwBase(r, -1);
    }
  }
}

// double -> T narrowing; relies on ModelConst's T members being one homogeneous leading block
template <typename T> inline void narrowModel(const ModelConst<double>& src, ModelConst<T>& dst) {
  typedef ModelConst<double> MCD;
  const size_t n = offsetof(MCD, iterations) / sizeof(double);
  const double* s = reinterpret_cast<const double*>(&src);
  T* d = reinterpret_cast<T*>(&dst);
  for (size_t i = 0; i < n; i++) d[i] = (T)s[i];
  dst.iterations = src.iterations; dst.ls_iterations = src.ls_iterations;
}

}  // namespace bb
