/*
 * ballbot_oracle.cpp -- fp64 single-env CPU ORACLE for the ballbot hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see ballbot_oracle.h).  PARITY UNPINNED: MuJoCo, `noise` and
 * numpy-quaternion are not available offline; this file restates their documented algorithms
 * for the single model ballbot_gym/models/ballbot.xml and the env logic of
 * ballbot_gym/envs/ballbot_env.py.  It is deliberately generic "MuJoCo-shaped" code
 * (body/dof tables, c-frame spatial vectors, dense efc_J) so that it shares no formulation
 * with the specialised CUDA engine it checks.
 *
 * Reference anchors (paths relative to /root/reference):
 *   model ................ ballbot_gym/models/ballbot.xml:3-5,23,35-93
 *   patched contact frame  tools/mujoco_fix.patch:9-18
 *   env step/obs/reset ... ballbot_gym/envs/ballbot_env.py:442-565,567-671,701-829,854-1036
 *   perlin terrain ....... ballbot_gym/terrain/perlin.py:8-74
 *   depth clip ........... ballbot_gym/sensors/rgbd.py:64-75
 *   rewards .............. ballbot_gym/rewards/directional.py:33-54
 */
#include "ballbot_oracle.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

typedef double R;
const R MINVAL = 1e-15;  // mjMINVAL
const int NB = 8, NQ = 17, NV = 15;
const int MAXCON = BBO_MAXCON, MAXEFC = 3 * BBO_MAXCON;
const int HN = BBO_HF_N;

// ----------------------------------------------------------------------------- small math
inline void v3set(R* r, R a, R b, R c) { r[0] = a; r[1] = b; r[2] = c; }
inline void v3cp(R* r, const R* a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
inline R dot3(const R* a, const R* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline void cross3(R* r, const R* a, const R* b) {
  R x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
inline R norm3(const R* a) { return std::sqrt(dot3(a, a)); }
inline R normalize3(R* a) {
  R n = norm3(a);
  if (n < MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; return 0; }
  a[0] /= n; a[1] /= n; a[2] /= n; return n;
}
inline void mulMatVec3(R* r, const R* m, const R* v) {  // r = m v (row-major 3x3)
  R x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  R y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  R z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
inline void mulMatTVec3(R* r, const R* m, const R* v) {  // r = m' v
  R x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  R y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  R z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
inline void mulMat3(R* r, const R* a, const R* b) {
  R t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  memcpy(r, t, sizeof(t));
}
inline void mulQuat(R* r, const R* a, const R* b) {
  R t[4] = {a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
            a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
            a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
            a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]};
  memcpy(r, t, sizeof(t));
}
inline void normalize4(R* q) {
  R n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  for (int i = 0; i < 4; i++) q[i] /= n;
}
inline void quat2mat(R* m, const R* q) {
  R q00 = q[0] * q[0], q11 = q[1] * q[1], q22 = q[2] * q[2], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[4] = q00 - q11 + q22 - q33; m[8] = q00 - q11 - q22 + q33;
  m[1] = 2 * (q[1] * q[2] - q[0] * q[3]); m[2] = 2 * (q[1] * q[3] + q[0] * q[2]);
  m[3] = 2 * (q[1] * q[2] + q[0] * q[3]); m[5] = 2 * (q[2] * q[3] - q[0] * q[1]);
  m[6] = 2 * (q[1] * q[3] - q[0] * q[2]); m[7] = 2 * (q[2] * q[3] + q[0] * q[1]);
}
inline void axisAngle2Quat(R* q, const R* axis, R angle) {
  if (angle == 0) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  R s = std::sin(angle * 0.5);
  q[0] = std::cos(angle * 0.5); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
// MJCF euler, default eulerseq "xyz" intrinsic, degrees  [3P-memory: user_objects ResolveOrientation]
inline void euler2quat(R* q, R ex, R ey, R ez) {
  const R d2r = M_PI / 180.0;
  R ax[3] = {1, 0, 0}, ay[3] = {0, 1, 0}, az[3] = {0, 0, 1}, qx[4], qy[4], qz[4];
  axisAngle2Quat(qx, ax, ex * d2r); axisAngle2Quat(qy, ay, ey * d2r); axisAngle2Quat(qz, az, ez * d2r);
  mulQuat(q, qx, qy); mulQuat(q, q, qz); normalize4(q);
}
// spatial cross products (vectors are [rot(3); lin(3)])  [3P-memory: mju_crossMotion / mju_crossForce]
inline void crossMotion(R* r, const R* vel, const R* v) {
  R a[3], b[3], c[3];
  cross3(a, vel, v); cross3(b, vel, v + 3); cross3(c, vel + 3, v);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
  r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
inline void crossForce(R* r, const R* vel, const R* f) {
  R a[3], b[3], c[3];
  cross3(a, vel, f); cross3(b, vel + 3, f + 3); cross3(c, vel, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
  r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}
// 10-number c-frame inertia times a motion vector [3P-memory: mju_mulInertVec]
inline void mulInertVec(R* r, const R* i, const R* v) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
inline R dot6(const R* a, const R* b) { return dot3(a, b) + dot3(a + 3, b + 3); }

// dense Cholesky (lower) of n x n, returns rank deficiency count
int cholFactor(R* A, int n) {
  int bad = 0;
  for (int j = 0; j < n; j++) {
    R s = A[j * n + j];
    for (int k = 0; k < j; k++) s -= A[j * n + k] * A[j * n + k];
    if (s < MINVAL) { s = MINVAL; bad++; }
    s = std::sqrt(s);
    A[j * n + j] = s;
    for (int i = j + 1; i < n; i++) {
      R t = A[i * n + j];
      for (int k = 0; k < j; k++) t -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = t / s;
    }
  }
  return bad;
}
void cholSolve(R* x, const R* L, const R* b, int n) {
  for (int i = 0; i < n; i++) {
    R s = b[i];
    for (int k = 0; k < i; k++) s -= L[i * n + k] * x[k];
    x[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    R s = x[i];
    for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
}

// ----------------------------------------------------------------------------- model
enum { JNONE = 0, JFREE = 1, JHINGE = 2 };
struct Geom {  // geoms used for collision/rendering
  int type;    // 0 sphere 1 capsule 2 cylinder
  int body;
  R pos[3], mat[9];  // in body frame
  R size[2];         // radius, half-length
};
struct Model {
  int parent[NB], rootid[NB], jtype[NB], qadr[NB], dadr[NB], dofnum[NB];
  R bpos[NB][3], bquat[NB][4], ipos[NB][3], inertia[NB][9], mass[NB];
  R jaxis[NB][3], janchor[NB][3];
  int dof_body[NV], dof_parent[NV];
  R armature[NV], damping[NV];
  R qpos0[NQ];
  Geom ball, wheel[3], tower, stick[2];
  R cam_pos[2][3], cam_mat[2][9];  // camera frames in the base body frame
  R invweight0[NB][2], meaninertia;
  // options (ballbot.xml:3-5; rest MuJoCo defaults)
  R timestep, gravity[3], tolerance, ls_tolerance, impratio;
  int iterations, ls_iterations;
  R solref[2], solimp[5];
  R fric_wheel[3], fric_hfield[3];  // (mu_t1, mu_t2) used: condim 3
  R hf_size[4];                     // 5 5 zscale base
};

void geomInertia(const Geom& g, R density, R* mass, R* Idiag) {
  R r = g.size[0], h = 2 * g.size[1];
  if (g.type == 0) {
    *mass = density * 4.0 / 3.0 * M_PI * r * r * r;
    Idiag[0] = Idiag[1] = Idiag[2] = 0.4 * (*mass) * r * r;
  } else if (g.type == 2) {
    *mass = density * M_PI * r * r * h;
    Idiag[0] = Idiag[1] = (*mass) * (3 * r * r + h * h) / 12.0;
    Idiag[2] = (*mass) * r * r / 2.0;
  } else {  // capsule: cylinder + two hemispheres
    R vc = M_PI * r * r * h, vs = 4.0 / 3.0 * M_PI * r * r * r;
    *mass = density * (vc + vs);
    R mc = density * vc, ms = density * vs;
    Idiag[0] = Idiag[1] = mc * (3 * r * r + h * h) / 12.0 + 0.4 * ms * r * r + ms * h * (3 * r + 2 * h) / 8.0;
    Idiag[2] = mc * r * r / 2.0 + 0.4 * ms * r * r;
  }
}
// accumulate a rigid piece (mass, com, inertia about its com in body frame) into a body
struct Piece { R m, c[3], I[9]; };
void pieceFromGeom(Piece& p, const Geom& g, R density) {
  R Id[3];
  geomInertia(g, density, &p.m, Id);
  v3cp(p.c, g.pos);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      R s = 0;
      for (int k = 0; k < 3; k++) s += g.mat[3 * i + k] * Id[k] * g.mat[3 * j + k];
      p.I[3 * i + j] = s;
    }
}
void pieceBox(Piece& p, const R* pos, R hx, R hy, R hz, R density) {
  p.m = density * 8 * hx * hy * hz;
  v3cp(p.c, pos);
  memset(p.I, 0, sizeof(p.I));
  p.I[0] = p.m / 3.0 * (hy * hy + hz * hz); p.I[4] = p.m / 3.0 * (hx * hx + hz * hz); p.I[8] = p.m / 3.0 * (hx * hx + hy * hy);
}
void bodyFromPieces(Model& m, int b, const std::vector<Piece>& ps) {
  R M = 0, c[3] = {0, 0, 0};
  for (auto& p : ps) { M += p.m; for (int k = 0; k < 3; k++) c[k] += p.m * p.c[k]; }
  for (int k = 0; k < 3; k++) c[k] /= M;
  R I[9] = {0};
  for (auto& p : ps) {
    R d[3] = {p.c[0] - c[0], p.c[1] - c[1], p.c[2] - c[2]};
    R dd = dot3(d, d);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) I[3 * i + j] += p.I[3 * i + j] + p.m * ((i == j ? dd : 0) - d[i] * d[j]);
  }
  m.mass[b] = M; v3cp(m.ipos[b], c); memcpy(m.inertia[b], I, sizeof(I));
}
void matFromZ(R* mat, const R* zdir) {  // any rotation whose third column is zdir (capsule symmetry)
  R z[3]; v3cp(z, zdir); normalize3(z);
  R t[3] = {1, 0, 0}; if (std::fabs(z[0]) > 0.9) v3set(t, 0, 1, 0);
  R x[3], y[3];
  cross3(y, z, t); normalize3(y); cross3(x, y, z);
  for (int i = 0; i < 3; i++) { mat[3 * i] = x[i]; mat[3 * i + 1] = y[i]; mat[3 * i + 2] = z[i]; }
}

struct Contact {
  R dist, pos[3], frame[9], friction[3], mu;
  int body1, body2, pair;  // pair: 0..2 wheel i, 3 hfield
  int efc_adr;
};
struct Data {
  R qpos[NQ], qvel[NV], qacc[NV], qacc_warmstart[NV], ctrl[3], time;
  R xpos[NB][3], xquat[NB][4], xmat[NB][9], xipos[NB][3], xanchor[NB][3], xaxis[NB][3];
  R subtree_com[NB][3], cinert[NB][10], cdof[NV][6], cdof_dot[NV][6], cvel[NB][6];
  R qM[NV * NV], qL[NV * NV];
  R qfrc_bias[NV], qfrc_passive[NV], qfrc_actuator[NV], qfrc_smooth[NV], qacc_smooth[NV];
  int ncon; Contact con[MAXCON];
  int nefc;
  R efc_J[MAXEFC * NV], efc_pos[MAXEFC], efc_D[MAXEFC], efc_R[MAXEFC], efc_aref[MAXEFC], efc_vel[MAXEFC];
  R efc_force[MAXEFC], efc_KBIP[MAXEFC][4], efc_diagApprox[MAXEFC];
  int efc_state[MAXEFC];  // 0 satisfied 1 quadratic 2 cone
  int solver_niter;
  std::vector<float> hfield;  // HN*HN
};

void kinematics(const Model& m, Data& d);
void comPos(const Model& m, Data& d);
void crb(const Model& m, Data& d);

void buildModel(Model& m) {
  memset(&m, 0, sizeof(m));
  m.timestep = 0.002; v3set(m.gravity, 0, 0, -9.81); m.tolerance = 1e-8; m.ls_tolerance = 0.01; m.impratio = 1;
  m.iterations = 100; m.ls_iterations = 50;
  m.solref[0] = 0.02; m.solref[1] = 1;
  R si[5] = {0.9, 0.95, 0.001, 0.5, 2}; memcpy(m.solimp, si, sizeof(si));
  // explicit pairs friction="0.001 1.0" (ballbot.xml:90-92); dynamic pairs: max of geom defaults (1, 0.005, 1e-4)
  v3set(m.fric_wheel, 0.001, 1.0, 0.005); v3set(m.fric_hfield, 1.0, 1.0, 0.005);
  m.hf_size[0] = 5; m.hf_size[1] = 5; m.hf_size[2] = 2.0; m.hf_size[3] = 0.1;
  int parent[NB] = {0, 0, 1, 1, 1, 1, 1, 0}, root[NB] = {0, 1, 1, 1, 1, 1, 1, 7};
  int jt[NB] = {JNONE, JFREE, JNONE, JNONE, JHINGE, JHINGE, JHINGE, JFREE};
  int qa[NB] = {-1, 0, -1, -1, 7, 8, 9, 10}, da[NB] = {-1, 0, -1, -1, 6, 7, 8, 9}, dn[NB] = {0, 6, 0, 0, 1, 1, 1, 6};
  for (int b = 0; b < NB; b++) {
    m.parent[b] = parent[b]; m.rootid[b] = root[b]; m.jtype[b] = jt[b]; m.qadr[b] = qa[b]; m.dadr[b] = da[b]; m.dofnum[b] = dn[b];
    m.bquat[b][0] = 1;
  }
  v3set(m.bpos[1], 0, 0, 0.24);
  v3set(m.bpos[2], 0.17, -0.01, -0.06); euler2quat(m.bquat[2], 180, -30, 0);
  v3set(m.bpos[3], -0.17, -0.01, -0.06); euler2quat(m.bquat[3], 180, 30, 0);
  for (int i = 0; i < 3; i++) {
    v3set(m.bpos[4 + i], 0, 0, -0.001); euler2quat(m.bquat[4 + i], 0, 0, 120.0 * i);
    v3set(m.jaxis[4 + i], -0.15316554764123935, -0.6903189805903613, -0.7071067953657663);
    normalize3(m.jaxis[4 + i]);
    v3set(m.janchor[4 + i], 0, 0, 0.0293);
  }
  v3set(m.bpos[7], 0, 0, 0.26);
  for (int k = 0; k < NV; k++) { m.armature[k] = 0; m.damping[k] = 0; }
  for (int k = 0; k < 6; k++) { m.dof_body[k] = 1; m.dof_parent[k] = k - 1; }
  for (int k = 6; k < 9; k++) { m.dof_body[k] = 4 + (k - 6); m.dof_parent[k] = 5; m.armature[k] = 0.005; m.damping[k] = 0.8; }
  for (int k = 9; k < 15; k++) { m.dof_body[k] = 7; m.dof_parent[k] = (k == 9 ? -1 : k - 1); }
  R q0[NQ] = {0, 0, 0.24, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.26, 1, 0, 0, 0}; memcpy(m.qpos0, q0, sizeof(q0));

  // geoms
  R eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  m.ball.type = 0; m.ball.body = 7; v3set(m.ball.pos, 0, 0, -0.14); memcpy(m.ball.mat, eye, sizeof(eye)); m.ball.size[0] = 0.09;
  m.tower.type = 2; m.tower.body = 1; v3set(m.tower.pos, 0, 0, 0.2); memcpy(m.tower.mat, eye, sizeof(eye)); m.tower.size[0] = 0.11; m.tower.size[1] = 0.14;
  for (int i = 0; i < 3; i++) {
    Geom& g = m.wheel[i]; g.type = 1; g.body = 4 + i; v3set(g.pos, -0.018, -0.08, -0.053);
    R q[4]; euler2quat(q, -45, 9, 0); quat2mat(g.mat, q); g.size[0] = 0.025; g.size[1] = 0.02;
  }
  for (int i = 0; i < 2; i++) {  // fromto capsules "0 0 0  -/+0.2 0 0", radius 0.01
    Geom& g = m.stick[i]; g.type = 1; g.body = 2 + i; R sx = (i == 0 ? -1.0 : 1.0);
    v3set(g.pos, 0.1 * sx, 0, 0); R z[3] = {sx, 0, 0}; matFromZ(g.mat, z); g.size[0] = 0.01; g.size[1] = 0.1;
  }
  // inertial frames from geoms (inertiafromgeom auto). cone meshes (cone.stl missing, density 1) omitted.
  Piece p; std::vector<Piece> ps;
  ps.clear(); pieceFromGeom(p, m.tower, 23.6); ps.push_back(p);
  { R bp[3] = {0, 0, 0.002}; pieceBox(p, bp, 0.1, 0.1, 0.1, 400.0); ps.push_back(p); }
  bodyFromPieces(m, 1, ps);
  for (int i = 0; i < 2; i++) { ps.clear(); pieceFromGeom(p, m.stick[i], 1000.0); ps.push_back(p); bodyFromPieces(m, 2 + i, ps); }
  for (int i = 0; i < 3; i++) { ps.clear(); pieceFromGeom(p, m.wheel[i], 620.0); ps.push_back(p); bodyFromPieces(m, 4 + i, ps); }
  ps.clear(); pieceFromGeom(p, m.ball, 55.0); ps.push_back(p); bodyFromPieces(m, 7, ps);

  // cameras: body frames 2,3 (euler 180 -/+30 0), camera euler 180 0 0 inside (ballbot.xml:44-54)
  for (int i = 0; i < 2; i++) {
    R qc[4], q[4]; euler2quat(qc, 180, 0, 0); mulQuat(q, m.bquat[2 + i], qc); quat2mat(m.cam_mat[i], q);
    v3cp(m.cam_pos[i], m.bpos[2 + i]);
  }

  // ---- set0: invweight0 and meaninertia at qpos0  [3P-memory: engine_setconst.c]
  Data* d = new Data();
  memset(d->qpos, 0, sizeof(d->qpos)); memcpy(d->qpos, m.qpos0, sizeof(m.qpos0));
  kinematics(m, *d); comPos(m, *d); crb(m, *d);
  R tr = 0; for (int i = 0; i < NV; i++) tr += d->qM[i * NV + i];
  m.meaninertia = tr / NV;
  memcpy(d->qL, d->qM, sizeof(d->qM)); cholFactor(d->qL, NV);
  for (int b = 1; b < NB; b++) {
    // jacobian of body com (6 x nv): rows 0-2 translational, 3-5 rotational
    R J[6][NV]; memset(J, 0, sizeof(J));
    int bb = b; while (m.dofnum[bb] == 0) bb = m.parent[bb];
    int dof = m.dadr[bb] + m.dofnum[bb] - 1;
    R off[3]; for (int k = 0; k < 3; k++) off[k] = d->xipos[b][k] - d->subtree_com[m.rootid[b]][k];
    for (; dof >= 0; dof = m.dof_parent[dof]) {
      R t[3]; cross3(t, d->cdof[dof], off);
      for (int k = 0; k < 3; k++) { J[k][dof] = d->cdof[dof][3 + k] + t[k]; J[3 + k][dof] = d->cdof[dof][k]; }
    }
    R A[6] = {0};
    for (int r = 0; r < 6; r++) { R x[NV]; cholSolve(x, d->qL, J[r], NV); for (int k = 0; k < NV; k++) A[r] += J[r][k] * x[k]; }
    m.invweight0[b][0] = std::fmax(MINVAL, (A[0] + A[1] + A[2]) / 3);
    m.invweight0[b][1] = std::fmax(MINVAL, (A[3] + A[4] + A[5]) / 3);
  }
  delete d;
}

// ----------------------------------------------------------------------------- position stage
void kinematics(const Model& m, Data& d) {  // [3P-memory: mj_kinematics]
  v3set(d.xpos[0], 0, 0, 0); d.xquat[0][0] = 1; d.xquat[0][1] = d.xquat[0][2] = d.xquat[0][3] = 0;
  quat2mat(d.xmat[0], d.xquat[0]);
  for (int b = 1; b < NB; b++) {
    int p = m.parent[b];
    if (m.jtype[b] == JFREE) {
      R* q = d.qpos + m.qadr[b];
      normalize4(q + 3);  // in-place, as mj_kinematics does
      v3cp(d.xpos[b], q); memcpy(d.xquat[b], q + 3, 4 * sizeof(R));
      v3cp(d.xanchor[b], d.xpos[b]);
    } else {
      R t[3]; mulMatVec3(t, d.xmat[p], m.bpos[b]);
      for (int k = 0; k < 3; k++) d.xpos[b][k] = d.xpos[p][k] + t[k];
      mulQuat(d.xquat[b], d.xquat[p], m.bquat[b]);
      if (m.jtype[b] == JHINGE) {
        R mat[9]; quat2mat(mat, d.xquat[b]);
        mulMatVec3(t, mat, m.janchor[b]);
        for (int k = 0; k < 3; k++) d.xanchor[b][k] = d.xpos[b][k] + t[k];
        mulMatVec3(d.xaxis[b], mat, m.jaxis[b]);
        R ql[4]; axisAngle2Quat(ql, m.jaxis[b], d.qpos[m.qadr[b]] - m.qpos0[m.qadr[b]]);
        mulQuat(d.xquat[b], d.xquat[b], ql);
        quat2mat(mat, d.xquat[b]); mulMatVec3(t, mat, m.janchor[b]);
        for (int k = 0; k < 3; k++) d.xpos[b][k] = d.xanchor[b][k] - t[k];
      }
    }
    normalize4(d.xquat[b]);
    quat2mat(d.xmat[b], d.xquat[b]);
    R t[3]; mulMatVec3(t, d.xmat[b], m.ipos[b]);
    for (int k = 0; k < 3; k++) d.xipos[b][k] = d.xpos[b][k] + t[k];
  }
}
void comPos(const Model& m, Data& d) {  // [3P-memory: mj_comPos]
  R smass[NB];
  for (int b = 0; b < NB; b++) { smass[b] = m.mass[b]; for (int k = 0; k < 3; k++) d.subtree_com[b][k] = m.mass[b] * d.xipos[b][k]; }
  for (int b = NB - 1; b > 0; b--) {
    int p = m.parent[b]; smass[p] += smass[b];
    for (int k = 0; k < 3; k++) d.subtree_com[p][k] += d.subtree_com[b][k];
  }
  for (int b = 0; b < NB; b++) {
    if (smass[b] < MINVAL) v3cp(d.subtree_com[b], d.xipos[b]);
    else for (int k = 0; k < 3; k++) d.subtree_com[b][k] /= smass[b];
  }
  for (int b = 1; b < NB; b++) {
    const R* com = d.subtree_com[m.rootid[b]];
    R dif[3] = {d.xipos[b][0] - com[0], d.xipos[b][1] - com[1], d.xipos[b][2] - com[2]};
    R tmp[9], I[9], mt[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) mt[3 * i + j] = d.xmat[b][3 * j + i];
    mulMat3(tmp, d.xmat[b], m.inertia[b]); mulMat3(I, tmp, mt);
    R ms = m.mass[b], dd = dot3(dif, dif);
    R* ci = d.cinert[b];
    ci[0] = I[0] + ms * (dd - dif[0] * dif[0]); ci[1] = I[4] + ms * (dd - dif[1] * dif[1]); ci[2] = I[8] + ms * (dd - dif[2] * dif[2]);
    ci[3] = I[1] - ms * dif[0] * dif[1]; ci[4] = I[2] - ms * dif[0] * dif[2]; ci[5] = I[5] - ms * dif[1] * dif[2];
    ci[6] = ms * dif[0]; ci[7] = ms * dif[1]; ci[8] = ms * dif[2]; ci[9] = ms;
    // cdof
    if (m.jtype[b] == JFREE) {
      int a = m.dadr[b];
      R off[3] = {com[0] - d.xpos[b][0], com[1] - d.xpos[b][1], com[2] - d.xpos[b][2]};
      for (int k = 0; k < 3; k++) {
        memset(d.cdof[a + k], 0, 6 * sizeof(R)); d.cdof[a + k][3 + k] = 1;
        R ax[3] = {d.xmat[b][k], d.xmat[b][3 + k], d.xmat[b][6 + k]};
        v3cp(d.cdof[a + 3 + k], ax); cross3(d.cdof[a + 3 + k] + 3, ax, off);
      }
    } else if (m.jtype[b] == JHINGE) {
      int a = m.dadr[b];
      R off[3] = {com[0] - d.xanchor[b][0], com[1] - d.xanchor[b][1], com[2] - d.xanchor[b][2]};
      v3cp(d.cdof[a], d.xaxis[b]); cross3(d.cdof[a] + 3, d.xaxis[b], off);
    }
  }
}
void crb(const Model& m, Data& d) {  // [3P-memory: mj_crb], dense symmetric qM
  R c[NB][10];
  for (int b = 0; b < NB; b++) for (int k = 0; k < 10; k++) c[b][k] = (b ? d.cinert[b][k] : 0);
  for (int b = NB - 1; b > 0; b--) if (m.parent[b] > 0) for (int k = 0; k < 10; k++) c[m.parent[b]][k] += c[b][k];
  memset(d.qM, 0, sizeof(d.qM));
  for (int i = 0; i < NV; i++) {
    R buf[6]; mulInertVec(buf, c[m.dof_body[i]], d.cdof[i]);
    d.qM[i * NV + i] = dot6(d.cdof[i], buf) + m.armature[i];
    for (int j = m.dof_parent[i]; j >= 0; j = m.dof_parent[j]) d.qM[i * NV + j] = d.qM[j * NV + i] = dot6(d.cdof[j], buf);
  }
}

// jacobian (3 x nv, translational) of a world point attached to body b  [3P-memory: mj_jac]
void jacPoint(const Model& m, const Data& d, int b, const R* point, R* jacp /*3*NV*/) {
  memset(jacp, 0, 3 * NV * sizeof(R));
  if (b == 0) return;
  int bb = b; while (bb && m.dofnum[bb] == 0) bb = m.parent[bb];
  if (!bb) return;
  const R* com = d.subtree_com[m.rootid[b]];
  R off[3] = {point[0] - com[0], point[1] - com[1], point[2] - com[2]};
  for (int dof = m.dadr[bb] + m.dofnum[bb] - 1; dof >= 0; dof = m.dof_parent[dof]) {
    R t[3]; cross3(t, d.cdof[dof], off);
    for (int k = 0; k < 3; k++) jacp[k * NV + dof] = d.cdof[dof][3 + k] + t[k];
  }
}

// ----------------------------------------------------------------------------- collision
void makeFrame(R* f) {  // [3P-memory: mju_makeFrame]
  normalize3(f);
  if (norm3(f + 3) < 0.5) { v3set(f + 3, 0, 0, 0); if (f[1] < 0.5 && f[1] > -0.5) f[4] = 1; else f[5] = 1; }
  R s = dot3(f, f + 3);
  for (int k = 0; k < 3; k++) f[3 + k] -= s * f[k];
  normalize3(f + 3);
  cross3(f + 6, f, f + 3);
}
// closest point on triangle abc to p (Ericson, Real-Time Collision Detection 5.1.5)
void closestPtTriangle(R* out, const R* p, const R* a, const R* b, const R* c) {
  R ab[3], ac[3], ap[3];
  for (int k = 0; k < 3; k++) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; ap[k] = p[k] - a[k]; }
  R d1 = dot3(ab, ap), d2 = dot3(ac, ap);
  if (d1 <= 0 && d2 <= 0) { v3cp(out, a); return; }
  R bp[3]; for (int k = 0; k < 3; k++) bp[k] = p[k] - b[k];
  R d3 = dot3(ab, bp), d4 = dot3(ac, bp);
  if (d3 >= 0 && d4 <= d3) { v3cp(out, b); return; }
  R vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) { R v = d1 / (d1 - d3); for (int k = 0; k < 3; k++) out[k] = a[k] + v * ab[k]; return; }
  R cp[3]; for (int k = 0; k < 3; k++) cp[k] = p[k] - c[k];
  R d5 = dot3(ab, cp), d6 = dot3(ac, cp);
  if (d6 >= 0 && d5 <= d6) { v3cp(out, c); return; }
  R vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { R w = d2 / (d2 - d6); for (int k = 0; k < 3; k++) out[k] = a[k] + w * ac[k]; return; }
  R va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    R w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    for (int k = 0; k < 3; k++) out[k] = b[k] + w * (c[k] - b[k]);
    return;
  }
  R den = 1.0 / (va + vb + vc), v = vb * den, w = vc * den;
  for (int k = 0; k < 3; k++) out[k] = a[k] + ab[k] * v + ac[k] * w;
}
// sphere (centre c, radius r) against the prism under top triangle (a,b,c3): exact closest feature of the
// continuous surface patch. Returns 1 and fills dist/normal/pos when penetrating (dist < 0).
int spherePrism(const R* c, R r, const R* a, const R* b, const R* c3, R* dist, R* nrm, R* pos) {
  // top-plane normal (pointing up)
  R e1[3], e2[3], n[3];
  for (int k = 0; k < 3; k++) { e1[k] = b[k] - a[k]; e2[k] = c3[k] - a[k]; }
  cross3(n, e1, e2); if (n[2] < 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
  normalize3(n);
  R q[3]; closestPtTriangle(q, c, a, b, c3);
  R dv[3] = {c[0] - q[0], c[1] - q[1], c[2] - q[2]};
  R ap[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
  R h = dot3(ap, n);  // signed height of the centre above the top plane
  R dl = norm3(dv);
  if (h < 0) {
    // centre below the top plane: only a contact if the centre is inside this prism's column
    R u = (e1[0] * e2[1] - e1[1] * e2[0]);
    R s = (ap[0] * e2[1] - ap[1] * e2[0]) / u, t = (e1[0] * ap[1] - e1[1] * ap[0]) / u;
    if (s < 0 || t < 0 || s + t > 1) return 0;
    *dist = h - r; v3cp(nrm, n);
    for (int k = 0; k < 3; k++) pos[k] = c[k] - n[k] * (r + 0.5 * (*dist));
    return 1;
  }
  if (dl >= r || dl < MINVAL) return 0;
  *dist = dl - r;
  for (int k = 0; k < 3; k++) nrm[k] = dv[k] / dl;
  for (int k = 0; k < 3; k++) pos[k] = q[k] + nrm[k] * 0.5 * (*dist);
  return 1;
}


// ---- additional geom pairs (ballbot.xml:41-69: every geom but `ballast` has contype = conaffinity = 1)
// closest points of segments p1-q1 and p2-q2 (Ericson, Real-Time Collision Detection 5.1.9)
R closestSegSeg(const R* p1, const R* q1, const R* p2, const R* q2, R* c1, R* c2) {
  R d1[3], d2[3], r[3];
  for (int k = 0; k < 3; k++) { d1[k] = q1[k] - p1[k]; d2[k] = q2[k] - p2[k]; r[k] = p1[k] - p2[k]; }
  R a = dot3(d1, d1), e = dot3(d2, d2), f = dot3(d2, r), s, t;
  if (a <= MINVAL && e <= MINVAL) { s = t = 0; }
  else if (a <= MINVAL) { s = 0; t = f / e; t = t < 0 ? 0 : (t > 1 ? 1 : t); }
  else {
    R c = dot3(d1, r);
    if (e <= MINVAL) { t = 0; s = -c / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
    else {
      R b = dot3(d1, d2), den = a * e - b * b;
      s = den > MINVAL ? (b * f - c * e) / den : 0; s = s < 0 ? 0 : (s > 1 ? 1 : s);
      t = (b * s + f) / e;
      if (t < 0) { t = 0; s = -c / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
      else if (t > 1) { t = 1; s = (b - c) / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
    }
  }
  for (int k = 0; k < 3; k++) { c1[k] = p1[k] + d1[k] * s; c2[k] = p2[k] + d2[k] * t; }
  R dv[3] = {c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]};
  return dot3(dv, dv);
}
// closest points between the segment p0-p1 and the triangle abc: ps on the segment, pt on the triangle; returns squared distance.
// Candidates in a fixed order (segment end points vs the triangle, then the segment vs the three edges); first minimum wins.
R closestSegTriangle(const R* p0, const R* p1, const R* a, const R* b, const R* c, R* ps, R* pt) {
  R best = 1e300, q[3], s1[3], s2[3];
  const R* ends[2] = {p0, p1};
  for (int i = 0; i < 2; i++) {
    closestPtTriangle(q, ends[i], a, b, c);
    R dv[3] = {ends[i][0] - q[0], ends[i][1] - q[1], ends[i][2] - q[2]}, d2 = dot3(dv, dv);
    if (d2 < best) { best = d2; v3cp(ps, ends[i]); v3cp(pt, q); }
  }
  const R* ed[3][2] = {{a, b}, {b, c}, {c, a}};
  for (int i = 0; i < 3; i++) {
    R d2 = closestSegSeg(p0, p1, ed[i][0], ed[i][1], s1, s2);
    if (d2 < best) { best = d2; v3cp(ps, s1); v3cp(pt, s2); }
  }
  return best;
}
// capsule (segment p0-p1, radius r) against the prism under the top triangle (a,b,c3); same conventions as spherePrism:
// exact closest feature of the top surface patch; a segment point under the top plane only counts inside the prism's column.
int capsulePrism(const R* p0, const R* p1, R r, const R* a, const R* b, const R* c3, R* dist, R* nrm, R* pos) {
  R e1[3], e2[3], n[3];
  for (int k = 0; k < 3; k++) { e1[k] = b[k] - a[k]; e2[k] = c3[k] - a[k]; }
  cross3(n, e1, e2); if (n[2] < 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
  normalize3(n);
  R ps[3], pt[3];
  closestSegTriangle(p0, p1, a, b, c3, ps, pt);
  // lowest segment point relative to the top plane
  R h0 = (p0[0] - a[0]) * n[0] + (p0[1] - a[1]) * n[1] + (p0[2] - a[2]) * n[2];
  R h1 = (p1[0] - a[0]) * n[0] + (p1[1] - a[1]) * n[1] + (p1[2] - a[2]) * n[2];
  const R* pl = h0 <= h1 ? p0 : p1; R hl = h0 <= h1 ? h0 : h1;
  if (hl < 0) {   // the segment dips under the top plane: deep contact at its lowest end, only inside this prism's column
    R ap[3] = {pl[0] - a[0], pl[1] - a[1], pl[2] - a[2]};
    R u = (e1[0] * e2[1] - e1[1] * e2[0]);
    R s = (ap[0] * e2[1] - ap[1] * e2[0]) / u, t = (e1[0] * ap[1] - e1[1] * ap[0]) / u;
    if (!(s < 0 || t < 0 || s + t > 1)) {
      *dist = hl - r; v3cp(nrm, n);
      for (int k = 0; k < 3; k++) pos[k] = pl[k] - n[k] * (r + 0.5 * (*dist));
      return 1;
    }
    R hs = (ps[0] - a[0]) * n[0] + (ps[1] - a[1]) * n[1] + (ps[2] - a[2]) * n[2];
    if (hs < 0) return 0;   // closest feature reached from below the plane: handled by the neighbouring prism
  }
  R dv[3] = {ps[0] - pt[0], ps[1] - pt[1], ps[2] - pt[2]}, dl = norm3(dv);
  if (dl >= r || dl < MINVAL) return 0;
  *dist = dl - r;
  for (int k = 0; k < 3; k++) nrm[k] = dv[k] / dl;
  for (int k = 0; k < 3; k++) pos[k] = pt[k] + nrm[k] * 0.5 * (*dist);
  return 1;
}
void geomWorld(const Data& d, const Geom& g, R* c, R* axis) {
  R t[3], gm[9];
  mulMatVec3(t, d.xmat[g.body], g.pos); for (int k = 0; k < 3; k++) c[k] = d.xpos[g.body][k] + t[k];
  mulMat3(gm, d.xmat[g.body], g.mat); v3set(axis, gm[2], gm[5], gm[8]);
}
// capsule geom vs the heightfield: sub-grid of the capsule's AABB, prisms in the scan order of the ball pair
void capsuleHfield(const Model& m, Data& d, const Geom& g, int pairId) {
  const R sx = m.hf_size[0], sy = m.hf_size[1], sz = m.hf_size[2], sb = m.hf_size[3];
  const int nrow = HN, ncol = HN; const float* data = d.hfield.data();
  R cc[3], ax[3]; geomWorld(d, g, cc, ax);
  const R rad = g.size[0], len = g.size[1];
  R p0[3], p1[3]; for (int k = 0; k < 3; k++) { p0[k] = cc[k] - ax[k] * len; p1[k] = cc[k] + ax[k] * len; }
  R lo[3], hi[3]; for (int k = 0; k < 3; k++) { lo[k] = std::fmin(p0[k], p1[k]) - rad; hi[k] = std::fmax(p0[k], p1[k]) + rad; }
  if (sx < lo[0] || -sx > hi[0] || sy < lo[1] || -sy > hi[1] || sz < lo[2] || -sb > hi[2]) return;
  int cmin = (int)std::floor((lo[0] + sx) / (2 * sx) * (ncol - 1)), cmax = (int)std::ceil((hi[0] + sx) / (2 * sx) * (ncol - 1));
  int rmin = (int)std::floor((lo[1] + sy) / (2 * sy) * (nrow - 1)), rmax = (int)std::ceil((hi[1] + sy) / (2 * sy) * (nrow - 1));
  if (cmin < 0) cmin = 0; if (rmin < 0) rmin = 0; if (cmax > ncol - 1) cmax = ncol - 1; if (rmax > nrow - 1) rmax = nrow - 1;
  const R dx = 2 * sx / (ncol - 1), dy = 2 * sy / (nrow - 1), zmin = lo[2];
  int cnt = 0; const int dr[2] = {1, 0};
  for (int r = rmin; r < rmax && cnt < 50; r++) {
    R v[3][3] = {{0}}; int nvert = 0;
    for (int c = cmin; c <= cmax && cnt < 50; c++)
      for (int k = 0; k < 2 && cnt < 50; k++) {
        v3cp(v[0], v[1]); v3cp(v[1], v[2]);
        v3set(v[2], dx * c - sx, dy * (r + dr[k]) - sy, (R)data[(r + dr[k]) * ncol + c] * sz);
        if (++nvert < 3) continue;
        if (v[0][2] < zmin && v[1][2] < zmin && v[2][2] < zmin) continue;
        R dist, nrm[3], pos[3];
        if (!capsulePrism(p0, p1, rad, v[0], v[1], v[2], &dist, nrm, pos)) continue;
        if (d.ncon >= MAXCON) continue;
        Contact& cn = d.con[d.ncon++];
        memset(&cn, 0, sizeof(cn));
        cn.dist = dist; v3cp(cn.pos, pos); v3cp(cn.frame, nrm); makeFrame(cn.frame);
        cn.body1 = 0; cn.body2 = g.body; cn.pair = pairId; v3cp(cn.friction, m.fric_hfield);
        cnt++;
      }
  }
}

int g_extra_pairs = 1;
void collision(const Model& m, Data& d) {
  d.ncon = 0;
  // ball sphere world pose
  R bc[3], t[3];
  mulMatVec3(t, d.xmat[7], m.ball.pos); for (int k = 0; k < 3; k++) bc[k] = d.xpos[7][k] + t[k];
  R br = m.ball.size[0];
  // --- explicit pairs the_ball x wheel_mesh_i: patched mjraw_SphereCapsule (mujoco_fix.patch:9-18)
  for (int i = 0; i < 3; i++) {
    const Geom& g = m.wheel[i]; int wb = g.body;
    R cc[3], gm[9], axis[3];
    mulMatVec3(t, d.xmat[wb], g.pos); for (int k = 0; k < 3; k++) cc[k] = d.xpos[wb][k] + t[k];
    mulMat3(gm, d.xmat[wb], g.mat); v3set(axis, gm[2], gm[5], gm[8]);
    R vec[3] = {bc[0] - cc[0], bc[1] - cc[1], bc[2] - cc[2]};
    R x = dot3(axis, vec), len = g.size[1];
    if (x > len) x = len; if (x < -len) x = -len;
    R np[3] = {cc[0] + axis[0] * x, cc[1] + axis[1] * x, cc[2] + axis[2] * x};
    R dif[3] = {np[0] - bc[0], np[1] - bc[1], np[2] - bc[2]};
    R cd = norm3(dif), mind = br + g.size[0];
    if (cd > mind) continue;  // margin 0
    Contact& c = d.con[d.ncon++];
    memset(&c, 0, sizeof(c));
    for (int k = 0; k < 3; k++) c.frame[k] = dif[k] / cd;
    c.dist = cd - mind;
    for (int k = 0; k < 3; k++) c.pos[k] = bc[k] + c.frame[k] * (br + 0.5 * c.dist);
    v3cp(c.frame + 3, axis);  // the patch
    makeFrame(c.frame);
    c.body1 = 7; c.body2 = wb; c.pair = i;
    v3cp(c.friction, m.fric_wheel);
  }
  // --- the_ball x terrain hfield: mjc_ConvexHField sub-grid of triangular prisms [3P-memory]
  {
    const R sx = m.hf_size[0], sy = m.hf_size[1], sz = m.hf_size[2], sb = m.hf_size[3];
    const int nrow = HN, ncol = HN;
    const float* data = d.hfield.data();
    bool skip = false;
    if (sx < bc[0] - br || -sx > bc[0] + br || sy < bc[1] - br || -sy > bc[1] + br) skip = true;
    if (sz < bc[2] - br || -sb > bc[2] + br) skip = true;
    if (!skip) {
      R xmin = bc[0] - br, xmax = bc[0] + br, ymin = bc[1] - br, ymax = bc[1] + br, zmin = bc[2] - br;
      int cmin = (int)std::floor((xmin + sx) / (2 * sx) * (ncol - 1)), cmax = (int)std::ceil((xmax + sx) / (2 * sx) * (ncol - 1));
      int rmin = (int)std::floor((ymin + sy) / (2 * sy) * (nrow - 1)), rmax = (int)std::ceil((ymax + sy) / (2 * sy) * (nrow - 1));
      if (cmin < 0) cmin = 0; if (rmin < 0) rmin = 0; if (cmax > ncol - 1) cmax = ncol - 1; if (rmax > nrow - 1) rmax = nrow - 1;
      R dx = 2 * sx / (ncol - 1), dy = 2 * sy / (nrow - 1);
      int cnt = 0; const int dr[2] = {1, 0};
      for (int r = rmin; r < rmax && cnt < 50; r++) {
        R v[3][3] = {{0}}; int nvert = 0;
        for (int c = cmin; c <= cmax && cnt < 50; c++)
          for (int k = 0; k < 2 && cnt < 50; k++) {
            // shift and append vertex (triangle strip)
            v3cp(v[0], v[1]); v3cp(v[1], v[2]);
            v3set(v[2], dx * c - sx, dy * (r + dr[k]) - sy, (R)data[(r + dr[k]) * ncol + c] * sz);
            if (++nvert < 3) continue;
            if (v[0][2] < zmin && v[1][2] < zmin && v[2][2] < zmin) continue;
            R dist, nrm[3], pos[3];
            if (!spherePrism(bc, br, v[0], v[1], v[2], &dist, nrm, pos)) continue;
            if (d.ncon >= MAXCON) continue;
            Contact& cn = d.con[d.ncon++];
            memset(&cn, 0, sizeof(cn));
            cn.dist = dist; v3cp(cn.pos, pos); v3cp(cn.frame, nrm); makeFrame(cn.frame);
            cn.body1 = 0; cn.body2 = 7; cn.pair = 3; v3cp(cn.friction, m.fric_hfield);
            cnt++;
          }
      }
    }
  }
  if (!g_extra_pairs) return;
  // --- dynamic pairs of the other colliding geoms (friction = element-wise max of the geom defaults, condim 3):
  // world heightfield x {wheel capsules, camera-stick capsules}; ball x {camera sticks (patched sphere-capsule), tower cylinder}.
  // Filtered by MuJoCo: same weld body (base/cam geoms among themselves), parent-child (base/cam x wheels). Wheel x wheel
  // capsules are ~0.09 m apart for every hinge angle; tower x heightfield and the cone meshes are not modelled (ORACLE_ASSUMPTIONS).
  for (int i = 0; i < 2; i++) capsuleHfield(m, d, m.stick[i], 5 + i);
  for (int i = 0; i < 3; i++) capsuleHfield(m, d, m.wheel[i], 7 + i);
  for (int i = 0; i < 2; i++) {   // ball x stick: mjraw_SphereCapsule with the patched frame
    const Geom& g = m.stick[i]; R cc[3], axis[3]; geomWorld(d, g, cc, axis);
    R vec[3] = {bc[0] - cc[0], bc[1] - cc[1], bc[2] - cc[2]};
    R x = dot3(axis, vec), len = g.size[1];
    if (x > len) x = len; if (x < -len) x = -len;
    R dif[3] = {cc[0] + axis[0] * x - bc[0], cc[1] + axis[1] * x - bc[1], cc[2] + axis[2] * x - bc[2]};
    R cd = norm3(dif), mind = br + g.size[0];
    if (cd > mind || d.ncon >= MAXCON) continue;
    Contact& c = d.con[d.ncon++];
    memset(&c, 0, sizeof(c));
    for (int k = 0; k < 3; k++) c.frame[k] = dif[k] / cd;
    c.dist = cd - mind;
    for (int k = 0; k < 3; k++) c.pos[k] = bc[k] + c.frame[k] * (br + 0.5 * c.dist);
    v3cp(c.frame + 3, axis); makeFrame(c.frame);
    c.body1 = 7; c.body2 = g.body; c.pair = 11 + i; v3cp(c.friction, m.fric_hfield);
  }
  {   // ball x tower: mjraw_SphereCylinder (side / cap / corner)
    const Geom& g = m.tower; R cc[3], axis[3]; geomWorld(d, g, cc, axis);
    R vec[3] = {bc[0] - cc[0], bc[1] - cc[1], bc[2] - cc[2]};
    R x = dot3(vec, axis), pp[3] = {vec[0] - axis[0] * x, vec[1] - axis[1] * x, vec[2] - axis[2] * x}, pp2 = dot3(pp, pp);
    R rad = g.size[0], hgt = g.size[1];
    bool side = std::fabs(x) < hgt, cap = pp2 < rad * rad;
    if (side && cap) { if (hgt - std::fabs(x) < rad - std::sqrt(pp2)) side = false; else cap = false; }
    R tgt[3], trad; bool hit = false; R nrm[3], dist = 0;
    if (side) { for (int k = 0; k < 3; k++) tgt[k] = cc[k] + axis[k] * x; trad = rad; }
    else if (cap) {   // plane of the nearer cap against the sphere
      R sg = x > 0 ? 1 : -1, pc[3]; for (int k = 0; k < 3; k++) pc[k] = cc[k] + sg * hgt * axis[k];
      R dd = sg * ((bc[0] - pc[0]) * axis[0] + (bc[1] - pc[1]) * axis[1] + (bc[2] - pc[2]) * axis[2]) - br;
      if (dd < 0 && d.ncon < MAXCON) { hit = true; dist = dd; for (int k = 0; k < 3; k++) nrm[k] = -sg * axis[k]; }
      trad = -1;
    } else {
      R pn = std::sqrt(pp2), sg = x > 0 ? 1 : -1;
      for (int k = 0; k < 3; k++) tgt[k] = cc[k] + pp[k] / pn * rad + sg * hgt * axis[k];
      trad = 0;
    }
    if (trad >= 0) {
      R dif[3] = {tgt[0] - bc[0], tgt[1] - bc[1], tgt[2] - bc[2]}, cd = norm3(dif);
      if (cd < br + trad && cd > MINVAL && d.ncon < MAXCON) { hit = true; dist = cd - br - trad; for (int k = 0; k < 3; k++) nrm[k] = dif[k] / cd; }
    }
    if (hit) {
      Contact& c = d.con[d.ncon++];
      memset(&c, 0, sizeof(c));
      c.dist = dist; v3cp(c.frame, nrm);
      for (int k = 0; k < 3; k++) c.pos[k] = bc[k] + nrm[k] * (br + 0.5 * dist);
      makeFrame(c.frame);
      c.body1 = 7; c.body2 = g.body; c.pair = 10; v3cp(c.friction, m.fric_hfield);
    }
  }
}

// ----------------------------------------------------------------------------- constraints
void makeConstraint(const Model& m, Data& d) {  // [3P-memory: mj_makeConstraint, mj_makeImpedance]
  d.nefc = 0;
  for (int ci = 0; ci < d.ncon; ci++) {
    Contact& c = d.con[ci];
    c.efc_adr = -1;
    if (c.dist >= 0) continue;  // includemargin = 0
    c.efc_adr = d.nefc;
    R j1[3 * NV], j2[3 * NV];
    jacPoint(m, d, c.body1, c.pos, j1); jacPoint(m, d, c.body2, c.pos, j2);
    R tran = m.invweight0[c.body1][0] + m.invweight0[c.body2][0];
    for (int r = 0; r < 3; r++) {
      int e = d.nefc++;
      for (int k = 0; k < NV; k++) {
        R s = 0; for (int a = 0; a < 3; a++) s += c.frame[3 * r + a] * (j2[a * NV + k] - j1[a * NV + k]);
        d.efc_J[e * NV + k] = s;
      }
      d.efc_pos[e] = (r == 0 ? c.dist : 0);
      d.efc_diagApprox[e] = tran;
    }
  }
  // impedance, R, KBIP
  for (int e = 0; e < d.nefc; e++) {
    R pos = d.efc_pos[e];
    R tc = std::fmax(m.solref[0], 2 * m.timestep), dr = m.solref[1];
    R d0 = m.solimp[0], dmax = m.solimp[1], width = m.solimp[2], mid = m.solimp[3], power = m.solimp[4];
    R x = std::fabs(pos) / width, imp, impP = 0;
    if (x >= 1) imp = dmax; else if (x <= 0) imp = d0;
    else {
      R y;
      if (power == 1) y = x;
      else if (x <= mid) y = std::pow(x, power) / std::pow(mid, power - 1);
      else y = 1 - std::pow(1 - x, power) / std::pow(1 - mid, power - 1);
      imp = d0 + y * (dmax - d0);
    }
    d.efc_R[e] = std::fmax(MINVAL, (1 - imp) * d.efc_diagApprox[e] / imp);
    d.efc_KBIP[e][0] = 1 / (dmax * dmax * tc * tc * dr * dr);
    d.efc_KBIP[e][1] = 2 / (dmax * tc);
    d.efc_KBIP[e][2] = imp; d.efc_KBIP[e][3] = impP;
  }
  // elliptic friction rows: R_t1 = R_n/impratio, R_tj = R_t1 mu1^2/muj^2, regularised cone mu
  for (int ci = 0; ci < d.ncon; ci++) {
    Contact& c = d.con[ci]; int i = c.efc_adr; if (i < 0) continue;
    d.efc_R[i + 1] = d.efc_R[i] / std::fmax(MINVAL, m.impratio);
    d.efc_R[i + 2] = d.efc_R[i + 1] * c.friction[0] * c.friction[0] / (c.friction[1] * c.friction[1]);
    c.mu = c.friction[0] * std::sqrt(d.efc_R[i + 1] / d.efc_R[i]);
  }
  for (int e = 0; e < d.nefc; e++) d.efc_D[e] = 1 / d.efc_R[e];
}
void referenceConstraint(Data& d) {  // aref = -B*vel - K*imp*(pos-margin)
  for (int e = 0; e < d.nefc; e++) {
    R v = 0; for (int k = 0; k < NV; k++) v += d.efc_J[e * NV + k] * d.qvel[k];
    d.efc_vel[e] = v;
    d.efc_aref[e] = -d.efc_KBIP[e][1] * v - d.efc_KBIP[e][0] * d.efc_KBIP[e][2] * d.efc_pos[e];
  }
}

// ----------------------------------------------------------------------------- velocity stage
void comVel(const Model& m, Data& d) {  // [3P-memory: mj_comVel]
  memset(d.cvel[0], 0, 6 * sizeof(R));
  for (int b = 1; b < NB; b++) {
    R cv[6]; memcpy(cv, d.cvel[m.parent[b]], sizeof(cv));
    int a = m.dadr[b];
    if (m.jtype[b] == JFREE) {
      for (int j = 0; j < 3; j++) { memset(d.cdof_dot[a + j], 0, 6 * sizeof(R)); for (int k = 0; k < 6; k++) cv[k] += d.cdof[a + j][k] * d.qvel[a + j]; }
      for (int j = 3; j < 6; j++) crossMotion(d.cdof_dot[a + j], cv, d.cdof[a + j]);
      for (int j = 3; j < 6; j++) for (int k = 0; k < 6; k++) cv[k] += d.cdof[a + j][k] * d.qvel[a + j];
    } else if (m.jtype[b] == JHINGE) {
      crossMotion(d.cdof_dot[a], cv, d.cdof[a]);
      for (int k = 0; k < 6; k++) cv[k] += d.cdof[a][k] * d.qvel[a];
    }
    memcpy(d.cvel[b], cv, sizeof(cv));
  }
}
void rneBias(const Model& m, Data& d) {  // [3P-memory: mj_rne(flg_acc=0)]
  R cacc[NB][6], cfrc[NB][6];
  memset(cacc, 0, sizeof(cacc)); memset(cfrc, 0, sizeof(cfrc));
  for (int k = 0; k < 3; k++) cacc[0][3 + k] = -m.gravity[k];
  for (int b = 1; b < NB; b++) {
    memcpy(cacc[b], cacc[m.parent[b]], 6 * sizeof(R));
    for (int j = 0; j < m.dofnum[b]; j++) { int a = m.dadr[b] + j; for (int k = 0; k < 6; k++) cacc[b][k] += d.cdof_dot[a][k] * d.qvel[a]; }
    R t[6], t1[6];
    mulInertVec(cfrc[b], d.cinert[b], cacc[b]);
    mulInertVec(t, d.cinert[b], d.cvel[b]); crossForce(t1, d.cvel[b], t);
    for (int k = 0; k < 6; k++) cfrc[b][k] += t1[k];
  }
  for (int b = NB - 1; b > 0; b--) if (m.parent[b] > 0) for (int k = 0; k < 6; k++) cfrc[m.parent[b]][k] += cfrc[b][k];
  for (int i = 0; i < NV; i++) d.qfrc_bias[i] = dot6(d.cdof[i], cfrc[m.dof_body[i]]);
}

// ----------------------------------------------------------------------------- Newton solver
struct Solver {
  const Model& m; Data& d; int nefc;
  R Ma[NV], jar[MAXEFC], grad[NV], Mgrad[NV], search[NV], Mv[NV], jv[MAXEFC];
  R H[NV * NV];
  R cost, gauss;
  R quad[MAXEFC][3], quadGauss[3];
  Solver(const Model& mm, Data& dd) : m(mm), d(dd), nefc(dd.nefc) {}

  void mulM(R* r, const R* v) { for (int i = 0; i < NV; i++) { R s = 0; for (int k = 0; k < NV; k++) s += d.qM[i * NV + k] * v[k]; r[i] = s; } }
  void mulJ(R* r, const R* v) { for (int e = 0; e < nefc; e++) { R s = 0; for (int k = 0; k < NV; k++) s += d.efc_J[e * NV + k] * v[k]; r[e] = s; } }

  // [3P-memory: mj_constraintUpdate, elliptic cones]; fills force/state, returns constraint cost; optional cone Hessians
  R constraintUpdate(const R* jr, R* force, int* state, R (*hcone)[9]) {
    R cost = 0;
    for (int ci = 0; ci < d.ncon; ci++) {
      const Contact& c = d.con[ci]; int i = c.efc_adr; if (i < 0) continue;
      R mu = c.mu, f1 = c.friction[0], f2 = c.friction[1];
      R U0 = jr[i] * mu, U1 = jr[i + 1] * f1, U2 = jr[i + 2] * f2;
      R N = U0, T = std::sqrt(U1 * U1 + U2 * U2);
      if (N >= mu * T || (T <= 0 && N >= 0)) {
        force[i] = force[i + 1] = force[i + 2] = 0; state[i] = state[i + 1] = state[i + 2] = 0;
      } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
        for (int j = 0; j < 3; j++) { force[i + j] = -d.efc_D[i + j] * jr[i + j]; cost += 0.5 * d.efc_D[i + j] * jr[i + j] * jr[i + j]; state[i + j] = 1; }
      } else {
        R Dm = d.efc_D[i] / (mu * mu * (1 + mu * mu));
        R NT = N - mu * T;
        cost += 0.5 * Dm * NT * NT;
        force[i] = -Dm * NT * mu;
        force[i + 1] = -force[i] / T * f1 * U1; force[i + 2] = -force[i] / T * f2 * U2;
        state[i] = state[i + 1] = state[i + 2] = 2;
        if (hcone) {
          R U[3] = {U0, U1, U2}, sc[3] = {mu, f1, f2}; R* h = hcone[ci];
          h[0] = 1;
          for (int j = 1; j < 3; j++) h[j] = h[3 * j] = -mu * U[j] / T;
          for (int k = 1; k < 3; k++) for (int j = 1; j < 3; j++) h[3 * k + j] = mu * N / (T * T * T) * U[k] * U[j];
          for (int j = 1; j < 3; j++) h[3 * j + j] += mu * mu - mu * N / T;
          for (int k = 0; k < 3; k++) for (int j = 0; j < 3; j++) h[3 * k + j] *= Dm * sc[k] * sc[j];
        }
      }
    }
    return cost;
  }
  R totalCost(const R* qa, R* force, int* state) {  // Gauss + constraint at arbitrary qacc (warmstart test)
    R ma[NV], jr[MAXEFC];
    mulM(ma, qa); mulJ(jr, qa); for (int e = 0; e < nefc; e++) jr[e] -= d.efc_aref[e];
    R g = 0; for (int i = 0; i < NV; i++) g += 0.5 * (ma[i] - d.qfrc_smooth[i]) * (qa[i] - d.qacc_smooth[i]);
    return g + constraintUpdate(jr, force, state, nullptr);
  }
  void update(bool first) {
    static thread_local R hcone[MAXCON][9];
    R cc = constraintUpdate(jar, d.efc_force, d.efc_state, hcone);
    gauss = 0; for (int i = 0; i < NV; i++) gauss += 0.5 * (Ma[i] - d.qfrc_smooth[i]) * (d.qacc[i] - d.qacc_smooth[i]);
    cost = gauss + cc;
    // Hessian H = M + J' diag(D quad) J + cone blocks, dense Cholesky
    memcpy(H, d.qM, sizeof(H));
    for (int e = 0; e < nefc; e++) if (d.efc_state[e] == 1) {
      const R* J = d.efc_J + e * NV; R D = d.efc_D[e];
      for (int a = 0; a < NV; a++) { if (J[a] == 0) continue; for (int b = 0; b < NV; b++) H[a * NV + b] += D * J[a] * J[b]; }
    }
    for (int ci = 0; ci < d.ncon; ci++) {
      int i = d.con[ci].efc_adr; if (i < 0 || d.efc_state[i] != 2) continue;
      for (int k = 0; k < 3; k++) for (int j = 0; j < 3; j++) {
        R h = hcone[ci][3 * k + j]; const R* Jk = d.efc_J + (i + k) * NV; const R* Jj = d.efc_J + (i + j) * NV;
        for (int a = 0; a < NV; a++) { if (Jk[a] == 0) continue; for (int b = 0; b < NV; b++) H[a * NV + b] += h * Jk[a] * Jj[b]; }
      }
    }
    cholFactor(H, NV);
    for (int i = 0; i < NV; i++) { R s = Ma[i] - d.qfrc_smooth[i]; for (int e = 0; e < nefc; e++) s -= d.efc_J[e * NV + i] * d.efc_force[e]; grad[i] = s; }
    cholSolve(Mgrad, H, grad, NV);
    (void)first;
  }
  struct Pt { R alpha, cost, d1, d2; };
  // [3P-memory: PrimalPrepare]
  void prepare() {
    mulM(Mv, search); mulJ(jv, search);
    quadGauss[0] = gauss; quadGauss[1] = 0; quadGauss[2] = 0;
    for (int i = 0; i < NV; i++) { quadGauss[1] += search[i] * (Ma[i] - d.qfrc_smooth[i]); quadGauss[2] += 0.5 * search[i] * Mv[i]; }
    for (int e = 0; e < nefc; e++) {
      quad[e][0] = 0.5 * d.efc_D[e] * jar[e] * jar[e]; quad[e][1] = d.efc_D[e] * jar[e] * jv[e]; quad[e][2] = 0.5 * d.efc_D[e] * jv[e] * jv[e];
    }
  }
  // [3P-memory: PrimalEval]
  Pt eval(R alpha) {
    Pt p; p.alpha = alpha;
    p.cost = quadGauss[0] + alpha * quadGauss[1] + alpha * alpha * quadGauss[2];
    p.d1 = quadGauss[1] + 2 * alpha * quadGauss[2]; p.d2 = 2 * quadGauss[2];
    for (int ci = 0; ci < d.ncon; ci++) {
      const Contact& c = d.con[ci]; int i = c.efc_adr; if (i < 0) continue;
      R mu = c.mu, f1 = c.friction[0], f2 = c.friction[1];
      R U0 = jar[i] * mu, V0 = jv[i] * mu;
      R u1 = jar[i + 1] * f1, u2 = jar[i + 2] * f2, v1 = jv[i + 1] * f1, v2 = jv[i + 2] * f2;
      R UU = u1 * u1 + u2 * u2, UV = u1 * v1 + u2 * v2, VV = v1 * v1 + v2 * v2;
      R Dm = d.efc_D[i] / (mu * mu * (1 + mu * mu));
      R q0 = quad[i][0] + quad[i + 1][0] + quad[i + 2][0], q1 = quad[i][1] + quad[i + 1][1] + quad[i + 2][1], q2 = quad[i][2] + quad[i + 1][2] + quad[i + 2][2];
      R N = U0 + alpha * V0, Tsqr = UU + alpha * (2 * UV + alpha * VV);
      bool bottom = false;
      if (Tsqr <= 0) { if (N < 0) bottom = true; }
      else {
        R T = std::sqrt(Tsqr);
        if (N >= mu * T) {}
        else if (mu * N + T <= 0) bottom = true;
        else {
          R N1 = V0, T1 = (UV + alpha * VV) / T, T2 = VV / T - (UV + alpha * VV) * T1 / (T * T);
          R NT = N - mu * T;
          p.cost += 0.5 * Dm * NT * NT;
          p.d1 += Dm * NT * (N1 - mu * T1);
          p.d2 += Dm * ((N1 - mu * T1) * (N1 - mu * T1) + NT * (-mu * T2));
        }
      }
      if (bottom) { p.cost += q0 + alpha * q1 + alpha * alpha * q2; p.d1 += q1 + 2 * alpha * q2; p.d2 += 2 * q2; }
    }
    if (p.d2 < MINVAL) p.d2 = MINVAL;
    return p;
  }
  int updateBracket(Pt& p, const Pt* cand, Pt& pnext) {
    int flag = 0;
    for (int i = 0; i < 3; i++) {
      if (p.d1 < 0 && cand[i].d1 < 0 && p.d1 < cand[i].d1) { p = cand[i]; flag = 1; }
      else if (p.d1 > 0 && cand[i].d1 > 0 && p.d1 > cand[i].d1) { p = cand[i]; flag = 2; }
    }
    if (flag) pnext = eval(p.alpha - p.d1 / p.d2);
    return flag;
  }
  // [3P-memory: PrimalSearch], exact 1-D Newton line search with bracketing
  R lineSearch(R scale) {
    R snorm = 0; for (int i = 0; i < NV; i++) snorm += search[i] * search[i]; snorm = std::sqrt(snorm);
    if (snorm < MINVAL) return 0;
    R gtol = m.tolerance * m.ls_tolerance * snorm / scale;
    prepare();
    int it = 0;
    Pt p0 = eval(0), p1 = eval(p0.alpha - p0.d1 / p0.d2), p2 = p0, pmid, p1next, p2next;
    if (p0.cost < p1.cost) p1 = p0;
    if (std::fabs(p1.d1) < gtol) return p1.alpha;
    R dir = (p1.d1 < 0 ? 1 : -1);
    bool p2update = false;
    while (p1.d1 * dir <= -gtol && it < m.ls_iterations) {
      p2 = p1; p2update = true;
      p1 = eval(p1.alpha - p1.d1 / p1.d2); it++;
      if (std::fabs(p1.d1) < gtol) return p1.alpha;
    }
    if (it >= m.ls_iterations) return p1.alpha;
    if (!p2update) return p1.alpha;
    p2next = p1; p1next = eval(p1.alpha - p1.d1 / p1.d2);
    while (it < m.ls_iterations) {
      pmid = eval(0.5 * (p1.alpha + p2.alpha)); it++;
      Pt cand[3] = {p1next, p2next, pmid};
      for (int i = 0; i < 3; i++) if (std::fabs(cand[i].d1) < gtol) return cand[i].alpha;
      int b1 = updateBracket(p1, cand, p1next), b2 = updateBracket(p2, cand, p2next);
      if (!b1 && !b2) return (pmid.cost < p0.cost) ? pmid.alpha : 0;
    }
    if (p1.cost <= p2.cost && p1.cost < p0.cost) return p1.alpha;
    if (p2.cost <= p1.cost && p2.cost < p0.cost) return p2.alpha;
    return 0;
  }
  void run() {  // [3P-memory: mj_solNewton via mj_solPrimal]
    // warmstart choice (mj_fwdConstraint/warmstart)
    {
      static thread_local R f[MAXEFC]; static thread_local int s[MAXEFC];
      R cw = totalCost(d.qacc_warmstart, f, s), cs = totalCost(d.qacc_smooth, f, s);
      memcpy(d.qacc, cw > cs ? d.qacc_smooth : d.qacc_warmstart, NV * sizeof(R));
    }
    mulM(Ma, d.qacc); mulJ(jar, d.qacc); for (int e = 0; e < nefc; e++) jar[e] -= d.efc_aref[e];
    update(true);
    for (int i = 0; i < NV; i++) search[i] = -Mgrad[i];
    R scale = 1 / (m.meaninertia * NV);
    int iter = 0;
    while (iter < m.iterations) {
      R alpha = lineSearch(scale);
      if (alpha == 0) break;
      for (int i = 0; i < NV; i++) { d.qacc[i] += alpha * search[i]; Ma[i] += alpha * Mv[i]; }
      for (int e = 0; e < nefc; e++) jar[e] += alpha * jv[e];
      R old = cost;
      update(false);
      R improvement = scale * (old - cost), gn = 0;
      for (int i = 0; i < NV; i++) gn += grad[i] * grad[i];
      R gradient = scale * std::sqrt(gn);
      iter++;
      if (improvement < m.tolerance || gradient < m.tolerance) break;
      for (int i = 0; i < NV; i++) search[i] = -Mgrad[i];
    }
    d.solver_niter = iter;
  }
};

// ----------------------------------------------------------------------------- forward / step
void forward(const Model& m, Data& d) {  // [3P-memory: mj_forward]
  kinematics(m, d); comPos(m, d); crb(m, d);
  memcpy(d.qL, d.qM, sizeof(d.qM)); cholFactor(d.qL, NV);
  collision(m, d); makeConstraint(m, d);
  comVel(m, d);
  for (int i = 0; i < NV; i++) d.qfrc_passive[i] = -m.damping[i] * d.qvel[i];
  referenceConstraint(d);
  rneBias(m, d);
  memset(d.qfrc_actuator, 0, sizeof(d.qfrc_actuator));
  for (int i = 0; i < 3; i++) { R c = d.ctrl[i]; if (c > 10) c = 10; if (c < -10) c = -10; d.qfrc_actuator[6 + i] = c; }
  for (int i = 0; i < NV; i++) d.qfrc_smooth[i] = d.qfrc_passive[i] - d.qfrc_bias[i] + d.qfrc_actuator[i];
  cholSolve(d.qacc_smooth, d.qL, d.qfrc_smooth, NV);
  d.solver_niter = 0;
  if (d.nefc == 0) { memcpy(d.qacc, d.qacc_smooth, sizeof(d.qacc)); memset(d.efc_force, 0, sizeof(d.efc_force)); }
  else { Solver s(m, d); s.run(); }
}
void integratePos(const Model& m, R* qpos, const R* qvel, R h) {  // [3P-memory: mj_integratePos]
  for (int b = 1; b < NB; b++) {
    if (m.jtype[b] == JFREE) {
      R* q = qpos + m.qadr[b]; const R* v = qvel + m.dadr[b];
      for (int k = 0; k < 3; k++) q[k] += h * v[k];
      R ax[3] = {v[3], v[4], v[5]}; R ang = h * normalize3(ax); R qr[4];
      axisAngle2Quat(qr, ax, ang); normalize4(q + 3); mulQuat(q + 3, q + 3, qr);
    } else if (m.jtype[b] == JHINGE) qpos[m.qadr[b]] += h * qvel[m.dadr[b]];
  }
}
bool badState(const Data& d) {
  for (int i = 0; i < NQ; i++) if (!(std::fabs(d.qpos[i]) < 1e10)) return true;
  for (int i = 0; i < NV; i++) if (!(std::fabs(d.qvel[i]) < 1e10)) return true;
  return false;
}
void mjStep(const Model& m, Data& d) {  // [3P-memory: mj_step with mj_RungeKutta(4)]
  const R A[3][3] = {{0.5, 0, 0}, {0, 0.5, 0}, {0, 0, 1}}, B[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6};
  const R h = m.timestep;
  R X[4][NQ + NV], F[4][NV], time0 = d.time;
  forward(m, d);
  memcpy(X[0], d.qpos, NQ * sizeof(R)); memcpy(X[0] + NQ, d.qvel, NV * sizeof(R)); memcpy(F[0], d.qacc, NV * sizeof(R));
  for (int i = 1; i < 4; i++) {
    R dv[NV], da[NV];
    for (int k = 0; k < NV; k++) { dv[k] = 0; da[k] = 0; for (int j = 0; j < i; j++) { dv[k] += A[i - 1][j] * X[j][NQ + k]; da[k] += A[i - 1][j] * F[j][k]; } }
    memcpy(X[i], X[0], sizeof(X[0]));
    integratePos(m, X[i], dv, h);
    for (int k = 0; k < NV; k++) X[i][NQ + k] += h * da[k];
    memcpy(d.qpos, X[i], NQ * sizeof(R)); memcpy(d.qvel, X[i] + NQ, NV * sizeof(R));
    forward(m, d);
    memcpy(X[i], d.qpos, NQ * sizeof(R));  // kinematics normalises quaternions in place
    memcpy(F[i], d.qacc, NV * sizeof(R));
  }
  R dv[NV], da[NV];
  for (int k = 0; k < NV; k++) { dv[k] = 0; da[k] = 0; for (int j = 0; j < 4; j++) { dv[k] += B[j] * X[j][NQ + k]; da[k] += B[j] * F[j][k]; } }
  memcpy(d.qpos, X[0], NQ * sizeof(R)); memcpy(d.qvel, X[0] + NQ, NV * sizeof(R));
  for (int k = 0; k < NV; k++) d.qvel[k] += h * da[k];
  integratePos(m, d.qpos, dv, h);
  d.time = time0 + h;
  memcpy(d.qacc_warmstart, d.qacc, NV * sizeof(R));  // last stage qacc (mj_advance)
}

// ----------------------------------------------------------------------------- simplex noise (noise._simplex restated)
const unsigned char PERM[256] = {
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
inline int perm(int i) { return PERM[i & 255]; }
const float GRAD4[32][4] = {
    {0, 1, 1, 1},  {0, 1, 1, -1},  {0, 1, -1, 1},  {0, 1, -1, -1},  {0, -1, 1, 1},  {0, -1, 1, -1},  {0, -1, -1, 1},  {0, -1, -1, -1},
    {1, 0, 1, 1},  {1, 0, 1, -1},  {1, 0, -1, 1},  {1, 0, -1, -1},  {-1, 0, 1, 1},  {-1, 0, 1, -1},  {-1, 0, -1, 1},  {-1, 0, -1, -1},
    {1, 1, 0, 1},  {1, 1, 0, -1},  {1, -1, 0, 1},  {1, -1, 0, -1},  {-1, 1, 0, 1},  {-1, 1, 0, -1},  {-1, -1, 0, 1},  {-1, -1, 0, -1},
    {1, 1, 1, 0},  {1, 1, -1, 0},  {1, -1, 1, 0},  {1, -1, -1, 0},  {-1, 1, 1, 0},  {-1, 1, -1, 0},  {-1, -1, 1, 0},  {-1, -1, -1, 0}};
float noise4(float x, float y, float z, float w) {
  const float F4 = 0.309016994f, G4 = 0.138196601f;
  float s = (x + y + z + w) * F4;
  float i = floorf(x + s), j = floorf(y + s), k = floorf(z + s), l = floorf(w + s);
  float t = (i + j + k + l) * G4;
  float x0 = x - (i - t), y0 = y - (j - t), z0 = z - (k - t), w0 = w - (l - t);
  // coordinate ranks (equivalent to Gustavson's 64-entry simplex lookup)
  int rx = (x0 > y0) + (x0 > z0) + (x0 > w0);
  int ry = !(x0 > y0) + (y0 > z0) + (y0 > w0);
  int rz = !(x0 > z0) + !(y0 > z0) + (z0 > w0);
  int rw = !(x0 > w0) + !(y0 > w0) + !(z0 > w0);
  int I = (int)i & 255, J = (int)j & 255, K = (int)k & 255, L = (int)l & 255;
  float total = 0;
  for (int c = 0; c < 5; c++) {
    int i1, j1, k1, l1;  // corner c steps along the c highest-ranked coordinates
    if (c == 0) { i1 = j1 = k1 = l1 = 0; }
    else if (c == 4) { i1 = j1 = k1 = l1 = 1; }
    else { int thr = 4 - c; i1 = rx >= thr; j1 = ry >= thr; k1 = rz >= thr; l1 = rw >= thr; }
    float xc = x0 - i1 + c * G4, yc = y0 - j1 + c * G4, zc = z0 - k1 + c * G4, wc = w0 - l1 + c * G4;
    float tt = 0.6f - xc * xc - yc * yc - zc * zc - wc * wc;
    if (tt >= 0.0f) {
      int gi = perm(I + i1 + perm(J + j1 + perm(K + k1 + perm(L + l1)))) & 0x1f;
      tt *= tt;
      total += tt * tt * (GRAD4[gi][0] * xc + GRAD4[gi][1] * yc + GRAD4[gi][2] * zc + GRAD4[gi][3] * wc);
    }
  }
  return 27.0f * total;
}

// 2-D simplex noise of noise._simplex (untiled branch of snoise2: terrain/gradient.py:74-80)  [3P-memory]
const float GRAD3[12][3] = {{1, 1, 0}, {-1, 1, 0}, {1, -1, 0}, {-1, -1, 0}, {1, 0, 1}, {-1, 0, 1}, {1, 0, -1}, {-1, 0, -1}, {0, 1, 1}, {0, -1, 1}, {0, 1, -1}, {0, -1, -1}};
float noise2(float x, float y) {
  const float F2 = 0.3660254037844386f, G2 = 0.21132486540518713f;
  float s = (x + y) * F2, i = floorf(x + s), j = floorf(y + s), t = (i + j) * G2;
  float xx[3], yy[3], n[3] = {0.f, 0.f, 0.f};
  xx[0] = x - (i - t); yy[0] = y - (j - t);
  int i1 = xx[0] > yy[0], j1 = xx[0] <= yy[0];
  xx[2] = xx[0] + G2 * 2.0f - 1.0f; yy[2] = yy[0] + G2 * 2.0f - 1.0f;
  xx[1] = xx[0] - i1 + G2; yy[1] = yy[0] - j1 + G2;
  int I = (int)i & 255, J = (int)j & 255;
  int g[3] = {perm(I + perm(J)) % 12, perm(I + i1 + perm(J + j1)) % 12, perm(I + 1 + perm(J + 1)) % 12};
  for (int c = 0; c < 3; c++) {
    float f = 0.5f - xx[c] * xx[c] - yy[c] * yy[c];
    if (f > 0) n[c] = f * f * f * f * (GRAD3[g[c]][0] * xx[c] + GRAD3[g[c]][1] * yy[c]);
  }
  return (n[0] + n[1] + n[2]) * 70.0f;
}
float fbm4(float x, float y, float z, float w, int octaves, float persistence, float lacunarity) {
  float freq = 1.0f, amp = 1.0f, mx = 1.0f, total = noise4(x, y, z, w);
  for (int i = 1; i < octaves; i++) {
    freq *= lacunarity; amp *= persistence; mx += amp;
    total += noise4(x * freq, y * freq, z * freq, w * freq) * amp;
  }
  return total / mx;
}

// ----------------------------------------------------------------------------- depth ray cast (CPU)
// Replaces mujoco.Renderer depth (sensors/rgbd.py:64-75): planar depth along the optical axis at pixel centres,
// fovy 90, clipped at 1.0. Scene: hfield + ball + wheels + tower + sticks (ballast is group 3: hidden; cone mesh missing).
struct Ray { R o[3], dvec[3]; };
R raySphere(const Ray& r, const R* c, R rad) {
  R oc[3] = {r.o[0] - c[0], r.o[1] - c[1], r.o[2] - c[2]};
  R a = dot3(r.dvec, r.dvec), b = dot3(oc, r.dvec), cc = dot3(oc, oc) - rad * rad;
  R disc = b * b - a * cc; if (disc < 0) return -1;
  R t = (-b - std::sqrt(disc)) / a;  // front face only
  return t;
}
R rayCylinderSide(const Ray& r, const R* c, const R* ax, R rad, R hl) {
  R oc[3] = {r.o[0] - c[0], r.o[1] - c[1], r.o[2] - c[2]};
  R od = dot3(oc, ax), dd = dot3(r.dvec, ax);
  R op[3], dp[3]; for (int k = 0; k < 3; k++) { op[k] = oc[k] - od * ax[k]; dp[k] = r.dvec[k] - dd * ax[k]; }
  R a = dot3(dp, dp), b = dot3(op, dp), cc = dot3(op, op) - rad * rad;
  if (a < 1e-18) return -1;
  R disc = b * b - a * cc; if (disc < 0) return -1;
  R t = (-b - std::sqrt(disc)) / a;
  R z = od + t * dd; if (z < -hl || z > hl) return -1;
  return t;
}
R rayCapsule(const Ray& r, const R* c, const R* ax, R rad, R hl) {
  R best = -1, t = rayCylinderSide(r, c, ax, rad, hl);
  if (t > 0) best = t;
  for (int s = -1; s <= 1; s += 2) {
    R e[3] = {c[0] + s * hl * ax[0], c[1] + s * hl * ax[1], c[2] + s * hl * ax[2]};
    t = raySphere(r, e, rad);
    if (t > 0) {
      R hp[3] = {r.o[0] + t * r.dvec[0] - c[0], r.o[1] + t * r.dvec[1] - c[1], r.o[2] + t * r.dvec[2] - c[2]};
      if (s * dot3(hp, ax) >= hl && (best < 0 || t < best)) best = t;
    }
  }
  return best;
}
R rayCylinder(const Ray& r, const R* c, const R* ax, R rad, R hl) {
  R best = -1, t = rayCylinderSide(r, c, ax, rad, hl);
  if (t > 0) best = t;
  R oc[3] = {r.o[0] - c[0], r.o[1] - c[1], r.o[2] - c[2]};
  R od = dot3(oc, ax), dd = dot3(r.dvec, ax);
  for (int s = -1; s <= 1; s += 2) {
    if (std::fabs(dd) < 1e-18 || s * dd >= 0) continue;  // front face only
    t = (s * hl - od) / dd; if (t <= 0) continue;
    R hp[3]; for (int k = 0; k < 3; k++) hp[k] = oc[k] + t * r.dvec[k] - (s * hl) * ax[k];
    if (dot3(hp, hp) <= rad * rad && (best < 0 || t < best)) best = t;
  }
  return best;
}
// ray vs heightfield surface: march cell by cell over the x-y grid (Amanatides-Woo), test both triangles per cell
R rayTri(const Ray& r, const R* a, const R* b, const R* c) {
  R e1[3], e2[3], p[3], tv[3], q[3];
  for (int k = 0; k < 3; k++) { e1[k] = b[k] - a[k]; e2[k] = c[k] - a[k]; }
  cross3(p, r.dvec, e2); R det = dot3(e1, p); if (std::fabs(det) < 1e-18) return -1;
  R inv = 1 / det; for (int k = 0; k < 3; k++) tv[k] = r.o[k] - a[k];
  R u = dot3(tv, p) * inv; if (u < 0 || u > 1) return -1;
  cross3(q, tv, e1); R v = dot3(r.dvec, q) * inv; if (v < 0 || u + v > 1) return -1;
  R t = dot3(e2, q) * inv; return t > 0 ? t : -1;
}
R rayHfield(const Ray& r, const float* data, const R* hs, R tmax) {
  const R sx = hs[0], sy = hs[1], sz = hs[2]; const int n = HN; const R dx = 2 * sx / (n - 1), dy = 2 * sy / (n - 1);
  // clip to the grid extent
  R t0 = 0, t1 = tmax;
  for (int ax = 0; ax < 2; ax++) {
    R lo = (ax ? -sy : -sx), hi = -lo, o = r.o[ax], dd = r.dvec[ax];
    if (std::fabs(dd) < 1e-18) { if (o < lo || o > hi) return -1; }
    else { R ta = (lo - o) / dd, tb = (hi - o) / dd; if (ta > tb) std::swap(ta, tb); if (ta > t0) t0 = ta; if (tb < t1) t1 = tb; }
  }
  if (t0 >= t1) return -1;
  R px = r.o[0] + (t0 + 1e-12) * r.dvec[0], py = r.o[1] + (t0 + 1e-12) * r.dvec[1];
  int cx = (int)std::floor((px + sx) / dx), cy = (int)std::floor((py + sy) / dy);
  if (cx < 0) cx = 0; if (cx > n - 2) cx = n - 2; if (cy < 0) cy = 0; if (cy > n - 2) cy = n - 2;
  int stx = r.dvec[0] > 0 ? 1 : -1, sty = r.dvec[1] > 0 ? 1 : -1;
  R tdx = std::fabs(r.dvec[0]) < 1e-18 ? 1e300 : dx / std::fabs(r.dvec[0]), tdy = std::fabs(r.dvec[1]) < 1e-18 ? 1e300 : dy / std::fabs(r.dvec[1]);
  R nbx = -sx + (cx + (stx > 0 ? 1 : 0)) * dx, nby = -sy + (cy + (sty > 0 ? 1 : 0)) * dy;
  R tmx = std::fabs(r.dvec[0]) < 1e-18 ? 1e300 : (nbx - r.o[0]) / r.dvec[0], tmy = std::fabs(r.dvec[1]) < 1e-18 ? 1e300 : (nby - r.o[1]) / r.dvec[1];
  R tcur = t0;
  for (int iter = 0; iter < 4 * n; iter++) {
    if (cx < 0 || cx > n - 2 || cy < 0 || cy > n - 2 || tcur > t1) return -1;
    R tnext = std::fmin(tmx, tmy);
    // vertices of the cell, split along (c,r)-(c+1,r+1) as in the prism construction
    R v00[3] = {-sx + cx * dx, -sy + cy * dy, (R)data[cy * n + cx] * sz};
    R v10[3] = {-sx + (cx + 1) * dx, -sy + cy * dy, (R)data[cy * n + cx + 1] * sz};
    R v01[3] = {-sx + cx * dx, -sy + (cy + 1) * dy, (R)data[(cy + 1) * n + cx] * sz};
    R v11[3] = {-sx + (cx + 1) * dx, -sy + (cy + 1) * dy, (R)data[(cy + 1) * n + cx + 1] * sz};
    R ta = rayTri(r, v01, v00, v11), tb = rayTri(r, v00, v11, v10);
    R best = -1; if (ta > 0) best = ta; if (tb > 0 && (best < 0 || tb < best)) best = tb;
    if (best > 0 && best <= tmax) return best;
    if (tmx < tmy) { cx += stx; tcur = tmx; tmx += tdx; } else { cy += sty; tcur = tmy; tmy += tdy; }
    (void)tnext;
  }
  return -1;
}

}  // namespace

// ============================================================================= env
struct bbo_env {
  bbo_config cfg;
  Model m;
  Data d;
  int step_counter;
  bool have_image; R image_ts;
  std::vector<float> img[2];
  int cam_refreshed;
};

namespace {

void renderCam(bbo_env* e, int cam, float* out) {
  const Model& m = e->m; const Data& d = e->d; const int H = e->cfg.im_h, W = e->cfg.im_w;
  R cpos[3], cmat[9], t[3];
  mulMatVec3(t, d.xmat[1], m.cam_pos[cam]); for (int k = 0; k < 3; k++) cpos[k] = d.xpos[1][k] + t[k];
  mulMat3(cmat, d.xmat[1], m.cam_mat[cam]);
  // scene primitives in world frame
  R ballc[3]; mulMatVec3(t, d.xmat[7], m.ball.pos); for (int k = 0; k < 3; k++) ballc[k] = d.xpos[7][k] + t[k];
  struct Prim { int type; R c[3], ax[3], rad, hl; }; Prim prims[6]; int np = 0;
  auto addGeom = [&](const Geom& g) {
    Prim& p = prims[np++]; p.type = g.type; p.rad = g.size[0]; p.hl = g.size[1];
    mulMatVec3(t, d.xmat[g.body], g.pos); for (int k = 0; k < 3; k++) p.c[k] = d.xpos[g.body][k] + t[k];
    R gm[9]; mulMat3(gm, d.xmat[g.body], g.mat); v3set(p.ax, gm[2], gm[5], gm[8]);
  };
  for (int i = 0; i < 3; i++) addGeom(m.wheel[i]);
  addGeom(m.tower); addGeom(m.stick[0]); addGeom(m.stick[1]);
  const R tmax = 1.0;  // depth >= 1 is clipped to 1 anyway (rgbd.py:74); fovy 90 => |dir| <= sqrt(3)
  for (int i = 0; i < H; i++)
    for (int j = 0; j < W; j++) {
      R xn = (2 * (j + 0.5) / W - 1) * ((R)W / H), yn = 1 - 2 * (i + 0.5) / H;  // tan(fovy/2)=1
      R dc[3] = {xn, yn, -1};
      Ray r; v3cp(r.o, cpos); mulMatVec3(r.dvec, cmat, dc);
      R best = tmax;
      R tt = raySphere(r, ballc, m.ball.size[0]); if (tt > 1e-4 && tt < best) best = tt;
      for (int p = 0; p < np; p++) {
        tt = prims[p].type == 1 ? rayCapsule(r, prims[p].c, prims[p].ax, prims[p].rad, prims[p].hl)
                                : rayCylinder(r, prims[p].c, prims[p].ax, prims[p].rad, prims[p].hl);
        if (tt > 1e-4 && tt < best) best = tt;
      }
      tt = rayHfield(r, d.hfield.data(), m.hf_size, best); if (tt > 1e-4 && tt < best) best = tt;
      out[i * W + j] = (float)(best >= 1.0 ? 1.0 : best);
    }
}

// ballbot_env.py:701-829
void getObs(bbo_env* e, const float* last_action, float* obs) {
  const Data& d = e->d;
  e->cam_refreshed = 0;
  if (e->cfg.cameras) {
    R delta = d.time - e->image_ts;
    if (!e->have_image || delta >= 1.0 / e->cfg.camera_frame_rate) {
      renderCam(e, 0, e->img[0].data()); renderCam(e, 1, e->img[1].data());
      e->have_image = true; e->image_ts = d.time; e->cam_refreshed = 1;
    }
  }
  // orientation: numpy-quaternion as_rotation_vector = 2*log(normalized q)  [3P-memory]
  const R* q = d.xquat[1];
  R qn = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  R w = q[0] / qn, x = q[1] / qn, y = q[2] / qn, z = q[3] / qn;
  R b = std::sqrt(x * x + y * y + z * z), rv[3] = {0, 0, 0};
  if (std::fabs(b) <= 1e-14 * std::fabs(w)) { if (w < 0) { rv[0] = 2 * M_PI; } }
  else { R f = 2 * std::atan2(b, w) / b; rv[0] = f * x; rv[1] = f * y; rv[2] = f * z; }
  for (int k = 0; k < 3; k++) obs[k] = (float)rv[k];
  auto clip2 = [](float v) { return v < -2.f ? -2.f : (v > 2.f ? 2.f : v); };
  // "angular_vel" = cvel[3:6] (really linear), "vel" = cvel[0:3] (really angular): ballbot_env.py:794-800
  for (int k = 0; k < 3; k++) obs[3 + k] = clip2((float)d.cvel[1][3 + k]);
  for (int k = 0; k < 3; k++) obs[6 + k] = clip2((float)d.cvel[1][k]);
  // motor_state = qvel[joint id 1..3] / max_wheel_velocity: ballbot_env.py:783-788
  for (int k = 0; k < 3; k++) obs[9 + k] = clip2((float)d.qvel[1 + k] / (float)e->cfg.max_wheel_velocity);
  for (int k = 0; k < 3; k++) obs[12 + k] = last_action[k];
  obs[15] = e->cfg.cameras ? (float)(d.time - e->image_ts) : 0.f;
}

}  // namespace

extern "C" {

void bbo_default_config(bbo_config* c) {
  c->max_ep_steps = 4000; c->max_allowed_tilt = 20.0; c->max_wheel_velocity = 10.0; c->camera_frame_rate = 90.0;
  c->reward_scale = 0.01; c->action_reg_coef = -0.0001; c->survival_bonus = 0.02; c->target_dir[0] = 0; c->target_dir[1] = 1;
  c->hfield_zscale = 2.0; c->cameras = 0; c->im_h = 64; c->im_w = 64;
  c->reward_type = 0; c->goal[0] = 0; c->goal[1] = 0; c->distance_scale = 1.0;
}
bbo_env* bbo_create(const bbo_config* cfg) {
  bbo_env* e = new bbo_env();
  { const char* x = getenv("BBO_EXTRA"); if (x) g_extra_pairs = x[0] != '0'; }
  e->cfg = *cfg; buildModel(e->m); e->m.hf_size[2] = cfg->hfield_zscale;
  memset(e->d.qpos, 0, sizeof(R) * NQ);
  e->d.hfield.assign(HN * HN, 0.f);
  e->img[0].assign(cfg->im_h * cfg->im_w, 1.f); e->img[1].assign(cfg->im_h * cfg->im_w, 1.f);
  bbo_reset(e, nullptr, nullptr);
  return e;
}
void bbo_destroy(bbo_env* e) { delete e; }

double bbo_spawn_offset(const float* hf, double zscale) {  // ballbot_env.py:540-565 (window rows/cols 140..151)
  float mx = -1e30f;
  for (int r = 140; r < 152; r++) for (int c = 140; c < 152; c++) { float v = hf ? hf[r * HN + c] : 0.f; if (v > mx) mx = v; }
  return (double)mx * zscale + 0.01;
}
static void resetData(bbo_env* e) {  // mj_resetData
  Data& d = e->d;
  memcpy(d.qpos, e->m.qpos0, sizeof(R) * NQ);
  memset(d.qvel, 0, sizeof(d.qvel)); memset(d.qacc, 0, sizeof(d.qacc)); memset(d.qacc_warmstart, 0, sizeof(d.qacc_warmstart));
  memset(d.ctrl, 0, sizeof(d.ctrl)); d.time = 0;
}
int bbo_set_hfield(bbo_env* e, const float* hf) {
  if (hf) memcpy(e->d.hfield.data(), hf, sizeof(float) * HN * HN); else std::fill(e->d.hfield.begin(), e->d.hfield.end(), 0.f);
  return 0;
}
int bbo_reset(bbo_env* e, const float* hf, float* obs) {
  bbo_set_hfield(e, hf);
  double off = bbo_spawn_offset(e->d.hfield.data(), e->m.hf_size[2]);
  resetData(e);
  e->d.qpos[2] += off; e->d.qpos[12] += off;
  forward(e->m, e->d);  // mj_forward (ballbot_env.py:620)
  e->step_counter = 0; e->have_image = false; e->image_ts = 0;
  float zero[3] = {0, 0, 0}, tmp[16];
  getObs(e, zero, obs ? obs : tmp);
  return 0;
}
int bbo_step(bbo_env* e, const float* a, float* obs, float* reward, uint8_t* terminated, uint8_t* failure, float* info4) {
  Data& d = e->d; const bbo_config& c = e->cfg;
  for (int k = 0; k < 3; k++) {  // ballbot_env.py:903-907
    R u = (R)a[k] * c.max_wheel_velocity;
    if (u > c.max_wheel_velocity) u = c.max_wheel_velocity; if (u < -c.max_wheel_velocity) u = -c.max_wheel_velocity;
    d.ctrl[k] = -u;
  }
  if (badState(d)) { resetData(e); }  // mj_checkPos/Vel reset semantics
  mjStep(e->m, d);
  getObs(e, a, obs);
  // reward, float32 arithmetic as NumPy>=2 does (ballbot_env.py:929-937)
  float r;
  if (c.reward_type == 1) {   // DistanceReward on pos2d (rewards/distance.py:33-50): float32 norm of (goal - pos2d)
    const float dx = (float)c.goal[0] - (float)d.xpos[1][0], dy = (float)c.goal[1] - (float)d.xpos[1][1];
    r = (-(float)c.distance_scale * sqrtf(dx * dx + dy * dy)) * (float)c.reward_scale;
  } else r = (obs[6] * (float)c.target_dir[0] + obs[7] * (float)c.target_dir[1]) * (float)c.reward_scale;
  float nrm = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  r += (float)c.action_reg_coef * (nrm * nrm);
  e->step_counter++;
  bool term = e->step_counter >= c.max_ep_steps, fail = false;
  // tilt from the float32 rotation vector (ballbot_env.py:989-1006): angle = acos(R[2][2])
  R rv[3] = {obs[0], obs[1], obs[2]}, th = norm3(rv), qx, qy;
  if (th < 1e-300) { qx = qy = 0; } else { R s = std::sin(0.5 * th) / th; qx = rv[0] * s; qy = rv[1] * s; }
  R angle = std::acos(1 - 2 * (qx * qx + qy * qy)) * 180.0 / M_PI;
  if (angle > c.max_allowed_tilt) { fail = true; term = true; } else r += (float)c.survival_bonus;
  *reward = r; *terminated = term; *failure = fail;
  if (info4) { info4[0] = (float)d.xpos[1][0]; info4[1] = (float)d.xpos[1][1]; info4[2] = (float)e->step_counter; info4[3] = (float)e->cam_refreshed; }
  return 0;
}
int bbo_get_depth(bbo_env* e, float* i0, float* i1) {
  size_t n = (size_t)e->cfg.im_h * e->cfg.im_w;
  memcpy(i0, e->img[0].data(), n * sizeof(float)); memcpy(i1, e->img[1].data(), n * sizeof(float));
  return 0;
}
int bbo_render_depth(bbo_env* e, int cam, float* img) { renderCam(e, cam, img); return 0; }
int bbo_get_state(bbo_env* e, double* qpos, double* qvel, double* warm, double* time) {
  if (qpos) memcpy(qpos, e->d.qpos, sizeof(R) * NQ); if (qvel) memcpy(qvel, e->d.qvel, sizeof(R) * NV);
  if (warm) memcpy(warm, e->d.qacc_warmstart, sizeof(R) * NV); if (time) *time = e->d.time;
  return 0;
}
int bbo_set_state(bbo_env* e, const double* qpos, const double* qvel, const double* warm, double time) {
  if (qpos) memcpy(e->d.qpos, qpos, sizeof(R) * NQ); if (qvel) memcpy(e->d.qvel, qvel, sizeof(R) * NV);
  if (warm) memcpy(e->d.qacc_warmstart, warm, sizeof(R) * NV); e->d.time = time;
  return 0;
}
int bbo_mj_step(bbo_env* e, const double* ctrl) { for (int k = 0; k < 3; k++) e->d.ctrl[k] = ctrl[k]; mjStep(e->m, e->d); return 0; }
int bbo_forward(bbo_env* e, const double* ctrl, double* qM, double* bias, double* qas, double* qacc, int* ncon, int* niter) {
  if (ctrl) for (int k = 0; k < 3; k++) e->d.ctrl[k] = ctrl[k];
  forward(e->m, e->d);
  if (qM) memcpy(qM, e->d.qM, sizeof(R) * NV * NV); if (bias) memcpy(bias, e->d.qfrc_bias, sizeof(R) * NV);
  if (qas) memcpy(qas, e->d.qacc_smooth, sizeof(R) * NV); if (qacc) memcpy(qacc, e->d.qacc, sizeof(R) * NV);
  if (ncon) *ncon = e->d.ncon; if (niter) *niter = e->d.solver_niter;
  return 0;
}
int bbo_get_contacts(bbo_env* e, int maxcon, double* dist, double* pos, double* frame, int* pair) {
  int n = e->d.ncon < maxcon ? e->d.ncon : maxcon;
  for (int i = 0; i < n; i++) {
    const Contact& c = e->d.con[i];
    if (dist) dist[i] = c.dist; if (pos) memcpy(pos + 3 * i, c.pos, 3 * sizeof(R));
    if (frame) memcpy(frame + 9 * i, c.frame, 9 * sizeof(R)); if (pair) pair[i] = c.pair;
  }
  return e->d.ncon;
}
int bbo_get_efc(bbo_env* e, int maxefc, double* J, double* aref, double* D, double* force) {
  int n = e->d.nefc < maxefc ? e->d.nefc : maxefc;
  if (J) memcpy(J, e->d.efc_J, sizeof(R) * NV * n); if (aref) memcpy(aref, e->d.efc_aref, sizeof(R) * n);
  if (D) memcpy(D, e->d.efc_D, sizeof(R) * n); if (force) memcpy(force, e->d.efc_force, sizeof(R) * n);
  return e->d.nefc;
}
int bbo_get_model(double* mass, double* ipos, double* inertia, double* invw, double* meaninertia) {
  static Model m; static bool built = false; if (!built) { buildModel(m); built = true; }
  if (mass) memcpy(mass, m.mass, sizeof(m.mass)); if (ipos) memcpy(ipos, m.ipos, sizeof(m.ipos));
  if (inertia) memcpy(inertia, m.inertia, sizeof(m.inertia)); if (invw) memcpy(invw, m.invweight0, sizeof(m.invweight0));
  if (meaninertia) *meaninertia = m.meaninertia;
  return 0;
}
int bbo_get_kin(bbo_env* e, double* xpos, double* xquat, double* cvel) {
  if (xpos) memcpy(xpos, e->d.xpos[1], 3 * sizeof(R)); if (xquat) memcpy(xquat, e->d.xquat[1], 4 * sizeof(R));
  if (cvel) memcpy(cvel, e->d.cvel[1], 6 * sizeof(R));
  return 0;
}

static int g_trig = -1;   // 1: fast_sin / fast_cos of noise/_noise.h (default), 0: libm sinf / cosf (BBO_TRIG=l, experiment record)
static inline float fast_sin_(float x) {
  volatile float z = (x + 25165824.0f);
  x = x - (z - 25165824.0f);
  float y = x - x * fabsf(x);
  return y * (3.1f + 3.6f * fabsf(y));
}
static inline float tsin(float x) { return g_trig ? fast_sin_(x) : sinf(x); }
static inline float tcos(float x) { return g_trig ? fast_sin_(x + 0.5f) : cosf(x); }
float bbo_snoise2_tiled(float x, float y, int octaves, float persistence, float lacunarity, float repeatx, float repeaty, int base) {
  // noise._simplex py_noise2, tiled branch (both repeats given) [3P-memory]
  if (g_trig < 0) { const char* e = getenv("BBO_TRIG"); g_trig = !(e && e[0] == 'l'); }
  float z = (float)base, w = z;
  float yf = (float)(y * 2.0 / repeaty), yr = (float)(repeaty * M_1_PI * 0.5);
  float vy = tsin(yf), vyz = tcos(yf);
  y = vy * yr; w += vyz * yr;
  float xf = (float)(x * 2.0 / repeatx), xr = (float)(repeatx * M_1_PI * 0.5);
  float vx = tsin(xf), vxz = tcos(xf);
  x = vx * xr; z += vxz * xr;
  return fbm4(x, y, z, w, octaves, persistence, lacunarity);
}
float bbo_snoise2(float x, float y, int octaves, float persistence, float lacunarity, int base) {
  // untiled branch of py_noise2: `base` is added to both coordinates of every octave
  const float z = (float)base;
  float freq = 1.0f, amp = 1.0f, mx = 1.0f, total = noise2(x + z, y + z);
  for (int i = 1; i < octaves; i++) {
    freq *= lacunarity; amp *= persistence; mx += amp;
    total += noise2(x * freq + z, y * freq + z) * amp;
  }
  return total / mx;
}
int bbo_perlin_terrain(int n, double scale, int octaves, double persistence, double lacunarity, double amplitude, int seed, float* out) {
  // terrain/perlin.py:51-74
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      float x = (float)(i / scale), y = (float)(j / scale);
      double nv = (double)bbo_snoise2_tiled(x, y, octaves, (float)persistence, (float)lacunarity, 1024.f, 1024.f, seed);
      double v = (nv + 1.0) / 2.0 * amplitude;
      if (v < 0) v = 0; if (v > 1) v = 1;
      out[i * n + j] = (float)v;
    }
  return 0;
}
}  // extern "C"
