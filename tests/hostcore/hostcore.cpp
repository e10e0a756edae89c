// tests/hostcore/hostcore.cpp -- TEST TOOL ONLY.
// Compiles the engine's per-env physics core (openballbot_rl_b200/csrc/bb_core.cuh, written __host__ __device__)
// for the CPU so that the `-m "not gpu"` suite can compare the *engine's own arithmetic* with the oracle without a
// GPU.  It is never loaded by the product package: the product path is the CUDA kernels and fails without them.
#include "../../openballbot_rl_b200/csrc/bb_model.h"

using namespace bb;

static int g_solver_fast = 0;   // hc_set_solver: 0 = reference iteration path, 1 = solver_mode 1 of the engine

template <typename T> static const ModelConst<T>& model() {
  static ModelConst<T> mc; static bool ok = false;
  if (!ok) { ModelConst<double> md; buildModelConst(md); narrowModel(md, mc); ok = true; }
  return mc;
}

template <typename T>
static int fwd(const double* qpos, const double* qvel, const double* ctrl, const double* warm, const float* hf, double zscale,
               double* qacc, double* M225, double* qfs, double* kin /*13*/, int* ncon, int* niter, double* cdist, double* cpos, double* cframe, int* ctype) {
  static thread_local Scratch<T> s;
  T q[NQ], v[NV], c[3], w[NV], a[NV];
  for (int i = 0; i < NQ; i++) q[i] = (T)qpos[i];
  for (int i = 0; i < NV; i++) { v[i] = (T)qvel[i]; w[i] = (T)warm[i]; }
  for (int i = 0; i < 3; i++) c[i] = (T)ctrl[i];
  KinOut<T> k;
  memset(&s, 0xFF, sizeof(s));   // poison: the GPU kernel's local scratch is uninitialised too
  forwardDynamics(model<T>(), q, v, c, w, hf, (T)zscale, s, a, &k, g_solver_fast != 0);
  for (int i = 0; i < NV; i++) qacc[i] = a[i];
  if (M225) for (int i = 0; i < NV; i++) for (int j = 0; j < NV; j++) M225[i * NV + j] = s.M[tidx(i, j)];
  if (qfs) for (int i = 0; i < NV; i++) qfs[i] = s.qfs[i];
  if (kin) { for (int i = 0; i < 4; i++) kin[i] = k.quatB[i]; for (int i = 0; i < 3; i++) { kin[4 + i] = k.cvel_ang[i]; kin[7 + i] = k.cvel_lin[i]; kin[10 + i] = k.posB[i]; } }
  if (ncon) *ncon = k.ncon; if (niter) *niter = k.niter;
  for (int i = 0; i < s.nc; i++) {
    if (cdist) cdist[i] = s.cDist[i];
    if (cpos) for (int j = 0; j < 3; j++) cpos[3 * i + j] = s.cP[i][j];
    if (cframe) for (int j = 0; j < 9; j++) cframe[9 * i + j] = s.cF[i][j];
    if (ctype) ctype[i] = s.ctype[i];
  }
  return s.nc;
}
template <typename T>
static void step(double* qpos, double* qvel, double* warm, const double* ctrl, const float* hf, double zscale, double* kin, int* ncon, int* niter) {
  static thread_local Scratch<T> s;
  T q[NQ], v[NV], c[3], w[NV];
  for (int i = 0; i < NQ; i++) q[i] = (T)qpos[i];
  for (int i = 0; i < NV; i++) { v[i] = (T)qvel[i]; w[i] = (T)warm[i]; }
  for (int i = 0; i < 3; i++) c[i] = (T)ctrl[i];
  KinOut<T> k;
  memset(&s, 0xFF, sizeof(s));
  rk4Step(model<T>(), q, v, w, c, hf, (T)zscale, s, &k, (T*)nullptr, g_solver_fast != 0);
  for (int i = 0; i < NQ; i++) qpos[i] = q[i];
  for (int i = 0; i < NV; i++) { qvel[i] = v[i]; warm[i] = w[i]; }
  if (kin) { for (int i = 0; i < 4; i++) kin[i] = k.quatB[i]; for (int i = 0; i < 3; i++) { kin[4 + i] = k.cvel_ang[i]; kin[7 + i] = k.cvel_lin[i]; kin[10 + i] = k.posB[i]; } }
  if (ncon) *ncon = k.ncon; if (niter) *niter = k.niter;
}

extern "C" {
void hc_set_solver(int fast) { g_solver_fast = fast; }
int hc_forward(int prec, const double* qpos, const double* qvel, const double* ctrl, const double* warm, const float* hf, double zscale,
               double* qacc, double* M225, double* qfs, double* kin, int* ncon, int* niter, double* cdist, double* cpos, double* cframe, int* ctype) {
  return prec == 32 ? fwd<float>(qpos, qvel, ctrl, warm, hf, zscale, qacc, M225, qfs, kin, ncon, niter, cdist, cpos, cframe, ctype)
                    : fwd<double>(qpos, qvel, ctrl, warm, hf, zscale, qacc, M225, qfs, kin, ncon, niter, cdist, cpos, cframe, ctype);
}
void hc_step(int prec, double* qpos, double* qvel, double* warm, const double* ctrl, const float* hf, double zscale, double* kin, int* ncon, int* niter) {
  if (prec == 32) step<float>(qpos, qvel, warm, ctrl, hf, zscale, kin, ncon, niter); else step<double>(qpos, qvel, warm, ctrl, hf, zscale, kin, ncon, niter);
}
void hc_model(double* dA12, double* meaninertia, double* masses3 /*m0,mw,mL*/, double* c0) {
  const ModelConst<double>& m = model<double>();
  for (int i = 0; i < NCT; i++) dA12[i] = m.dA[i];
  *meaninertia = m.meaninertia; masses3[0] = m.m0; masses3[1] = m.mw; masses3[2] = m.mL;
  for (int i = 0; i < 3; i++) c0[i] = m.c0[i];
}
}
#ifdef BB_STATS
extern "C" long hc_ls_evals() { return bb_stats_ls_evals; }
extern "C" long hc_newton_iters() { return bb_stats_newton; }
extern "C" void hc_stats_reset() { bb_stats_ls_evals = 0; bb_stats_newton = 0; }
#endif
