// bb_rollout.cu -- device-resident rollout post-processing for the GPU VecEnv (sm_100a), engine-independent C ABI.
//
// Replaces RolloutBuffer.compute_returns_and_advantage of the reference's learner (Stable-Baselines3 2.6.0
// DictRolloutBuffer, driven by PPO.collect_rollouts: ballbot_rl/training/train.py:126-141, 284), which runs as a Python
// loop over the horizon on host numpy arrays after every rollout.  Here the [T, N] reward / value / done tensors never
// leave HBM: one thread per env walks its column backwards (the recursion is sequential in t, independent across envs),
// loads are coalesced across envs and issued UNROLL steps ahead of the dependent arithmetic.  HBM-bound:
// 17 algorithmic bytes per transition (reward 4 + value 4 + done 1 read, advantage 4 + return 4 written).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ballbot_b200.h"

namespace {

constexpr int GAE_UNROLL = 8;

__global__ void __launch_bounds__(128) k_gae(int T, int N, const float* __restrict__ rewards, const float* __restrict__ values,
                                             const uint8_t* __restrict__ dones, float gamma, float lam, float* __restrict__ adv,
                                             float* __restrict__ ret) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float last = 0.f;
  float vnext = values[(size_t)T * N + i];          // bootstrap value of the observation after the last step
  int t = T - 1;
  for (; t >= GAE_UNROLL - 1; t -= GAE_UNROLL) {
    float r[GAE_UNROLL], v[GAE_UNROLL]; uint8_t d[GAE_UNROLL];
#pragma unroll
    for (int k = 0; k < GAE_UNROLL; k++) {
      const size_t o = (size_t)(t - k) * N + i;
      r[k] = rewards[o]; v[k] = values[o]; d[k] = dones[o];
    }
#pragma unroll
    for (int k = 0; k < GAE_UNROLL; k++) {
      const size_t o = (size_t)(t - k) * N + i;
      const float nonterminal = d[k] ? 0.f : 1.f;
      const float delta = r[k] + gamma * vnext * nonterminal - v[k];
      last = delta + gamma * lam * nonterminal * last;
      adv[o] = last; ret[o] = last + v[k];
      vnext = v[k];
    }
  }
  for (; t >= 0; t--) {
    const size_t o = (size_t)t * N + i;
    const float v = values[o], nonterminal = dones[o] ? 0.f : 1.f;
    const float delta = rewards[o] + gamma * vnext * nonterminal - v;
    last = delta + gamma * lam * nonterminal * last;
    adv[o] = last; ret[o] = last + v;
    vnext = v;
  }
}

}  // namespace

extern "C" int bb_gae(const float* rewards_dev, const float* values_dev, const uint8_t* dones_dev, int32_t T, int32_t N, float gamma,
                      float gae_lambda, float* advantages_dev, float* returns_dev, void* stream) {
  if (T < 0 || N < 0) return BB_ERR_INVALID;
  if (T == 0 || N == 0) return BB_OK;               // empty rollout: nothing to do (the pointers may be NULL)
  if (!rewards_dev || !values_dev || !dones_dev || !advantages_dev || !returns_dev) return BB_ERR_INVALID;
  k_gae<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(T, N, rewards_dev, values_dev, dones_dev, gamma, gae_lambda, advantages_dev, returns_dev);
  return cudaGetLastError() == cudaSuccess ? BB_OK : BB_ERR_CUDA;
}

// ---- fp64 FMA peak of this device (profiling aid): 8 independent DFMA chains per thread, 148 x 8 CTAs of 256 threads.
// The step kernels are latency-bound fp64 code, so the HBM roofline says little about them; bench.py divides the fp64 flop
// count of the step kernels (ncu: dfma x 2 + dmul + dadd) by this measured figure for roofline.compute.
namespace {
__global__ void __launch_bounds__(256) k_fp64_peak(int iters, double a, double b, double* out) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456) out[0] = s;     // never true: keeps the chains alive
}
}  // namespace

extern "C" int bb_fp64_peak(int32_t device, double* tflops) {
  if (!tflops) return BB_ERR_INVALID;
  int prev = -1, ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device) return BB_ERR_NO_DEVICE;
  cudaGetDevice(&prev); cudaSetDevice(device);
  double* out = nullptr; cudaEvent_t e0, e1;
  int rc = BB_OK;
  if (cudaMalloc(&out, sizeof(double)) != cudaSuccess) rc = BB_ERR_CUDA;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * 8, iters = 4096;
  double best = 0;
  for (int rep = 0; rep < 5 && rc == BB_OK; rep++) {       // first repetition = warm-up
    cudaEventRecord(e0);
    k_fp64_peak<<<grid, 256>>>(iters, 0.999999, 1e-6, out);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { rc = BB_ERR_CUDA; break; }
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 8 * 16 * (double)iters * 256.0 * grid;
    if (rep > 0 && fl / (ms * 1e-3) / 1e12 > best) best = fl / (ms * 1e-3) / 1e12;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  if (prev >= 0) cudaSetDevice(prev);
  *tflops = best;
  return rc;
}

// ---- fused PPO optimiser step on flat buffers ("fusing the gradient all-reduce with the AdamW update").
// The learner keeps every trainable parameter as a view of ONE flat fp32 buffer and autograd accumulates into ONE flat gradient
// buffer whose tail carries {sum of KL * samples, samples} of the minibatch; a single NCCL all-reduce (SUM) of that buffer is
// followed by this step.  Everything SB3 does on the host between all-reduce and optimizer.step happens here on the device:
//   * mean over the global minibatch (divide by the reduced sample count)
//   * target-KL early stop (SB3 PPO.train: approx_kl > 1.5 * target_kl ends the iteration BEFORE the step): sets a sticky flag,
//     a set flag turns this and every later step of the iteration into a no-op -- no host synchronisation per minibatch
//   * clip_grad_norm_(max_grad_norm), AdamW (decoupled weight decay, bias correction), step counter
// ctrl (device double[8]): [0] stop flag, [1] step count, [2] last global grad norm, [3] last KL, [4] updates applied,
// [5..7] running sums of the logged losses (added by the host side through the tail of the gradient buffer).
namespace {
__global__ void __launch_bounds__(256) k_grad_sumsq(const float* __restrict__ g, int n, double* __restrict__ out) {
  double acc = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { const double v = g[i]; acc += v * v; }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) { double s = 0; for (int k = 0; k < 8; k++) s += sh[k]; atomicAdd(out, s); }
}
__global__ void __launch_bounds__(256) k_adamw_flat(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                    int n, float lr, float b1, float b2, float eps, float wd, float max_norm, float kl_limit,
                                                    const double* __restrict__ sumsq, double* __restrict__ ctrl) {
  // tail of g: g[n] = sum over ranks of (KL * samples), g[n + 1] = samples
  const double cnt = g[n + 1] > 0.f ? (double)g[n + 1] : 1.0;
  const double kl = (double)g[n] / cnt;
  const bool stop = ctrl[0] != 0.0 || (kl_limit > 0.f && kl > (double)kl_limit);
  const double step = ctrl[1] + 1.0;
  const double gn = sqrt(*sumsq) / cnt;                              // norm of the mean gradient
  const float scale = (float)((max_norm > 0.f && gn > (double)max_norm ? (double)max_norm / (gn + 1e-6) : 1.0) / cnt);
  const float c1 = 1.f - powf(b1, (float)step), c2 = 1.f - powf(b2, (float)step);
  if (!stop) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const float gi = g[i] * scale;
      const float mi = b1 * m[i] + (1.f - b1) * gi, vi = b2 * v[i] + (1.f - b2) * gi * gi;
      m[i] = mi; v[i] = vi;
      float w = p[i];
      w -= lr * wd * w;                                               // decoupled weight decay (torch.optim.AdamW)
      w -= lr * (mi / c1) / (sqrtf(vi / c2) + eps);
      p[i] = w;
    }
  }
  // the flag / counters are advanced by the single-thread follow-up kernel k_adamw_commit (every block of this one reads them)
}
__global__ void k_adamw_commit(const float* __restrict__ g, int n, float kl_limit, const double* __restrict__ sumsq, double* __restrict__ ctrl) {
  const double cnt = g[n + 1] > 0.f ? (double)g[n + 1] : 1.0;
  const double kl = (double)g[n] / cnt;
  const bool stop = ctrl[0] != 0.0 || (kl_limit > 0.f && kl > (double)kl_limit);
  ctrl[2] = sqrt(*sumsq) / cnt; ctrl[3] = kl;
  if (stop) ctrl[0] = 1.0; else { ctrl[1] += 1.0; ctrl[4] += 1.0; }
}
}  // namespace

extern "C" int bb_adamw_step(float* param_dev, const float* grad_dev, float* m_dev, float* v_dev, int32_t n, float lr, float beta1, float beta2,
                             float eps, float weight_decay, float max_grad_norm, float kl_limit, double* ctrl_dev, double* scratch_dev, void* stream) {
  if (!param_dev || !grad_dev || !m_dev || !v_dev || !ctrl_dev || !scratch_dev || n < 1) return BB_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(scratch_dev, 0, sizeof(double), s) != cudaSuccess) return BB_ERR_CUDA;
  const int blocks = n < 148 * 256 * 4 ? (n + 255) / 256 : 148 * 4;
  k_grad_sumsq<<<blocks, 256, 0, s>>>(grad_dev, n, scratch_dev);
  k_adamw_flat<<<blocks, 256, 0, s>>>(param_dev, grad_dev, m_dev, v_dev, n, lr, beta1, beta2, eps, weight_decay, max_grad_norm, kl_limit, scratch_dev, ctrl_dev);
  k_adamw_commit<<<1, 1, 0, s>>>(grad_dev, n, kl_limit, scratch_dev, ctrl_dev);
  return cudaGetLastError() == cudaSuccess ? BB_OK : BB_ERR_CUDA;
}
