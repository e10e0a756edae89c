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
