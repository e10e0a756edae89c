// bb_torch_ops.cpp -- PyTorch extension over the C ABI (include/ballbot_b200.h): the hot-path entry points as dispatcher ops
// `torch.ops.ballbot.{step, reset, add_reward, gae}` taking / mutating CUDA tensors in place, without copies.
//
// "PyTorch extension wraps these as torch.ops.ballbot.* taking/returning CUDA tensors".  The ops run on the
// current CUDA stream of the tensors' device (c10::cuda::getCurrentCUDAStream) under a device guard, so they compose with
// torch streams, CUDA-graph capture (torch.cuda.graph) and the dispatcher like any other CUDA op.  The engine handle is the
// bb_engine* as an int (created / destroyed through the C ABI; ctypes stays the binding for non-torch hosts).
// Host code only: the kernels live in libballbot_b200.so, which this library links against ($ORIGIN rpath).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/types.h>

#include "../../include/ballbot_b200.h"

namespace {

bb_engine* handle(int64_t h) {
  TORCH_CHECK(h != 0, "ballbot: null engine handle");
  return reinterpret_cast<bb_engine*>(static_cast<intptr_t>(h));
}
void check(int rc, bb_engine* e, const char* what) { TORCH_CHECK(rc == BB_OK, what, " failed (", rc, "): ", bb_last_error(e)); }

template <typename T> T* ptr(const at::Tensor& t, at::ScalarType st, int64_t numel, const char* name, bool optional = false) {
  if (!t.defined() || t.numel() == 0) { TORCH_CHECK(optional, "ballbot: output tensor '", name, "' is required"); return nullptr; }
  TORCH_CHECK(t.is_cuda() && t.is_contiguous() && t.scalar_type() == st, "ballbot: '", name, "' must be a contiguous CUDA tensor of the documented dtype");
  TORCH_CHECK(numel < 0 || t.numel() == numel, "ballbot: '", name, "' has ", t.numel(), " elements, expected ", numel);
  return reinterpret_cast<T*>(t.data_ptr());
}

// outs: the 16 tensors of bb_io in declaration order (orientation, angular_vel, vel, motor_state, actions, rel_image_ts,
// rgbd_0, rgbd_1, reward, terminated, failure, pos2d, terminal_obs, episode_return, episode_length, status); optional ones may be empty
bb_io makeIo(bb_engine* e, at::TensorList o) {
  TORCH_CHECK(o.size() == 16, "ballbot: expected the 16 bb_io tensors, got ", o.size());
  const int64_t N = bb_num_envs(e);
  bb_io io;
  io.orientation = ptr<float>(o[0], at::kFloat, 3 * N, "orientation"); io.angular_vel = ptr<float>(o[1], at::kFloat, 3 * N, "angular_vel");
  io.vel = ptr<float>(o[2], at::kFloat, 3 * N, "vel"); io.motor_state = ptr<float>(o[3], at::kFloat, 3 * N, "motor_state");
  io.actions = ptr<float>(o[4], at::kFloat, 3 * N, "actions"); io.rel_image_ts = ptr<float>(o[5], at::kFloat, N, "relative_image_timestamp");
  io.rgbd_0 = ptr<float>(o[6], at::kFloat, -1, "rgbd_0", true); io.rgbd_1 = ptr<float>(o[7], at::kFloat, -1, "rgbd_1", true);
  io.reward = ptr<float>(o[8], at::kFloat, N, "reward"); io.terminated = ptr<uint8_t>(o[9], at::kByte, N, "terminated");
  io.failure = ptr<uint8_t>(o[10], at::kByte, N, "failure"); io.pos2d = ptr<float>(o[11], at::kFloat, 2 * N, "pos2d");
  io.terminal_obs = ptr<float>(o[12], at::kFloat, 16 * N, "terminal_obs", true); io.episode_return = ptr<float>(o[13], at::kFloat, N, "episode_return", true);
  io.episode_length = ptr<int32_t>(o[14], at::kInt, N, "episode_length", true); io.status = ptr<int32_t>(o[15], at::kInt, N, "status", true);
  return io;
}

// replaces VecEnv.step_async + step_wait (ballbot_env.py:854-1036 x N); outputs are written in place, stream-ordered, no sync
void step(int64_t h, const at::Tensor& actions, at::TensorList outs) {
  bb_engine* e = handle(h);
  const c10::cuda::CUDAGuard guard(actions.device());
  const bb_io io = makeIo(e, outs);
  const float* a = ptr<float>(actions, at::kFloat, 3 * (int64_t)bb_num_envs(e), "actions");
  check(bb_step(e, a, &io, at::cuda::getCurrentCUDAStream().stream()), e, "bb_step");
}
// replaces VecEnv.reset / BBotSimulation.reset (ballbot_env.py:567-671): mask uint8[N] and terrain seeds int32[N] are optional
void reset(int64_t h, const c10::optional<at::Tensor>& mask, const c10::optional<at::Tensor>& seeds, at::TensorList outs) {
  bb_engine* e = handle(h);
  TORCH_CHECK(outs.size() == 16, "ballbot: expected the 16 bb_io tensors");
  const c10::cuda::CUDAGuard guard(outs[0].device());
  const bb_io io = makeIo(e, outs);
  const int64_t N = bb_num_envs(e);
  const uint8_t* m = mask.has_value() ? ptr<uint8_t>(*mask, at::kByte, N, "mask") : nullptr;
  const int32_t* s = seeds.has_value() ? ptr<int32_t>(*seeds, at::kInt, N, "seeds") : nullptr;
  check(bb_reset(e, m, s, &io, at::cuda::getCurrentCUDAStream().stream()), e, "bb_reset");
}
void add_reward(int64_t h, const at::Tensor& term, at::TensorList outs) {
  bb_engine* e = handle(h);
  const c10::cuda::CUDAGuard guard(term.device());
  const bb_io io = makeIo(e, outs);
  check(bb_add_reward(e, ptr<float>(term, at::kFloat, (int64_t)bb_num_envs(e), "term"), &io, at::cuda::getCurrentCUDAStream().stream()), e, "bb_add_reward");
}
// GAE over device-resident [T, N] rollout tensors (SB3 RolloutBuffer.compute_returns_and_advantage)
std::tuple<at::Tensor, at::Tensor> gae(const at::Tensor& rewards, const at::Tensor& values, const at::Tensor& dones, double gamma, double lam) {
  TORCH_CHECK(rewards.dim() == 2 && values.dim() == 2 && values.size(0) == rewards.size(0) + 1 && values.size(1) == rewards.size(1), "ballbot::gae: rewards [T,N], values [T+1,N]");
  const c10::cuda::CUDAGuard guard(rewards.device());
  const int64_t T = rewards.size(0), N = rewards.size(1);
  at::Tensor adv = at::empty_like(rewards), ret = at::empty_like(rewards);
  if (T * N == 0) return {adv, ret};
  const int rc = bb_gae(ptr<float>(rewards, at::kFloat, T * N, "rewards"), ptr<float>(values, at::kFloat, (T + 1) * N, "values"), ptr<uint8_t>(dones, at::kByte, T * N, "dones"),
                        (int32_t)T, (int32_t)N, (float)gamma, (float)lam, adv.data_ptr<float>(), ret.data_ptr<float>(), at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == BB_OK, "bb_gae failed (", rc, ")");
  return {adv, ret};
}

}  // namespace

TORCH_LIBRARY(ballbot, m) {
  m.def("step(int engine, Tensor actions, Tensor(a!)[] outs) -> ()");
  m.def("reset(int engine, Tensor? mask, Tensor? seeds, Tensor(a!)[] outs) -> ()");
  m.def("add_reward(int engine, Tensor term, Tensor(a!)[] outs) -> ()");
  m.def("gae(Tensor rewards, Tensor values, Tensor dones, float gamma, float gae_lambda) -> (Tensor, Tensor)");
}
TORCH_LIBRARY_IMPL(ballbot, CUDA, m) {
  m.impl("step", &step);
  m.impl("reset", &reset);
  m.impl("add_reward", &add_reward);
  m.impl("gae", &gae);
}
