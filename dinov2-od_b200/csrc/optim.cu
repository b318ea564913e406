// Fused gradient-norm clip + Adam over ONE flat fp32 buffer.
//
// The reference runs clip_grad_norm_(model.parameters(), 1.0) and Adam(lr, weight_decay -- L2 added to
// the gradient, not AdamW) over 59-85 small tensors every micro-step (train.py:1000-1004, 1104-1110):
// launch-bound.  Here the trainable parameters, their gradients (dino_detector/parallel.py,
// the buffer that is also all-reduced) and the Adam moments are flat fp32 buffers:
//   dod_sumsq      norm2 += sum g^2            (one pass, HBM-bound)
//   dod_adam_step  clip coefficient from the device-resident norm, L2 decay, moments, update
// No host sync: the norm never leaves the device.
#include "common.cuh"
#include "../../include/dod.h"

namespace dod {
void count_launch(int n = 1);
namespace {

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ out) {
  float acc = 0.f;
  for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < n; i += int64_t(gridDim.x) * 256) {
    const float v = g[i];
    acc = fmaf(v, v, acc);
  }
  acc = warp_sum(acc);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    atomicAdd(out, s);
  }
}

__global__ void counter_add_kernel(int64_t* c, int n, int64_t delta) {
  if (int(threadIdx.x) < n) c[threadIdx.x] += delta;
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, int64_t n, const float* __restrict__ sumsq, float max_norm, float lr,
            float beta1, float beta2, float eps, float weight_decay, float bias_c1, float bias_c2_sqrt,
            const int64_t* __restrict__ step_ptr) {
  if (step_ptr != nullptr) {  // bias corrections from the device-resident step counter
    const float t = float(*step_ptr);
    bias_c1 = 1.0f - powf(beta1, t);
    bias_c2_sqrt = sqrtf(1.0f - powf(beta2, t));
  }
  // torch.nn.utils.clip_grad_norm_: coef = clamp(max_norm / (total_norm + 1e-6), max=1)
  float coef = 1.0f;
  if (max_norm > 0.f) coef = fminf(max_norm / (sqrtf(sumsq[0]) + 1e-6f), 1.0f);
  for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < n; i += int64_t(gridDim.x) * 256) {
    const float pi = p[i];
    float gi = fmaf(weight_decay, pi, g[i] * coef);          // L2 in the gradient (torch Adam)
    const float mi = fmaf(beta1, m[i], (1.0f - beta1) * gi);
    const float vi = fmaf(beta2, v[i], (1.0f - beta2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bias_c2_sqrt + eps;
    p[i] = pi - (lr / bias_c1) * (mi / denom);
  }
}

}  // namespace
}  // namespace dod

extern "C" int32_t dod_sumsq(const dod_sumsq_args* a, dod_stream_t stream_) {
  using namespace dod;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->x && a->out && a->n >= 0, "dod_sumsq: bad arguments");
  if (a->n == 0) return DOD_OK;
  int64_t blocks = (a->n + 255) / 256;
  if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
  sumsq_kernel<<<unsigned(blocks), 256, 0, stream>>>(a->x, a->n, a->out);
  int rc = check_cuda(cudaGetLastError(), "sumsq_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}

extern "C" int32_t dod_adam_step(const dod_adam_args* a, dod_stream_t stream_) {
  using namespace dod;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->param && a->grad && a->exp_avg && a->exp_avg_sq, "dod_adam_step: null pointer");
  DOD_REQUIRE(a->n > 0 && (a->step >= 1 || a->step_ptr), "dod_adam_step: n > 0 and step >= 1 required");
  DOD_REQUIRE(a->max_grad_norm <= 0.f || a->grad_sumsq, "dod_adam_step: clipping needs grad_sumsq");
  const float bc1 = 1.0f - powf(a->beta1, float(a->step >= 1 ? a->step : 1));
  const float bc2 = 1.0f - powf(a->beta2, float(a->step >= 1 ? a->step : 1));
  int64_t blocks = (a->n + 255) / 256;
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  adam_kernel<<<unsigned(blocks), 256, 0, stream>>>(a->param, a->grad, a->exp_avg, a->exp_avg_sq, a->n,
                                                   a->grad_sumsq, a->max_grad_norm, a->lr, a->beta1,
                                                   a->beta2, a->eps, a->weight_decay, bc1, sqrtf(bc2), a->step_ptr);
  int rc = check_cuda(cudaGetLastError(), "adam_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}

extern "C" int32_t dod_counter_add(const dod_counter_add_args* a, dod_stream_t stream_) {
  using namespace dod;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->counters && a->n > 0 && a->n <= 32, "dod_counter_add: 1..32 counters");
  counter_add_kernel<<<1, 32, 0, stream>>>(a->counters, int(a->n), a->delta);
  int rc = check_cuda(cudaGetLastError(), "counter_add_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}
