// LayerNorm kernels (HBM-bound): one warp per row, 16-byte vector loads,
// fp32 statistics via warp shuffles, two-pass (mean, then centred variance) on
// register-resident data so the row is read from HBM exactly once.
//
// dod_layernorm      : y = LN(x)                 (HF Dinov2 norm1/norm2/layernorm, eps 1e-6)
// dod_add_layernorm  : y = LN(x + r)             (decoder post-norm, eps 1e-5)
#include "common.cuh"
#include "../../include/dod.h"

namespace dod {
void count_launch(int n = 1);
namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kMaxVec = 12;  // 12 * 4 floats * 32 lanes = 1536 columns max

template <bool X_BF16>
__device__ __forceinline__ void load_row(const void* x, int64_t row_off, int d, int lane,
                                         float4 (&v)[kMaxVec]) {
  const int nvec = d >> 2;
  if (X_BF16) {
    const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x) + row_off;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const uint2 raw = *reinterpret_cast<const uint2*>(xp + c * 4);
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
        const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
        v[i] = make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
      } else {
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  } else {
    const float* xp = reinterpret_cast<const float*>(x) + row_off;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int c = lane + i * 32;
      v[i] = c < nvec ? *reinterpret_cast<const float4*>(xp + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

__device__ __forceinline__ void normalize_store(float4 (&v)[kMaxVec], int d, int lane, float eps,
                                                const float* gamma, const float* beta,
                                                float* yf, __nv_bfloat16* yb) {
  const int nvec = d >> 2;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / float(d);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, e = v[i].z - mean, f = v[i].w - mean;
      q += (a * a + b * b) + (e * e + f * f);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / float(d) + eps);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
      const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + c);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + bt.x;
      o.y = (v[i].y - mean) * rstd * g.y + bt.y;
      o.z = (v[i].z - mean) * rstd * g.z + bt.z;
      o.w = (v[i].w - mean) * rstd * g.w + bt.w;
      if (yf) *reinterpret_cast<float4*>(yf + c * 4) = o;
      if (yb) {
        uint2 pk;
        pk.x = pack_bf16x2(o.x, o.y);
        pk.y = pack_bf16x2(o.z, o.w);
        *reinterpret_cast<uint2*>(yb + c * 4) = pk;
      }
    }
  }
}

template <bool X_BF16>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
layernorm_kernel(const void* __restrict__ x, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float* __restrict__ yf,
                 __nv_bfloat16* __restrict__ yb, int64_t rows, int d, int64_t ldx, int64_t ldy,
                 float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  float4 v[kMaxVec];
  load_row<X_BF16>(x, row * ldx, d, lane, v);
  normalize_store(v, d, lane, eps, gamma, beta, yf ? yf + row * ldy : nullptr,
                  yb ? yb + row * ldy : nullptr);
}

template <bool R_BF16>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
add_layernorm_kernel(const float* __restrict__ x, const void* __restrict__ r,
                     const float* __restrict__ gamma, const float* __restrict__ beta,
                     float* __restrict__ yf, __nv_bfloat16* __restrict__ yb, int64_t rows, int d,
                     float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  float4 v[kMaxVec], w[kMaxVec];
  load_row<false>(x, row * d, d, lane, v);
  load_row<R_BF16>(r, row * d, d, lane, w);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    v[i].x += w[i].x; v[i].y += w[i].y; v[i].z += w[i].z; v[i].w += w[i].w;
  }
  normalize_store(v, d, lane, eps, gamma, beta, yf ? yf + row * d : nullptr,
                  yb ? yb + row * d : nullptr);
}

}  // namespace
__global__ void ln_rstd_kernel(const float2* __restrict__ stats, float* __restrict__ rstd, int64_t rows,
                               int slots, float inv_dim, float eps, float* __restrict__ max_ratio) {
  const int64_t row = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  float ratio = 0.0f;
  if (row < rows) {
    float s1 = 0.0f, s2 = 0.0f;
    for (int i = 0; i < slots; ++i) {
      const float2 v = __ldg(stats + int64_t(i) * rows + row);
      s1 += v.x;
      s2 += v.y;
    }
    const float mean = s1 * inv_dim;
    const float r = rsqrtf(fmaxf(s2 * inv_dim - mean * mean, 0.0f) + eps);
    rstd[row] = r;
    ratio = fabsf(mean) * r;
  }
  if (max_ratio != nullptr) {
    // non-negative floats order like their bit patterns: one integer atomicMax per warp
    ratio = warp_max(ratio);
    if ((threadIdx.x & 31) == 0 && ratio > 0.0f && isfinite(ratio))
      atomicMax(reinterpret_cast<unsigned int*>(max_ratio), __float_as_uint(ratio));
  }
}

}  // namespace dod

extern "C" int32_t dod_ln_rstd(const dod_ln_rstd_args* a, dod_stream_t stream_) {
  using namespace dod;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->row_stats && a->rstd, "dod_ln_rstd: null pointer");
  DOD_REQUIRE(a->rows >= 0 && a->slots > 0 && a->slots <= 4096 && a->dim > 0, "dod_ln_rstd: bad shape");
  DOD_REQUIRE((uintptr_t(a->row_stats) & 7) == 0, "dod_ln_rstd: row_stats must be 8-byte aligned");
  if (a->rows == 0) return DOD_OK;
  ln_rstd_kernel<<<unsigned((a->rows + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const float2*>(a->row_stats), a->rstd, a->rows, int(a->slots), 1.0f / float(a->dim), a->eps,
      a->max_mean_ratio);
  int rc = check_cuda(cudaGetLastError(), "ln_rstd_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}

extern "C" int32_t dod_layernorm(const dod_layernorm_args* a, dod_stream_t stream_) {
  using namespace dod;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->x && a->gamma && a->beta && a->y, "dod_layernorm: null pointer");
  DOD_REQUIRE(a->rows >= 0 && a->d > 0, "dod_layernorm: bad shape");
  DOD_REQUIRE(a->d % 4 == 0 && a->d <= kMaxVec * 128, "dod_layernorm: d must be a multiple of 4 and <= %d",
              kMaxVec * 128);
  DOD_REQUIRE(a->ldx % 4 == 0 && a->ldy % 4 == 0 && a->ldx >= a->d && a->ldy >= a->d,
              "dod_layernorm: ldx/ldy must be multiples of 4 and >= d");
  DOD_REQUIRE((uintptr_t(a->x) & 15) == 0 && (uintptr_t(a->y) & 15) == 0 &&
                  (uintptr_t(a->gamma) & 15) == 0 && (uintptr_t(a->beta) & 15) == 0 &&
                  (!a->y2 || (uintptr_t(a->y2) & 15) == 0),
              "dod_layernorm: pointers must be 16-byte aligned");
  if (a->rows == 0) return DOD_OK;
  float* yf = nullptr;
  __nv_bfloat16* yb = nullptr;
  if (a->y_dtype == DOD_F32) {
    yf = reinterpret_cast<float*>(a->y);
    yb = reinterpret_cast<__nv_bfloat16*>(a->y2);
  } else {
    yb = reinterpret_cast<__nv_bfloat16*>(a->y);
    yf = reinterpret_cast<float*>(a->y2);
  }
  const unsigned grid = unsigned((a->rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
  if (a->x_dtype == DOD_BF16)
    layernorm_kernel<true><<<grid, kWarpsPerBlock * 32, 0, stream>>>(
        a->x, a->gamma, a->beta, yf, yb, a->rows, int(a->d), a->ldx, a->ldy, a->eps);
  else
    layernorm_kernel<false><<<grid, kWarpsPerBlock * 32, 0, stream>>>(
        a->x, a->gamma, a->beta, yf, yb, a->rows, int(a->d), a->ldx, a->ldy, a->eps);
  int rc = check_cuda(cudaGetLastError(), "layernorm_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}

extern "C" int32_t dod_add_layernorm(const dod_add_layernorm_args* a, dod_stream_t stream_) {
  using namespace dod;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->x && a->r && a->gamma && a->beta && (a->y || a->y_bf16),
              "dod_add_layernorm: null pointer");
  DOD_REQUIRE(a->rows >= 0 && a->d > 0 && a->d % 4 == 0 && a->d <= kMaxVec * 128,
              "dod_add_layernorm: d must be a multiple of 4 and <= %d", kMaxVec * 128);
  if (a->rows == 0) return DOD_OK;
  const unsigned grid = unsigned((a->rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
  if (a->r_dtype == DOD_BF16)
    add_layernorm_kernel<true><<<grid, kWarpsPerBlock * 32, 0, stream>>>(
        a->x, a->r, a->gamma, a->beta, a->y, reinterpret_cast<__nv_bfloat16*>(a->y_bf16), a->rows,
        int(a->d), a->eps);
  else
    add_layernorm_kernel<false><<<grid, kWarpsPerBlock * 32, 0, stream>>>(
        a->x, a->r, a->gamma, a->beta, a->y, reinterpret_cast<__nv_bfloat16*>(a->y_bf16), a->rows,
        int(a->d), a->eps);
  int rc = check_cuda(cudaGetLastError(), "add_layernorm_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}
