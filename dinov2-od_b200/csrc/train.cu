// Backward-pass / training-only kernels of libdod (HBM- or latency-bound helpers around the
// tcgen05 GEMM, which does every dense contraction of the backward pass as well).
//
//  dod_transpose_bf16    batched [R, C] -> [C, R] (operands of the M-reduction GEMMs)
//  dod_lowrank_wgrad     out[c, j] += alpha * sum_m big[m, c] * small[m, j]   (LoRA dA / dB, r <= 64)
//  dod_colsum            out[c] += sum_m x[m, c]                               (bias gradients)
//  dod_layernorm_bwd     dx (+ residual-path gradient), optional dgamma / dbeta
//  dod_eltwise           casts, LayerScale, GELU / ReLU / SwiGLU / sigmoid derivatives, dropout
//  dod_softmax_rows(_bwd) materialised-probability attention used by the two LoRA blocks and the
//                        decoder in training (forward saves P, backward is four batched GEMMs)
//  dod_deform_sample_bwd gradient of the bilinear sampling (reference deformable_attention.py:100-178)
#include "common.cuh"
#include "../../include/dod.h"

namespace dod {
void count_launch(int n = 1);
namespace {

__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float ldv(const void* p, int64_t i, int dt) {
  return dt == DOD_F32 ? reinterpret_cast<const float*>(p)[i]
                       : bf2f(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void stv(void* p, int64_t i, int dt, float v) {
  if (dt == DOD_F32) reinterpret_cast<float*>(p)[i] = v;
  else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

// ------------------------------------------------------------------ transpose
__global__ void __launch_bounds__(256)
transpose_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int rows,
                 int cols, int64_t ld_in, int64_t ld_out, int64_t bs_in, int64_t bs_out, int inner,
                 int64_t is_in) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int b = blockIdx.z;  // outer * inner + inner index; the output is contiguous over both
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const __nv_bfloat16* ip = in + int64_t(b / inner) * bs_in + int64_t(b % inner) * is_in;
  __nv_bfloat16* op = out + int64_t(b) * bs_out;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
  for (int r = ty; r < 64; r += 4) {
    const int rr = r0 + r, cc = c0 + tx;
    tile[r][tx] = (rr < rows && cc < cols) ? ip[int64_t(rr) * ld_in + cc] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  for (int c = ty; c < 64; c += 4) {
    const int cc = c0 + c, rr = r0 + tx;
    if (cc < cols && rr < rows) op[int64_t(cc) * ld_out + rr] = tile[tx][c];
  }
}

// ------------------------------------------------------------- low-rank wgrad
// out[c, j] (or out[j, c]) += alpha * sum_m big[m, c] * small[m, j],  j < r <= 64.
// grid.x = column blocks of 256, grid.y = row chunks; thread = one column, r accumulators.
template <int R>
__global__ void __launch_bounds__(256)
lowrank_wgrad_kernel(const __nv_bfloat16* __restrict__ big, const __nv_bfloat16* __restrict__ small,
                     float* __restrict__ out, int64_t m, int cols, int r, int64_t ld_big,
                     int64_t ld_small, int64_t ldo, int transposed, float alpha, int rows_per_cta) {
  __shared__ float ssm[32][R];
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int64_t m0 = int64_t(blockIdx.y) * rows_per_cta;
  const int64_t m1 = min(m, m0 + rows_per_cta);
  float acc[R];
#pragma unroll
  for (int j = 0; j < R; ++j) acc[j] = 0.f;
  for (int64_t mm = m0; mm < m1; mm += 32) {
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * R; i += 256) {
      const int rr = i / R, j = i - rr * R;
      ssm[rr][j] = (mm + rr < m1 && j < r) ? bf2f(small[(mm + rr) * ld_small + j]) : 0.f;
    }
    __syncthreads();
    if (c < cols) {
      const int nrow = (m1 - mm) < 32 ? int(m1 - mm) : 32;
      for (int rr = 0; rr < nrow; ++rr) {
        const float v = bf2f(big[(mm + rr) * ld_big + c]);
#pragma unroll
        for (int j = 0; j < R; ++j) acc[j] = fmaf(v, ssm[rr][j], acc[j]);
      }
    }
  }
  if (c < cols) {
#pragma unroll
    for (int j = 0; j < R; ++j)
      if (j < r) atomicAdd(transposed ? out + int64_t(j) * ldo + c : out + int64_t(c) * ldo + j, alpha * acc[j]);
  }
}

// Vectorised form: a warp covers 32 * CPL adjacent columns of one row with ONE load instruction (CPL bf16 per
// lane, CPL * R = 64 accumulators per thread), the eight warps of a CTA take different rows, partial sums meet
// in shared memory.  The thread-per-column kernel above loads 2 bytes per thread per row and ran at ~1 TB/s
// (L/14 LoRA blocks: 140 us for a 90 MB operand); it stays as the fallback for unaligned operands.
template <int CPL, int R>
__global__ void __launch_bounds__(256, 2)
lowrank_wgrad_vec_kernel(const __nv_bfloat16* __restrict__ big, const __nv_bfloat16* __restrict__ small,
                         float* __restrict__ out, int64_t m, int cols, int r, int64_t ld_big, int64_t ld_small,
                         int64_t ldo, int transposed, float alpha, int rows_per_cta) {
  static_assert(CPL * R == 64, "64 accumulators per thread");
  constexpr int UNR = R <= 8 ? 4 : 2;
  __shared__ float sacc[CPL * R * 32];  // [i][j][lane]: conflict-free shared-memory atomics
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 32 * CPL + lane * CPL;
  const bool col_ok = c + CPL <= cols;
  for (int i = threadIdx.x; i < CPL * R * 32; i += 256) sacc[i] = 0.f;
  __syncthreads();
  float acc[CPL][R];
#pragma unroll
  for (int i = 0; i < CPL; ++i)
#pragma unroll
    for (int j = 0; j < R; ++j) acc[i][j] = 0.f;
  const int64_t m0 = int64_t(blockIdx.y) * rows_per_cta;
  const int64_t m1 = min(m, m0 + rows_per_cta);
  if (col_ok) {
    for (int64_t mb = m0 + warp; mb < m1; mb += 8 * UNR) {
      uint32_t vraw[UNR][(CPL + 1) / 2];
      uint4 sraw[UNR][R / 8];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int64_t mm = mb + 8 * u;
        if (mm < m1) {
          const __nv_bfloat16* bp = big + mm * ld_big + c;
          if constexpr (CPL == 8) {
            const uint4 q = *reinterpret_cast<const uint4*>(bp);
            vraw[u][0] = q.x; vraw[u][1] = q.y; vraw[u][2] = q.z; vraw[u][3] = q.w;
          } else if constexpr (CPL == 4) {
            const uint2 q = *reinterpret_cast<const uint2*>(bp);
            vraw[u][0] = q.x; vraw[u][1] = q.y;
          } else if constexpr (CPL == 2) {
            vraw[u][0] = *reinterpret_cast<const uint32_t*>(bp);
          } else {
            vraw[u][0] = *reinterpret_cast<const uint16_t*>(bp);
          }
          const uint4* sp = reinterpret_cast<const uint4*>(small + mm * ld_small);
#pragma unroll
          for (int q = 0; q < R / 8; ++q) sraw[u][q] = __ldg(sp + q);
        } else {
#pragma unroll
          for (int q = 0; q < (CPL + 1) / 2; ++q) vraw[u][q] = 0u;
#pragma unroll
          for (int q = 0; q < R / 8; ++q) sraw[u][q] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        float sv[R];
#pragma unroll
        for (int q = 0; q < R / 8; ++q) {
          const uint32_t w4[4] = {sraw[u][q].x, sraw[u][q].y, sraw[u][q].z, sraw[u][q].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            sv[q * 8 + 2 * e] = __uint_as_float(w4[e] << 16);
            sv[q * 8 + 2 * e + 1] = __uint_as_float(w4[e] & 0xffff0000u);
          }
        }
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
          const uint32_t w2 = vraw[u][i >> 1];
          const float v = __uint_as_float((i & 1) ? (w2 & 0xffff0000u) : (w2 << 16));
#pragma unroll
          for (int j = 0; j < R; ++j) acc[i][j] = fmaf(v, sv[j], acc[i][j]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < CPL; ++i)
#pragma unroll
      for (int j = 0; j < R; ++j) atomicAdd(&sacc[(i * R + j) * 32 + lane], acc[i][j]);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < CPL * R * 32; idx += 256) {
    const int ln = idx & 31, ij = idx >> 5;
    const int i = ij / R, j = ij - i * R;
    const int col = blockIdx.x * 32 * CPL + ln * CPL + i;
    if (col < cols && j < r)
      atomicAdd(transposed ? out + int64_t(j) * ldo + col : out + int64_t(col) * ldo + j, alpha * sacc[idx]);
  }
}

__global__ void __launch_bounds__(256)
colsum_kernel(const void* __restrict__ x, int dt, float* __restrict__ out, int64_t m, int cols,
              int64_t ld, int rows_per_cta) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  const int64_t m0 = int64_t(blockIdx.y) * rows_per_cta;
  const int64_t m1 = min(m, m0 + rows_per_cta);
  float acc = 0.f;
  for (int64_t mm = m0; mm < m1; ++mm) acc += ldv(x, mm * ld + c, dt);
  atomicAdd(out + c, acc);
}

// --------------------------------------------------------------- LayerNorm bwd
// y = (x - mean) * rstd * gamma + beta.  dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),
// g = dy * gamma.  One warp per row; statistics recomputed from x (fp32).
constexpr int kLnMaxVec = 12;
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const void* __restrict__ dy, int dy_dt, const float* __restrict__ x,
                     const float* __restrict__ gamma, const float* __restrict__ dres,
                     float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     int64_t rows, int d, float eps) {
  const int lane = threadIdx.x & 31;
  const int nvec = d >> 2;
  // dgamma / dbeta (decoder LayerNorms): per-CTA partial sums in shared memory, one global atomic per column
  // and CTA at the end.  One global atomic per element and row (1600 rows x 768 columns x 2 on the same 1536
  // addresses) serialised in L2 and made these small launches as slow as the [43840, 1024] ones.
  extern __shared__ float s_part[];  // [2][d] when dgamma != nullptr
  if (dgamma) {
    for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) s_part[i] = 0.f;
    __syncthreads();
  }
  for (int64_t row = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5); row < rows; row += int64_t(gridDim.x) * 8) {
  float4 xv[kLnMaxVec], gv[kLnMaxVec];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      xv[i] = *reinterpret_cast<const float4*>(x + row * d + c * 4);
      s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
    } else xv[i] = make_float4(0, 0, 0, 0);
  }
  const float mean = warp_sum(s) / float(d);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
      q += (xv[i].x * xv[i].x + xv[i].y * xv[i].y) + (xv[i].z * xv[i].z + xv[i].w * xv[i].w);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / float(d) + eps);
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;  // xhat
      // one 16-byte (f32) / 8-byte (bf16) load per lane: four scalar loads per lane moved 2-4 bytes per
      // instruction and held the kernel at 3.8 TB/s
      float4 dyv;
      if (dy_dt == DOD_F32) {
        dyv = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + row * d + c * 4);
      } else {
        const uint2 rw = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy) + row * d + c * 4);
        dyv = make_float4(__uint_as_float(rw.x << 16), __uint_as_float(rw.x & 0xffff0000u),
                          __uint_as_float(rw.y << 16), __uint_as_float(rw.y & 0xffff0000u));
      }
      if (dgamma) {
        atomicAdd(s_part + c * 4 + 0, dyv.x * xv[i].x);
        atomicAdd(s_part + c * 4 + 1, dyv.y * xv[i].y);
        atomicAdd(s_part + c * 4 + 2, dyv.z * xv[i].z);
        atomicAdd(s_part + c * 4 + 3, dyv.w * xv[i].w);
        atomicAdd(s_part + d + c * 4 + 0, dyv.x);
        atomicAdd(s_part + d + c * 4 + 1, dyv.y);
        atomicAdd(s_part + d + c * 4 + 2, dyv.z);
        atomicAdd(s_part + d + c * 4 + 3, dyv.w);
      }
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
      gv[i] = make_float4(dyv.x * g.x, dyv.y * g.y, dyv.z * g.z, dyv.w * g.w);
      sg += (gv[i].x + gv[i].y) + (gv[i].z + gv[i].w);
      sgx += (gv[i].x * xv[i].x + gv[i].y * xv[i].y) + (gv[i].z * xv[i].z + gv[i].w * xv[i].w);
    }
  }
  const float mg = warp_sum(sg) / float(d), mgx = warp_sum(sgx) / float(d);
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      float4 o;
      o.x = rstd * (gv[i].x - mg - xv[i].x * mgx);
      o.y = rstd * (gv[i].y - mg - xv[i].y * mgx);
      o.z = rstd * (gv[i].z - mg - xv[i].z * mgx);
      o.w = rstd * (gv[i].w - mg - xv[i].w * mgx);
      if (dres) {
        const float4 r = *reinterpret_cast<const float4*>(dres + row * d + c * 4);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      *reinterpret_cast<float4*>(dx + row * d + c * 4) = o;
    }
  }
  }  // rows of this warp
  if (dgamma) {
    __syncthreads();
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
      atomicAdd(dgamma + i, s_part[i]);
      atomicAdd(dbeta + i, s_part[d + i]);
    }
  }
}

// ------------------------------------------------------------------ elementwise
__device__ __forceinline__ float gelu_f(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}
// bf16 outputs: erf(z) ~ tanh(z (c0 + c1 z^2 + c2 z^4 + c3 z^6)) with one MUFU.TANH (the form the GEMM epilogue
// uses for the frozen blocks, gemm.cu gelu_bf16_x2: max |err| 5.5e-5 + 2^-11 relative from tanh.approx, under an
// eighth of the bf16 rounding of the result) and exp(-x^2/2) with one MUFU.EX2.  erff + expf cost ~50 issue
// slots per element and made GELU_BWD issue-bound (367 us for [43840, 4096] against 165 us of HBM time).
__device__ __forceinline__ float erf_tanh_form(float x) {  // erf(x / sqrt 2)
  const float s = x * x;
  float q = fmaf(s, 8.668819692e-06f, -3.982046110e-04f);
  q = fmaf(q, s, 3.698001081e-02f);
  q = fmaf(q, s, 7.976396815e-01f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(q * x));
  return t;
}
__device__ __forceinline__ float gelu_fast(float x) {
  const float hx = 0.5f * x;
  return fmaf(hx, erf_tanh_form(x), hx);
}
__device__ __forceinline__ float gelu_grad_fast(float x) {
  const float cdf = fmaf(0.5f, erf_tanh_form(x), 0.5f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-0.72134752044448170f * x * x));
  return fmaf(x * 0.3989422804014327f, e, cdf);
}
__device__ __forceinline__ uint32_t hash32(uint64_t v) {  // splitmix-style counter hash
  v += 0x9E3779B97F4A7C15ull;
  v = (v ^ (v >> 30)) * 0xBF58476D1CE4E5B9ull;
  v = (v ^ (v >> 27)) * 0x94D049BB133111EBull;
  return uint32_t((v ^ (v >> 31)) >> 32);
}

// 8 consecutive elements of a row per thread: 16-byte (bf16) / 2 x 16-byte (f32) accesses when the
// caller's pointers and leading dimensions allow it (VEC), scalar accesses otherwise.
template <bool VEC>
__device__ __forceinline__ void ld8(const void* p, int64_t off, int dt, int n, float (&v)[8]) {
  if (VEC) {
    if (dt == DOD_BF16) {
      const uint4 raw = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p) + off);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
      }
    } else {
      const float4 lo = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + off);
      const float4 hi = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + off + 4);
      v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
      v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = i < n ? ldv(p, off + i, dt) : 0.f;
  }
}
template <bool VEC>
__device__ __forceinline__ void st8(void* p, int64_t off, int dt, int n, const float (&v)[8]) {
  if (VEC) {
    if (dt == DOD_BF16) {
      uint4 raw;
      raw.x = pack_bf16x2(v[0], v[1]);
      raw.y = pack_bf16x2(v[2], v[3]);
      raw.z = pack_bf16x2(v[4], v[5]);
      raw.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p) + off) = raw;
    } else {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + off) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + off + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < n) stv(p, off + i, dt, v[i]);
  }
}

template <bool VEC>
__global__ void __launch_bounds__(256)
eltwise_kernel(int mode, const void* __restrict__ a, int a_dt, const void* __restrict__ b, int b_dt,
               const float* __restrict__ vec, void* __restrict__ out, int out_dt, void* __restrict__ out2,
               int out2_dt, int64_t rows, int cols, int64_t ld_a, int64_t ld_b, int64_t ld_out,
               float p0, uint64_t seed, const int64_t* __restrict__ seed_ptr) {
  const int gpr = (cols + 7) >> 3;  // 8-element groups per row
  const int64_t g = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (g >= rows * gpr) return;
  const int64_t r = g / gpr;
  const int c = int(g - r * gpr) << 3;
  const int n = min(8, cols - c);
  float x[8], y[8], o[8];
  ld8<VEC>(a, r * ld_a + c, a_dt, n, x);
  switch (mode) {
    case DOD_ELT_CAST:
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = x[i];
      break;
    case DOD_ELT_SCALE_COLS:  // out = a * vec[c]
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = i < n ? x[i] * vec[c + i] : 0.f;
      break;
    case DOD_ELT_ADD:
      ld8<VEC>(b, r * ld_b + c, b_dt, n, y);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = x[i] + y[i];
      break;
    case DOD_ELT_GELU_FWD:
      if (out_dt == DOD_BF16) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = gelu_fast(x[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = gelu_f(x[i]);
      }
      break;
    case DOD_ELT_GELU_BWD:  // a = upstream grad, b = pre-activation
      ld8<VEC>(b, r * ld_b + c, b_dt, n, y);
      if (out_dt == DOD_BF16) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = x[i] * gelu_grad_fast(y[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = x[i] * gelu_grad(y[i]);
      }
      break;
    case DOD_ELT_RELU_BWD:  // a = upstream grad, b = activation output
      ld8<VEC>(b, r * ld_b + c, b_dt, n, y);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = y[i] > 0.f ? x[i] : 0.f;
      break;
    case DOD_ELT_SIGMOID_BWD:  // a = upstream grad, b = sigmoid output
      ld8<VEC>(b, r * ld_b + c, b_dt, n, y);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = x[i] * y[i] * (1.0f - y[i]);
      break;
    case DOD_ELT_SWIGLU_FWD:  // a = [rows, 2*cols] (gate | linear), out = silu(gate) * linear
      ld8<VEC>(a, r * ld_a + cols + c, a_dt, n, y);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = x[i] / (1.0f + expf(-x[i])) * y[i];
      break;
    case DOD_ELT_SWIGLU_BWD: {  // a = upstream [rows, cols], b = pre-activation [rows, 2*cols]; out [rows, 2*cols]
      float gt[8], o2[8];
      ld8<VEC>(b, r * ld_b + c, b_dt, n, gt);
      ld8<VEC>(b, r * ld_b + cols + c, b_dt, n, y);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float sg = 1.0f / (1.0f + expf(-gt[i]));
        o[i] = x[i] * y[i] * (sg * (1.0f + gt[i] * (1.0f - sg)));
        o2[i] = x[i] * gt[i] * sg;
      }
      st8<VEC>(out, r * ld_out + cols + c, out_dt, n, o2);
      break;
    }
    case DOD_ELT_DROPOUT: {  // out = a * keep / (1 - p); the mask is a pure function of (seed, element index)
      if (seed_ptr != nullptr) seed += uint64_t(*seed_ptr) << 44;
      const uint64_t i0 = uint64_t(r) * uint64_t(cols) + uint64_t(c);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        o[i] = (hash32(seed + i0 + i) >> 8) * (1.0f / 16777216.0f) >= p0 ? x[i] * (1.0f / (1.0f - p0)) : 0.f;
      if (out2) st8<VEC>(out2, r * ld_out + c, out2_dt, n, o);
      break;
    }
    case DOD_ELT_AXPBY:
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = vec[0] * x[i];
      if (b) {
        ld8<VEC>(b, r * ld_b + c, b_dt, n, y);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += vec[1] * y[i];
      }
      break;
    default: return;
  }
  st8<VEC>(out, r * ld_out + c, out_dt, n, o);
}

// ------------------------------------------------------------------ softmax rows
// P[row, :n] = softmax(scale * S[row, :n]) (bf16, zero padded to ldp); optional prob dropout.
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const void* __restrict__ s, int s_dt, __nv_bfloat16* __restrict__ p, int64_t rows,
                    int n, int64_t lds, int64_t ldp, float scale, float drop_p, uint64_t seed,
                    const int64_t* __restrict__ seed_ptr) {
  if (seed_ptr != nullptr) seed += uint64_t(*seed_ptr) << 44;
  const int lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float mx = -INFINITY;
  for (int j = lane; j < n; j += 32) mx = fmaxf(mx, ldv(s, row * lds + j, s_dt) * scale);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < n; j += 32) sum += expf(ldv(s, row * lds + j, s_dt) * scale - mx);
  const float inv = 1.0f / warp_sum(sum);
  for (int j = lane; j < ldp; j += 32) {
    float v = 0.f;
    if (j < n) {
      v = expf(ldv(s, row * lds + j, s_dt) * scale - mx) * inv;
      if (drop_p > 0.f)
        v = (hash32(seed + uint64_t(row) * uint64_t(ldp) + j) >> 8) * (1.0f / 16777216.0f) >= drop_p
                ? v / (1.0f - drop_p) : 0.f;
    }
    p[row * ldp + j] = __float2bfloat16_rn(v);
  }
}

// dS = scale * P * (dP' - sum_k P dP'), P = probabilities BEFORE dropout; with attention dropout
// dP' = dP * keep / (1 - p) where keep is the mask softmax_rows_kernel drew for (seed, row, j).
__global__ void __launch_bounds__(256)
softmax_bwd_rows_kernel(const __nv_bfloat16* __restrict__ p, const void* __restrict__ dp, int dp_dt,
                        __nv_bfloat16* __restrict__ ds, int64_t rows, int n, int64_t ldp, int64_t lddp,
                        int64_t ldds, float scale, float drop_p, uint64_t seed,
                        const int64_t* __restrict__ seed_ptr) {
  if (seed_ptr != nullptr) seed += uint64_t(*seed_ptr) << 44;
  const int lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float keep_scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  auto dpe = [&](int j) {
    float v = ldv(dp, row * lddp + j, dp_dt);
    if (drop_p > 0.f)
      v = (hash32(seed + uint64_t(row) * uint64_t(ldp) + j) >> 8) * (1.0f / 16777216.0f) >= drop_p ? v * keep_scale : 0.f;
    return v;
  };
  float dot = 0.f;
  for (int j = lane; j < n; j += 32) dot += bf2f(p[row * ldp + j]) * dpe(j);
  dot = warp_sum(dot);
  for (int j = lane; j < ldds; j += 32) {
    float v = 0.f;
    if (j < n) v = scale * bf2f(p[row * ldp + j]) * (dpe(j) - dot);
    ds[row * ldds + j] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------- deformable sampling bwd
// Gradients of dod_deform_sample w.r.t. value (atomicAdd, fp32), offsets, logits and the raw
// reference-point logits.  One CTA per (image, query); thread = channel; per-(head, point)
// reductions over the head's channels go through shared-memory atomics.
__global__ void __launch_bounds__(512) deform_sample_bwd_kernel(const void* __restrict__ value, int v_dt,
                                         const float* __restrict__ ref, const float* __restrict__ offs,
                                         const float* __restrict__ logits, const void* __restrict__ dout,
                                         int do_dt, float* __restrict__ dvalue, float* __restrict__ dq,
                                         int queries, int heads, int points, int dh, int gh, int gw,
                                         int64_t ldv_, int64_t ldref, int64_t ldoffs, int64_t ldlog,
                                         int64_t lddo, int64_t lddv, int64_t lddq, int ref_is_logit) {
  extern __shared__ float sh[];  // [heads*points] dA (grad wrt attention weight), [heads*points*2] dloc
  const int hp = heads * points;
  float* s_dw = sh;
  float* s_dl = sh + hp;
  for (int i = threadIdx.x; i < 3 * hp; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int64_t row = blockIdx.x;
  const int b = int(row / queries);
  const int d_model = heads * dh;
  const int64_t hw = int64_t(gh) * gw;
  float rx = ref[row * ldref + 0], ry = ref[row * ldref + 1];
  if (ref_is_logit) {
    rx = 1.0f / (1.0f + expf(-rx));
    ry = 1.0f / (1.0f + expf(-ry));
  }
  for (int c = threadIdx.x; c < d_model; c += blockDim.x) {
    const int h = c / dh;
    const float* lg = logits + row * ldlog + h * points;
    const float* of = offs + row * ldoffs + h * points * 2;
    float mx = -INFINITY;
    for (int p = 0; p < points; ++p) mx = fmaxf(mx, lg[p]);
    float den = 0.f;
    for (int p = 0; p < points; ++p) den += expf(lg[p] - mx);
    const float g = ldv(dout, row * lddo + c, do_dt);
    for (int p = 0; p < points; ++p) {
      const float wgt = expf(lg[p] - mx) / den;
      const float ux = rx + of[2 * p + 0], uy = ry + of[2 * p + 1];
      const float lx = fminf(fmaxf(ux, 0.f), 1.f), ly = fminf(fmaxf(uy, 0.f), 1.f);
      const float sx = lx * float(gw - 1), sy = ly * float(gh - 1);
      int x0 = int(floorf(sx)), y0 = int(floorf(sy));
      int x1 = x0 + 1, y1 = y0 + 1;
      x0 = min(max(x0, 0), gw - 1); x1 = min(max(x1, 0), gw - 1);
      y0 = min(max(y0, 0), gh - 1); y1 = min(max(y1, 0), gh - 1);
      const float wx1 = sx - float(x0), wx0 = 1.0f - wx1;
      const float wy1 = sy - float(y0), wy0 = 1.0f - wy1;
      const int64_t i00 = int64_t(y0) * gw + x0, i01 = int64_t(y1) * gw + x0;
      const int64_t i10 = int64_t(y0) * gw + x1, i11 = int64_t(y1) * gw + x1;
      const int64_t vb = int64_t(b) * hw;
      const float v00 = ldv(value, (vb + i00) * ldv_ + c, v_dt), v01 = ldv(value, (vb + i01) * ldv_ + c, v_dt);
      const float v10 = ldv(value, (vb + i10) * ldv_ + c, v_dt), v11 = ldv(value, (vb + i11) * ldv_ + c, v_dt);
      const float samp = v00 * (wx0 * wy0) + v01 * (wx0 * wy1) + v10 * (wx1 * wy0) + v11 * (wx1 * wy1);
      // d value
      const float gw_ = g * wgt;
      atomicAdd(dvalue + (vb + i00) * lddv + c, gw_ * wx0 * wy0);
      atomicAdd(dvalue + (vb + i01) * lddv + c, gw_ * wx0 * wy1);
      atomicAdd(dvalue + (vb + i10) * lddv + c, gw_ * wx1 * wy0);
      atomicAdd(dvalue + (vb + i11) * lddv + c, gw_ * wx1 * wy1);
      // d attention weight (pre-softmax handled below)
      atomicAdd(&s_dw[h * points + p], g * samp);
      // d location: samp depends on sx through wx1 (wx0 = 1 - wx1); the integer corners are
      // piecewise constant.  d samp / d sx = (v10 - v00) wy0 + (v11 - v01) wy1, likewise for sy.
      const float dsx = (v10 - v00) * wy0 + (v11 - v01) * wy1;
      const float dsy = (v01 - v00) * wx0 + (v11 - v10) * wx1;
      // clamp(., 0, 1) passes the gradient only strictly inside (torch.clamp: inclusive bounds pass)
      const float px = (ux >= 0.f && ux <= 1.f) ? float(gw - 1) : 0.f;
      const float py = (uy >= 0.f && uy <= 1.f) ? float(gh - 1) : 0.f;
      atomicAdd(&s_dl[(h * points + p) * 2 + 0], gw_ * dsx * px);
      atomicAdd(&s_dl[(h * points + p) * 2 + 1], gw_ * dsy * py);
    }
  }
  __syncthreads();
  // softmax backward over points, offsets and reference point
  for (int h = threadIdx.x; h < heads; h += blockDim.x) {
    const float* lg = logits + row * ldlog + h * points;
    float mx = -INFINITY;
    for (int p = 0; p < points; ++p) mx = fmaxf(mx, lg[p]);
    float den = 0.f;
    for (int p = 0; p < points; ++p) den += expf(lg[p] - mx);
    float dot = 0.f;
    for (int p = 0; p < points; ++p) dot += expf(lg[p] - mx) / den * s_dw[h * points + p];
    for (int p = 0; p < points; ++p) {
      const float w = expf(lg[p] - mx) / den;
      // dq row layout = the fused query projection: [offsets 2hp | logits hp | ref 2]
      dq[row * lddq + 2 * hp + h * points + p] = w * (s_dw[h * points + p] - dot);
      dq[row * lddq + (h * points + p) * 2 + 0] = s_dl[(h * points + p) * 2 + 0];
      dq[row * lddq + (h * points + p) * 2 + 1] = s_dl[(h * points + p) * 2 + 1];
    }
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    float acc = 0.f;
    for (int i = 0; i < hp; ++i) acc += s_dl[i * 2 + threadIdx.x];
    const float r = threadIdx.x == 0 ? rx : ry;
    dq[row * lddq + 3 * hp + threadIdx.x] = ref_is_logit ? acc * r * (1.0f - r) : acc;
  }
}


// Vectorised form for bf16 values: a thread owns EIGHT adjacent channels of one (image, query) row (one 16-byte
// gather per corner, the per-head scalars evaluated once per group), d value leaves as two 16-byte vector
// reductions per corner (red.global.add.v4.f32) instead of eight scalar atomics, and the per-(head, point)
// sums for d weight / d location are formed WITHOUT atomics: every thread sums its eight channels in order,
// writes the partial to shared memory, and one thread per (row, head, point) adds the head's dh / 8 partials in
// order.  The gradients of the position projections (sums of large cancelling terms) are therefore the same
// bits on every run -- the thread-per-channel kernel above adds 96 channels per head through shared-memory
// atomics in arrival order.
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(256)
deform_sample_bwd_vec_kernel(const __nv_bfloat16* __restrict__ value, const float* __restrict__ ref,
                             const float* __restrict__ offs, const float* __restrict__ logits,
                             const void* __restrict__ dout, int do_dt, float* __restrict__ dvalue,
                             float* __restrict__ dq, int64_t rows, int queries, int heads, int points, int dh, int gh,
                             int gw, int64_t ldv_, int64_t ldref, int64_t ldoffs, int64_t ldlog, int64_t lddo,
                             int64_t lddv, int64_t lddq, int ref_is_logit) {
  extern __shared__ float sh[];
  const int groups = heads * dh / 8, gph = dh / 8;  // 8-channel groups per row / per head
  const int hp = heads * points;
  const int rows_per_cta = blockDim.x / groups;
  float* s_part = sh;                                         // [rows_per_cta][groups][points][3]
  float* s_dw = s_part + rows_per_cta * groups * points * 3;  // [rows_per_cta][hp]
  float* s_dl = s_dw + rows_per_cta * hp;                     // [rows_per_cta][hp][2]
  const int r_in = threadIdx.x / groups;
  const int grp = threadIdx.x - r_in * groups;
  const int64_t row = int64_t(blockIdx.x) * rows_per_cta + r_in;
  const bool live = r_in < rows_per_cta && row < rows;
  float rx = 0.f, ry = 0.f;
  if (live) {
    rx = ref[row * ldref + 0];
    ry = ref[row * ldref + 1];
    if (ref_is_logit) {
      rx = 1.0f / (1.0f + expf(-rx));
      ry = 1.0f / (1.0f + expf(-ry));
    }
    const int c = grp * 8, h = c / dh;
    const int b = int(row / queries);
    const int64_t hw = int64_t(gh) * gw;
    const float* lg = logits + row * ldlog + h * points;
    const float* of = offs + row * ldoffs + h * points * 2;
    float mx = -INFINITY;
    for (int p = 0; p < points; ++p) mx = fmaxf(mx, lg[p]);
    float den = 0.f;
    for (int p = 0; p < points; ++p) den += expf(lg[p] - mx);
    float g[8];
    if (do_dt == DOD_F32) {
      const float4 g0 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dout) + row * lddo + c);
      const float4 g1 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dout) + row * lddo + c + 4);
      g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
    } else {
      const uint4 gq = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(dout) + row * lddo + c);
      const uint32_t gwd[4] = {gq.x, gq.y, gq.z, gq.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        g[2 * i] = __uint_as_float(gwd[i] << 16);
        g[2 * i + 1] = __uint_as_float(gwd[i] & 0xffff0000u);
      }
    }
    for (int p = 0; p < points; ++p) {
      const float wgt = expf(lg[p] - mx) / den;
      const float ux = rx + of[2 * p + 0], uy = ry + of[2 * p + 1];
      const float lx = fminf(fmaxf(ux, 0.f), 1.f), ly = fminf(fmaxf(uy, 0.f), 1.f);
      const float sx = lx * float(gw - 1), sy = ly * float(gh - 1);
      int x0 = int(floorf(sx)), y0 = int(floorf(sy));
      int x1 = x0 + 1, y1 = y0 + 1;
      x0 = min(max(x0, 0), gw - 1); x1 = min(max(x1, 0), gw - 1);
      y0 = min(max(y0, 0), gh - 1); y1 = min(max(y1, 0), gh - 1);
      const float wx1 = sx - float(x0), wx0 = 1.0f - wx1;
      const float wy1 = sy - float(y0), wy0 = 1.0f - wy1;
      const int64_t vb = int64_t(b) * hw;
      const int64_t i00 = vb + int64_t(y0) * gw + x0, i01 = vb + int64_t(y1) * gw + x0;
      const int64_t i10 = vb + int64_t(y0) * gw + x1, i11 = vb + int64_t(y1) * gw + x1;
      const uint4 q00 = __ldg(reinterpret_cast<const uint4*>(value + i00 * ldv_ + c));
      const uint4 q01 = __ldg(reinterpret_cast<const uint4*>(value + i01 * ldv_ + c));
      const uint4 q10 = __ldg(reinterpret_cast<const uint4*>(value + i10 * ldv_ + c));
      const uint4 q11 = __ldg(reinterpret_cast<const uint4*>(value + i11 * ldv_ + c));
      const uint32_t a00[4] = {q00.x, q00.y, q00.z, q00.w}, a01[4] = {q01.x, q01.y, q01.z, q01.w};
      const uint32_t a10[4] = {q10.x, q10.y, q10.z, q10.w}, a11[4] = {q11.x, q11.y, q11.z, q11.w};
      // clamp(., 0, 1) passes the gradient only inside the closed interval (torch.clamp)
      const float px = (ux >= 0.f && ux <= 1.f) ? float(gw - 1) : 0.f;
      const float py = (uy >= 0.f && uy <= 1.f) ? float(gh - 1) : 0.f;
      float d00[8], d01[8], d10[8], d11[8];
      float dw = 0.f, dlx = 0.f, dly = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int w = i >> 1;
        const bool hi = i & 1;
        const float v00 = __uint_as_float(hi ? (a00[w] & 0xffff0000u) : (a00[w] << 16));
        const float v01 = __uint_as_float(hi ? (a01[w] & 0xffff0000u) : (a01[w] << 16));
        const float v10 = __uint_as_float(hi ? (a10[w] & 0xffff0000u) : (a10[w] << 16));
        const float v11 = __uint_as_float(hi ? (a11[w] & 0xffff0000u) : (a11[w] << 16));
        const float samp = v00 * (wx0 * wy0) + v01 * (wx0 * wy1) + v10 * (wx1 * wy0) + v11 * (wx1 * wy1);
        const float gw_ = g[i] * wgt;
        d00[i] = gw_ * wx0 * wy0;
        d01[i] = gw_ * wx0 * wy1;
        d10[i] = gw_ * wx1 * wy0;
        d11[i] = gw_ * wx1 * wy1;
        dw += g[i] * samp;
        const float dsx = (v10 - v00) * wy0 + (v11 - v01) * wy1;
        const float dsy = (v01 - v00) * wx0 + (v11 - v10) * wx1;
        dlx += gw_ * dsx * px;
        dly += gw_ * dsy * py;
      }
      float* o00 = dvalue + i00 * lddv + c;
      float* o01 = dvalue + i01 * lddv + c;
      float* o10 = dvalue + i10 * lddv + c;
      float* o11 = dvalue + i11 * lddv + c;
      red_add_v4(o00, d00[0], d00[1], d00[2], d00[3]); red_add_v4(o00 + 4, d00[4], d00[5], d00[6], d00[7]);
      red_add_v4(o01, d01[0], d01[1], d01[2], d01[3]); red_add_v4(o01 + 4, d01[4], d01[5], d01[6], d01[7]);
      red_add_v4(o10, d10[0], d10[1], d10[2], d10[3]); red_add_v4(o10 + 4, d10[4], d10[5], d10[6], d10[7]);
      red_add_v4(o11, d11[0], d11[1], d11[2], d11[3]); red_add_v4(o11 + 4, d11[4], d11[5], d11[6], d11[7]);
      float* sp = s_part + ((r_in * groups + grp) * points + p) * 3;
      sp[0] = dw;
      sp[1] = dlx;
      sp[2] = dly;
    }
  }
  __syncthreads();
  // one thread per (row, head, point): the head's partials in group order
  for (int i = threadIdx.x; i < rows_per_cta * hp; i += blockDim.x) {
    const int r = i / hp, hpi = i - r * hp, h = hpi / points, p = hpi - h * points;
    float dw = 0.f, dlx = 0.f, dly = 0.f;
    for (int gi = 0; gi < gph; ++gi) {
      const float* sp = s_part + ((r * groups + h * gph + gi) * points + p) * 3;
      dw += sp[0];
      dlx += sp[1];
      dly += sp[2];
    }
    s_dw[r * hp + hpi] = dw;
    s_dl[(r * hp + hpi) * 2 + 0] = dlx;
    s_dl[(r * hp + hpi) * 2 + 1] = dly;
  }
  __syncthreads();
  // softmax backward over the points, offsets (dq row layout = the fused query projection:
  // [offsets 2hp | logits hp | ref 2]), then the reference point
  for (int i = threadIdx.x; i < rows_per_cta * heads; i += blockDim.x) {
    const int r = i / heads, h = i - r * heads;
    const int64_t rw = int64_t(blockIdx.x) * rows_per_cta + r;
    if (rw >= rows) continue;
    const float* lg = logits + rw * ldlog + h * points;
    const float* dwp = s_dw + r * hp + h * points;
    float mx = -INFINITY;
    for (int p = 0; p < points; ++p) mx = fmaxf(mx, lg[p]);
    float den = 0.f;
    for (int p = 0; p < points; ++p) den += expf(lg[p] - mx);
    float dot = 0.f;
    for (int p = 0; p < points; ++p) dot += expf(lg[p] - mx) / den * dwp[p];
    for (int p = 0; p < points; ++p) {
      const float w = expf(lg[p] - mx) / den;
      dq[rw * lddq + 2 * hp + h * points + p] = w * (dwp[p] - dot);
      dq[rw * lddq + (h * points + p) * 2 + 0] = s_dl[(r * hp + h * points + p) * 2 + 0];
      dq[rw * lddq + (h * points + p) * 2 + 1] = s_dl[(r * hp + h * points + p) * 2 + 1];
    }
  }
  for (int i = threadIdx.x; i < rows_per_cta * 2; i += blockDim.x) {
    const int r = i >> 1, xy = i & 1;
    const int64_t rw = int64_t(blockIdx.x) * rows_per_cta + r;
    if (rw >= rows) continue;
    float acc = 0.f;
    for (int k = 0; k < hp; ++k) acc += s_dl[(r * hp + k) * 2 + xy];
    float rr = ref[rw * ldref + xy];
    if (ref_is_logit) rr = 1.0f / (1.0f + expf(-rr));
    dq[rw * lddq + 3 * hp + xy] = ref_is_logit ? acc * rr * (1.0f - rr) : acc;
  }
}

}  // namespace
}  // namespace dod

using namespace dod;

extern "C" int32_t dod_transpose_bf16(const dod_transpose_args* a, dod_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->in && a->out, "dod_transpose_bf16: null pointer");
  DOD_REQUIRE(a->rows > 0 && a->cols > 0 && a->batch > 0 && a->batch <= 65535 && a->ld_in >= a->cols &&
                  a->ld_out >= a->rows,
              "dod_transpose_bf16: bad shape");
  const int64_t inner = a->batch_inner > 1 ? a->batch_inner : 1;
  DOD_REQUIRE(a->batch * inner <= 65535, "dod_transpose_bf16: too many batches");
  dim3 grid(unsigned((a->cols + 63) / 64), unsigned((a->rows + 63) / 64), unsigned(a->batch * inner));
  DOD_REQUIRE(grid.y <= 65535, "dod_transpose_bf16: too many rows");
  transpose_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16*)a->in, (__nv_bfloat16*)a->out,
                                            int(a->rows), int(a->cols), a->ld_in, a->ld_out,
                                            a->batch_stride_in, a->batch_stride_out, int(inner),
                                            a->inner_stride_in);
  int rc = check_cuda(cudaGetLastError(), "transpose_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}

extern "C" int32_t dod_lowrank_wgrad(const dod_lowrank_wgrad_args* a, dod_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->big && a->small && a->out, "dod_lowrank_wgrad: null pointer");
  DOD_REQUIRE(a->m > 0 && a->cols > 0 && a->r > 0 && a->r <= 64, "dod_lowrank_wgrad: need 0 < r <= 64");
  {
    // vectorised kernel, eight columns of `small` per pass (wider variants -- 4 x 16, 2 x 32 accumulators per
    // thread -- measured slower than repeated passes: 864 us vs ~320 us for r = 24 over 3072 columns).  Needs
    // 16-byte-aligned rows of both operands; `small` must be stored at least 8 * ceil(r / 8) wide (the LoRA
    // activations are stored 64 wide).
    const int rp = (a->r + 7) / 8 * 8;
    const bool aligned = a->ld_small % 8 == 0 && a->ld_small >= rp && (uintptr_t(a->small) & 15) == 0 &&
                         a->cols % 8 == 0 && a->ld_big % 8 == 0 && (uintptr_t(a->big) & 15) == 0;
    if (aligned) {
      // rows per CTA: enough CTAs for two per SM, few enough that the final global atomics stay cheap
      const int col_blocks = int((a->cols + 255) / 256);
      const int64_t chunks = (2 * 148 + col_blocks - 1) / col_blocks;
      int64_t rows_per_cta = (a->m + chunks - 1) / chunks;
      rows_per_cta = (rows_per_cta + 63) / 64 * 64;
      dim3 grid(unsigned(col_blocks), unsigned((a->m + rows_per_cta - 1) / rows_per_cta));
      for (int j0 = 0; j0 < a->r; j0 += 8) {
        const int rj = a->r - j0 < 8 ? a->r - j0 : 8;
        float* outp = a->out + (a->transposed ? int64_t(j0) * a->ldo : int64_t(j0));
        lowrank_wgrad_vec_kernel<8, 8><<<grid, 256, 0, stream>>>(
            (const __nv_bfloat16*)a->big, (const __nv_bfloat16*)a->small + j0, outp, a->m, int(a->cols), rj, a->ld_big,
            a->ld_small, a->ldo, a->transposed, a->alpha, int(rows_per_cta));
        int rc = check_cuda(cudaGetLastError(), "lowrank_wgrad_vec_kernel launch");
        if (rc != 0) return rc;
        count_launch();
      }
      return DOD_OK;
    }
  }
  const int rows_per_cta = 512;
  dim3 grid(unsigned((a->cols + 255) / 256), unsigned((a->m + rows_per_cta - 1) / rows_per_cta));
  DOD_REQUIRE(grid.y <= 65535, "dod_lowrank_wgrad: too many rows");
#define DOD_LRW(R)                                                                                   \
  lowrank_wgrad_kernel<R><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)a->big,                   \
                                                    (const __nv_bfloat16*)a->small, a->out, a->m,   \
                                                    int(a->cols), int(a->r), a->ld_big, a->ld_small, \
                                                    a->ldo, a->transposed, a->alpha, rows_per_cta)
  if (a->r <= 8) DOD_LRW(8);
  else if (a->r <= 16) DOD_LRW(16);
  else if (a->r <= 32) DOD_LRW(32);
  else DOD_LRW(64);
#undef DOD_LRW
  int rc = check_cuda(cudaGetLastError(), "lowrank_wgrad_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}

extern "C" int32_t dod_colsum(const dod_colsum_args* a, dod_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->x && a->out, "dod_colsum: null pointer");
  DOD_REQUIRE(a->m > 0 && a->cols > 0 && a->ld >= a->cols, "dod_colsum: bad shape");
  const int rows_per_cta = 256;
  dim3 grid(unsigned((a->cols + 255) / 256), unsigned((a->m + rows_per_cta - 1) / rows_per_cta));
  DOD_REQUIRE(grid.y <= 65535, "dod_colsum: too many rows");
  colsum_kernel<<<grid, 256, 0, stream>>>(a->x, a->x_dtype, a->out, a->m, int(a->cols), a->ld, rows_per_cta);
  int rc = check_cuda(cudaGetLastError(), "colsum_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}

extern "C" int32_t dod_layernorm_bwd(const dod_layernorm_bwd_args* a, dod_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->dy && a->x && a->gamma && a->dx, "dod_layernorm_bwd: null pointer");
  DOD_REQUIRE(a->rows > 0 && a->d > 0 && a->d % 4 == 0 && a->d <= kLnMaxVec * 128,
              "dod_layernorm_bwd: d must be a multiple of 4 and <= %d", kLnMaxVec * 128);
  DOD_REQUIRE((a->dgamma == nullptr) == (a->dbeta == nullptr), "dod_layernorm_bwd: dgamma/dbeta go together");
  DOD_REQUIRE((uintptr_t(a->dy) & 15) == 0 && (uintptr_t(a->x) & 15) == 0 && (uintptr_t(a->dx) & 15) == 0 &&
                  (!a->dres || (uintptr_t(a->dres) & 15) == 0),
              "dod_layernorm_bwd: dy / x / dx / dres must be 16-byte aligned");
  // with dgamma / dbeta: few persistent CTAs (each ends with 2 d global atomics); without: one row per warp
  const int64_t ctas = (a->rows + 7) / 8;
  const unsigned grid = unsigned(a->dgamma && ctas > 148 ? 148 : ctas);
  const size_t smem = a->dgamma ? size_t(2 * a->d) * sizeof(float) : 0;
  layernorm_bwd_kernel<<<grid, 256, smem, stream>>>(
      a->dy, a->dy_dtype, a->x, a->gamma, a->dres, a->dx, a->dgamma, a->dbeta, a->rows, int(a->d), a->eps);
  int rc = check_cuda(cudaGetLastError(), "layernorm_bwd_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}

extern "C" int32_t dod_eltwise(const dod_eltwise_args* a, dod_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->a && a->out, "dod_eltwise: null pointer");
  DOD_REQUIRE(a->mode >= DOD_ELT_CAST && a->mode <= DOD_ELT_AXPBY, "dod_eltwise: bad mode");
  DOD_REQUIRE(a->rows >= 0 && a->cols > 0, "dod_eltwise: bad shape");
  if (a->rows == 0) return DOD_OK;
  const int64_t total = a->rows * ((a->cols + 7) / 8);
  DOD_REQUIRE((total + 255) / 256 < (int64_t(1) << 31), "dod_eltwise: too many elements");
  // 16-byte accesses need every operand row to start 16-byte aligned and whole 8-element groups
  auto vec_ok = [](const void* p, int64_t ld, int dt) {
    return p == nullptr || ((uintptr_t(p) & 15) == 0 && ld % (dt == DOD_BF16 ? 8 : 4) == 0);
  };
  const bool vec = a->cols % 8 == 0 && vec_ok(a->a, a->ld_a, a->a_dtype) && vec_ok(a->b, a->ld_b, a->b_dtype) &&
                   vec_ok(a->out, a->ld_out, a->out_dtype) && vec_ok(a->out2, a->ld_out, a->out2_dtype);
  auto kern = vec ? eltwise_kernel<true> : eltwise_kernel<false>;
  kern<<<unsigned((total + 255) / 256), 256, 0, stream>>>(
      a->mode, a->a, a->a_dtype, a->b, a->b_dtype, a->vec, a->out, a->out_dtype, a->out2, a->out2_dtype,
      a->rows, int(a->cols), a->ld_a, a->ld_b, a->ld_out, a->p0, uint64_t(a->seed), a->seed_ptr);
  int rc = check_cuda(cudaGetLastError(), "eltwise_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}

extern "C" int32_t dod_softmax_rows(const dod_softmax_rows_args* a, dod_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->s && a->p, "dod_softmax_rows: null pointer");
  DOD_REQUIRE(a->rows > 0 && a->n > 0 && a->lds >= a->n && a->ldp >= a->n, "dod_softmax_rows: bad shape");
  softmax_rows_kernel<<<unsigned((a->rows + 7) / 8), 256, 0, stream>>>(
      a->s, a->s_dtype, (__nv_bfloat16*)a->p, a->rows, int(a->n), a->lds, a->ldp, a->scale, a->drop_p,
      uint64_t(a->seed), a->seed_ptr);
  int rc = check_cuda(cudaGetLastError(), "softmax_rows_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}

extern "C" int32_t dod_softmax_bwd_rows(const dod_softmax_bwd_rows_args* a, dod_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->p && a->dp && a->ds, "dod_softmax_bwd_rows: null pointer");
  DOD_REQUIRE(a->rows > 0 && a->n > 0 && a->ldp >= a->n && a->lddp >= a->n && a->ldds >= a->n,
              "dod_softmax_bwd_rows: bad shape");
  softmax_bwd_rows_kernel<<<unsigned((a->rows + 7) / 8), 256, 0, stream>>>(
      (const __nv_bfloat16*)a->p, a->dp, a->dp_dtype, (__nv_bfloat16*)a->ds, a->rows, int(a->n), a->ldp,
      a->lddp, a->ldds, a->scale, a->drop_p, uint64_t(a->seed), a->seed_ptr);
  int rc = check_cuda(cudaGetLastError(), "softmax_bwd_rows_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}

extern "C" int32_t dod_deform_sample_bwd(const dod_deform_sample_bwd_args* a, dod_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->value && a->ref && a->offs && a->logits && a->dout && a->dvalue && a->dqproj,
              "dod_deform_sample_bwd: null pointer");
  DOD_REQUIRE(a->batch > 0 && a->queries > 0 && a->heads > 0 && a->points > 0 && a->head_dim > 0 &&
                  a->grid_h > 0 && a->grid_w > 0,
              "dod_deform_sample_bwd: bad shape");
  const int hp = int(a->heads * a->points);
  DOD_REQUIRE(a->lddq >= 3 * hp + 2, "dod_deform_sample_bwd: dqproj rows must hold 3*H*P + 2 columns");
  const int d_model = int(a->heads * a->head_dim);
  {
    // vectorised kernel: bf16 values, 8-channel groups inside one head, 16-byte-aligned rows of value / dout / dvalue
    const char* e_vec = getenv("DOD_DEFORM_VEC");  // 0: the thread-per-channel kernel (A/B, tests)
    const int groups = d_model / 8;
    const bool do_f32 = a->dout_dtype == DOD_F32;
    if (!(e_vec != nullptr && e_vec[0] == '0') && a->value_dtype == DOD_BF16 && a->head_dim % 8 == 0 &&
        a->ldv % 8 == 0 && (reinterpret_cast<uintptr_t>(a->value) & 15) == 0 && a->lddv % 4 == 0 &&
        (reinterpret_cast<uintptr_t>(a->dvalue) & 15) == 0 && a->lddo % (do_f32 ? 4 : 8) == 0 &&
        (reinterpret_cast<uintptr_t>(a->dout) & 15) == 0 && groups >= 1 && groups <= 256) {
      const int rows_per_cta = 256 / groups;
      const int64_t rows = a->batch * a->queries;
      const unsigned vgrid = unsigned((rows + rows_per_cta - 1) / rows_per_cta);
      const size_t smem = size_t(rows_per_cta) * (size_t(groups) * a->points * 3 + 3 * hp) * sizeof(float);
      if (smem <= 48 * 1024) {
        deform_sample_bwd_vec_kernel<<<vgrid, rows_per_cta * groups, smem, stream>>>(
            (const __nv_bfloat16*)a->value, a->ref, a->offs, a->logits, a->dout, a->dout_dtype, a->dvalue, a->dqproj,
            rows, int(a->queries), int(a->heads), int(a->points), int(a->head_dim), int(a->grid_h), int(a->grid_w),
            a->ldv, a->ldref, a->ldoffs, a->ldlog, a->lddo, a->lddv, a->lddq, a->ref_is_logit);
        int rc = check_cuda(cudaGetLastError(), "deform_sample_bwd_vec_kernel launch");
        if (rc == 0) count_launch();
        return rc;
      }
    }
  }
  const int threads = d_model >= 512 ? 512 : ((d_model + 31) / 32) * 32;
  deform_sample_bwd_kernel<<<unsigned(a->batch * a->queries), threads, 3 * hp * sizeof(float), stream>>>(
      a->value, a->value_dtype, a->ref, a->offs, a->logits, a->dout, a->dout_dtype, a->dvalue, a->dqproj,
      int(a->queries), int(a->heads), int(a->points), int(a->head_dim), int(a->grid_h), int(a->grid_w),
      a->ldv, a->ldref, a->ldoffs, a->ldlog, a->lddo, a->lddv, a->lddq, a->ref_is_logit);
  int rc = check_cuda(cudaGetLastError(), "deform_sample_bwd_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}
