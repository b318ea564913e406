// dod_gemm_bf16 — persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   out = residual + scale[n] * act( A[M,K] . W[N,K]^T (+ A2 . W2^T) + bias[n] )
//
// Layout / pipeline
//   * A and W are both K-major (row-major activations, nn.Linear weights), so
//     both operands are TMA-loaded as [rows x 64] bf16 boxes with the 128-byte
//     swizzle and consumed by tcgen05.mma straight from shared memory.
//   * CTA tile 128 x BN (BN = 256 / 128 / 64), BK = 64; 4..8 smem stages.
//   * warp 0: TMA producer, warp 1: MMA issuer (one lane), warps 2..5: epilogue
//     (TMEM -> registers -> fused bias/act/LayerScale/residual -> global).
//   * two TMEM accumulator stages, so the epilogue of tile i overlaps the
//     mainloop of tile i+1; grid = min(tiles, #SM) persistent CTAs; tiles are
//     walked N-fastest so CTAs running together share A rows through L2.
//   * the optional second K segment (LoRA: A2 = x.A^T, W2 = alpha*B) is simply
//     more k-blocks accumulated into the same TMEM tile.
//
// Replaces cuBLAS calls behind nn.Linear / Conv2d in the reference path (see
// include/dod.h for file:line citations).

#include "common.cuh"
#include "../../include/dod.h"

namespace dod {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kThreads = 192;

struct GemmParams {
  int M, N, K1blocks, K2blocks;
  int tiles_m, tiles_n;
  const float* bias;
  const float* scale;
  const float* residual;
  int64_t ldr;
  void* out;
  int64_t ldo;
  int act;
  int out_f32;
  int patch_rows;
};

template <int BN>
struct SmemLayout {
  static constexpr int kStageA = BM * BK * 2;
  static constexpr int kStageB = BN * BK * 2;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kBarBytes = 1024;
  static constexpr int kTotal = kStages * kStage + kBarBytes + 1024 /*align slack*/;
};

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
            const __grid_constant__ CUtensorMap tm_a2, const __grid_constant__ CUtensorMap tm_w2,
            const GemmParams p) {
  using L = SmemLayout<BN>;
  constexpr int kStages = L::kStages;
  constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : 2 * BN;  // two accumulator stages

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * L::kStage);
  uint64_t* full = bars;                    // [kStages]
  uint64_t* empty = bars + kStages;         // [kStages]
  uint64_t* tfull = bars + 2 * kStages;     // [2]
  uint64_t* tempty = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_a);
    prefetch_tmap(&tm_w);
    if (p.K2blocks) {
      prefetch_tmap(&tm_a2);
      prefetch_tmap(&tm_w2);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.tiles_m * p.tiles_n;
  const int kblocks = p.K1blocks + p.K2blocks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mb = tile / p.tiles_n, nb = tile % p.tiles_n;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = stage_base + s * L::kStage;
          uint8_t* sb = sa + L::kStageA;
          mbar_expect_tx(&full[s], L::kStage);
          if (kb < p.K1blocks) {
            tma_load_2d(sa, &tm_a, &full[s], kb * BK, mb * BM);
            tma_load_2d(sb, &tm_w, &full[s], kb * BK, nb * BN);
          } else {
            tma_load_2d(sa, &tm_a2, &full[s], (kb - p.K1blocks) * BK, mb * BM);
            tma_load_2d(sb, &tm_w2, &full[s], (kb - p.K1blocks) * BK, nb * BN);
          }
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, false, false);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + s * L::kStage);
          const uint32_t sb = sa + L::kStageA;
          const uint64_t da = make_sdesc_sw128(sa, 16, 1024);
          const uint64_t db = make_sdesc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 B inside the 128-B swizzle row: +2 in (addr >> 4)
            umma_ss(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc,
                    (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[s]);  // frees the smem stage when these MMAs retire
          if (kb == kblocks - 1) umma_commit(&tfull[acc]);
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int mb = tile / p.tiles_n, nb = tile % p.tiles_n;
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      mbar_wait(&tfull[acc], acc_ph);
      tc_fence_after();
      const int m = mb * BM + quad * 32 + lane;
      const bool row_ok = m < p.M;
      int64_t out_row = m, res_row = m;
      if (p.patch_rows > 0) {
        out_row = int64_t(m) + m / p.patch_rows + 1;
        res_row = 1 + m % p.patch_rows;
      }
      const uint32_t t_row = tmem_base + (uint32_t(quad * 32) << 16) + acc * BN;

      if (p.act == DOD_ACT_SWIGLU) {
        // columns [0,BN/2) of the tile are gates, [BN/2,BN) the linear halves.
        constexpr int H = BN / 2;
#pragma unroll 1
        for (int c = 0; c < H / 32; ++c) {
          uint32_t g[32], u[32];
          tmem_ld_32x32(t_row + c * 32, g);
          tmem_ld_32x32(t_row + H + c * 32, u);
          tmem_ld_wait();
          if (c == H / 32 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
          }
          const int n_in = nb * BN + c * 32;           // gate column in W space
          const int n_out = nb * H + c * 32;           // output column
          if (row_ok && n_out < p.N / 2) {
            float r[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float gv = __uint_as_float(g[j]), uv = __uint_as_float(u[j]);
              if (p.bias) {
                gv += __ldg(p.bias + n_in + j);
                uv += __ldg(p.bias + n_in + H + j);
              }
              r[j] = silu(gv) * uv;
            }
            if (p.out_f32) {
              float* o = reinterpret_cast<float*>(p.out) + out_row * p.ldo + n_out;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(o + j) = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
            } else {
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + out_row * p.ldo + n_out;
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 v;
                v.x = pack_bf16x2(r[j], r[j + 1]);
                v.y = pack_bf16x2(r[j + 2], r[j + 3]);
                v.z = pack_bf16x2(r[j + 4], r[j + 5]);
                v.w = pack_bf16x2(r[j + 6], r[j + 7]);
                *reinterpret_cast<uint4*>(o + j) = v;
              }
            }
          }
        }
        continue;
      }

#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + c * 32, v);
        tmem_ld_wait();
        if (c == BN / 32 - 1) {
          // accumulator fully read: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        const int n0 = nb * BN + c * 32;
        if (!row_ok || n0 >= p.N) continue;
        float r[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __uint_as_float(v[j]);
        const bool full_chunk = (n0 + 32 <= p.N);
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (full_chunk || n0 + j < p.N) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
              r[j] += b.x; r[j + 1] += b.y; r[j + 2] += b.z; r[j + 3] += b.w;
            }
          }
        }
        if (p.act == DOD_ACT_GELU_ERF) {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = gelu_erf(r[j]);
        } else if (p.act == DOD_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = fmaxf(r[j], 0.0f);
        }
        if (p.scale) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (full_chunk || n0 + j < p.N) {
              const float4 s = __ldg(reinterpret_cast<const float4*>(p.scale + n0 + j));
              r[j] *= s.x; r[j + 1] *= s.y; r[j + 2] *= s.z; r[j + 3] *= s.w;
            }
          }
        }
        if (p.residual) {
          const float* rp = reinterpret_cast<const float*>(p.residual) + res_row * p.ldr + n0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (full_chunk || n0 + j < p.N) {
              const float4 q = *reinterpret_cast<const float4*>(rp + j);
              r[j] += q.x; r[j + 1] += q.y; r[j + 2] += q.z; r[j + 3] += q.w;
            }
          }
        }
        if (p.out_f32) {
          float* o = reinterpret_cast<float*>(p.out) + out_row * p.ldo + n0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (full_chunk || n0 + j < p.N)
              *reinterpret_cast<float4*>(o + j) = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
          }
        } else {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + out_row * p.ldo + n0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (full_chunk || n0 + j < p.N) {
              uint4 q;
              q.x = pack_bf16x2(r[j], r[j + 1]);
              q.y = pack_bf16x2(r[j + 2], r[j + 3]);
              q.z = pack_bf16x2(r[j + 4], r[j + 5]);
              q.w = pack_bf16x2(r[j + 6], r[j + 7]);
              *reinterpret_cast<uint4*>(o + j) = q;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

template <int BN>
int launch(const dod_gemm_args& a, cudaStream_t stream) {
  using L = SmemLayout<BN>;
  static bool attr_set = false;  // benign race: idempotent attribute
  if (!attr_set) {
    DOD_CUDA_OK(cudaFuncSetAttribute(gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     L::kTotal));
    attr_set = true;
  }
  CUtensorMap tm_a, tm_w, tm_a2, tm_w2;
  if (int rc = make_tmap_2d(&tm_a, a.a, 2, a.m, a.k, a.lda, BM, BK)) return rc;
  if (int rc = make_tmap_2d(&tm_w, a.w, 2, a.n, a.k, a.ldw, BN, BK)) return rc;
  if (a.a2) {
    if (int rc = make_tmap_2d(&tm_a2, a.a2, 2, a.m, a.k2, a.lda2, BM, BK)) return rc;
    if (int rc = make_tmap_2d(&tm_w2, a.w2, 2, a.n, a.k2, a.ldw2, BN, BK)) return rc;
  } else {
    tm_a2 = tm_a;
    tm_w2 = tm_w;
  }
  GemmParams p;
  p.M = int(a.m);
  p.N = int(a.n);
  p.K1blocks = int((a.k + BK - 1) / BK);
  p.K2blocks = a.a2 ? int((a.k2 + BK - 1) / BK) : 0;
  p.tiles_m = int((a.m + BM - 1) / BM);
  p.tiles_n = int((a.n + BN - 1) / BN);
  p.bias = a.bias;
  p.scale = a.scale;
  p.residual = reinterpret_cast<const float*>(a.residual);
  p.ldr = a.ldr;
  p.out = a.out;
  p.ldo = a.ldo;
  p.act = a.act;
  p.out_f32 = a.out_dtype == DOD_F32;
  p.patch_rows = a.patch_rows;
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_kernel<BN><<<grid, kThreads, L::kTotal, stream>>>(tm_a, tm_w, tm_a2, tm_w2, p);
  return check_cuda(cudaGetLastError(), "gemm_kernel launch");
}

}  // namespace

void count_launch(int n = 1);

}  // namespace dod

extern "C" int32_t dod_gemm_bf16(const dod_gemm_args* a, dod_stream_t stream_) {
  using namespace dod;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->a && a->w && a->out, "dod_gemm_bf16: null pointer");
  DOD_REQUIRE(a->m > 0 && a->n > 0 && a->k > 0, "dod_gemm_bf16: empty problem m=%lld n=%lld k=%lld",
              (long long)a->m, (long long)a->n, (long long)a->k);
  DOD_REQUIRE(a->m < (1ll << 31) && a->n < (1ll << 31) && a->k < (1ll << 31),
              "dod_gemm_bf16: dimension too large");
  DOD_REQUIRE(a->lda % 8 == 0 && a->ldw % 8 == 0 && a->lda >= a->k && a->ldw >= a->k,
              "dod_gemm_bf16: lda/ldw must be >= k and multiples of 8 (16-byte TMA strides)");
  DOD_REQUIRE((uintptr_t(a->a) & 15) == 0 && (uintptr_t(a->w) & 15) == 0 &&
                  (uintptr_t(a->out) & 15) == 0,
              "dod_gemm_bf16: a/w/out must be 16-byte aligned");
  DOD_REQUIRE(a->n % 8 == 0, "dod_gemm_bf16: n must be a multiple of 8 (got %lld)", (long long)a->n);
  if (a->a2 || a->w2) {
    DOD_REQUIRE(a->a2 && a->w2 && a->k2 > 0 && a->lda2 % 8 == 0 && a->ldw2 % 8 == 0 &&
                    a->lda2 >= a->k2 && a->ldw2 >= a->k2,
                "dod_gemm_bf16: bad second K segment");
    DOD_REQUIRE((uintptr_t(a->a2) & 15) == 0 && (uintptr_t(a->w2) & 15) == 0,
                "dod_gemm_bf16: a2/w2 must be 16-byte aligned");
  }
  DOD_REQUIRE(a->act >= DOD_ACT_NONE && a->act <= DOD_ACT_SWIGLU, "dod_gemm_bf16: bad act");
  DOD_REQUIRE(a->out_dtype == DOD_BF16 || a->out_dtype == DOD_F32, "dod_gemm_bf16: bad out_dtype");
  const int64_t out_cols = a->act == DOD_ACT_SWIGLU ? a->n / 2 : a->n;
  DOD_REQUIRE(a->ldo >= out_cols && a->ldo % (a->out_dtype == DOD_F32 ? 4 : 8) == 0,
              "dod_gemm_bf16: ldo must cover the output row and keep 16-byte alignment");
  if (a->residual)
    DOD_REQUIRE(a->ldr % 4 == 0 && (uintptr_t(a->residual) & 15) == 0,
                "dod_gemm_bf16: residual must be 16-byte aligned with ldr %% 4 == 0");
  if (a->bias) DOD_REQUIRE((uintptr_t(a->bias) & 15) == 0, "dod_gemm_bf16: bias alignment");
  if (a->scale) DOD_REQUIRE((uintptr_t(a->scale) & 15) == 0, "dod_gemm_bf16: scale alignment");
  int rc;
  if (a->act == DOD_ACT_SWIGLU) {
    DOD_REQUIRE(a->n % 256 == 0 && !a->scale && !a->residual && a->patch_rows == 0,
                "dod_gemm_bf16: SWIGLU needs n %% 256 == 0 and no scale/residual");
    rc = launch<256>(*a, stream);
  } else if (a->n > 128) {
    rc = launch<256>(*a, stream);
  } else if (a->n > 64) {
    rc = launch<128>(*a, stream);
  } else {
    rc = launch<64>(*a, stream);
  }
  if (rc == 0) count_launch();
  return rc;
}
