// dod_gemm_bf16 — persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   out = residual + scale[n] * act( A[M,K] . W[N,K]^T (+ A2 . W2^T) + bias[n] )
//
// Layout / pipeline
//   * A and W are both K-major (row-major activations, nn.Linear weights), so
//     both operands are TMA-loaded as [rows x 64] bf16 boxes with the 128-byte
//     swizzle and consumed by tcgen05.mma straight from shared memory.
//   * CTA tile 128 x BN (BN = 256 / 128 / 64), BK = 64; 3..6 smem stages.
//   * warp 0: TMA producer (one lane), warp 1: MMA issuer (converged warp, elected
//     tcgen05 instructions), warps 2..9: epilogue.
//   * two TMEM accumulator stages, so the epilogue of tile i overlaps the
//     mainloop of tile i+1; grid = min(tiles, #SM) persistent CTAs; tiles are
//     walked N-fastest so CTAs running together share A rows through L2.
//   * the optional second K segment (LoRA: A2 = x.A^T, W2 = alpha*B) is simply
//     more k-blocks accumulated into the same TMEM tile.
//   * epilogue: each TMEM lane quadrant (32 rows) is served by TWO warps that
//     alternate over 64-byte-wide column chunks.  A chunk goes TMEM -> registers
//     (one row per thread) -> bias / act / LayerScale / +residual -> 64B-swizzled
//     shared memory -> TMA store, so global writes are full coalesced lines; the
//     fp32 residual chunk arrives the same way through a TMA load that is
//     prefetched one chunk ahead.  (The first version stored rows straight from
//     registers: 32 lanes x 16 B scattered over 32 rows per instruction made the
//     K=768 GEMMs epilogue-bound at 16-43 % tensor-pipe utilisation, see
//     profiles/r01_summary.md.)
//   * CTA-pair kernel (gemm2_kernel, M >= 512 and N >= 256): without a residual the warp stores PAIRS of
//     adjacent chunks as one 4 KB box of 128-byte rows; with the fp32 residual it runs an in-place ring of
//     three chunk buffers per warp (TMA load -> add -> TMA store from the same buffer) that is filled two
//     chunks ahead across tile boundaries.
//   * LayerNorm folded across two GEMMs (include/dod.h): the residual epilogue can also emit bf16(out) and
//     per-row partial sums of out / out^2 (producer); any TMA-store epilogue can multiply the accumulator
//     by a per-row scale loaded one tile ahead (consumer: rstd, with gamma and the mean removal folded
//     into the weights).
//   * the patch-embedding row map (patch_rows > 0) cannot be expressed as a TMA
//     box and keeps the direct register -> global path (0.4 % of the FLOPs).
//
// Replaces cuBLAS calls behind nn.Linear / Conv2d in the reference path (see
// include/dod.h for file:line citations).

#include <cstdlib>

#include "common.cuh"
#include "../../include/dod.h"

namespace dod {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + kEpiWarps * 32;
constexpr int kChunkBytes = 32 * 64;  // one epilogue chunk buffer: 32 rows x 64 B
#ifdef DOD_GEMM_NARROW
constexpr bool kWideStores = false;  // A/B builds (tools/build_variant.sh narrow gemm.cu -DDOD_GEMM_NARROW)
#else
constexpr bool kWideStores = true;   // CTA-pair kernel, no residual: 128-byte-row TMA stores
#endif

struct GemmParams {
  int M, N, K1blocks, K2blocks;
  int tiles_m, tiles_n, batch, batch_inner;
  const float* bias;
  const float* scale;
  const float* residual;
  int64_t ldr;
  void* out;
  int64_t ldo;
  int act;
  int out_f32;
  int patch_rows;
  int direct;  // 1: register -> global epilogue (patch rows / shapes TMA cannot store)
  int a_trans, w_trans;  // operand stored [K, M] / [K, N]: MN-major smem tiles (64-column blocks 8 KB apart)
  int w_shared;          // batched problem whose W (and residual) is the same for every batch entry
  // LayerNorm folded across two GEMMs (include/dod.h): producer outputs / consumer inputs
  __nv_bfloat16* out2;
  int64_t ldo2;
  float* stats_out;
  const float* row_scale;  // consumer: f32 [M], accumulator of row m is multiplied by it before bias / act
};

// consumer side of the folded LayerNorm: rstd of this thread's row.  Requested one TILE ahead: the fc1 GEMM
// is epilogue-bound, so a load issued at the start of a tile's epilogue is an L2 round trip (~1000 clk) on the
// critical path of every tile (ncu: long-scoreboard stalls +40 %, fc1 +14 %)
__device__ __forceinline__ float load_row_scale(const GemmParams& p, int row) {
  float v = 0.0f;
  if (row < p.M) asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p.row_scale + row));
  return v;
}

// MN-major operand tile in shared memory: [MN / 64 blocks][64 K rows][64 MN elements = 128 B],
// 128B-swizzled; one TMA box per block.  tcgen05 descriptor: LBO = 8192 (between 64-wide MN blocks),
// SBO = 1024 (between 8-row K groups); a 16-deep K step advances the start address by 16 * 128 B.
constexpr int kMnBlockBytes = 64 * BK * 2;

// two batch levels (outer x inner) of a problem; a level that is not used still needs a valid
// (16-byte multiple, non-zero) TMA stride
struct BatchDims {
  uint64_t outer, inner, os_a, os_w, os_o, is_a, is_w, is_o;
  explicit BatchDims(const dod_gemm_args& a) {
    outer = a.batch > 1 ? uint64_t(a.batch) : 1;
    inner = a.batch_inner > 1 ? uint64_t(a.batch_inner) : 1;
    const int64_t n_out = a.act == DOD_ACT_SWIGLU ? a.n / 2 : a.n;
    (void)n_out;
    is_a = inner > 1 ? uint64_t(a.inner_stride_a) : uint64_t(a.m) * a.lda;
    is_w = inner > 1 ? uint64_t(a.inner_stride_w) : uint64_t(a.n) * a.ldw;
    is_o = inner > 1 ? uint64_t(a.inner_stride_out) : uint64_t(a.m) * a.ldo;
    os_a = outer > 1 ? uint64_t(a.batch_stride_a) : uint64_t(a.m) * a.lda;
    os_w = outer > 1 ? uint64_t(a.batch_stride_w) : uint64_t(a.n) * a.ldw;
    os_o = outer > 1 ? uint64_t(a.batch_stride_out) : uint64_t(a.m) * a.ldo;
  }
};

template <int BN, bool RES>
struct SmemLayout {
  static constexpr int kStageA = BM * BK * 2;
  static constexpr int kStageB = BN * BK * 2;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kStages = (BN == 256) ? (RES ? 3 : 4) : (BN == 128 ? 5 : 6);
  static constexpr int kEpiBufs = RES ? 4 : 2;  // per warp: 2 out (+ 2 residual)
  static constexpr int kEpiBytes = kEpiWarps * kEpiBufs * kChunkBytes;
  static constexpr int kBarBytes = 512;
  static constexpr int kTotal = kStages * kStage + kEpiBytes + kBarBytes + 1024 /*align slack*/;
  static_assert(kTotal <= 232448, "shared memory budget exceeded");
};

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// GELU for bf16 outputs.  0.5 x (1 + erf(x/sqrt2)) with
//   erf(z) ~ tanh(a),  a = z (c0 + c1 z^2 + c2 z^4 + c3 z^6)      (max |err| 5.5e-5, monotone in z)
// i.e. gelu(x) = hx + hx * tanh(x * P(x^2)), hx = x/2 (constants below have 1/sqrt2 folded in).  Formula
// error 8.5e-5 abs; tanh.approx.f32 (one MUFU op, rel. error 2^-11) adds <= 2.5e-4 |x| -- together
// under 1/8 of the bf16 rounding of the result.  Two packed elements per instruction (Blackwell
// f32x2 FMA/MUL) and ONE MUFU op per element: ~3.5 issue slots per element instead of ~33 for erff,
// which kept the fc1 epilogue above the MMA time per tile (profiles/r01_summary.md; an exp2 + rcp
// sigmoid form with two MUFU ops per element still left fc1 at 59 % tensor-pipe).  fp32 outputs
// (fp32 mode) keep erff.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float2 gelu_bf16_x2(float2 x) {
  const float2 s = __fmul2_rn(x, x);
  float2 q = __ffma2_rn(s, make_float2(8.668819692e-06f, 8.668819692e-06f),
                        make_float2(-3.982046110e-04f, -3.982046110e-04f));
  q = __ffma2_rn(q, s, make_float2(3.698001081e-02f, 3.698001081e-02f));
  q = __ffma2_rn(q, s, make_float2(7.976396815e-01f, 7.976396815e-01f));
  const float2 a = __fmul2_rn(q, x);
  const float2 hx = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  return __ffma2_rn(hx, make_float2(tanh_approx(a.x), tanh_approx(a.y)), hx);
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

template <int COLS>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[COLS]) {
  if constexpr (COLS == 16) tmem_ld_32x16(taddr, v);
  else tmem_ld_32x32(taddr, v);
}

// byte offset of logical 16-byte chunk j of row r inside a 64B-swizzled [32 x 64 B] buffer
__device__ __forceinline__ uint32_t sw64(int r, int j) { return r * 64 + ((j ^ ((r >> 1) & 3)) << 4); }

// Epilogue of one tile through shared memory, in chunks of COLS output columns
// (16 fp32 or 32 bf16 = 64 bytes per row).
// WIDE: the warp handles PAIRS of adjacent chunks and stores them with one TMA box of 128-byte rows (4 KB,
// 128B-swizzled) instead of two boxes of 64-byte rows.  The TMA unit's cost is per ROW, not per byte (~2.7
// clk per row for 64 B and 128 B rows alike -- the figure that also paces the mainloop's operand loads), so
// halving the epilogue's row count shortens every tile (profiles/r01_summary.md).
template <int BN, bool RES, bool OUT_F32, bool WIDE = false>
__device__ __forceinline__ void epilogue_tile_tma(const GemmParams& p, const CUtensorMap* tm_out,
                                                  const CUtensorMap* tm_res, uint32_t t_row, int bz,
                                                  int mb, int nb, int quad, int half, int lane,
                                                  uint8_t* out_buf, uint8_t* res_buf,
                                                  uint64_t* res_full, uint32_t& out_cnt,
                                                  uint32_t& res_issue, uint32_t& res_wait,
                                                  uint32_t tempty_addr, uint64_t* tfull_bar,
                                                  uint32_t tfull_phase, float ln_rstd = 0.0f) {
  static_assert(!(RES && WIDE), "the wide store path has no residual prefetch");
  constexpr int COLS = OUT_F32 ? 16 : 32;
  constexpr int STEP = WIDE ? 4 : 2;  // chunks between the starts of this warp's consecutive groups
  constexpr int SUBS = WIDE ? 2 : 1;  // chunks per group (= per TMA store)
  const bool swiglu = p.act == DOD_ACT_SWIGLU;
  const int width = swiglu ? BN / 2 : BN;  // output columns of this tile
  const int nch = width / COLS;
  const int row0 = mb * BM + quad * 32;
  const int col_base = nb * width;
  const int n_out = swiglu ? p.N / 2 : p.N;
  if constexpr (RES) {
    if (half < nch) {
      if (lane == 0) {
        uint64_t* bar = &res_full[res_issue & 1];
        mbar_expect_tx(bar, kChunkBytes);
        tma_load_2d(res_buf + (res_issue & 1) * kChunkBytes, tm_res, bar, col_base + half * COLS, row0);
      }
      ++res_issue;
    }
  }

  // the accumulator is awaited HERE, after the caller requested the row statistics of the folded LayerNorm
  mbar_wait(tfull_bar, tfull_phase);
  tc_fence_after();

  bool group_ok = false;  // WIDE: the group's first chunk is inside N and M (a store will be issued)
  uint8_t* ob = out_buf;
#pragma unroll 1
  for (int c0 = half * SUBS; c0 < nch; c0 += STEP)
#pragma unroll 1
  for (int sub = 0; sub < SUBS; ++sub) {
    const int c = c0 + sub;
    const int n0 = col_base + c * COLS;  // first output column of the chunk
    if constexpr (RES) {
      if (c + 2 < nch) {
        if (lane == 0) {
          uint64_t* bar = &res_full[res_issue & 1];
          mbar_expect_tx(bar, kChunkBytes);
          tma_load_2d(res_buf + (res_issue & 1) * kChunkBytes, tm_res, bar, n0 + 2 * COLS, row0);
        }
        ++res_issue;
      }
    }
    uint32_t v[COLS];
    uint32_t u[COLS];
    tmem_ld_cols<COLS>(t_row + c * COLS, v);
    if (swiglu) tmem_ld_cols<COLS>(t_row + BN / 2 + c * COLS, u);
    tmem_ld_wait();
    if (c0 + STEP >= nch && sub == SUBS - 1) {
      // this warp has read its share of the accumulator: hand the TMEM stage back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_addr);
    }
    float r[COLS];
#pragma unroll
    for (int j = 0; j < COLS; ++j) r[j] = __uint_as_float(v[j]);
    const bool in_range = n0 < n_out;  // chunk entirely outside N: nothing to compute or store
    if (swiglu) {
      // accumulator columns [0, BN/2) are gates, [BN/2, BN) the linear halves (interleaved W rows)
      const int nw = nb * BN + c * COLS;
#pragma unroll
      for (int j = 0; j < COLS; ++j) {
        float gv = r[j], uv = __uint_as_float(u[j]);
        if (p.row_scale) {
          gv *= ln_rstd;
          uv *= ln_rstd;
        }
        if (p.bias) {
          gv += __ldg(p.bias + nw + j);
          uv += __ldg(p.bias + nw + BN / 2 + j);
        }
        r[j] = silu(gv) * uv;
      }
    } else if (in_range) {
      if (p.row_scale) {
        // folded LayerNorm: the rows of W'' sum to zero, so acc is already (h - mean) . (gamma W)^T
#pragma unroll
        for (int j = 0; j < COLS; j += 4) {
          if (n0 + j < p.N) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
            r[j] = fmaf(r[j], ln_rstd, b.x);
            r[j + 1] = fmaf(r[j + 1], ln_rstd, b.y);
            r[j + 2] = fmaf(r[j + 2], ln_rstd, b.z);
            r[j + 3] = fmaf(r[j + 3], ln_rstd, b.w);
          }
        }
      } else if (p.bias) {
#pragma unroll
        for (int j = 0; j < COLS; j += 4) {
          if (n0 + j < p.N) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
            r[j] += b.x; r[j + 1] += b.y; r[j + 2] += b.z; r[j + 3] += b.w;
          }
        }
      }
      if (p.act == DOD_ACT_GELU_ERF) {
        if constexpr (OUT_F32) {
#pragma unroll
          for (int j = 0; j < COLS; ++j) r[j] = gelu_erf(r[j]);
        } else {
#pragma unroll
          for (int j = 0; j < COLS; j += 2) {
            const float2 g = gelu_bf16_x2(make_float2(r[j], r[j + 1]));
            r[j] = g.x;
            r[j + 1] = g.y;
          }
        }
      } else if (p.act == DOD_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < COLS; ++j) r[j] = fmaxf(r[j], 0.0f);
      }
      if (p.scale) {
#pragma unroll
        for (int j = 0; j < COLS; j += 4) {
          if (n0 + j < p.N) {
            const float4 s = __ldg(reinterpret_cast<const float4*>(p.scale + n0 + j));
            r[j] *= s.x; r[j + 1] *= s.y; r[j + 2] *= s.z; r[j + 3] *= s.w;
          }
        }
      }
    }
    if constexpr (RES) {
      mbar_wait(&res_full[res_wait & 1], (res_wait >> 1) & 1);
      const uint8_t* rb = res_buf + (res_wait & 1) * kChunkBytes;
      ++res_wait;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 q = *reinterpret_cast<const float4*>(rb + sw64(lane, j));
        r[4 * j] += q.x; r[4 * j + 1] += q.y; r[4 * j + 2] += q.z; r[4 * j + 3] += q.w;
      }
    }
    // Nothing to store (chunk past N, or this warp's rows past M): leave the staging buffers alone.
    // Rotating them without committing a store let a later chunk overwrite a buffer whose TMA store
    // was still in flight (wait_group.read<1> only covers COMMITTED groups): the last valid chunk of a
    // mostly-empty N tile was then written as zeros.
    if constexpr (!WIDE) {
      if (!in_range || row0 >= p.M) continue;
      // the store issued two chunks ago read this buffer: wait until it has been drained
      if (lane == 0) tma_store_wait_read<1>();
      __syncwarp();
      ob = out_buf + (out_cnt & 1) * kChunkBytes;
      ++out_cnt;
      if constexpr (OUT_F32) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(ob + sw64(lane, j)) =
              make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 q;
          q.x = pack_bf16x2(r[8 * j], r[8 * j + 1]);
          q.y = pack_bf16x2(r[8 * j + 2], r[8 * j + 3]);
          q.z = pack_bf16x2(r[8 * j + 4], r[8 * j + 5]);
          q.w = pack_bf16x2(r[8 * j + 6], r[8 * j + 7]);
          *reinterpret_cast<uint4*>(ob + sw64(lane, j)) = q;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(tm_out, ob, n0, row0, bz % p.batch_inner, bz / p.batch_inner);
        tma_store_commit();
      }
    } else {
      if (sub == 0) {
        group_ok = in_range && row0 < p.M;
        if (group_ok) {
          // the store issued two groups ago read this 4 KB buffer: wait until it has been drained
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          ob = out_buf + (out_cnt & 1) * (2 * kChunkBytes);
          ++out_cnt;
        }
      }
      if (group_ok && in_range) {
        // 128-byte rows, 128B swizzle: 16-byte piece jj of row r lives at r * 128 + ((jj ^ (r & 7)) << 4);
        // this chunk fills pieces sub * 4 .. sub * 4 + 3 of the warp's 32 rows
        uint8_t* rowp = ob + lane * 128;
        const int x = lane & 7;
        if constexpr (OUT_F32) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(rowp + (((sub * 4 + j) ^ x) << 4)) =
                make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 q;
            q.x = pack_bf16x2(r[8 * j], r[8 * j + 1]);
            q.y = pack_bf16x2(r[8 * j + 2], r[8 * j + 3]);
            q.z = pack_bf16x2(r[8 * j + 4], r[8 * j + 5]);
            q.w = pack_bf16x2(r[8 * j + 6], r[8 * j + 7]);
            *reinterpret_cast<uint4*>(rowp + (((sub * 4 + j) ^ x) << 4)) = q;
          }
        }
      }
      if (sub == SUBS - 1 && group_ok) {
        // columns of the group past N (second chunk out of range) are clipped by the TMA store
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(tm_out, ob, col_base + c0 * COLS, row0, bz % p.batch_inner, bz / p.batch_inner);
          tma_store_commit();
        }
      }
    }
  }
}

// Residual epilogue of the CTA-pair kernel (fp32 output, BN = 256): the fp32 residual is the larger HBM
// stream of the K = 768 projection GEMM, and one 2 KB chunk in flight per warp (16 KB per SM) did not
// cover DRAM latency (proj ran at 4.1 TB/s of HBM traffic).  Each warp therefore runs a ring of
// kResRing chunk buffers that is used IN PLACE: TMA load of the residual chunk -> add accumulator ->
// written back to the same swizzled buffer -> TMA store from it.  The load for chunk k + kResRing - 1 is
// issued when chunk k has been stored, into the buffer of chunk k - 1 once its store has been read, and
// the stream of chunks continues ACROSS the warp's tiles, so kResRing - 1 chunks (4 KB per warp, 32 KB
// per SM) are always in flight, also while the warp waits for the next accumulator.  The fourth 2 KB of
// the warp's staging area holds two 1 KB buffers for the optional bf16 copy of the output (folded
// LayerNorm, include/dod.h), stored by TMA in the same bulk group as the fp32 chunk.
constexpr int kResRing = 3;
constexpr int kHalfChunkBytes = kChunkBytes / 2;  // bf16 copy of a 16-column chunk: 32 rows x 32 B

struct ResStream {
  const CUtensorMap* tm;
  uint8_t* buf;    // kResRing chunk buffers of this warp
  uint8_t* buf16;  // two half-chunk buffers (bf16 copy)
  uint64_t* bars;  // kResRing full barriers of this warp
  int tile, c, num_tiles, num_pairs, tiles_n, tiles_per_batch, rank, half;
  int row0, col_base;
  uint32_t issued;

  // the residual is shared by every batch entry (position embedding of the patch GEMM): rows within the batch
  __device__ __forceinline__ void locate(int quad) {
    const int tl = tile % tiles_per_batch;
    row0 = ((tl / tiles_n) * 2 + rank) * BM + quad * 32;
    col_base = (tl % tiles_n) * 256;
  }
  // whole warp calls; lane 0 issues the load of the next chunk of the stream
  __device__ __forceinline__ void issue(int lane, int quad) {
    if (tile >= num_tiles) return;
    if (lane == 0) {
      const uint32_t slot = issued % kResRing;
      mbar_expect_tx(&bars[slot], kChunkBytes);
      tma_load_2d(buf + slot * kChunkBytes, tm, &bars[slot], col_base + c * 16, row0);
    }
    ++issued;
    c += 2;
    if (c >= 16) {
      c = half;
      tile += num_pairs;
      locate(quad);
    }
  }
};

__device__ __forceinline__ void epilogue_tile_res_ring(const GemmParams& p, const CUtensorMap* tm_out,
                                                       const CUtensorMap* tm_out2, uint32_t t_row, int bz, int mb, int nb, int quad, int half,
                                                       int lane, ResStream& rs, uint32_t& res_wait,
                                                       uint32_t tempty_addr) {
  constexpr int COLS = 16, NCH = 256 / COLS;
  const int row0 = mb * BM + quad * 32;
  float st1 = 0.0f, st2 = 0.0f;  // folded LayerNorm: sums of h and h^2 over this warp's 128 columns
#pragma unroll 1
  for (int c = half; c < NCH; c += 2) {
    const int n0 = nb * 256 + c * COLS;
    uint32_t v[COLS];
    tmem_ld_32x16(t_row + c * COLS, v);
    tmem_ld_wait();
    if (c + 2 >= NCH) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_addr);
    }
    float r[COLS];
#pragma unroll
    for (int j = 0; j < COLS; ++j) r[j] = __uint_as_float(v[j]);
    const bool in_range = n0 < p.N;
    if (in_range) {
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < COLS; j += 4) {
          if (n0 + j < p.N) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
            r[j] += b.x; r[j + 1] += b.y; r[j + 2] += b.z; r[j + 3] += b.w;
          }
        }
      }
      if (p.act == DOD_ACT_GELU_ERF) {
#pragma unroll
        for (int j = 0; j < COLS; ++j) r[j] = gelu_erf(r[j]);
      } else if (p.act == DOD_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < COLS; ++j) r[j] = fmaxf(r[j], 0.0f);
      }
      if (p.scale) {
#pragma unroll
        for (int j = 0; j < COLS; j += 4) {
          if (n0 + j < p.N) {
            const float4 s = __ldg(reinterpret_cast<const float4*>(p.scale + n0 + j));
            r[j] *= s.x; r[j + 1] *= s.y; r[j + 2] *= s.z; r[j + 3] *= s.w;
          }
        }
      }
    }
    const uint32_t slot = res_wait % kResRing;
    mbar_wait(&rs.bars[slot], (res_wait / kResRing) & 1);
    ++res_wait;
    uint8_t* rb = rs.buf + slot * kChunkBytes;
    // every lane reads and rewrites only its own row of the buffer: no cross-lane hazard
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float4* q = reinterpret_cast<float4*>(rb + sw64(lane, j));
      const float4 x = *q;
      r[4 * j] += x.x; r[4 * j + 1] += x.y; r[4 * j + 2] += x.z; r[4 * j + 3] += x.w;
      *q = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    }
    uint8_t* hb = rs.buf16 + (res_wait & 1) * kHalfChunkBytes;
    if (p.out2 != nullptr && in_range) {
      // bf16 copy of the new residual stream (the next projection's A operand) + LayerNorm partial sums;
      // 32 B per row, 32B-swizzled: 16-byte piece j of row r at r * 32 + ((j ^ ((r >> 2) & 1)) << 4)
      if (row0 + lane < p.M) {
#pragma unroll
        for (int j = 0; j < COLS; ++j) {
          st1 += r[j];
          st2 = fmaf(r[j], r[j], st2);
        }
      }
      const int sw = (lane >> 2) & 1;
      *reinterpret_cast<uint4*>(hb + lane * 32 + (sw << 4)) =
          make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]),
                     pack_bf16x2(r[6], r[7]));
      *reinterpret_cast<uint4*>(hb + lane * 32 + ((sw ^ 1) << 4)) =
          make_uint4(pack_bf16x2(r[8], r[9]), pack_bf16x2(r[10], r[11]), pack_bf16x2(r[12], r[13]),
                     pack_bf16x2(r[14], r[15]));
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (in_range && row0 < p.M) {
        tma_store_4d(tm_out, rb, n0, row0, bz % p.batch_inner, bz / p.batch_inner);
        if (p.out2 != nullptr) tma_store_2d(tm_out2, hb, n0, row0);
      }
      tma_store_commit();  // one group per chunk (possibly empty) keeps the wait_group arithmetic uniform
      // the next load goes into the buffer of the PREVIOUS chunk: its store must have been read
      tma_store_wait_read<1>();
    }
    __syncwarp();
    rs.issue(lane, quad);
  }
  if (p.stats_out != nullptr && row0 + lane < p.M)
    reinterpret_cast<float2*>(p.stats_out)[int64_t(nb * 2 + half) * p.M + row0 + lane] = make_float2(st1, st2);
}

// Register -> global epilogue (one row per thread).  Kept for the patch-embedding row map.
template <int BN>
__device__ __forceinline__ void epilogue_tile_direct(const GemmParams& p, uint32_t t_row, int mb,
                                                     int nb, int quad, int half, int lane,
                                                     uint32_t tempty_addr) {
  const int m = mb * BM + quad * 32 + lane;
  const bool row_ok = m < p.M;
  int64_t out_row = m, res_row = m;
  if (p.patch_rows > 0) {
    out_row = int64_t(m) + m / p.patch_rows + 1;
    res_row = 1 + m % p.patch_rows;
  }
  constexpr int NCH = BN / 32;
#pragma unroll 1
  for (int c = half; c < NCH; c += 2) {
    uint32_t v[32];
    tmem_ld_32x32(t_row + c * 32, v);
    tmem_ld_wait();
    if (c + 2 >= NCH) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_addr);
    }
    const int n0 = nb * BN + c * 32;
    if (!row_ok || n0 >= p.N) continue;
    float r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = __uint_as_float(v[j]);
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (n0 + j < p.N) {
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
        if (n0 + j < p.N) {
          const float4 s = __ldg(reinterpret_cast<const float4*>(p.scale + n0 + j));
          r[j] *= s.x; r[j + 1] *= s.y; r[j + 2] *= s.z; r[j + 3] *= s.w;
        }
      }
    }
    if (p.residual) {
      const float* rp = p.residual + res_row * p.ldr + n0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (n0 + j < p.N) {
          const float4 q = *reinterpret_cast<const float4*>(rp + j);
          r[j] += q.x; r[j + 1] += q.y; r[j + 2] += q.z; r[j + 3] += q.w;
        }
      }
    }
    if (p.out_f32) {
      float* o = reinterpret_cast<float*>(p.out) + out_row * p.ldo + n0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (n0 + j < p.N)
          *reinterpret_cast<float4*>(o + j) = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
      }
    } else {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + out_row * p.ldo + n0;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        if (n0 + j < p.N) {
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

// TRANS: the launch may carry transposed operand views (a_trans / w_trans); false compiles the
// K-major-only kernel whose issue loops have nothing to decide at run time (5 % faster on the
// inference GEMMs than testing the flags there)
template <int BN, bool RES, bool TRANS>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
            const __grid_constant__ CUtensorMap tm_a2, const __grid_constant__ CUtensorMap tm_w2,
            const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_res,
            const GemmParams p) {
  using L = SmemLayout<BN, RES>;
  constexpr int kStages = L::kStages;
  constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : 2 * BN;  // two accumulator stages

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* epi_base = smem + kStages * L::kStage;  // 1 KB aligned (stage sizes are multiples of 1 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_base + L::kEpiBytes);
  uint64_t* full = bars;                        // [kStages]
  uint64_t* empty = bars + kStages;             // [kStages]
  uint64_t* tfull = bars + 2 * kStages;         // [2]
  uint64_t* tempty = bars + 2 * kStages + 2;    // [2]
  uint64_t* res_bars = bars + 2 * kStages + 4;  // [kEpiWarps][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bars + 2 * kEpiWarps);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_a);
    prefetch_tmap(&tm_w);
    if (p.K2blocks) {
      prefetch_tmap(&tm_a2);
      prefetch_tmap(&tm_w2);
    }
    if (!p.direct) prefetch_tmap(&tm_out);
    if (RES) prefetch_tmap(&tm_res);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], kEpiWarps);
    }
    for (int s = 0; s < 2 * kEpiWarps; ++s) mbar_init(&res_bars[s], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_batch = p.tiles_m * p.tiles_n;
  const int num_tiles = tiles_per_batch * p.batch;
  const int kblocks = p.K1blocks + p.K2blocks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int bz = tile / tiles_per_batch, tl = tile % tiles_per_batch;
        const int mb = tl / p.tiles_n, nb = tl % p.tiles_n;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = stage_base + s * L::kStage;
          uint8_t* sb = sa + L::kStageA;
          mbar_expect_tx(&full[s], L::kStage);
          if (kb < p.K1blocks) {
            const int bi = bz % p.batch_inner, bo = bz / p.batch_inner;
            if (TRANS && p.a_trans) {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i)
                tma_load_4d(sa + i * kMnBlockBytes, &tm_a, &full[s], mb * BM + i * 64, kb * BK, bi, bo);
            } else {
              tma_load_4d(sa, &tm_a, &full[s], kb * BK, mb * BM, bi, bo);
            }
            if (TRANS && p.w_trans) {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i)
                tma_load_4d(sb + i * kMnBlockBytes, &tm_w, &full[s], nb * BN + i * 64, kb * BK, bi, bo);
            } else {
              tma_load_4d(sb, &tm_w, &full[s], kb * BK, nb * BN, bi, bo);
            }
          } else {
            tma_load_2d(sa, &tm_a2, &full[s], (kb - p.K1blocks) * BK, mb * BM);
            tma_load_2d(sb, &tm_w2, &full[s], (kb - p.K1blocks) * BK, nb * BN);
          }
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp converged, tcgen05 instructions elected) ==========
    {
      const bool at = TRANS && p.a_trans != 0, wt = TRANS && p.w_trans != 0;
      const uint32_t idesc = make_idesc_bf16(BM, BN, at, wt);
      const uint32_t a_lbo = at ? kMnBlockBytes : 16, b_lbo = wt ? kMnBlockBytes : 16;
      // descriptor start-address step (in 16-byte units) of one 16-deep K slice
      const uint64_t a_step = at ? 16 * 128 / 16 : 2, b_step = wt ? 16 * 128 / 16 : 2;
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
          const uint64_t da = make_sdesc_sw128(sa, a_lbo, 1024);
          const uint64_t db = make_sdesc_sw128(sb, b_lbo, 1024);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // K-major: advance 16 bf16 = 32 B inside the 128-B swizzle row: +2 in (addr >> 4)
              umma_ss(d_tmem, da + a_step * k, db + b_step * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty[s]);  // frees the smem stage when these MMAs retire
            if (kb == kblocks - 1) umma_commit(&tfull[acc]);
          }
          __syncwarp();
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int quad = warp & 3;         // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;  // which of the quadrant's two warps
    const int ew = warp - 2;
    uint8_t* out_buf = epi_base + ew * L::kEpiBufs * kChunkBytes;
    uint8_t* res_buf = out_buf + 2 * kChunkBytes;
    uint64_t* res_full = res_bars + 2 * ew;
    uint32_t out_cnt = 0, res_issue = 0, res_wait = 0;
    int it = 0;
    float scale_next = 0.0f;  // row_scale of this thread's row in the NEXT tile (batch == 1 with row_scale)
    if (p.row_scale && int(blockIdx.x) < num_tiles)
      scale_next = load_row_scale(p, (int(blockIdx.x) / p.tiles_n) * BM + quad * 32 + lane);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int bz = tile / tiles_per_batch, tl = tile % tiles_per_batch;
      const int mb = tl / p.tiles_n, nb = tl % p.tiles_n;
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      const float ln_rstd = scale_next;
      if (p.row_scale && tile + int(gridDim.x) < num_tiles)
        scale_next = load_row_scale(p, ((tile + int(gridDim.x)) / p.tiles_n) * BM + quad * 32 + lane);
      const uint32_t t_row = tmem_base + (uint32_t(quad * 32) << 16) + acc * BN;
      if (p.direct) {
        mbar_wait(&tfull[acc], acc_ph);
        tc_fence_after();
        epilogue_tile_direct<BN>(p, t_row, mb, nb, quad, half, lane, smem_u32(&tempty[acc]));
      } else if (p.out_f32) {
        epilogue_tile_tma<BN, RES, true>(p, &tm_out, &tm_res, t_row, bz, mb, nb, quad, half, lane, out_buf,
                                         res_buf, res_full, out_cnt, res_issue, res_wait,
                                         smem_u32(&tempty[acc]), &tfull[acc], acc_ph, ln_rstd);
      } else {
        epilogue_tile_tma<BN, false, false>(p, &tm_out, &tm_res, t_row, bz, mb, nb, quad, half, lane,
                                            out_buf, res_buf, res_full, out_cnt, res_issue, res_wait,
                                            smem_u32(&tempty[acc]), &tfull[acc], acc_ph, ln_rstd);
      }
    }
    if (lane == 0) tma_store_wait<0>();  // all bulk stores of this warp are complete
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}


// ---------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256
// tile.  Each CTA holds its own 128 rows of A and HALF of the W tile (128 of the 256 N rows), so
// the shared-memory traffic per SM drops from (A 4 KB + B 8 KB) to (A 4 KB + B 4 KB) per 128-cycle
// MMA and the TMA fill from 48 KB to 32 KB per k-block: at 128 x 256 per single CTA the two together
// exceed the 128 B/clk shared-memory port (profiles/r01_summary.md: tensor pipe stuck at 64-70 %).
//   * both CTAs' producers load with the cta_group::2 TMA form that credits the LEADER's full barrier;
//     the leader arms it with the bytes of both CTAs and issues tcgen05.mma.cta_group::2 (M = 256);
//   * tcgen05.commit multicasts to the empty / accumulator-full barriers of both CTAs;
//   * each CTA's 8 epilogue warps drain their own 128 TMEM lanes exactly like the 1-CTA kernel and
//     arrive on the leader's accumulator-empty barrier (remote mbarrier arrive for the peer).
template <bool RES>
struct SmemLayout2 {
  static constexpr int kStageA = BM * BK * 2;   // this CTA's 128 rows of A
  static constexpr int kStageB = 128 * BK * 2;  // this CTA's half of the 256 W rows
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kStages = (RES || kWideStores) ? 5 : 6;
  static constexpr int kEpiBufs = RES ? kResRing + 1 : (kWideStores ? 4 : 2);  // per warp (2 KB units): residual ring used in place
                                                           // + two half-size bf16 buffers, or two 4 KB
                                                           // buffers of 128-byte rows (wide stores)
  static constexpr int kResBars = RES ? kResRing : 2;   // per warp
  static constexpr int kEpiBytes = kEpiWarps * kEpiBufs * kChunkBytes;
  static constexpr int kBarBytes = 512;
  static_assert((2 * kStages + 4 + kEpiWarps * kResBars) * 8 + 4 <= kBarBytes, "barrier area");
  static constexpr int kTotal = kStages * kStage + kEpiBytes + kBarBytes + 1024;
  static_assert(kTotal <= 232448, "shared memory budget exceeded");
};

template <bool RES, bool TRANS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
             const __grid_constant__ CUtensorMap tm_a2, const __grid_constant__ CUtensorMap tm_w2,
             const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_res,
             const __grid_constant__ CUtensorMap tm_out2, const GemmParams p) {
  using L = SmemLayout2<RES>;
  constexpr int BN = 256;
  constexpr int kStages = L::kStages;
  constexpr uint32_t kTmemCols = 512;  // two accumulator stages of 256 columns (per CTA: 128 lanes)

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* epi_base = smem + kStages * L::kStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_base + L::kEpiBytes);
  uint64_t* full = bars;                        // [kStages]  used in the leader CTA only
  uint64_t* empty = bars + kStages;             // [kStages]  one per CTA (multicast commit)
  uint64_t* tfull = bars + 2 * kStages;         // [2]        one per CTA (multicast commit)
  uint64_t* tempty = bars + 2 * kStages + 2;    // [2]        leader only, 2 x kEpiWarps arrivals
  uint64_t* res_bars = bars + 2 * kStages + 4;  // [kEpiWarps][kResBars]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bars + L::kResBars * kEpiWarps);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_a);
    prefetch_tmap(&tm_w);
    if (p.K2blocks) {
      prefetch_tmap(&tm_a2);
      prefetch_tmap(&tm_w2);
    }
    prefetch_tmap(&tm_out);
    if (RES) prefetch_tmap(&tm_res);
    if (RES && p.out2 != nullptr) prefetch_tmap(&tm_out2);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 2 * kEpiWarps);
    }
    for (int s = 0; s < L::kResBars * kEpiWarps; ++s) mbar_init(&res_bars[s], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm<kTmemCols>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();  // barrier inits + TMEM allocation of both CTAs visible cluster-wide
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_batch = p.tiles_m * p.tiles_n;  // tiles_m counts 256-row pair tiles here
  const int num_tiles = tiles_per_batch * p.batch;
  const int kblocks = p.K1blocks + p.K2blocks;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int bz = tile / tiles_per_batch, tl = tile % tiles_per_batch;
        const int mb = tl / p.tiles_n, nb = tl % p.tiles_n;
        const int row_a = mb * 256 + int(rank) * 128;
        const int row_w = nb * 256 + int(rank) * 128;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = stage_base + s * L::kStage;
          uint8_t* sb = sa + L::kStageA;
          const uint32_t full_leader = map_to_cta(&full[s], 0);
          if (rank == 0) mbar_expect_tx(&full[s], 2 * L::kStage);  // bytes of both CTAs
          if (kb < p.K1blocks) {
            const int bi = bz % p.batch_inner, bo = bz / p.batch_inner;
            const int wbi = p.w_shared ? 0 : bi, wbo = p.w_shared ? 0 : bo;
            if (TRANS && p.a_trans) {
              tma_load_4d_2sm(sa, &tm_a, full_leader, row_a, kb * BK, bi, bo);
              tma_load_4d_2sm(sa + kMnBlockBytes, &tm_a, full_leader, row_a + 64, kb * BK, bi, bo);
            } else {
              tma_load_4d_2sm(sa, &tm_a, full_leader, kb * BK, row_a, bi, bo);
            }
            if (TRANS && p.w_trans) {
              tma_load_4d_2sm(sb, &tm_w, full_leader, row_w, kb * BK, wbi, wbo);
              tma_load_4d_2sm(sb + kMnBlockBytes, &tm_w, full_leader, row_w + 64, kb * BK, wbi, wbo);
            } else {
              tma_load_4d_2sm(sb, &tm_w, full_leader, kb * BK, row_w, wbi, wbo);
            }
          } else {
            tma_load_3d_2sm(sa, &tm_a2, full_leader, (kb - p.K1blocks) * BK, row_a, 0);
            tma_load_3d_2sm(sb, &tm_w2, full_leader, (kb - p.K1blocks) * BK, row_w, 0);
          }
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only; whole warp converged, tcgen05 elected) ======
    if (rank == 0) {
      const bool at = TRANS && p.a_trans != 0, wt = TRANS && p.w_trans != 0;
      const uint32_t idesc = make_idesc_bf16(256, BN, at, wt);
      const uint32_t a_lbo = at ? kMnBlockBytes : 16, b_lbo = wt ? kMnBlockBytes : 16;
      const uint64_t a_step = at ? 16 * 128 / 16 : 2, b_step = wt ? 16 * 128 / 16 : 2;
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
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
          const uint64_t da = make_sdesc_sw128(sa, a_lbo, 1024);
          const uint64_t db = make_sdesc_sw128(sb, b_lbo, 1024);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_ss_2sm(d_tmem, da + a_step * k, db + b_step * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2sm(&empty[s], 3);  // frees the stage in BOTH CTAs
            if (kb == kblocks - 1) umma_commit_2sm(&tfull[acc], 3);
          }
          __syncwarp();
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9 of both CTAs) =====================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int ew = warp - 2;
    uint8_t* out_buf = epi_base + ew * L::kEpiBufs * kChunkBytes;
    uint64_t* res_full = res_bars + L::kResBars * ew;
    uint32_t out_cnt = 0, res_issue = 0, res_wait = 0;
    const uint32_t tempty_leader[2] = {map_to_cta(&tempty[0], 0), map_to_cta(&tempty[1], 0)};
    ResStream rs;
    if constexpr (RES) {
      // launch_bn sends residual problems here only with fp32 output and batch == 1
      rs.tm = &tm_res;
      rs.buf = out_buf;
      rs.buf16 = out_buf + kResRing * kChunkBytes;
      rs.bars = res_full;
      rs.tile = pair;
      rs.c = half;
      rs.num_tiles = num_tiles;
      rs.num_pairs = num_pairs;
      rs.tiles_n = p.tiles_n;
      rs.tiles_per_batch = tiles_per_batch;
      rs.rank = int(rank);
      rs.half = half;
      rs.issued = 0;
      rs.locate(quad);
#pragma unroll
      for (int i = 0; i < kResRing - 1; ++i) rs.issue(lane, quad);
    }
    int it = 0;
    float scale_next = 0.0f;  // row_scale of this thread's row in the NEXT tile (batch == 1 with row_scale)
    if (!RES && p.row_scale && pair < num_tiles)
      scale_next = load_row_scale(p, ((pair / p.tiles_n) * 2 + int(rank)) * BM + quad * 32 + lane);
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      const int bz = tile / tiles_per_batch, tl = tile % tiles_per_batch;
      const int mb = (tl / p.tiles_n) * 2 + int(rank);  // 128-row block of THIS CTA
      const int nb = tl % p.tiles_n;
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      const float ln_rstd = scale_next;
      if (!RES && p.row_scale && tile + num_pairs < num_tiles)
        scale_next = load_row_scale(p, (((tile + num_pairs) / p.tiles_n) * 2 + int(rank)) * BM + quad * 32 + lane);
      const uint32_t t_row = tmem_base + (uint32_t(quad * 32) << 16) + acc * BN;
      if constexpr (RES) {
        mbar_wait(&tfull[acc], acc_ph);
        tc_fence_after();
        epilogue_tile_res_ring(p, &tm_out, &tm_out2, t_row, bz, mb, nb, quad, half, lane, rs, res_wait,
                               tempty_leader[acc]);
      } else if (p.out_f32) {
        epilogue_tile_tma<BN, false, true, kWideStores>(p, &tm_out, &tm_res, t_row, bz, mb, nb, quad, half, lane, out_buf,
                                                 out_buf, res_full, out_cnt, res_issue, res_wait,
                                                 tempty_leader[acc], &tfull[acc], acc_ph, ln_rstd);
      } else {
        epilogue_tile_tma<BN, false, false, kWideStores>(p, &tm_out, &tm_res, t_row, bz, mb, nb, quad, half, lane,
                                                  out_buf, out_buf, res_full, out_cnt, res_issue, res_wait,
                                                  tempty_leader[acc], &tfull[acc], acc_ph, ln_rstd);
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  cluster_sync_all();  // nobody exits (or frees TMEM) while the peer may still touch its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<kTmemCols>(tmem_base);
  }
}

void fill_ln_params(GemmParams& p, const dod_gemm_args& a) {
  p.out2 = reinterpret_cast<__nv_bfloat16*>(a.out_bf16);
  p.ldo2 = a.ldo_bf16;
  p.stats_out = a.row_stats_out;
  p.row_scale = a.row_scale;
}

template <bool RES, bool TRANS>
int launch2(const dod_gemm_args& a, cudaStream_t stream) {
  using L = SmemLayout2<RES>;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    DOD_CUDA_OK(cudaFuncSetAttribute(gemm2_kernel<RES, TRANS>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
  }
  CUtensorMap tm_a, tm_w, tm_a2, tm_w2, tm_out, tm_res;
  const BatchDims bd(a);
  if (int rc = a.a_trans ? make_tmap_4d(&tm_a, a.a, 2, bd.outer, bd.inner, a.k, a.m, bd.os_a, bd.is_a, a.lda, BK, 64, 128)
                         : make_tmap_4d(&tm_a, a.a, 2, bd.outer, bd.inner, a.m, a.k, bd.os_a, bd.is_a, a.lda, 128, BK, 128))
    return rc;
  // batch_stride_w == 0 on a batched problem: one W (and one residual) for every batch entry
  const bool w_shared = a.batch > 1 && a.batch_stride_w == 0;
  const uint64_t w_outer = w_shared ? 1 : bd.outer, w_inner = w_shared ? 1 : bd.inner;
  const uint64_t w_os = w_shared ? uint64_t(a.n) * a.ldw : bd.os_w, w_is = w_shared ? uint64_t(a.n) * a.ldw : bd.is_w;
  if (int rc = a.w_trans ? make_tmap_4d(&tm_w, a.w, 2, w_outer, w_inner, a.k, a.n, w_os, w_is, a.ldw, BK, 64, 128)
                         : make_tmap_4d(&tm_w, a.w, 2, w_outer, w_inner, a.n, a.k, w_os, w_is, a.ldw, 128, BK, 128))
    return rc;
  if (a.a2) {
    if (int rc = make_tmap_3d(&tm_a2, a.a2, 2, 1, a.m, a.k2, uint64_t(a.m) * a.lda2, a.lda2, 128, BK, 128)) return rc;
    if (int rc = make_tmap_3d(&tm_w2, a.w2, 2, 1, a.n, a.k2, uint64_t(a.n) * a.ldw2, a.ldw2, 128, BK, 128)) return rc;
  } else {
    tm_a2 = tm_a;
    tm_w2 = tm_w;
  }
  const bool out_f32 = a.out_dtype == DOD_F32;
  const int64_t n_out = a.act == DOD_ACT_SWIGLU ? a.n / 2 : a.n;
  // residual epilogue: 64-byte rows (in-place ring); otherwise 128-byte rows (wide stores)
  if (int rc = (RES || !kWideStores)
                   ? make_tmap_4d(&tm_out, a.out, out_f32 ? 4 : 2, bd.outer, bd.inner, a.m, n_out, bd.os_o, bd.is_o, a.ldo,
                                  32, out_f32 ? 16 : 32, 64)
                   : make_tmap_4d(&tm_out, a.out, out_f32 ? 4 : 2, bd.outer, bd.inner, a.m, n_out, bd.os_o, bd.is_o,
                                  a.ldo, 32, out_f32 ? 32 : 64, 128))
    return rc;
  if (RES) {
    if (int rc = make_tmap_2d(&tm_res, a.residual, 4, a.m, a.n, a.ldr, 32, 16, 64)) return rc;
  } else {
    tm_res = tm_a;
  }
  CUtensorMap tm_out2 = tm_a;
  if (RES && a.out_bf16)
    if (int rc = make_tmap_2d(&tm_out2, a.out_bf16, 2, a.m, a.n, a.ldo_bf16, 32, 16, 32)) return rc;
  GemmParams p;
  p.M = int(a.m);
  p.N = int(a.n);
  p.K1blocks = int((a.k + BK - 1) / BK);
  p.K2blocks = a.a2 ? int((a.k2 + BK - 1) / BK) : 0;
  p.tiles_m = int((a.m + 255) / 256);
  p.tiles_n = int((a.n + 255) / 256);
  p.batch = int(bd.outer * bd.inner);
  p.batch_inner = int(bd.inner);
  p.bias = a.bias;
  p.scale = a.scale;
  p.residual = reinterpret_cast<const float*>(a.residual);
  p.ldr = a.ldr;
  p.out = a.out;
  p.ldo = a.ldo;
  p.act = a.act;
  p.out_f32 = out_f32;
  p.patch_rows = 0;
  p.direct = 0;
  p.a_trans = a.a_trans;
  p.w_trans = a.w_trans;
  fill_ln_params(p, a);
  p.w_shared = w_shared ? 1 : 0;
  const int tiles = p.tiles_m * p.tiles_n * p.batch;
  int max_pairs = num_sms() / 2;
  if (const char* e = getenv("DOD_GEMM_MAX_PAIRS")) {  // developer knob: how the mainloop scales with fewer SMs
    const int v = atoi(e);
    if (v > 0 && v < max_pairs) max_pairs = v;
  }
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  gemm2_kernel<RES, TRANS><<<2 * pairs, kThreads, L::kTotal, stream>>>(tm_a, tm_w, tm_a2, tm_w2, tm_out, tm_res,
                                                                       tm_out2, p);
  return check_cuda(cudaGetLastError(), "gemm2_kernel launch");
}

template <int BN, bool RES, bool TRANS>
int launch(const dod_gemm_args& a, cudaStream_t stream, bool direct) {
  using L = SmemLayout<BN, RES>;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    DOD_CUDA_OK(cudaFuncSetAttribute(gemm_kernel<BN, RES, TRANS>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
  }
  CUtensorMap tm_a, tm_w, tm_a2, tm_w2, tm_out, tm_res;
  const BatchDims bd(a);
  if (int rc = a.a_trans ? make_tmap_4d(&tm_a, a.a, 2, bd.outer, bd.inner, a.k, a.m, bd.os_a, bd.is_a, a.lda, BK, 64, 128)
                         : make_tmap_4d(&tm_a, a.a, 2, bd.outer, bd.inner, a.m, a.k, bd.os_a, bd.is_a, a.lda, BM, BK, 128))
    return rc;
  if (int rc = a.w_trans ? make_tmap_4d(&tm_w, a.w, 2, bd.outer, bd.inner, a.k, a.n, bd.os_w, bd.is_w, a.ldw, BK, 64, 128)
                         : make_tmap_4d(&tm_w, a.w, 2, bd.outer, bd.inner, a.n, a.k, bd.os_w, bd.is_w, a.ldw, BN, BK, 128))
    return rc;
  if (a.a2) {
    if (int rc = make_tmap_2d(&tm_a2, a.a2, 2, a.m, a.k2, a.lda2, BM, BK)) return rc;
    if (int rc = make_tmap_2d(&tm_w2, a.w2, 2, a.n, a.k2, a.ldw2, BN, BK)) return rc;
  } else {
    tm_a2 = tm_a;
    tm_w2 = tm_w;
  }
  const bool out_f32 = a.out_dtype == DOD_F32;
  const int64_t n_out = a.act == DOD_ACT_SWIGLU ? a.n / 2 : a.n;
  if (!direct) {
    if (int rc = make_tmap_4d(&tm_out, a.out, out_f32 ? 4 : 2, bd.outer, bd.inner, a.m, n_out, bd.os_o,
                              bd.is_o, a.ldo, 32, out_f32 ? 16 : 32, 64))
      return rc;
  } else {
    tm_out = tm_a;
  }
  if (RES && !direct) {
    if (int rc = make_tmap_2d(&tm_res, a.residual, 4, a.m, a.n, a.ldr, 32, 16, 64)) return rc;
  } else {
    tm_res = tm_a;
  }
  GemmParams p;
  p.M = int(a.m);
  p.N = int(a.n);
  p.K1blocks = int((a.k + BK - 1) / BK);
  p.K2blocks = a.a2 ? int((a.k2 + BK - 1) / BK) : 0;
  p.tiles_m = int((a.m + BM - 1) / BM);
  p.tiles_n = int((a.n + BN - 1) / BN);
  p.batch = int(bd.outer * bd.inner);
  p.batch_inner = int(bd.inner);
  p.bias = a.bias;
  p.scale = a.scale;
  p.residual = reinterpret_cast<const float*>(a.residual);
  p.ldr = a.ldr;
  p.out = a.out;
  p.ldo = a.ldo;
  p.act = a.act;
  p.out_f32 = out_f32;
  p.patch_rows = a.patch_rows;
  p.direct = direct;
  p.a_trans = a.a_trans;
  p.w_trans = a.w_trans;
  fill_ln_params(p, a);
  p.w_shared = 0;
  const int tiles = p.tiles_m * p.tiles_n * p.batch;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_kernel<BN, RES, TRANS><<<grid, kThreads, L::kTotal, stream>>>(tm_a, tm_w, tm_a2, tm_w2, tm_out, tm_res, p);
  return check_cuda(cudaGetLastError(), "gemm_kernel launch");
}

// DOD_GEMM_2CTA=0 disables the CTA-pair kernel (A/B measurements)
bool use_pair_kernel() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DOD_GEMM_2CTA");
    v = e ? atoi(e) : 1;
  }
  return v != 0;
}

template <int BN>
int launch_bn(const dod_gemm_args& a, cudaStream_t stream) {
  // TMA epilogue needs: no row remap, residual only together with fp32 output
  const bool direct = a.patch_rows > 0 || (a.residual && a.out_dtype != DOD_F32);
  const bool trans = a.a_trans || a.w_trans;
  // the folded-LayerNorm producer (out_bf16 / row_stats_out) exists only in the pair kernel's residual epilogue:
  // it takes any M (rows past M are zero-filled by the TMA loads and clipped by the stores), so that whether a
  // LayerNorm is folded never depends on how many token rows a batch -- or a data-parallel shard of it -- has
  if (BN == 256 && !direct && (a.m >= 512 || a.out_bf16) && a.n >= 256 && use_pair_kernel()) {
    if (trans) return a.residual ? launch2<true, true>(a, stream) : launch2<false, true>(a, stream);
    return a.residual ? launch2<true, false>(a, stream) : launch2<false, false>(a, stream);
  }
  if (trans) {
    if (a.residual && !direct) return launch<BN, true, true>(a, stream, false);
    return launch<BN, false, true>(a, stream, direct);
  }
  if (a.residual && !direct) return launch<BN, true, false>(a, stream, false);
  return launch<BN, false, false>(a, stream, direct);
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
  DOD_REQUIRE(a->lda % 8 == 0 && a->ldw % 8 == 0 && a->lda >= (a->a_trans ? a->m : a->k) &&
                  a->ldw >= (a->w_trans ? a->n : a->k),
              "dod_gemm_bf16: lda/ldw must cover the stored row (k, or m / n for transposed views) and be "
              "multiples of 8 (16-byte TMA strides)");
  DOD_REQUIRE(!(a->a_trans || a->w_trans) || (!a->a2 && !a->w2),
              "dod_gemm_bf16: transposed operand views take no second K segment");
  DOD_REQUIRE((uintptr_t(a->a) & 15) == 0 && (uintptr_t(a->w) & 15) == 0 &&
                  (uintptr_t(a->out) & 15) == 0,
              "dod_gemm_bf16: a/w/out must be 16-byte aligned");
  // bias / scale are read as float4 and the direct epilogue stores whole vectors: those need n % 8 == 0.
  // The TMA-store epilogue clips at the tensor edge, so bias-free problems (attention scores) take any n.
  DOD_REQUIRE(a->n % 8 == 0 || (!a->bias && !a->scale && a->patch_rows == 0 &&
                                !(a->residual && a->out_dtype != DOD_F32) && a->act != DOD_ACT_SWIGLU),
              "dod_gemm_bf16: n must be a multiple of 8 when bias/scale/patch rows are used (got %lld)",
              (long long)a->n);
  if (a->a2 || a->w2) {
    DOD_REQUIRE(a->a2 && a->w2 && a->k2 > 0 && a->lda2 % 8 == 0 && a->ldw2 % 8 == 0 &&
                    a->lda2 >= a->k2 && a->ldw2 >= a->k2,
                "dod_gemm_bf16: bad second K segment");
    DOD_REQUIRE((uintptr_t(a->a2) & 15) == 0 && (uintptr_t(a->w2) & 15) == 0,
                "dod_gemm_bf16: a2/w2 must be 16-byte aligned");
  }
  DOD_REQUIRE(a->act >= DOD_ACT_NONE && a->act <= DOD_ACT_SWIGLU, "dod_gemm_bf16: bad act");
  if (a->out_bf16 || a->row_stats_out) {
    // folded LayerNorm, producer side: only the CTA-pair kernel's residual epilogue implements it
    DOD_REQUIRE(a->out_bf16 && a->row_stats_out, "dod_gemm_bf16: out_bf16 and row_stats_out come together");
    DOD_REQUIRE(a->residual && a->out_dtype == DOD_F32 && a->n >= 256 && a->n % 16 == 0 &&
                    a->patch_rows == 0 && a->batch <= 1 && a->batch_inner <= 1 && use_pair_kernel(),
                "dod_gemm_bf16: out_bf16 / row_stats_out need a residual GEMM with fp32 output, "
                "n >= 256, n %% 16 == 0");
    DOD_REQUIRE(a->ldo_bf16 >= a->n && a->ldo_bf16 % 8 == 0 && (uintptr_t(a->out_bf16) & 15) == 0 &&
                    (uintptr_t(a->row_stats_out) & 7) == 0,
                "dod_gemm_bf16: out_bf16 needs ldo_bf16 >= n, ldo_bf16 %% 8 == 0 and 16-byte alignment");
  }
  if (a->row_scale) {
    DOD_REQUIRE(a->bias, "dod_gemm_bf16: row_scale needs a bias (folded LayerNorm: b + W beta)");
    DOD_REQUIRE(a->patch_rows == 0 && !(a->residual && a->out_dtype != DOD_F32) && a->batch <= 1 &&
                    a->batch_inner <= 1 && a->n % 8 == 0 && (uintptr_t(a->row_scale) & 3) == 0,
                "dod_gemm_bf16: row_scale is not available with patch rows / bf16 residual output / batches");
  }
  if (a->batch_inner > 1) {
    DOD_REQUIRE(a->inner_stride_a % 8 == 0 && a->inner_stride_w % 8 == 0 &&
                    a->inner_stride_out % (a->out_dtype == DOD_F32 ? 4 : 8) == 0 && a->inner_stride_a > 0 &&
                    a->inner_stride_w > 0 && a->inner_stride_out > 0,
                "dod_gemm_bf16: inner batch strides must be positive multiples of 16 bytes");
    DOD_REQUIRE(!a->residual && !a->a2 && a->patch_rows == 0,
                "dod_gemm_bf16: batched problems take no residual / second K segment / patch rows");
  }
  if (a->batch > 1) {
    // batch_stride_w == 0: W -- and the fp32 residual, if any -- is shared by every batch entry (the patch
    // embedding: one GEMM per image so that its token rows form a TMA box).  CTA-pair kernel only.
    const bool shared = a->batch_stride_w == 0;
    if (shared)
      DOD_REQUIRE(a->m >= 512 && a->n >= 256 && a->batch_inner <= 1 && !a->a_trans && !a->w_trans &&
                      (!a->residual || a->out_dtype == DOD_F32) && use_pair_kernel(),
                  "dod_gemm_bf16: a shared W (batch_stride_w == 0) needs m >= 512, n >= 256, one batch level");
    DOD_REQUIRE((shared || !a->residual) && !a->a2 && a->patch_rows == 0,
                "dod_gemm_bf16: batched problems take no second K segment / patch rows, and a residual only "
                "together with a shared W");
    DOD_REQUIRE(a->batch_stride_a % 8 == 0 && a->batch_stride_w % 8 == 0 &&
                    a->batch_stride_out % (a->out_dtype == DOD_F32 ? 4 : 8) == 0 &&
                    a->batch_stride_a > 0 && a->batch_stride_w >= 0 && a->batch_stride_out > 0,
                "dod_gemm_bf16: batch strides must be positive multiples of 16 bytes");
    DOD_REQUIRE(a->batch * (a->batch_inner > 1 ? a->batch_inner : 1) * ((a->m + 127) / 128) *
                        ((a->n + 63) / 64) < (1ll << 31),
                "dod_gemm_bf16: too many tiles");
  }
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
    rc = launch_bn<256>(*a, stream);
  } else if (a->n > 128) {
    rc = launch_bn<256>(*a, stream);
  } else if (a->n > 64) {
    rc = launch_bn<128>(*a, stream);
  } else {
    rc = launch_bn<64>(*a, stream);
  }
  if (rc == 0) count_launch();
  return rc;
}
