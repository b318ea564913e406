// dod_fmha_fwd — fused multi-head self-attention, head dim 64, for sm_100a.
//
// One CTA per (batch, head, 128-query tile).  Flash-style single pass over the
// keys in tiles of 128:
//     S = Q.K^T           tcgen05.mma  (SS: Q, K from 128B-swizzled smem)   -> TMEM
//     P = exp2(S*c - m)   4 softmax warps, one thread per query row (all 128 key columns of the
//                         tile in registers: no row-maximum exchange), 168 registers per thread
//     O += P.V            tcgen05.mma  (TS: P from TMEM as bf16, V MN-major smem)
// O stays in TMEM for the whole pass; the running max is only advanced (and O
// rescaled through tcgen05.ld/st) when it grows by more than 2^8, so the common
// iteration touches O not at all.  K and V are double-buffered TMA rings fed by two
// single-thread issue warps (Q/K loads + Q.K^T, V loads + P.V); the S_{j+1} MMA is
// issued as soon as S_j has been pulled into registers so the tensor pipe runs under
// the softmax.  Two CTAs are resident per SM (87 KB smem, 256 TMEM columns, 192 threads each).
// Round 1 ran two threads per row (256 softmax threads at 96 registers, 605 us on the bench shape);
// thread-per-row removes the shared-memory exchange and halves the warps that pay the per-tile barrier
// round trips (warps issue in order): 567 us; second half of the row loaded under the first half's
// maximum, P-store completion under the next score load: 563 us (cuDNN SDPA: 491 us).
// The XU pipe (16 ex2/clk/SM) sets the floor: 1045 clk per 128x128 tile per SM measured for the bare
// instruction mix (tools/ubench/softmax_mix.cu); this kernel runs at ~1300 clk per tile plus the
// prologue / epilogue of each CTA.  profiles/r01_summary.md and r02_summary.md hold the measured phase
// timelines and the variants that were tried and dropped (polynomial exp2 offload, 3 CTAs per SM,
// two query tiles per CTA, a persistent software-pipelined single-stream kernel).
// Round 2, measured on top of it and dropped (profiles/r02_summary.md): a persistent form of this kernel (two
// CTAs per SM walking the (image, head, query tile) items: 588 us -- the co-resident CTA already fills the
// prologue / epilogue of its neighbour), skipping the warps whose rows lie beyond the sequence (563 us) and the
// exp2 of key chunks beyond it (617 us: the branch splits the MUFU block), reducing the row maximum on the side
// and applying it one tile late (644 / 685 us).  ncu: XU pipe 62 %, issue slots 41 %, stall reasons wait /
// long scoreboard -- the in-order chain of two independently phased softmax warps per scheduler is the bound.
// fmha64_kernel at the end of this file (64-key tiles, four CTAs per SM) is the opt-in form for short sequences.
// Optionally writes the per-row log-sum-exp consumed by dod_fmha_bwd.
//
// Replaces F.scaled_dot_product_attention behind HF Dinov2SelfAttention
// (transformers modeling_dinov2.py:215-229); scale 1/sqrt(64), non-causal,
// no mask, dropout 0.

#include <cstdlib>

#include "common.cuh"
#include "../../include/dod.h"

namespace dod {
void count_launch(int n = 1);
namespace {

constexpr int kD = 64;        // head dim
constexpr int kTile = 128;    // query rows per CTA == key rows per tile
constexpr int kTileBytes = kTile * kD * 2;  // 16 KB
constexpr int kKVStages = 2;
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kColS = 0, kColP = 128, kColO = 192;
constexpr int kSoftmaxThreads = 128;            // one thread per query row
constexpr int kThreads = kSoftmaxThreads + 64;  // + Q.K^T issue warp + P.V issue warp
constexpr int kChunks = kTile / 32;             // 32-column register chunks of a score row
constexpr int kSmemBytes = kTileBytes * (1 + 2 * kKVStages) + 256 + 1024;

// raw MUFU.EX2 (flush-to-zero): exp2f() wraps it in a denormal-range test + two multiplies per
// element, which doubled the issue slots of the softmax loop
__device__ __forceinline__ float ex2_approx(float x) {
#ifdef DOD_FMHA_NOEXP
  // developer experiment (tools/build_variant.sh noexp attention.cu -DDOD_FMHA_NOEXP): wrong results, shows the
  // kernel's time with the MUFU work removed
  return fmaf(x, 0.001f, 1.0f);
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}

#ifdef DOD_FMHA_TRACE
// developer instrumentation (tools/build_variant.sh trace attention.cu -DDOD_FMHA_TRACE): SM-clock
// stamps at the hand-off points of one CTA, read back with dod_debug_fmha_trace()
__device__ long long g_trace[3 * 16 * 8];
#define TRACE_ON (blockIdx.x == 3 && blockIdx.y == 5 && blockIdx.z == 20)
#define TRACE(slot, j, pt, dep)                                                        \
  do {                                                                                 \
    if (TRACE_ON && lane == 0) {                                                       \
      long long t_;                                                                    \
      asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) : "r"(uint32_t(dep)) : "memory"); \
      g_trace[((slot) * 16 + (j)) * 8 + (pt)] = t_;                                    \
    }                                                                                  \
  } while (0)
#else
#define TRACE(slot, j, pt, dep) do { } while (0)
#endif

struct FmhaParams {
  int seq, heads;
  int q_off, k_off, v_off;
  float scale_log2;  // scale * log2(e)
  __nv_bfloat16* ctx;
  int64_t ldo;
  float* lse;  // optional [B, heads, S]
};

// compiler fence for registers filled by an asynchronous tcgen05.ld whose wait is not adjacent to it: uses placed
// after this cannot be scheduled before the tcgen05.wait::ld that precedes it (volatile asms keep their order)
__device__ __forceinline__ void regs_ready32(uint32_t (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 32; i += 16)
    asm volatile("" : "+r"(v[i]), "+r"(v[i + 1]), "+r"(v[i + 2]), "+r"(v[i + 3]), "+r"(v[i + 4]), "+r"(v[i + 5]),
                      "+r"(v[i + 6]), "+r"(v[i + 7]), "+r"(v[i + 8]), "+r"(v[i + 9]), "+r"(v[i + 10]),
                      "+r"(v[i + 11]), "+r"(v[i + 12]), "+r"(v[i + 13]), "+r"(v[i + 14]), "+r"(v[i + 15]));
}


__global__ void __launch_bounds__(kThreads, 2)
fmha_kernel(const __grid_constant__ CUtensorMap tm_qkv, const FmhaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kTileBytes;
  uint8_t* sV = sK + kKVStages * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kKVStages * kTileBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;   // [2]
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* s_free = bars + 10;
  uint64_t* p_full = bars + 11;
  uint64_t* o_full = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int n_kv = (p.seq + kTile - 1) / kTile;
  constexpr int kMmaWarp = kSoftmaxThreads / 32;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_qkv);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKVStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    mbar_init(s_full, 1);
    // one elected lane per softmax warp arrives (after __syncwarp): 256 per-thread arrivals on one
    // shared-memory word serialise (~32 cycles per warp) and sat on the critical path of every tile
    mbar_init(s_free, kSoftmaxThreads / 32);
    mbar_init(p_full, kSoftmaxThreads / 32);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kMmaWarp) {
    if (lane == 0) {
      // ---------------- Q/K loads + Q.K^T issue (single thread) ----------------
      // Two issuing threads (this one and the P.V thread below) because one thread issuing all
      // 12 MMAs of an iteration between its barrier waits took ~1750 of the ~2500 clocks an
      // iteration lasts (tools/fmha_trace.py) and delayed S_{j+1} behind P.V_j.
      constexpr uint32_t idesc_s = make_idesc_bf16(kTile, kTile, false, false);  // Q.K^T
      const int qc = p.q_off + head * kD, kc = p.k_off + head * kD;
      mbar_expect_tx(q_full, kTileBytes);
      tma_load_3d(sQ, &tm_qkv, q_full, qc, q_tile * kTile, b);
      for (int j = 0; j < kKVStages && j < n_kv; ++j) {
        mbar_expect_tx(&k_full[j], kTileBytes);
        tma_load_3d(sK + j * kTileBytes, &tm_qkv, &k_full[j], kc, j * kTile, b);
      }
      const uint64_t dq = make_sdesc_sw128(smem_u32(sQ), 16, 1024);
      auto issue_s = [&](int j) {
        const int s = j % kKVStages;
        mbar_wait(&k_full[s], (j / kKVStages) & 1);
        tc_fence_after();
        const uint64_t dk = make_sdesc_sw128(smem_u32(sK + s * kTileBytes), 16, 1024);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          umma_ss(tmem + kColS, dq + uint64_t(2 * k), dk + uint64_t(2 * k), idesc_s, k != 0);
        umma_commit(s_full);
        umma_commit(&k_empty[s]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j + 1 < n_kv; ++j) {
        const int s = j % kKVStages;
        TRACE(2, j, 0, 0);
        mbar_wait(s_free, j & 1);  // S_j is in registers: S columns reusable
        TRACE(2, j, 1, 0);
        issue_s(j + 1);
        TRACE(2, j, 2, 0);
        // S_j retired before s_full(j) fired, so K stage s is free: refill with K_{j+2}
        if (j + kKVStages < n_kv) {
          mbar_wait(&k_empty[s], (j / kKVStages) & 1);
          mbar_expect_tx(&k_full[s], kTileBytes);
          tma_load_3d(sK + s * kTileBytes, &tm_qkv, &k_full[s], kc, (j + kKVStages) * kTile, b);
        }
      }
    }
  } else if (warp == kMmaWarp + 1) {
    if (lane == 0) {
      // ---------------- V loads + P.V issue (single thread) ----------------
      constexpr uint32_t idesc_o = make_idesc_bf16(kTile, kD, false, true);  // P.V (V MN-major)
      const int vc = p.v_off + head * kD;
      for (int j = 0; j < kKVStages && j < n_kv; ++j) {
        mbar_expect_tx(&v_full[j], kTileBytes);
        tma_load_3d(sV + j * kTileBytes, &tm_qkv, &v_full[j], vc, j * kTile, b);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % kKVStages;
        mbar_wait(&v_full[s], (j / kKVStages) & 1);
        mbar_wait(p_full, j & 1);  // P_j stored (and O rescaled if needed)
        tc_fence_after();
        TRACE(2, j, 3, 0);
        {
          // V tile [128 keys x 64 d], d contiguous: MN-major B operand.  One MMA
          // consumes 16 keys = 16 rows of 128 B = 2 swizzle atoms (SBO 1024).
          const uint32_t sv = smem_u32(sV + s * kTileBytes);
#pragma unroll
          for (int k = 0; k < kTile / 16; ++k) {
            const uint64_t dv = make_sdesc_sw128(sv + k * 16 * 128, 16, 1024);
            umma_ts(tmem + kColO, tmem + kColP + 8 * k, dv, idesc_o, (j | k) != 0);
          }
        }
        umma_commit(o_full);
        umma_commit(&v_empty[s]);
        TRACE(2, j, 4, 0);
        // this thread is idle until P_{j+1}: wait for P.V_j to retire and refill its V stage
        if (j + kKVStages < n_kv) {
          mbar_wait(&v_empty[s], (j / kKVStages) & 1);
          mbar_expect_tx(&v_full[s], kTileBytes);
          tma_load_3d(sV + s * kTileBytes, &tm_qkv, &v_full[s], vc, (j + kKVStages) * kTile, b);
        }
        TRACE(2, j, 5, 0);
      }
    }
  } else {
    // ---------------- softmax warps: one thread per query row ----------------
    const int quad = warp & 3;         // TMEM lane quadrant
    const int row = quad * 32 + lane;  // row within the tile == TMEM lane
    const uint32_t t_lane = tmem + (uint32_t(quad * 32) << 16);
    float m_used = -INFINITY;  // in log2 units (already scaled)
    float l = 0.0f;            // row sum

    // s_ready / o_ready: the barrier was already seen complete by an early non-blocking test issued
    // in the middle of the previous / this exp2 phase (the MUFU pipe is the limiter there, the ~100-cycle
    // barrier instruction hides under it)
    bool s_ready = false;
    for (int j = 0; j < n_kv; ++j) {
      TRACE(0, j, 0, 0);
      if (!s_ready) mbar_wait(s_full, j & 1);
      tc_fence_after();
      bool o_ready = j == 0;
      s_ready = false;
      TRACE(0, j, 1, 0);
      const int valid = p.seq - j * kTile;  // keys valid in this tile
      // S is read from TMEM once and held in 128 registers (two-pass and polynomial-exp2 variants were
      // measured slower: profiles/r01_summary.md, r02_summary.md), in two groups of two 32-column chunks
      uint32_t sraw[kChunks][32];
      constexpr int kFirst = kChunks / 2;
#pragma unroll
      for (int c = 0; c < kFirst; ++c) tmem_ld_32x32(t_lane + kColS + c * 32, sraw[c]);
      if (j > 0) {
        // P of the previous tile: its tcgen05.st was issued at the end of that iteration; the completion wait
        // and the p_full arrival sit HERE, under the latency of the score load just issued
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        mbar_arrive_elect_addr(smem_u32(p_full));
      }
      tmem_ld_wait();
      // the second half of the row is loaded while the maximum of the first half is reduced
#pragma unroll
      for (int c = kFirst; c < kChunks; ++c) tmem_ld_32x32(t_lane + kColS + c * 32, sraw[c]);
      float mx0 = -INFINITY, mx1 = -INFINITY;
      auto mask_max = [&](int c0, int c1) {
        if (valid < kTile) {
          // tail tile only (1370 = 10*128 + 90): keys beyond the sequence get -inf
#pragma unroll
          for (int c = c0; c < c1; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i >= valid) sraw[c][i] = 0xff800000u;
        }
#pragma unroll
        for (int c = c0; c < c1; ++c)
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            mx0 = fmaxf(mx0, __uint_as_float(sraw[c][i]));
            mx1 = fmaxf(mx1, __uint_as_float(sraw[c][i + 1]));
          }
      };
      mask_max(0, kFirst);
      tmem_ld_wait();
#pragma unroll
      for (int c = kFirst; c < kChunks; ++c) regs_ready32(sraw[c]);
      tc_fence_before();
      __syncwarp();
      mbar_arrive_elect_addr(smem_u32(s_free));  // S_j is in registers: the next Q.K^T may overwrite it
      mask_max(kFirst, kChunks);
      TRACE(0, j, 2, __float_as_uint(fmaxf(mx0, mx1)));
      const float m_new = fmaxf(m_used, fmaxf(mx0, mx1) * p.scale_log2);
      // lazy max: keep the stale max while it is within 2^8 of the true one
      const bool bump = (m_new - m_used) > 8.0f;
      float alpha = 1.0f;
      if (bump) {
        alpha = exp2f(m_used - m_new);  // 0 on the first tile (m_used = -inf)
        m_used = m_new;
      }
      float2 sum2 = make_float2(0.0f, 0.0f);
      const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
      const float2 nm2 = make_float2(-m_used, -m_used);
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(sraw[c][i]), __uint_as_float(sraw[c][i + 1])), sc2, nm2);
          const float2 e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
          sum2 = __fadd2_rn(sum2, e);
          // packed bf16 pairs are written over score registers already consumed (pair k of the row lands in
          // word k: chunk c fills the lower / upper half of chunk c / 2), so P is stored from aligned 32-register
          // blocks without a second array
          sraw[c >> 1][(c & 1) * 16 + (i >> 1)] = pack_bf16x2(e.x, e.y);
        }
        if (c == kChunks / 2 - 1) {
          if (j > 0) o_ready = mbar_test_wait(o_full, (j - 1) & 1);
          if (j + 1 < n_kv) s_ready = mbar_test_wait(s_full, (j + 1) & 1);
        }
      }
      l = l * alpha + (sum2.x + sum2.y);
      TRACE(0, j, 4, sraw[0][31] ^ sraw[0][15] ^ __float_as_uint(l));

      if (j > 0) {
        if (!o_ready) mbar_wait(o_full, (j - 1) & 1);  // P.V of the previous tile retired
        tc_fence_after();
        TRACE(0, j, 5, 0);
        if (__any_sync(0xffffffffu, bump)) {
          // rescale the running O row (rare after the first tiles)
#pragma unroll
          for (int oc = 0; oc < kD / 32; ++oc) {
            uint32_t o[32];
            tmem_ld_32x32(t_lane + kColO + oc * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32(t_lane + kColO + oc * 32, o);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < kChunks / 2; ++c) tmem_st_32x32(t_lane + kColP + c * 32, sraw[c]);
      TRACE(0, j, 6, 0);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    mbar_arrive_elect_addr(smem_u32(p_full));  // P of the last tile

    // ---- epilogue: O / l -> ctx ----
    const float inv_l = 1.0f / l;
    mbar_wait(o_full, (n_kv - 1) & 1);
    tc_fence_after();
    const int q_row = q_tile * kTile + row;
    // log-sum-exp of the scaled scores in the log2 domain (m_used is the possibly stale maximum the
    // probabilities were formed against, so m_used + log2(sum) is exact): saved for dod_fmha_bwd
    if (p.lse != nullptr && q_row < p.seq)
      p.lse[(int64_t(b) * p.heads + head) * p.seq + q_row] = m_used + log2f(l);
    __nv_bfloat16* dst = p.ctx + (int64_t(b) * p.seq + q_row) * p.ldo + head * kD;
#pragma unroll
    for (int oc = 0; oc < kD / 32; ++oc) {
      uint32_t o[32];
      tmem_ld_32x32(t_lane + kColO + oc * 32, o);
      tmem_ld_wait();
      if (q_row < p.seq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
          v.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
          v.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
          v.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + oc * 32 + i) = v;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem);
  }
}



// ---------------------------------------------------------------------------------------------------------
// fmha64_kernel: the same pass with 64-key tiles and FOUR CTAs per SM.
//
// The 128-key kernel above is bound by the in-order chain of its softmax warps, not by a pipe (ncu: XU 62 %,
// issue slots 41 %, two independently phased softmax warps per scheduler; removing 2-5 % of the exp2 work
// does not move it, adding 64 ALU instructions per tile costs 15 %).  Here a thread still owns a query row
// but only 64 score columns at a time (~90 registers), P is written over the consumed S columns in TMEM
// (128 TMEM columns per CTA: S / P 64 + O 64) and ONE thread issues loads and both MMAs, so four CTAs
// (4 x 160 threads, 4 x 48 KB smem, 4 x 128 TMEM columns) are resident and every scheduler has four
// independently phased softmax warps feeding the XU pipe.  S_{j+1} is issued right behind P.V_j by the same
// thread (tcgen05.mma executes in issue order, so it cannot overwrite P_j before P.V_j has read it); the
// bubble this leaves in one CTA's chain is covered by the other three.
// ---------------------------------------------------------------------------------------------------------
constexpr int kKV = 64;                       // keys per tile
constexpr int kKVBytes = kKV * kD * 2;        // 8 KB
constexpr int kThreads64 = kSoftmaxThreads + 32;
constexpr uint32_t kTmemCols64 = 128;
constexpr uint32_t kColS64 = 0, kColO64 = 64;  // P (bf16, 32 columns) is written over S
constexpr int kSmemBytes64 = kTileBytes + 2 * kKVStages * kKVBytes + 256 + 1024;

__global__ void __launch_bounds__(kThreads64, 4)
fmha64_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, const FmhaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kTileBytes;
  uint8_t* sV = sK + kKVStages * kKVBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kKVStages * kKVBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;   // [2]
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* p_full = bars + 10;
  uint64_t* o_full = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int n_kv = (p.seq + kKV - 1) / kKV;
  constexpr int kIssueWarp = kSoftmaxThreads / 32;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_q);
    prefetch_tmap(&tm_kv);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKVStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, kSoftmaxThreads / 32);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == kIssueWarp) tmem_alloc<kTmemCols64>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kIssueWarp) {
    if (lane == 0) {
      // ---------------- loads + both MMAs (single thread) ----------------
      constexpr uint32_t idesc_s = make_idesc_bf16(kTile, kKV, false, false);  // Q.K^T
      constexpr uint32_t idesc_o = make_idesc_bf16(kTile, kD, false, true);    // P.V (V MN-major)
      const int qc = p.q_off + head * kD, kc = p.k_off + head * kD, vc = p.v_off + head * kD;
      mbar_expect_tx(q_full, kTileBytes);
      tma_load_3d(sQ, &tm_q, q_full, qc, q_tile * kTile, b);
      for (int j = 0; j < kKVStages && j < n_kv; ++j) {
        mbar_expect_tx(&k_full[j], kKVBytes);
        tma_load_3d(sK + j * kKVBytes, &tm_kv, &k_full[j], kc, j * kKV, b);
      }
      for (int j = 0; j < kKVStages && j < n_kv; ++j) {
        mbar_expect_tx(&v_full[j], kKVBytes);
        tma_load_3d(sV + j * kKVBytes, &tm_kv, &v_full[j], vc, j * kKV, b);
      }
      const uint64_t dq = make_sdesc_sw128(smem_u32(sQ), 16, 1024);
      auto issue_s = [&](int j) {
        const int s = j % kKVStages;
        mbar_wait(&k_full[s], (j / kKVStages) & 1);
        tc_fence_after();
        const uint64_t dk = make_sdesc_sw128(smem_u32(sK + s * kKVBytes), 16, 1024);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          umma_ss(tmem + kColS64, dq + uint64_t(2 * k), dk + uint64_t(2 * k), idesc_s, k != 0);
        umma_commit(s_full);
        umma_commit(&k_empty[s]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % kKVStages;
        mbar_wait(&v_full[s], (j / kKVStages) & 1);
        mbar_wait(p_full, j & 1);  // P_j stored over S_j (and O rescaled if needed)
        tc_fence_after();
        {
          const uint32_t sv = smem_u32(sV + s * kKVBytes);
#pragma unroll
          for (int k = 0; k < kKV / 16; ++k) {
            const uint64_t dv = make_sdesc_sw128(sv + k * 16 * 128, 16, 1024);
            umma_ts(tmem + kColO64, tmem + kColS64 + 8 * k, dv, idesc_o, (j | k) != 0);
          }
        }
        umma_commit(o_full);
        umma_commit(&v_empty[s]);
        // the next scores go over P_j: issued behind P.V_j, executed behind it
        if (j + 1 < n_kv) issue_s(j + 1);
        if (j + kKVStages < n_kv) {
          // K_j retired before s_full(j) fired; V_j retires with P.V_j
          mbar_wait(&k_empty[s], (j / kKVStages) & 1);
          mbar_expect_tx(&k_full[s], kKVBytes);
          tma_load_3d(sK + s * kKVBytes, &tm_kv, &k_full[s], kc, (j + kKVStages) * kKV, b);
          mbar_wait(&v_empty[s], (j / kKVStages) & 1);
          mbar_expect_tx(&v_full[s], kKVBytes);
          tma_load_3d(sV + s * kKVBytes, &tm_kv, &v_full[s], vc, (j + kKVStages) * kKV, b);
        }
      }
    }
  } else {
    // ---------------- softmax warps: one thread per query row, 64 keys per tile ----------------
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem + (uint32_t(quad * 32) << 16);
    const uint32_t a_s_full = smem_u32(s_full), a_p_full = smem_u32(p_full), a_o_full = smem_u32(o_full);
    float m_used = -INFINITY;  // in log2 units (already scaled)
    float l = 0.0f;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait_addr(a_s_full, j & 1);
      tc_fence_after();
      const int valid = p.seq - j * kKV;
      uint32_t sraw[2][32];
      tmem_ld_32x32(t_lane + kColS64, sraw[0]);
      tmem_ld_32x32(t_lane + kColS64 + 32, sraw[1]);
      tmem_ld_wait();
      if (valid < kKV) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= valid) sraw[c][i] = 0xff800000u;
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          mx0 = fmaxf(mx0, __uint_as_float(sraw[c][i]));
          mx1 = fmaxf(mx1, __uint_as_float(sraw[c][i + 1]));
        }
      const float m_new = fmaxf(m_used, fmaxf(mx0, mx1) * p.scale_log2);
      const bool bump = (m_new - m_used) > 8.0f;  // lazy maximum, as in the 128-key kernel
      float alpha = 1.0f;
      if (bump) {
        alpha = exp2f(m_used - m_new);
        m_used = m_new;
      }
      float2 sum2 = make_float2(0.0f, 0.0f);
      const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
      const float2 nm2 = make_float2(-m_used, -m_used);
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(sraw[c][i]), __uint_as_float(sraw[c][i + 1])), sc2, nm2);
          const float2 e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
          sum2 = __fadd2_rn(sum2, e);
          sraw[0][c * 16 + (i >> 1)] = pack_bf16x2(e.x, e.y);  // pair k of the row -> word k (already consumed)
        }
      l = l * alpha + (sum2.x + sum2.y);
      if (j > 0 && __any_sync(0xffffffffu, bump)) {
        mbar_wait_addr(a_o_full, (j - 1) & 1);  // P.V_{j-1} retired (it precedes S_j in the pipe: no real wait)
        tc_fence_after();
#pragma unroll
        for (int oc = 0; oc < kD / 32; ++oc) {
          uint32_t o[32];
          tmem_ld_32x32(t_lane + kColO64 + oc * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32(t_lane + kColO64 + oc * 32, o);
        }
      }
      tmem_st_32x32(t_lane + kColS64, sraw[0]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      mbar_arrive_elect_addr(a_p_full);
    }
    // ---- epilogue: O / l -> ctx ----
    const float inv_l = 1.0f / l;
    mbar_wait_addr(a_o_full, (n_kv - 1) & 1);
    tc_fence_after();
    const int q_row = q_tile * kTile + row;
    if (p.lse != nullptr && q_row < p.seq)
      p.lse[(int64_t(b) * p.heads + head) * p.seq + q_row] = m_used + log2f(l);
    __nv_bfloat16* dst = p.ctx + (int64_t(b) * p.seq + q_row) * p.ldo + head * kD;
#pragma unroll
    for (int oc = 0; oc < kD / 32; ++oc) {
      uint32_t o[32];
      tmem_ld_32x32(t_lane + kColO64 + oc * 32, o);
      tmem_ld_wait();
      if (q_row < p.seq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
          v.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
          v.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
          v.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + oc * 32 + i) = v;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kIssueWarp) {
    tc_fence_after();
    tmem_dealloc<kTmemCols64>(tmem);
  }
}

}  // namespace
}  // namespace dod

#ifdef DOD_FMHA_TRACE
extern "C" DOD_API int32_t dod_debug_fmha_trace(long long* host) {
  return int32_t(cudaMemcpyFromSymbol(host, dod::g_trace, sizeof(long long) * 3 * 16 * 8));
}
#endif

extern "C" int32_t dod_fmha_fwd(const dod_fmha_args* a, dod_stream_t stream_) {
  using namespace dod;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->qkv && a->ctx, "dod_fmha_fwd: null pointer");
  DOD_REQUIRE(a->batch > 0 && a->seq > 0 && a->heads > 0, "dod_fmha_fwd: empty problem");
  DOD_REQUIRE(a->batch <= 65535 && a->heads <= 65535, "dod_fmha_fwd: batch/heads exceed grid limits");
  DOD_REQUIRE(a->ld % 8 == 0 && a->ldo % 8 == 0, "dod_fmha_fwd: ld/ldo must be multiples of 8");
  DOD_REQUIRE((uintptr_t(a->qkv) & 15) == 0 && (uintptr_t(a->ctx) & 15) == 0,
              "dod_fmha_fwd: qkv/ctx must be 16-byte aligned");
  DOD_REQUIRE(a->q_off % 8 == 0 && a->k_off % 8 == 0 && a->v_off % 8 == 0,
              "dod_fmha_fwd: q/k/v column offsets must be multiples of 8");
  DOD_REQUIRE(a->q_off + a->heads * kD <= a->ld && a->k_off + a->heads * kD <= a->ld &&
                  a->v_off + a->heads * kD <= a->ld && a->heads * kD <= a->ldo,
              "dod_fmha_fwd: head slices exceed the row");
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    DOD_CUDA_OK(cudaFuncSetAttribute(fmha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  }
  CUtensorMap tm;
  if (int rc = make_tmap_3d(&tm, a->qkv, 2, a->batch, a->seq, a->ld, a->seq * a->ld, a->ld, kTile, kD))
    return rc;
  FmhaParams p;
  p.seq = int(a->seq);
  p.heads = int(a->heads);
  p.q_off = int(a->q_off);
  p.k_off = int(a->k_off);
  p.v_off = int(a->v_off);
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.ctx = reinterpret_cast<__nv_bfloat16*>(a->ctx);
  p.ldo = a->ldo;
  p.lse = a->lse;
  dim3 grid((a->seq + kTile - 1) / kTile, a->heads, a->batch);
  // DOD_FMHA64=1 selects the 4-CTA/SM 64-key kernel: faster on short sequences (64 x 12 heads: 21.5 vs 28.1 us at
  // 100 tokens, 59 vs 69 us at 257), level from ~384 tokens, slower at 1370 (590 vs 559 us).  It is opt-in: the
  // recorded parity margins (profiles/r02_parity_margins.json) belong to the 128-key kernel, and the 3-block g/14
  // case with its degenerate (1, 257) sampling grid sits within 10 % of the 2e-2 bar with either kernel.
  const char* e64 = getenv("DOD_FMHA64");
  const bool use64 = e64 != nullptr && e64[0] == '1';
  if (use64) {
    static PerDeviceOnce attr64_once;
    if (attr64_once.first()) {
      DOD_CUDA_OK(cudaFuncSetAttribute(fmha64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes64));
    }
    CUtensorMap tm_kv;
    if (int rc = make_tmap_3d(&tm_kv, a->qkv, 2, a->batch, a->seq, a->ld, a->seq * a->ld, a->ld, kKV, kD))
      return rc;
    fmha64_kernel<<<grid, kThreads64, kSmemBytes64, stream>>>(tm, tm_kv, p);
    int rc = check_cuda(cudaGetLastError(), "fmha64_kernel launch");
    if (rc == 0) count_launch();
    return rc;
  }
  fmha_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tm, p);
  int rc = check_cuda(cudaGetLastError(), "fmha_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}
