// dod_fmha_fwd — fused multi-head self-attention, head dim 64, for sm_100a.
//
// One CTA per (batch, head, 128-query tile).  Flash-style single pass over the
// keys in tiles of 128:
//     S = Q.K^T           tcgen05.mma  (SS: Q, K from 128B-swizzled smem)   -> TMEM
//     P = exp2(S*c - m)   8 softmax warps, two threads per query row (64 key columns each,
//                         row maximum exchanged through shared memory), S held in registers
//     O += P.V            tcgen05.mma  (TS: P from TMEM as bf16, V MN-major smem)
// O stays in TMEM for the whole pass; the running max is only advanced (and O
// rescaled through tcgen05.ld/st) when it grows by more than 2^8, so the common
// iteration touches O not at all.  K and V are double-buffered TMA rings fed by two
// single-thread issue warps (Q/K loads + Q.K^T, V loads + P.V); the S_{j+1} MMA is
// issued as soon as S_j has been pulled into registers so the tensor pipe runs under
// the softmax.  Two CTAs are resident per SM (87 KB smem, 256 TMEM columns each).
// The kernel is bound by the XU pipe (16 ex2/clk/SM): see profiles/r01_summary.md for
// the measured phase timeline and the variants that were tried and dropped.
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
constexpr int kSoftmaxThreads = 256;  // two threads per query row (64 key columns each)
constexpr int kThreads = kSoftmaxThreads + 64;  // + Q.K^T issue warp + P.V issue warp
constexpr int kXchgBytes = 2 * 2 * kTile * 4 + 2 * kTile * 4;  // max exchange (double buffered) + sums
constexpr int kSmemBytes = kTileBytes * (1 + 2 * kKVStages) + 256 + kXchgBytes + 1024;

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

__device__ __forceinline__ void pair_sync(int quad) {
  // the two warps that share a TMEM lane quadrant (rows quad*32 .. +31)
  asm volatile("bar.sync %0, 64;" ::"r"(quad + 1) : "memory");
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
  float* xmax = reinterpret_cast<float*>(bars + 32);  // [2 (tile parity)][2 (half)][128 rows]
  float* xsum = xmax + 2 * 2 * kTile;                 // [2 (half)][128 rows]

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
    // ---------------- softmax warps: two threads per query row ----------------
    const int quad = warp & 3;   // TMEM lane quadrant
    const int half = warp >> 2;  // key columns [half*64, +64) of every tile; O columns [half*32, +32)
    const int row = quad * 32 + lane;  // row within the tile == TMEM lane
    const uint32_t t_lane = tmem + (uint32_t(quad * 32) << 16);
    float m_used = -INFINITY;  // in log2 units (already scaled)
    float l = 0.0f;            // partial row sum over this thread's columns
    const uint32_t xmax_addr = smem_u32(xmax);

    // s_ready / o_ready: the barrier was already seen complete by an early non-blocking test issued
    // in the middle of the previous / this exp2 phase (the MUFU pipe is the limiter there, the ~100-cycle
    // barrier instruction hides under it)
    bool s_ready = false;
    for (int j = 0; j < n_kv; ++j) {
      TRACE(half, j, 0, 0);
      if (!s_ready) mbar_wait(s_full, j & 1);
      tc_fence_after();
      bool o_ready = j == 0;
      s_ready = false;
      TRACE(half, j, 1, 0);
      const int valid = p.seq - j * kTile - half * 64;  // keys valid in this thread's 64 columns
      // S is read from TMEM once and held in 64 registers (two-pass and polynomial-exp2 variants were
      // measured slower: profiles/r01_summary.md)
      uint32_t sraw[2][32];
      tmem_ld_32x32(t_lane + kColS + half * 64, sraw[0]);
      tmem_ld_32x32(t_lane + kColS + half * 64 + 32, sraw[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      mbar_arrive_elect_addr(smem_u32(s_free));  // S_j is in registers: the next Q.K^T may overwrite it
      if (valid < 64) {
        // tail tile only (1370 = 10*128 + 90): keys beyond the sequence get -inf
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
      // row maximum over both halves: exchange through shared memory (double buffered by parity)
      const uint32_t xm = xmax_addr + uint32_t((j & 1) * 2 * kTile) * 4u;
      TRACE(half, j, 2, __float_as_uint(fmaxf(mx0, mx1)));
      sts_f32(xm + uint32_t(half * kTile + row) * 4u, fmaxf(mx0, mx1));
      pair_sync(quad);
      TRACE(half, j, 3, 0);
      const float mx = fmaxf(fmaxf(mx0, mx1), lds_f32(xm + uint32_t((half ^ 1) * kTile + row) * 4u));
      const float m_new = fmaxf(m_used, mx * p.scale_log2);
      // lazy max: keep the stale max while it is within 2^8 of the true one
      const bool bump = (m_new - m_used) > 8.0f;
      float alpha = 1.0f;
      if (bump) {
        alpha = exp2f(m_used - m_new);  // 0 on the first tile (m_used = -inf)
        m_used = m_new;
      }
      float2 sum2 = make_float2(0.0f, 0.0f);
      uint32_t pk[32];
      const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
      const float2 nm2 = make_float2(-m_used, -m_used);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(sraw[c][i]), __uint_as_float(sraw[c][i + 1])), sc2, nm2);
          const float2 e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
          sum2 = __fadd2_rn(sum2, e);
          pk[c * 16 + (i >> 1)] = pack_bf16x2(e.x, e.y);
        }
        if (c == 0) {
          if (j > 0) o_ready = mbar_test_wait(o_full, (j - 1) & 1);
          if (j + 1 < n_kv) s_ready = mbar_test_wait(s_full, (j + 1) & 1);
        }
      }
      l = l * alpha + (sum2.x + sum2.y);
      TRACE(half, j, 4, pk[31] ^ pk[15] ^ __float_as_uint(l));

      if (j > 0) {
        if (!o_ready) mbar_wait(o_full, (j - 1) & 1);  // P.V of the previous tile retired
        tc_fence_after();
        TRACE(half, j, 5, 0);
        if (__any_sync(0xffffffffu, bump)) {
          // rescale this thread's half of the running O row (rare after the first tiles)
          uint32_t o[32];
          tmem_ld_32x32(t_lane + kColO + half * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32(t_lane + kColO + half * 32, o);
        }
      }
      tmem_st_32x32(t_lane + kColP + half * 32, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      mbar_arrive_elect_addr(smem_u32(p_full));
      TRACE(half, j, 6, 0);
    }

    // ---- epilogue: O / l -> ctx ----
    xsum[half * kTile + row] = l;
    pair_sync(quad);
    const float l_row = l + xsum[(half ^ 1) * kTile + row];
    const float inv_l = 1.0f / l_row;
    mbar_wait(o_full, (n_kv - 1) & 1);
    tc_fence_after();
    const int q_row = q_tile * kTile + row;
    // log-sum-exp of the scaled scores in the log2 domain (m_used is the possibly stale maximum the
    // probabilities were formed against, so m_used + log2(sum) is exact): saved for dod_fmha_bwd
    if (p.lse != nullptr && half == 0 && q_row < p.seq)
      p.lse[(int64_t(b) * p.heads + head) * p.seq + q_row] = m_used + log2f(l_row);
    __nv_bfloat16* dst = p.ctx + (int64_t(b) * p.seq + q_row) * p.ldo + head * kD + half * 32;
    uint32_t o[32];
    tmem_ld_32x32(t_lane + kColO + half * 32, o);
    tmem_ld_wait();
    if (q_row < p.seq) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
        v.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
        v.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
        v.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
        *reinterpret_cast<uint4*>(dst + i) = v;
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


// =====================================================================================================
// fmha2_kernel -- persistent, software-pipelined variant (default; DOD_FMHA_V1=1 selects the kernel above).
//
// What the phase traces of fmha_kernel showed (profiles/r01_summary.md, r02_summary.md): the MUFU pipe (the
// floor of this problem at head dim 64: 1045 clk per 128x128 tile per SM, tools/ubench/softmax_mix.cu) was
// busy only ~79 % of the time because every tile's chain "wait S -> tcgen05.ld -> row max -> exchange -> exp2
// -> store P" is serial inside a CTA, and the ~3300 clk prologue / epilogue of each (image, head, q-tile) CTA
// is exposed.  Here ONE persistent CTA per SM streams work items, and the softmax warps run a register
// pipeline with no bubble between tiles or items:
//   * S, P and O are all double-buffered in TMEM (2 x 128 + 2 x 64 + 2 x 64 = 512 columns, the whole TMEM of
//     the SM).  Q.K^T of tile t+1 is issued as soon as the softmax warps have pulled S_{t-1} into registers
//     (in the middle of tile t-2), i.e. more than a tile time before it is needed; P.V of tile t reads P_t
//     from its own buffer, so the two MMA chains never wait for each other.
//   * 8 softmax warps (two threads per row, 64 key columns each): while the exp2 of tile t runs (MUFU-bound)
//     the same warps tcgen05.ld the raw scores of tile t+1 into a second register set, reduce their row
//     maximum (FMNMX3 in the MUFU shadow) and post it to the partner thread through shared memory, so at the
//     top of tile t+1 the maximum is already known.
//   * 4 epilogue warps normalise O by the row sum and store ctx (and the log-sum-exp) of item i while the
//     softmax warps are already in item i+1; a loader warp (TMA: Q double buffer, K / V rings of 3) and two
//     MMA-issuing warps (Q.K^T and P.V) complete the CTA.  setmaxnreg moves registers from those warps to
//     the softmax warps (two 64-register score sets live).
// =====================================================================================================
constexpr int kS2 = 2;                        // S buffers and P buffers in TMEM
constexpr int kKV2 = 3;                       // K and V ring depth
constexpr uint32_t kCol2S = 0;                // S_b at kCol2S + 128 b   (fp32 scores)
constexpr uint32_t kCol2P = 256;              // P_b at kCol2P + 64 b    (bf16 pairs)
constexpr uint32_t kCol2O = 384;              // O_b at kCol2O + 64 b
constexpr int kSoftmaxWarps2 = 8, kEpiWarps2 = 4;
constexpr int kLoadWarp2 = 12, kQkWarp2 = 13, kPvWarp2 = 14;
constexpr int kThreads2 = 16 * 32;            // 4 warpgroups (the 16th warp only takes part in setmaxnreg)
constexpr int kRegsSoftmax = 200, kRegsEpi = 80, kRegsMisc = 32;
static_assert(2 * 128 * kRegsSoftmax + 128 * kRegsEpi + 128 * kRegsMisc <= 65536, "register file");
constexpr int kSmem2Tiles = kTileBytes * (2 + 2 * kKV2);
constexpr int kSmem2Floats = 2 * 2 * kTile /*xmax*/ + 2 * 2 * kTile /*lsum*/ + 2 * kTile /*mref*/;
constexpr int kSmem2Bytes = kSmem2Tiles + 512 /*barriers*/ + kSmem2Floats * 4 + 1024;

#ifdef DOD_FMHA_TRACE
// fmha2 trace: g_trace2[slot][tile - kTraceT0][point]; slot 0 / 1 = softmax warp 0 / 4, 2 = Q.K^T thread, 3 = P.V thread
constexpr int kTraceT0 = 22, kTraceN = 14, kTracePts = 12;
__device__ long long g_trace2[4 * kTraceN * kTracePts];
#define TRACE2(slot, t, pt)                                                                  \
  do {                                                                                       \
    if (blockIdx.x == 7 && (t) >= kTraceT0 && (t) < kTraceT0 + kTraceN) {                    \
      long long t_;                                                                          \
      asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_)::"memory");                           \
      g_trace2[((slot) * kTraceN + ((t) - kTraceT0)) * kTracePts + (pt)] = t_;               \
    }                                                                                        \
  } while (0)
#define TRACE2S(c, t, pt)                                                                    \
  do {                                                                                       \
    if ((c).quad == 0 && (c).lane == 0) TRACE2((c).half, t, pt);                             \
  } while (0)
#else
#define TRACE2(slot, t, pt) do { } while (0)
#define TRACE2S(c, t, pt) do { } while (0)
#endif

template <int kRegs>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }

// compiler fence for registers filled by an asynchronous tcgen05.ld: uses placed after this cannot be
// scheduled before the tcgen05.wait::ld that precedes it (volatile asms keep their order)
__device__ __forceinline__ void regs_ready(uint32_t (&v)[64]) {
#pragma unroll
  for (int i = 0; i < 64; i += 16)
    asm volatile("" : "+r"(v[i]), "+r"(v[i + 1]), "+r"(v[i + 2]), "+r"(v[i + 3]), "+r"(v[i + 4]), "+r"(v[i + 5]),
                      "+r"(v[i + 6]), "+r"(v[i + 7]), "+r"(v[i + 8]), "+r"(v[i + 9]), "+r"(v[i + 10]),
                      "+r"(v[i + 11]), "+r"(v[i + 12]), "+r"(v[i + 13]), "+r"(v[i + 14]), "+r"(v[i + 15]));
}

struct Fmha2Params {
  FmhaParams f;
  int n_qt, n_items;  // q tiles per (image, head); work items = batch * heads * n_qt
};

struct SoftmaxState {
  float m_used, l;
  int pending;  // S / P buffer whose tcgen05.st of P is in flight and whose p_full arrival is still owed, or -1
};

struct SoftmaxCtx {
  uint32_t tmem, t_lane;
  uint64_t *s_full, *s_free, *p_full, *pv_done, *l_ready, *l_free;
  uint64_t *x_mine, *x_peer;  // [2 (tile parity)]: "row maxima of this half posted" barriers of this warp / its partner
  float *xmax, *lsum, *mref;
  int quad, half, row, lane;
  int n_kv, seq;
  float scale_log2;
};

__device__ __forceinline__ void mbar_arrive_elect(uint64_t* bar) { mbar_arrive_elect_addr(smem_u32(bar)); }

// exp2 of 16 raw scores cur[i0 .. i0+15] against the reference maximum, packed bf16 pairs written to cur[i0/2 ..]
// (the packed values trail the reads: index i / 2 <= i)
__device__ __forceinline__ void exp_group(uint32_t (&cur)[64], int i0, float2 sc2, float2 nm2, float2& sum2) {
#pragma unroll
  for (int i = i0; i < i0 + 16; i += 2) {
    const float2 x = __ffma2_rn(make_float2(__uint_as_float(cur[i]), __uint_as_float(cur[i + 1])), sc2, nm2);
    const float2 e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
    sum2 = __fadd2_rn(sum2, e);
    cur[i >> 1] = pack_bf16x2(e.x, e.y);
  }
}

// the P store of the previous tile has been issued; complete it and tell the P.V warp
__device__ __forceinline__ void finish_pending(const SoftmaxCtx& c, SoftmaxState& st) {
  if (st.pending >= 0) {
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    mbar_arrive_elect(&c.p_full[st.pending]);
    st.pending = -1;
  }
}

// One tile of the softmax pipeline.  `cur` holds the raw scores of tile t (this thread's 64 columns), already
// masked, and `mx` their row maximum over BOTH halves; `nxt` receives the scores of tile t + 1.  Returns the row
// maximum of tile t + 1.  Everything that is not exp2 (the next tile's TMEM load, its masking and row maximum,
// the P store and its completion, barrier traffic) is placed BETWEEN groups of 16 exp2 so that it issues in the
// shadow of the MUFU pipe; HAS_NEXT / MASK_NEXT are compile-time so that the tile body is straight-line code.
template <bool HAS_NEXT, bool MASK_NEXT>
__device__ __forceinline__ float softmax_tile(const SoftmaxCtx& c, SoftmaxState& st, uint32_t (&cur)[64],
                                              uint32_t (&nxt)[64], float mx, int t, int j, int item_local,
                                              int j_next) {
  const int buf = t & 1;
  if (j == 0) {
    st.m_used = -INFINITY;
    st.l = 0.0f;
  }
  const float m_new = fmaxf(st.m_used, mx * c.scale_log2);
  // lazy max: keep the stale max while it is within 2^8 of the true one
  const bool bump = (m_new - st.m_used) > 8.0f;
  float alpha = 1.0f;
  if (bump) {
    alpha = exp2f(st.m_used - m_new);  // 0 on the first tile of an item (m_used = -inf)
    st.m_used = m_new;
  }
  float2 sum2 = make_float2(0.0f, 0.0f);
  const float2 sc2 = make_float2(c.scale_log2, c.scale_log2);
  const float2 nm2 = make_float2(-st.m_used, -st.m_used);
  const int nbuf = (t + 1) & 1;
  TRACE2S(c, t, 0);
  // barrier tests are issued one exp2 group before their result is needed (a SYNCS round trip is ~200 clk and
  // the warp issues in order: consumed at once it would stall the MUFU stream of this warp)
  bool s_ok = true;
  if constexpr (HAS_NEXT) s_ok = mbar_test_wait(&c.s_full[nbuf], ((t + 1) >> 1) & 1);
  exp_group(cur, 0, sc2, nm2, sum2);
  TRACE2S(c, t, 1);
  if constexpr (HAS_NEXT) {
    // scores of tile t + 1: asynchronous TMEM -> registers (Q.K^T of tile t+1 retired a tile time ago)
    if (!s_ok) mbar_wait(&c.s_full[nbuf], ((t + 1) >> 1) & 1);
    tc_fence_after();
    tmem_ld_32x32(c.t_lane + kCol2S + nbuf * 128 + c.half * 64, *reinterpret_cast<uint32_t(*)[32]>(&nxt[0]));
    tmem_ld_32x32(c.t_lane + kCol2S + nbuf * 128 + c.half * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&nxt[32]));
  }
  TRACE2S(c, t, 2);
  exp_group(cur, 16, sc2, nm2, sum2);
  TRACE2S(c, t, 3);
  finish_pending(c, st);  // P of tile t-1 (its store was issued at the end of that tile)
  TRACE2S(c, t, 4);
  exp_group(cur, 32, sc2, nm2, sum2);
  TRACE2S(c, t, 5);
  float mx0 = -INFINITY, mx1 = -INFINITY;
  bool x_ok = true;
  // the P buffer of this tile was last read by P.V of tile t-2 (retired long ago: tested here, consumed below)
  const bool pv_ok = t < 2 || mbar_test_wait(&c.pv_done[buf], ((t >> 1) - 1) & 1);
  if constexpr (HAS_NEXT) {
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    mbar_arrive_elect(&c.s_free[nbuf]);  // S_{t+1} is in registers: Q.K^T of tile t+3 may overwrite it
    regs_ready(nxt);
    if constexpr (MASK_NEXT) {
      const int valid = c.seq - j_next * kTile - c.half * 64;
#pragma unroll
      for (int i = 0; i < 64; ++i)
        if (i >= valid) nxt[i] = 0xff800000u;
    }
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
      mx0 = fmaxf(mx0, fmaxf(__uint_as_float(nxt[i]), __uint_as_float(nxt[i + 1])));
      mx1 = fmaxf(mx1, fmaxf(__uint_as_float(nxt[i + 2]), __uint_as_float(nxt[i + 3])));
    }
    TRACE2S(c, t, 6);
    c.xmax[((t + 1) & 1) * 2 * kTile + c.half * kTile + c.row] = fmaxf(mx0, mx1);
    __syncwarp();
    mbar_arrive_elect(&c.x_mine[(t + 1) & 1]);
    TRACE2S(c, t, 7);
    x_ok = mbar_test_wait(&c.x_peer[(t + 1) & 1], ((t + 1) >> 1) & 1);
  }
  TRACE2S(c, t, 8);
  exp_group(cur, 48, sc2, nm2, sum2);
  TRACE2S(c, t, 9);
  st.l = st.l * alpha + (sum2.x + sum2.y);
  const int obuf = item_local & 1;
  if (j > 0 && __any_sync(0xffffffffu, bump)) {
    // rescale this thread's half of the running O row (rare after the first tiles): P.V of tile t-1 must
    // have retired, and P.V of tile t is only issued after this warp's p_full arrival
    mbar_wait(&c.pv_done[(t - 1) & 1], ((t - 1) >> 1) & 1);
    tc_fence_after();
    uint32_t o[32];
    tmem_ld_32x32(c.t_lane + kCol2O + obuf * 64 + c.half * 32, o);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
    tmem_st_32x32(c.t_lane + kCol2O + obuf * 64 + c.half * 32, o);
  }
  // P_t (this thread's 32 packed columns) into its own buffer.  The completion wait and the p_full arrival
  // are deferred into the next tile (finish_pending), under its exp2 groups.
  if (!pv_ok) mbar_wait(&c.pv_done[buf], ((t >> 1) - 1) & 1);
  tmem_st_32x32(c.t_lane + kCol2P + buf * 64 + c.half * 32, *reinterpret_cast<uint32_t(*)[32]>(&cur[0]));
  st.pending = buf;
  if (j == c.n_kv - 1) {
    // item finished: hand the row sum / reference maximum to the epilogue warps
    mbar_wait(&c.l_free[obuf], ((item_local >> 1) & 1) ^ 1);
    c.lsum[obuf * 2 * kTile + c.half * kTile + c.row] = st.l;
    if (c.half == 0) c.mref[obuf * kTile + c.row] = st.m_used;
    __syncwarp();
    mbar_arrive_elect(&c.l_ready[obuf]);  // mbarrier arrive has release semantics (CTA scope)
  }
  float mx_next = -INFINITY;
  if constexpr (HAS_NEXT) {
    // the partner posted its half's maxima of tile t+1 in the middle of ITS tile t: normally long done.
    // use n of slot (t+1)&1 is n = (t+1)/2, parity n & 1
    TRACE2S(c, t, 10);
    if (!x_ok) mbar_wait(&c.x_peer[(t + 1) & 1], ((t + 1) >> 1) & 1);
    TRACE2S(c, t, 11);
    mx_next = fmaxf(fmaxf(mx0, mx1), c.xmax[((t + 1) & 1) * 2 * kTile + (c.half ^ 1) * kTile + c.row]);
  } else {
    finish_pending(c, st);  // last tile of this CTA
  }
  return mx_next;
}

__global__ void __launch_bounds__(kThreads2, 1)
fmha2_kernel(const __grid_constant__ CUtensorMap tm_qkv, const Fmha2Params pp) {
  const FmhaParams& p = pp.f;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                             // [2]
  uint8_t* sK = sQ + 2 * kTileBytes;              // [kKV2]
  uint8_t* sV = sK + kKV2 * kTileBytes;           // [kKV2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kKV2 * kTileBytes);
  uint64_t* q_full = bars;                        // [2]
  uint64_t* q_empty = q_full + 2;                 // [2]
  uint64_t* k_full = q_empty + 2;                 // [kKV2]
  uint64_t* k_empty = k_full + kKV2;
  uint64_t* v_full = k_empty + kKV2;
  uint64_t* v_empty = v_full + kKV2;
  uint64_t* s_full = v_empty + kKV2;              // [kS2]
  uint64_t* s_free = s_full + kS2;                // [kS2]
  uint64_t* p_full = s_free + kS2;                // [kS2]
  uint64_t* pv_done = p_full + kS2;               // [kS2]
  uint64_t* o_ready = pv_done + kS2;              // [2]
  uint64_t* o_free = o_ready + 2;                 // [2]
  uint64_t* l_ready = o_free + 2;                 // [2]
  uint64_t* l_free = l_ready + 2;                 // [2]
  uint64_t* x_post = l_free + 2;                  // [2 (half)][4 (quad)][2 (tile parity)]: row maxima posted
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_post + 16);
  static_assert((4 + 4 * kKV2 + 4 * kS2 + 8 + 16) * 8 + 4 <= 512, "barrier area");
  float* xmax = reinterpret_cast<float*>(bars + 64);  // [2 (tile parity)][2 (half)][128]
  float* lsum = xmax + 2 * 2 * kTile;                 // [2 (item parity)][2 (half)][128]
  float* mref = lsum + 2 * 2 * kTile;                 // [2 (item parity)][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_kv = (p.seq + kTile - 1) / kTile;
  const int G = gridDim.x;
  const int my_items = (pp.n_items - int(blockIdx.x) + G - 1) / G;  // items blockIdx.x, + G, + 2G, ...

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_qkv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&o_ready[i], 1);
      mbar_init(&o_free[i], kEpiWarps2);
      mbar_init(&l_ready[i], kSoftmaxWarps2);
      mbar_init(&l_free[i], kEpiWarps2);
    }
    for (int i = 0; i < kKV2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < kS2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], kSoftmaxWarps2);
      mbar_init(&p_full[i], kSoftmaxWarps2);
      mbar_init(&pv_done[i], 1);
    }
    for (int i = 0; i < 16; ++i) mbar_init(&x_post[i], 1);
    fence_barrier_init();
  }
  if (warp == kQkWarp2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  auto item_coords = [&](int il, int& b, int& head, int& qt) {
    const int item = int(blockIdx.x) + il * G;
    qt = item % pp.n_qt;
    const int bh = item / pp.n_qt;
    head = bh % p.heads;
    b = bh / p.heads;
  };

  if (warp >= kLoadWarp2) {
    reg_dec<kRegsMisc>();
    if (warp == kLoadWarp2) {
      if (lane == 0) {
        // ---------------- TMA loader: Q_i, then K_j / V_j of the item, rings bounded by the empty barriers ----
        int kt = 0;
        for (int il = 0; il < my_items; ++il) {
          int b, head, qt;
          item_coords(il, b, head, qt);
          const int qc = p.q_off + head * kD, kc = p.k_off + head * kD, vc = p.v_off + head * kD;
          mbar_wait(&q_empty[il & 1], ((il >> 1) & 1) ^ 1);
          mbar_expect_tx(&q_full[il & 1], kTileBytes);
          tma_load_3d(sQ + (il & 1) * kTileBytes, &tm_qkv, &q_full[il & 1], qc, qt * kTile, b);
          for (int j = 0; j < n_kv; ++j, ++kt) {
            const int s = kt % kKV2;
            const uint32_t ph = ((kt / kKV2) & 1) ^ 1;
            mbar_wait(&k_empty[s], ph);
            mbar_expect_tx(&k_full[s], kTileBytes);
            tma_load_3d(sK + s * kTileBytes, &tm_qkv, &k_full[s], kc, j * kTile, b);
            mbar_wait(&v_empty[s], ph);
            mbar_expect_tx(&v_full[s], kTileBytes);
            tma_load_3d(sV + s * kTileBytes, &tm_qkv, &v_full[s], vc, j * kTile, b);
          }
        }
      }
    } else if (warp == kQkWarp2) {
      if (lane == 0) {
        // ---------------- Q.K^T issue: S_t <- Q_i K_j^T as soon as S_{t-2} has been pulled into registers ----
        constexpr uint32_t idesc_s = make_idesc_bf16(kTile, kTile, false, false);
        int t = 0;
        for (int il = 0; il < my_items; ++il) {
          mbar_wait(&q_full[il & 1], (il >> 1) & 1);
          const uint64_t dq = make_sdesc_sw128(smem_u32(sQ + (il & 1) * kTileBytes), 16, 1024);
          for (int j = 0; j < n_kv; ++j, ++t) {
            const int s = t % kKV2, buf = t % kS2;
            TRACE2(2, t, 0);
            mbar_wait(&k_full[s], (t / kKV2) & 1);
            TRACE2(2, t, 1);
            if (t >= kS2) mbar_wait(&s_free[buf], ((t / kS2) - 1) & 1);
            TRACE2(2, t, 2);
            tc_fence_after();
            const uint64_t dk = make_sdesc_sw128(smem_u32(sK + s * kTileBytes), 16, 1024);
#pragma unroll
            for (int k = 0; k < kD / 16; ++k)
              umma_ss(tmem + kCol2S + buf * 128, dq + uint64_t(2 * k), dk + uint64_t(2 * k), idesc_s, k != 0);
            umma_commit(&s_full[buf]);
            umma_commit(&k_empty[s]);
            if (j == n_kv - 1) umma_commit(&q_empty[il & 1]);
            TRACE2(2, t, 3);
          }
        }
      }
    } else if (warp == kPvWarp2) {
      if (lane == 0) {
        // ---------------- P.V issue: O_i += P_t V_j ----------------
        constexpr uint32_t idesc_o = make_idesc_bf16(kTile, kD, false, true);  // V MN-major
        int t = 0;
        for (int il = 0; il < my_items; ++il) {
          const int obuf = il & 1;
          for (int j = 0; j < n_kv; ++j, ++t) {
            const int s = t % kKV2, buf = t % kS2;
            TRACE2(3, t, 0);
            mbar_wait(&v_full[s], (t / kKV2) & 1);
            if (j == 0) mbar_wait(&o_free[obuf], ((il >> 1) & 1) ^ 1);  // epilogue of item il-2 has read O
            TRACE2(3, t, 1);
            mbar_wait(&p_full[buf], (t / kS2) & 1);
            TRACE2(3, t, 2);
            tc_fence_after();
            const uint32_t sv = smem_u32(sV + s * kTileBytes);
#pragma unroll
            for (int k = 0; k < kTile / 16; ++k) {
              const uint64_t dv = make_sdesc_sw128(sv + k * 16 * 128, 16, 1024);
              umma_ts(tmem + kCol2O + obuf * 64, tmem + kCol2P + buf * 64 + 8 * k, dv, idesc_o, (j | k) != 0);
            }
            umma_commit(&pv_done[buf]);
            umma_commit(&v_empty[s]);
            if (j == n_kv - 1) umma_commit(&o_ready[obuf]);
            TRACE2(3, t, 3);
          }
        }
      }
    }
  } else if (warp >= kSoftmaxWarps2) {
    // ---------------- epilogue warps: ctx = O / l (and the log-sum-exp), one thread per query row ----------------
    reg_dec<kRegsEpi>();
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem + (uint32_t(quad * 32) << 16);
    for (int il = 0; il < my_items; ++il) {
      int b, head, qt;
      item_coords(il, b, head, qt);
      const int obuf = il & 1;
      const uint32_t ph = (il >> 1) & 1;
      mbar_wait(&l_ready[obuf], ph);
      const float l_row = lsum[obuf * 2 * kTile + row] + lsum[obuf * 2 * kTile + kTile + row];
      const float m_row = mref[obuf * kTile + row];
      __syncwarp();
      if (lane == 0) mbar_arrive(&l_free[obuf]);
      mbar_wait(&o_ready[obuf], ph);
      tc_fence_after();
      uint32_t o[64];
      tmem_ld_32x32(t_lane + kCol2O + obuf * 64, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
      tmem_ld_32x32(t_lane + kCol2O + obuf * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[obuf]);
      const float inv_l = 1.0f / l_row;
      const int q_row = qt * kTile + row;
      if (q_row < p.seq) {
        // log-sum-exp of the scaled scores in the log2 domain (m_row is the possibly stale maximum the
        // probabilities were formed against, so m_row + log2(sum) is exact): saved for dod_fmha_bwd
        if (p.lse != nullptr) p.lse[(int64_t(b) * p.heads + head) * p.seq + q_row] = m_row + log2f(l_row);
        __nv_bfloat16* dst = p.ctx + (int64_t(b) * p.seq + q_row) * p.ldo + head * kD;
#pragma unroll
        for (int i = 0; i < 64; i += 8) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
          v.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
          v.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
          v.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + i) = v;
        }
      }
    }
  } else {
    // ---------------- softmax warps: two threads per query row, register pipeline over tiles and items ----------------
    reg_inc<kRegsSoftmax>();
    SoftmaxCtx c;
    c.tmem = tmem;
    c.quad = warp & 3;
    c.half = warp >> 2;
    c.lane = lane;
    c.row = c.quad * 32 + lane;
    c.t_lane = tmem + (uint32_t(c.quad * 32) << 16);
    c.s_full = s_full; c.s_free = s_free; c.p_full = p_full; c.pv_done = pv_done; c.l_ready = l_ready; c.l_free = l_free;
    c.xmax = xmax; c.lsum = lsum; c.mref = mref;
    c.x_mine = x_post + (c.half * 4 + c.quad) * 2;
    c.x_peer = x_post + ((c.half ^ 1) * 4 + c.quad) * 2;
    c.n_kv = n_kv; c.seq = p.seq; c.scale_log2 = p.scale_log2;
    const int total = my_items * n_kv;
    SoftmaxState st;
    st.m_used = -INFINITY;
    st.l = 0.0f;
    st.pending = -1;
    uint32_t ra[64], rb[64];
    float mx = -INFINITY;
    if (total > 0) {
      // very first tile of this CTA: nothing to overlap it with
      mbar_wait(&s_full[0], 0);
      tc_fence_after();
      tmem_ld_32x32(c.t_lane + kCol2S + c.half * 64, *reinterpret_cast<uint32_t(*)[32]>(&ra[0]));
      tmem_ld_32x32(c.t_lane + kCol2S + c.half * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&ra[32]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      mbar_arrive_elect(&s_free[0]);
      regs_ready(ra);
      const int valid = p.seq - c.half * 64;
      float m0 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        if (i >= valid) ra[i] = 0xff800000u;
        m0 = fmaxf(m0, __uint_as_float(ra[i]));
      }
      xmax[c.half * kTile + c.row] = m0;
      __syncwarp();
      mbar_arrive_elect(&c.x_mine[0]);
      mbar_wait(&c.x_peer[0], 0);
      mx = fmaxf(m0, xmax[(c.half ^ 1) * kTile + c.row]);
    }
    int il = 0, j = 0;
    // tile t lives in ra for even t, in rb for odd t; the body is instantiated for (has next tile, next tile is
    // the masked tail tile of an item) so that each instance is branch-free around its exp2 groups
    const bool ragged = (p.seq % kTile) != 0;
#define DOD_FMHA2_STEP(CUR, NXT, T)                                                                      \
  {                                                                                                      \
    const int jn = (j + 1 == n_kv) ? 0 : j + 1;                                                          \
    if ((T) + 1 >= total) mx = softmax_tile<false, false>(c, st, CUR, NXT, mx, (T), j, il, jn);          \
    else if (ragged && jn == n_kv - 1) mx = softmax_tile<true, true>(c, st, CUR, NXT, mx, (T), j, il, jn); \
    else mx = softmax_tile<true, false>(c, st, CUR, NXT, mx, (T), j, il, jn);                            \
    il += (jn == 0);                                                                                     \
    j = jn;                                                                                              \
  }
    for (int t = 0; t < total; t += 2) {
      DOD_FMHA2_STEP(ra, rb, t)
      if (t + 1 < total) DOD_FMHA2_STEP(rb, ra, t + 1)
    }
#undef DOD_FMHA2_STEP
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kQkWarp2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

}  // namespace
}  // namespace dod

#ifdef DOD_FMHA_TRACE
extern "C" DOD_API int32_t dod_debug_fmha2_trace(long long* host) {
  return int32_t(cudaMemcpyFromSymbol(host, dod::g_trace2, sizeof(dod::g_trace2)));
}
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
  static int use_v1 = -1;
  if (use_v1 < 0) {
    const char* e = getenv("DOD_FMHA_V2");  // persistent variant: measured slower (profiles/r02_summary.md)
    use_v1 = (e && atoi(e) != 0) ? 0 : 1;
  }
  if (use_v1) {
    dim3 grid((a->seq + kTile - 1) / kTile, a->heads, a->batch);
    fmha_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tm, p);
  } else {
    static PerDeviceOnce attr2_once;
    if (attr2_once.first())
      DOD_CUDA_OK(cudaFuncSetAttribute(fmha2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem2Bytes));
    Fmha2Params pp;
    pp.f = p;
    pp.n_qt = int((a->seq + kTile - 1) / kTile);
    const int64_t items = a->batch * a->heads * int64_t(pp.n_qt);
    DOD_REQUIRE(items < (1ll << 31), "dod_fmha_fwd: too many work items");
    pp.n_items = int(items);
    const int grid = pp.n_items < num_sms() ? pp.n_items : num_sms();
    fmha2_kernel<<<grid, kThreads2, kSmem2Bytes, stream>>>(tm, pp);
  }
  int rc = check_cuda(cudaGetLastError(), "fmha_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}
