// dod_fmha_bwd — fused backward of the head-dim-64 self-attention (dod_fmha_fwd) for sm_100a.
//
// Autograd of SDPA behind HF Dinov2SelfAttention (transformers modeling_dinov2.py:215-229) for the
// LoRA-wrapped encoder layers (reference models/dinov2_backbone.py:45-51; loss.backward() at
// train.py:1101).  Nothing of size [S, S] touches HBM: the probabilities are recomputed from Q, K
// and the forward's per-row log-sum-exp.
//
// One CTA per (image, head, 128-key tile j); it walks the query tiles i and keeps dK_j, dV_j in TMEM:
//     S  = Q_i K_j^T                      tcgen05.mma SS (K-major A, B)            -> TMEM
//     dP = dO_i V_j^T                     tcgen05.mma SS                            -> TMEM
//     P  = exp2(S c - lse_i)              8 softmax warps (two threads per query row), bf16 -> smem
//     dS = P o (dP - D_i) * scale                                                   bf16 -> smem
//     dV_j += P^T dO_i,  dK_j += dS^T Q_i   A = the P / dS tile read MN-major, B = the dO / Q tile
//                                           read MN-major (same shared-memory bytes as above)
//     dQ_i  = dS K_j                      A = dS K-major, B = K_j MN-major -> TMEM -> fp32 smem ->
//                                         TMA reduce-add into the fp32 dQ accumulator in HBM
// D_i = rowsum(dO_i o O_i) comes from dod_fmha_bwd_prep (HBM-bound, one pass over dO and O).
// TMEM: S 128 + dP 128 + dV 64 + dK 64 + dQ 64 = 448 columns; shared memory 192 KB: one CTA per SM.

#include "common.cuh"
#include "../../include/dod.h"

namespace dod {
void count_launch(int n = 1);
namespace {

constexpr int kD = 64;
constexpr int kT = 128;
constexpr int kTileBytes = kT * kD * 2;  // 16 KB: one [128 x 64] bf16 tile
constexpr int kBlkBytes = kT * 128;      // 16 KB: one 64-column block of a [128 x 128] bf16 tile
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColS = 0, kColdP = 128, kColdV = 256, kColdK = 320, kColdQ = 384;
constexpr int kSoftmaxWarps = 8;
constexpr int kThreads = 64 + kSoftmaxWarps * 32;
// K, V, Q[2], dO[2], P (2 blocks), dS (2 blocks), dQ staging (2 boxes of [128 x 32] f32)
constexpr int kSmemBytes = 6 * kTileBytes + 2 * 2 * kBlkBytes + 2 * kBlkBytes + 256 + 1024;

struct BwdParams {
  int seq, heads;
  int q_off, k_off, v_off;
  float scale, scale_log2;
  const float* lse;    // [B, H, S] log2-domain log-sum-exp of the scaled scores (forward)
  const float* dsum;   // [B, H, S] rowsum(dO o O)
  __nv_bfloat16* dqkv; // [B*S, ld_dqkv]: dK at k_off, dV at v_off (head h at + h*64)
  int64_t ld_dqkv;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, const void* smem_src, int32_t c0,
                                                  int32_t c1, int32_t c2) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(m)),
      "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void softmax_bar(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kSoftmaxWarps * 32) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
fmha_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ CUtensorMap tm_dq, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + kTileBytes;
  uint8_t* sQ = sV + kTileBytes;        // [2]
  uint8_t* sdO = sQ + 2 * kTileBytes;   // [2]
  uint8_t* sP = sdO + 2 * kTileBytes;   // [2 key blocks][128 q rows][128 B]
  uint8_t* sdS = sP + 2 * kBlkBytes;    // same layout
  uint8_t* sdQ = sdS + 2 * kBlkBytes;   // [2 boxes][128 q rows][32 f32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdQ + 2 * kBlkBytes);
  uint64_t* kv_full = bars + 0;
  uint64_t* q_full = bars + 1;     // [2]
  uint64_t* q_empty = bars + 3;    // [2]
  uint64_t* sdp_full = bars + 5;   // S and dP of this step are in TMEM
  uint64_t* pds_full = bars + 6;   // P and dS of this step are in shared memory (8 warp arrivals)
  uint64_t* pds_empty = bars + 7;  // the three products reading P / dS retired
  uint64_t* dq_full = bars + 8;    // dQ_i (and every earlier MMA) retired
  uint64_t* dq_free = bars + 9;    // dQ columns drained to registers (8 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x;  // key tile
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int n_q = (p.seq + kT - 1) / kT;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_qkv);
    prefetch_tmap(&tm_do);
    prefetch_tmap(&tm_dq);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, kSoftmaxWarps);
    mbar_init(pds_empty, 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, kSoftmaxWarps);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      mbar_expect_tx(kv_full, 2 * kTileBytes);
      tma_load_3d(sK, &tm_qkv, kv_full, p.k_off + head * kD, j * kT, b);
      tma_load_3d(sV, &tm_qkv, kv_full, p.v_off + head * kD, j * kT, b);
      for (int i = 0; i < n_q; ++i) {
        const int s = i & 1;
        mbar_wait(&q_empty[s], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[s], 2 * kTileBytes);
        tma_load_3d(sQ + s * kTileBytes, &tm_qkv, &q_full[s], p.q_off + head * kD, i * kT, b);
        tma_load_3d(sdO + s * kTileBytes, &tm_do, &q_full[s], head * kD, i * kT, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      constexpr uint32_t idesc_s = make_idesc_bf16(kT, kT, false, false);   // S, dP: K-major A and B
      constexpr uint32_t idesc_kv = make_idesc_bf16(kT, kD, true, true);    // dV, dK: MN-major A and B
      constexpr uint32_t idesc_q = make_idesc_bf16(kT, kD, false, true);    // dQ: K-major A, MN-major B
      const uint64_t d_k = make_sdesc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t d_v = make_sdesc_sw128(smem_u32(sV), 16, 1024);
      // [128 x 128] tiles as MN-major A (M = keys): two 64-key blocks kBlkBytes apart (LBO)
      const uint64_t d_pT = make_sdesc_sw128(smem_u32(sP), kBlkBytes, 1024);
      const uint64_t d_dsT = make_sdesc_sw128(smem_u32(sdS), kBlkBytes, 1024);
      mbar_wait(kv_full, 0);
      for (int i = 0; i < n_q; ++i) {
        const int s = i & 1;
        const uint64_t d_q = make_sdesc_sw128(smem_u32(sQ + s * kTileBytes), 16, 1024);
        const uint64_t d_do = make_sdesc_sw128(smem_u32(sdO + s * kTileBytes), 16, 1024);
        mbar_wait(&q_full[s], (i >> 1) & 1);
        tc_fence_after();
        // S / dP columns are free: the softmax warps read step i-1 out before pds_full(i-1) fired
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) umma_ss(tmem + kColS, d_q + uint64_t(2 * k), d_k + uint64_t(2 * k), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) umma_ss(tmem + kColdP, d_do + uint64_t(2 * k), d_v + uint64_t(2 * k), idesc_s, k != 0);
        umma_commit(sdp_full);
        mbar_wait(pds_full, i & 1);
        tc_fence_after();
        // reduction over the 128 query rows: 8 slices of 16 rows = 2048 B of every operand
#pragma unroll
        for (int k = 0; k < kT / 16; ++k)
          umma_ss(tmem + kColdV, d_pT + uint64_t(128 * k), d_do + uint64_t(128 * k), idesc_kv, (i | k) != 0);
#pragma unroll
        for (int k = 0; k < kT / 16; ++k)
          umma_ss(tmem + kColdK, d_dsT + uint64_t(128 * k), d_q + uint64_t(128 * k), idesc_kv, (i | k) != 0);
        umma_commit(&q_empty[s]);  // Q_i, dO_i are free once these retire
        if (i > 0) {
          mbar_wait(dq_free, (i - 1) & 1);
          tc_fence_after();
        }
        // dQ_i = dS K_j: reduction over the 128 keys; A K-major (4 slices of 32 B inside each 64-key
        // block), B = K_j MN-major (16 key rows = 2048 B per slice)
#pragma unroll
        for (int k = 0; k < kT / 16; ++k) {
          const uint64_t d_ds = make_sdesc_sw128(smem_u32(sdS + (k >> 2) * kBlkBytes), 16, 1024) + uint64_t(2 * (k & 3));
          umma_ss(tmem + kColdQ, d_ds, d_k + uint64_t(128 * k), idesc_q, k != 0);
        }
        umma_commit(dq_full);
        umma_commit(pds_empty);
      }
    }
  } else {
    // ---------------- softmax / epilogue warps: two threads per row ----------------
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;  // S / dP columns [half*64, +64); dQ, dK, dV columns [half*32, +32)
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem + (uint32_t(quad * 32) << 16);
    const int bh = b * p.heads + head;
    const int valid_keys = p.seq - j * kT - half * 64;  // keys of this thread's 64 columns inside the sequence
    uint8_t* p_row = sP + half * kBlkBytes + row * 128;
    uint8_t* ds_row = sdS + half * kBlkBytes + row * 128;
    uint8_t* dq_row = sdQ + half * kBlkBytes + row * 128;
    const bool issuer = threadIdx.x == 64;

    for (int i = 0; i < n_q; ++i) {
      const int q_idx = i * kT + row;
      float lse = INFINITY, dsum = 0.0f;  // rows past the sequence: P = exp2(-inf) = 0
      if (q_idx < p.seq) {
        lse = p.lse[int64_t(bh) * p.seq + q_idx];
        dsum = p.dsum[int64_t(bh) * p.seq + q_idx];
      }
      mbar_wait(sdp_full, i & 1);
      tc_fence_after();
      if (i > 0) mbar_wait(pds_empty, (i - 1) & 1);  // the products of step i-1 no longer read P / dS
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(t_lane + kColS + half * 64 + c * 32, sv);
        tmem_ld_32x32(t_lane + kColdP + half * 64 + c * 32, dv);
        tmem_ld_wait();
        uint32_t pk[16], dk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float p0 = ex2_approx(fmaf(__uint_as_float(sv[e]), p.scale_log2, -lse));
          float p1 = ex2_approx(fmaf(__uint_as_float(sv[e + 1]), p.scale_log2, -lse));
          if (c * 32 + e >= valid_keys) p0 = 0.0f;
          if (c * 32 + e + 1 >= valid_keys) p1 = 0.0f;
          const float d0 = p0 * (__uint_as_float(dv[e]) - dsum) * p.scale;
          const float d1 = p1 * (__uint_as_float(dv[e + 1]) - dsum) * p.scale;
          pk[e >> 1] = pack_bf16x2(p0, p1);
          dk[e >> 1] = pack_bf16x2(d0, d1);
        }
        // 32 keys = 4 chunks of 16 B; 128B swizzle: chunk ^= row & 7
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int off = (((c * 4 + ch) ^ (row & 7)) << 4);
          *reinterpret_cast<uint4*>(p_row + off) = make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
          *reinterpret_cast<uint4*>(ds_row + off) = make_uint4(dk[4 * ch], dk[4 * ch + 1], dk[4 * ch + 2], dk[4 * ch + 3]);
        }
      }
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the MMA's async-proxy reads
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);

      // ---- drain dQ_i: TMEM -> fp32 staging -> TMA reduce-add into the dQ accumulator ----
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      uint32_t q[32];
      tmem_ld_32x32(t_lane + kColdQ + half * 32, q);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
      if (issuer) tma_store_wait_read<0>();  // the previous step's reduce has read the staging tile
      softmax_bar(2);
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
        *reinterpret_cast<uint4*>(dq_row + ((ch ^ (row & 7)) << 4)) = make_uint4(q[4 * ch], q[4 * ch + 1], q[4 * ch + 2], q[4 * ch + 3]);
      fence_proxy_async_smem();
      softmax_bar(1);
      if (issuer) {
        tma_reduce_add_3d(&tm_dq, sdQ, head * kD, i * kT, b);
        tma_reduce_add_3d(&tm_dq, sdQ + kBlkBytes, head * kD + 32, i * kT, b);
        tma_store_commit();
      }
    }

    // ---- dK_j, dV_j: every MMA retired before the last dq_full fired ----
    const int key = j * kT + row;
    __nv_bfloat16* base = p.dqkv + (int64_t(b) * p.seq + key) * p.ld_dqkv + head * kD + half * 32;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      uint32_t o[32];
      tmem_ld_32x32(t_lane + (which == 0 ? kColdK : kColdV) + half * 32, o);  // warp-collective: all lanes
      tmem_ld_wait();
      if (key < p.seq) {
        __nv_bfloat16* dst = base + (which == 0 ? p.k_off : p.v_off);
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[e]), __uint_as_float(o[e + 1]));
          v.y = pack_bf16x2(__uint_as_float(o[e + 2]), __uint_as_float(o[e + 3]));
          v.z = pack_bf16x2(__uint_as_float(o[e + 4]), __uint_as_float(o[e + 5]));
          v.w = pack_bf16x2(__uint_as_float(o[e + 6]), __uint_as_float(o[e + 7]));
          *reinterpret_cast<uint4*>(dst + e) = v;
        }
      }
    }
    if (issuer) tma_store_wait<0>();  // shared memory must outlive the bulk reduce
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem);
  }
}

// D[b, h, s] = sum_d dO[b, s, h*64 + d] * O[b, s, h*64 + d]; one thread per (row, head)
__global__ void fmha_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                                     float* __restrict__ dsum, int64_t batch, int seq, int heads, int64_t ldo,
                                     int64_t lddo) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = batch * seq * heads;
  if (idx >= total) return;
  const int h = int(idx % heads);
  const int64_t r = idx / heads;  // b * seq + s
  const uint4* po = reinterpret_cast<const uint4*>(o + r * ldo + h * kD);
  const uint4* pd = reinterpret_cast<const uint4*>(d_o + r * lddo + h * kD);
  float acc = 0.0f;
#pragma unroll
  for (int i = 0; i < kD / 8; ++i) {
    const uint4 a = po[i], c = pd[i];
    const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* c2 = reinterpret_cast<const __nv_bfloat162*>(&c);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 x = __bfloat1622float2(a2[e]), y = __bfloat1622float2(c2[e]);
      acc = fmaf(x.x, y.x, acc);
      acc = fmaf(x.y, y.y, acc);
    }
  }
  const int64_t bi = r / seq;
  const int s = int(r - bi * seq);
  dsum[(bi * heads + h) * seq + s] = acc;
}

}  // namespace
}  // namespace dod

extern "C" int32_t dod_fmha_bwd(const dod_fmha_bwd_args* a, dod_stream_t stream_) {
  using namespace dod;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->qkv && a->ctx && a->dctx && a->lse && a->dsum && a->dq_acc && a->dqkv,
              "dod_fmha_bwd: null pointer");
  DOD_REQUIRE(a->batch > 0 && a->seq > 0 && a->heads > 0, "dod_fmha_bwd: empty problem");
  DOD_REQUIRE(a->batch <= 65535 && a->heads <= 65535, "dod_fmha_bwd: batch/heads exceed grid limits");
  DOD_REQUIRE(a->ld % 8 == 0 && a->ldo % 8 == 0 && a->lddo % 8 == 0 && a->ld_dqkv % 8 == 0 && a->ld_dq % 4 == 0,
              "dod_fmha_bwd: leading dimensions must keep 16-byte alignment");
  DOD_REQUIRE((uintptr_t(a->qkv) & 15) == 0 && (uintptr_t(a->ctx) & 15) == 0 && (uintptr_t(a->dctx) & 15) == 0 &&
                  (uintptr_t(a->dq_acc) & 15) == 0 && (uintptr_t(a->dqkv) & 15) == 0,
              "dod_fmha_bwd: buffers must be 16-byte aligned");
  DOD_REQUIRE(a->q_off % 8 == 0 && a->k_off % 8 == 0 && a->v_off % 8 == 0,
              "dod_fmha_bwd: q/k/v column offsets must be multiples of 8");
  const int64_t hd = a->heads * kD;
  DOD_REQUIRE(a->q_off + hd <= a->ld && a->k_off + hd <= a->ld && a->v_off + hd <= a->ld && hd <= a->ldo &&
                  hd <= a->lddo && hd <= a->ld_dq && a->k_off + hd <= a->ld_dqkv && a->v_off + hd <= a->ld_dqkv,
              "dod_fmha_bwd: head slices exceed the row");
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    DOD_CUDA_OK(cudaFuncSetAttribute(fmha_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  }
  {
    const int64_t total = a->batch * a->seq * a->heads;
    fmha_bwd_prep_kernel<<<unsigned((total + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(a->ctx), reinterpret_cast<const __nv_bfloat16*>(a->dctx), a->dsum,
        a->batch, int(a->seq), int(a->heads), a->ldo, a->lddo);
    if (int rc = check_cuda(cudaGetLastError(), "fmha_bwd_prep_kernel launch")) return rc;
  }
  CUtensorMap tm_qkv, tm_do, tm_dq;
  if (int rc = make_tmap_3d(&tm_qkv, a->qkv, 2, a->batch, a->seq, a->ld, a->seq * a->ld, a->ld, kT, kD)) return rc;
  if (int rc = make_tmap_3d(&tm_do, a->dctx, 2, a->batch, a->seq, a->lddo, a->seq * a->lddo, a->lddo, kT, kD)) return rc;
  if (int rc = make_tmap_3d(&tm_dq, a->dq_acc, 4, a->batch, a->seq, a->ld_dq, a->seq * a->ld_dq, a->ld_dq, kT, 32)) return rc;
  BwdParams p;
  p.seq = int(a->seq);
  p.heads = int(a->heads);
  p.q_off = int(a->q_off);
  p.k_off = int(a->k_off);
  p.v_off = int(a->v_off);
  p.scale = a->scale;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.lse = a->lse;
  p.dsum = a->dsum;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(a->dqkv);
  p.ld_dqkv = a->ld_dqkv;
  dim3 grid(unsigned((a->seq + kT - 1) / kT), unsigned(a->heads), unsigned(a->batch));
  fmha_bwd_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tm_qkv, tm_do, tm_dq, p);
  int rc = check_cuda(cudaGetLastError(), "fmha_bwd_kernel launch");
  if (rc == 0) count_launch(2);
  return rc;
}
