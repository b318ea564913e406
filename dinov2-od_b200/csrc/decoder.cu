// Decoder-side kernels of libdod: few-query attention, "deformable" bilinear
// sampling and the small row utilities around the head GEMMs.  All of these are
// latency / L2-bound (B*Q <= a few thousand rows); the dense projections around
// them go through dod_gemm_bf16.
//
//  dod_mha_small      nn.MultiheadAttention core for Lq <= ~100 queries
//                     (reference deformable_attention.py:232-233, detr_decoder.py:29-35)
//  dod_deform_sample  reference deformable_attention.py:100-178 (the 4-deep python
//                     loop with .item()), one CTA per (image, query)
//  dod_rowcopy / dod_broadcast_rows / dod_cast_pad_bf16 / dod_split3_bf16
#include "common.cuh"
#include "../../include/dod.h"

namespace dod {
void count_launch(int n = 1);
namespace {

// ---------------------------------------------------------------------------
// few-query multi-head attention (generic head dim, fp32 math)
// ---------------------------------------------------------------------------
constexpr int kQT = 16;        // queries per CTA
constexpr int kMhaThreads = 256;

template <typename T>
__device__ __forceinline__ float ld1(const T* p);
template <>
__device__ __forceinline__ float ld1<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld1<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T>
__device__ __forceinline__ void st1(T* p, float v);
template <>
__device__ __forceinline__ void st1<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void st1<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

// Self-attention among the queries (Lq, Lk <= 128): the whole (image, head) problem lives in shared
// memory.  One CTA per (image, head); Q (pre-scaled), K, V rows staged as fp32 with an odd row pitch,
// scores by thread per (query, key) pair, softmax by warp per row, output by thread per (query, channel).
// The key-per-thread kernel below leaves 80 % of its threads idle at Lk = 50 (185 us per decoder layer at
// batch 64; this one: a few microseconds).
// 8 consecutive elements of a row -> fp32 (16-byte load for bf16 when VEC, else scalar loads)
template <typename T, bool VEC>
__device__ __forceinline__ void ld_row8(const T* p, float (&v)[8]) {
  if constexpr (VEC && sizeof(T) == 2) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = ld1(p + i);
  }
}

// VEC: dh % 8 == 0 and 16-byte-aligned rows.  Scores and outputs are register-tiled (4 queries x 2 keys,
// 4 queries x 1 channel per thread): the first version read two shared-memory words per FMA and spent 100 us
// per decoder layer at batch 64 on 0.5 GFLOP.
template <typename T, bool VEC>
__global__ void __launch_bounds__(kMhaThreads)
mha_tiny_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                T* __restrict__ out, int lq, int lk, int dh, int64_t ldq, int64_t ldk, int64_t ldv,
                int64_t ldo, float scale) {
  extern __shared__ float sm[];
  // VEC (bf16 inputs): K and V stay bf16 in shared memory (row pitch dh + 2 halves = an odd number of words,
  // conflict-free for consecutive rows) and are widened on use -- the same fp32 values, so the same bits out --
  // which takes the CTA from 68 to 49 KB at 50 x 50 x 96: four CTAs per SM instead of three, and the 512
  // (image, head) problems of a 64-image batch fit ONE wave (592 slots) instead of 1.15 (56 -> ~35 us per layer).
  constexpr bool K16 = VEC && sizeof(T) == 2;
  const int pitch = dh | 1, lkp = lk | 1;
  const int pitch16 = dh + 2;      // halves
  float* sq = sm;                  // [lq][pitch]
  float* sk = sq + lq * pitch;     // [lk][pitch] f32, or [lk][pitch16] bf16 followed by V in the same form
  float* sv = K16 ? sk + (lk * pitch16 + 1) / 2 : sk + lk * pitch;
  float* sp = K16 ? sv + (lk * pitch16 + 1) / 2 : sv + lk * pitch;  // [lq][lkp]
  const uint16_t* sk16 = reinterpret_cast<const uint16_t*>(sk);
  const uint16_t* sv16 = reinterpret_cast<const uint16_t*>(sv);
  auto kval = [&](int r, int d) -> float {
    if constexpr (K16) return __uint_as_float(uint32_t(sk16[r * pitch16 + d]) << 16);
    else return sk[r * pitch + d];
  };
  auto vval = [&](int r, int d) -> float {
    if constexpr (K16) return __uint_as_float(uint32_t(sv16[r * pitch16 + d]) << 16);
    else return sv[r * pitch + d];
  };
  const int h = blockIdx.x, b = blockIdx.y;
  const int t = threadIdx.x;
  if constexpr (VEC) {
    const int dv = dh >> 3;
    for (int i = t; i < lq * dv; i += kMhaThreads) {
      const int r = i / dv, c = (i - r * dv) << 3;
      float x[8];
      ld_row8<T, true>(q + (int64_t(b) * lq + r) * ldq + h * dh + c, x);
#pragma unroll
      for (int e = 0; e < 8; ++e) sq[r * pitch + c + e] = x[e] * scale;
    }
    for (int i = t; i < lk * dv; i += kMhaThreads) {
      const int r = i / dv, c = (i - r * dv) << 3;
      if constexpr (K16) {
        // rows start word-aligned (pitch16 even), c is a multiple of 8: four 32-bit stores per operand
        const uint4 kq = *reinterpret_cast<const uint4*>(k + (int64_t(b) * lk + r) * ldk + h * dh + c);
        const uint4 vq = *reinterpret_cast<const uint4*>(v + (int64_t(b) * lk + r) * ldv + h * dh + c);
        uint32_t* kd = reinterpret_cast<uint32_t*>(sk) + ((r * pitch16 + c) >> 1);
        uint32_t* vd = reinterpret_cast<uint32_t*>(sv) + ((r * pitch16 + c) >> 1);
        kd[0] = kq.x; kd[1] = kq.y; kd[2] = kq.z; kd[3] = kq.w;
        vd[0] = vq.x; vd[1] = vq.y; vd[2] = vq.z; vd[3] = vq.w;
      } else {
        float x[8], y[8];
        ld_row8<T, true>(k + (int64_t(b) * lk + r) * ldk + h * dh + c, x);
        ld_row8<T, true>(v + (int64_t(b) * lk + r) * ldv + h * dh + c, y);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          sk[r * pitch + c + e] = x[e];
          sv[r * pitch + c + e] = y[e];
        }
      }
    }
  } else {
    for (int i = t; i < lq * dh; i += kMhaThreads) {
      const int r = i / dh, d = i - r * dh;
      sq[r * pitch + d] = ld1(q + (int64_t(b) * lq + r) * ldq + h * dh + d) * scale;
    }
    for (int i = t; i < lk * dh; i += kMhaThreads) {
      const int r = i / dh, d = i - r * dh;
      sk[r * pitch + d] = ld1(k + (int64_t(b) * lk + r) * ldk + h * dh + d);
      sv[r * pitch + d] = ld1(v + (int64_t(b) * lk + r) * ldv + h * dh + d);
    }
  }
  __syncthreads();
  // scores: thread = 4 queries x 2 keys (keys kb and kb + nkb: consecutive threads read consecutive K rows,
  // odd pitch -> no bank conflicts; the Q words are warp-wide broadcasts)
  const int nqb = (lq + 3) >> 2, nkb = (lk + 1) >> 1;
  for (int item = t; item < nqb * nkb; item += kMhaThreads) {
    const int qb = item / nkb, kb = item - qb * nkb;
    const int q0 = qb << 2, k0 = kb, k1 = kb + nkb;
    const float* a0 = sq + min(q0, lq - 1) * pitch;
    const float* a1 = sq + min(q0 + 1, lq - 1) * pitch;
    const float* a2 = sq + min(q0 + 2, lq - 1) * pitch;
    const float* a3 = sq + min(q0 + 3, lq - 1) * pitch;
    const int r0 = k0, r1 = min(k1, lk - 1);
    float acc[4][2] = {};
    for (int d = 0; d < dh; ++d) {
      const float x0 = kval(r0, d), x1 = kval(r1, d);
      const float y0 = a0[d], y1 = a1[d], y2 = a2[d], y3 = a3[d];
      acc[0][0] = fmaf(y0, x0, acc[0][0]); acc[0][1] = fmaf(y0, x1, acc[0][1]);
      acc[1][0] = fmaf(y1, x0, acc[1][0]); acc[1][1] = fmaf(y1, x1, acc[1][1]);
      acc[2][0] = fmaf(y2, x0, acc[2][0]); acc[2][1] = fmaf(y2, x1, acc[2][1]);
      acc[3][0] = fmaf(y3, x0, acc[3][0]); acc[3][1] = fmaf(y3, x1, acc[3][1]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (q0 + i < lq) {
        sp[(q0 + i) * lkp + k0] = acc[i][0];
        if (k1 < lk) sp[(q0 + i) * lkp + k1] = acc[i][1];
      }
    }
  }
  __syncthreads();
  const int warp = t >> 5, lane = t & 31;
  for (int qi = warp; qi < lq; qi += kMhaThreads / 32) {
    float* row = sp + qi * lkp;
    float mx = -INFINITY;
    for (int j = lane; j < lk; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    float s_ = 0.f;
    for (int j = lane; j < lk; j += 32) {
      const float e = __expf(row[j] - mx);
      row[j] = e;
      s_ += e;
    }
    const float inv = 1.0f / warp_sum(s_);
    for (int j = lane; j < lk; j += 32) row[j] *= inv;
  }
  __syncthreads();
  // output: thread = 4 queries x 1 channel (consecutive threads: consecutive channels of V, P words broadcast)
  for (int item = t; item < nqb * dh; item += kMhaThreads) {
    const int qb = item / dh, d = item - qb * dh;
    const int q0 = qb << 2;
    const float* p0 = sp + min(q0, lq - 1) * lkp;
    const float* p1 = sp + min(q0 + 1, lq - 1) * lkp;
    const float* p2 = sp + min(q0 + 2, lq - 1) * lkp;
    const float* p3 = sp + min(q0 + 3, lq - 1) * lkp;
    float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
    for (int j = 0; j < lk; ++j) {
      const float x = vval(j, d);
      o0 = fmaf(p0[j], x, o0);
      o1 = fmaf(p1[j], x, o1);
      o2 = fmaf(p2[j], x, o2);
      o3 = fmaf(p3[j], x, o3);
    }
    T* op = out + (int64_t(b) * lq + q0) * ldo + h * dh + d;
    st1(op, o0);
    if (q0 + 1 < lq) st1(op + ldo, o1);
    if (q0 + 2 < lq) st1(op + 2 * ldo, o2);
    if (q0 + 3 < lq) st1(op + 3 * ldo, o3);
  }
}

template <typename T>
__global__ void __launch_bounds__(kMhaThreads)
mha_small_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                 T* __restrict__ out, int lq, int lk, int dh, int64_t ldq, int64_t ldk, int64_t ldv,
                 int64_t ldo, float scale, int lk_pad) {
  extern __shared__ float sm[];
  float* sq = sm;                 // [kQT][dh]
  float* sp = sm + kQT * dh;      // [kQT][lk_pad]
  float* sinv = sp + kQT * lk_pad;  // [kQT]

  const int q0 = blockIdx.x * kQT;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int nq = min(kQT, lq - q0);
  const int t = threadIdx.x;

  for (int i = t; i < kQT * dh; i += kMhaThreads) {
    const int qi = i / dh, d = i - qi * dh;
    sq[i] = qi < nq ? ld1(q + (int64_t(b) * lq + q0 + qi) * ldq + h * dh + d) * scale : 0.f;
  }
  __syncthreads();

  // scores: one key per thread, kQT running dot products
  for (int key = t; key < lk; key += kMhaThreads) {
    const T* kp = k + (int64_t(b) * lk + key) * ldk + h * dh;
    float acc[kQT];
#pragma unroll
    for (int i = 0; i < kQT; ++i) acc[i] = 0.f;
    for (int d = 0; d < dh; ++d) {
      const float kv = ld1(kp + d);
#pragma unroll
      for (int i = 0; i < kQT; ++i) acc[i] = fmaf(sq[i * dh + d], kv, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < kQT; ++i) sp[i * lk_pad + key] = acc[i];
  }
  __syncthreads();

  // softmax per query row: one warp per row
  const int warp = t >> 5, lane = t & 31;
  for (int qi = warp; qi < nq; qi += kMhaThreads / 32) {
    float* row = sp + qi * lk_pad;
    float mx = -INFINITY;
    for (int j = lane; j < lk; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int j = lane; j < lk; j += 32) {
      const float e = __expf(row[j] - mx);
      row[j] = e;
      s += e;
    }
    s = warp_sum(s);
    if (lane == 0) sinv[qi] = 1.0f / s;
  }
  __syncthreads();

  // out[q][d] = sum_k p[q][k] v[k][d]; thread -> (d, query group)
  // slot s of group g owns query row g + s * ngroups
  const int ngroups = kMhaThreads / dh;  // >= 1 (dh <= kMhaThreads)
  if (t < ngroups * dh) {
    const int d = t % dh, g = t / dh;
    float acc[kQT];
#pragma unroll
    for (int s = 0; s < kQT; ++s) acc[s] = 0.f;
    const T* vp = v + int64_t(b) * lk * ldv + h * dh + d;
    const int nslots = (nq - g + ngroups - 1) / ngroups;  // rows owned by this group (may be <= 0)
    for (int key = 0; key < lk; ++key) {
      const float vv = ld1(vp + int64_t(key) * ldv);
#pragma unroll
      for (int s = 0; s < kQT; ++s)
        if (s < nslots) acc[s] = fmaf(sp[(g + s * ngroups) * lk_pad + key], vv, acc[s]);
    }
#pragma unroll
    for (int s = 0; s < kQT; ++s)
      if (s < nslots) {
        const int qi = g + s * ngroups;
        st1(out + (int64_t(b) * lq + q0 + qi) * ldo + h * dh + d, acc[s] * sinv[qi]);
      }
  }
}

// ---------------------------------------------------------------------------
// deformable sampling: one CTA per (image, query), one thread per channel
// ---------------------------------------------------------------------------
template <typename TV, typename TO>
__global__ void deform_sample_kernel(const TV* __restrict__ value, const float* __restrict__ ref,
                                     const float* __restrict__ offs, const float* __restrict__ logits,
                                     TO* __restrict__ out, int queries, int heads, int points,
                                     int dh, int gh, int gw, int64_t ldv, int64_t ldref, int64_t ldoffs,
                                     int64_t ldlog, int64_t ldo, int ref_is_logit) {
  const int64_t row = blockIdx.x;  // b * Q + q
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
    float acc = 0.f;
    for (int p = 0; p < points; ++p) {
      const float wgt = expf(lg[p] - mx) / den;
      const float lx = fminf(fmaxf(rx + of[2 * p + 0], 0.f), 1.f);
      const float ly = fminf(fmaxf(ry + of[2 * p + 1], 0.f), 1.f);
      const float sx = lx * float(gw - 1), sy = ly * float(gh - 1);
      int x0 = int(floorf(sx)), y0 = int(floorf(sy));
      int x1 = x0 + 1, y1 = y0 + 1;
      x0 = min(max(x0, 0), gw - 1);
      x1 = min(max(x1, 0), gw - 1);
      y0 = min(max(y0, 0), gh - 1);
      y1 = min(max(y1, 0), gh - 1);
      const float wx1 = sx - float(x0), wx0 = 1.0f - wx1;
      const float wy1 = sy - float(y0), wy0 = 1.0f - wy1;
      const TV* vb = value + int64_t(b) * hw * ldv + c;
      const float v00 = ld1(vb + (int64_t(y0) * gw + x0) * ldv);
      const float v01 = ld1(vb + (int64_t(y1) * gw + x0) * ldv);
      const float v10 = ld1(vb + (int64_t(y0) * gw + x1) * ldv);
      const float v11 = ld1(vb + (int64_t(y1) * gw + x1) * ldv);
      float s = __fmul_rn(v00, __fmul_rn(wx0, wy0));
      s = __fadd_rn(s, __fmul_rn(v01, __fmul_rn(wx0, wy1)));
      s = __fadd_rn(s, __fmul_rn(v10, __fmul_rn(wx1, wy0)));
      s = __fadd_rn(s, __fmul_rn(v11, __fmul_rn(wx1, wy1)));
      acc = __fadd_rn(acc, __fmul_rn(s, wgt));
    }
    st1(out + row * ldo + c, acc);
  }
}

// Vectorised form for bf16 values: a thread owns EIGHT adjacent channels of one (image, query) row (one 16-byte
// gather per corner), so the per-head scalars -- softmax over the points, clamped sampling position, corner
// indices and weights -- are evaluated once per eight channels instead of once per channel, and several rows
// share a CTA.  The per-channel arithmetic (products, order of the four corner terms and of the points) is the
// one of deform_sample_kernel: results are bit-identical (tests/test_kernels_gpu.py).
template <typename TO>
__global__ void __launch_bounds__(256)
deform_sample_vec_kernel(const __nv_bfloat16* __restrict__ value, const float* __restrict__ ref,
                         const float* __restrict__ offs, const float* __restrict__ logits, TO* __restrict__ out,
                         int64_t rows, int queries, int heads, int points, int dh, int gh, int gw, int64_t ldv,
                         int64_t ldref, int64_t ldoffs, int64_t ldlog, int64_t ldo, int ref_is_logit) {
  const int groups = heads * dh / 8;                       // 8-channel groups per row
  const int rows_per_cta = blockDim.x / groups;
  const int r_in = threadIdx.x / groups;
  if (r_in >= rows_per_cta) return;
  const int64_t row = int64_t(blockIdx.x) * rows_per_cta + r_in;
  if (row >= rows) return;
  const int c = (threadIdx.x - r_in * groups) * 8;
  const int b = int(row / queries);
  const int64_t hw = int64_t(gh) * gw;
  float rx = ref[row * ldref + 0], ry = ref[row * ldref + 1];
  if (ref_is_logit) {
    rx = 1.0f / (1.0f + expf(-rx));
    ry = 1.0f / (1.0f + expf(-ry));
  }
  const int h = c / dh;
  const float* lg = logits + row * ldlog + h * points;
  const float* of = offs + row * ldoffs + h * points * 2;
  float mx = -INFINITY;
  for (int p = 0; p < points; ++p) mx = fmaxf(mx, lg[p]);
  float den = 0.f;
  for (int p = 0; p < points; ++p) den += expf(lg[p] - mx);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int p = 0; p < points; ++p) {
    const float wgt = expf(lg[p] - mx) / den;
    const float lx = fminf(fmaxf(rx + of[2 * p + 0], 0.f), 1.f);
    const float ly = fminf(fmaxf(ry + of[2 * p + 1], 0.f), 1.f);
    const float sx = lx * float(gw - 1), sy = ly * float(gh - 1);
    int x0 = int(floorf(sx)), y0 = int(floorf(sy));
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 = min(max(x0, 0), gw - 1);
    x1 = min(max(x1, 0), gw - 1);
    y0 = min(max(y0, 0), gh - 1);
    y1 = min(max(y1, 0), gh - 1);
    const float wx1 = sx - float(x0), wx0 = 1.0f - wx1;
    const float wy1 = sy - float(y0), wy0 = 1.0f - wy1;
    const float w00 = __fmul_rn(wx0, wy0), w01 = __fmul_rn(wx0, wy1), w10 = __fmul_rn(wx1, wy0),
                w11 = __fmul_rn(wx1, wy1);
    const __nv_bfloat16* vb = value + int64_t(b) * hw * ldv + c;
    const uint4 q00 = __ldg(reinterpret_cast<const uint4*>(vb + (int64_t(y0) * gw + x0) * ldv));
    const uint4 q01 = __ldg(reinterpret_cast<const uint4*>(vb + (int64_t(y1) * gw + x0) * ldv));
    const uint4 q10 = __ldg(reinterpret_cast<const uint4*>(vb + (int64_t(y0) * gw + x1) * ldv));
    const uint4 q11 = __ldg(reinterpret_cast<const uint4*>(vb + (int64_t(y1) * gw + x1) * ldv));
    const uint32_t a00[4] = {q00.x, q00.y, q00.z, q00.w}, a01[4] = {q01.x, q01.y, q01.z, q01.w};
    const uint32_t a10[4] = {q10.x, q10.y, q10.z, q10.w}, a11[4] = {q11.x, q11.y, q11.z, q11.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int w = i >> 1;
      const bool hi = i & 1;
      const float v00 = __uint_as_float(hi ? (a00[w] & 0xffff0000u) : (a00[w] << 16));
      const float v01 = __uint_as_float(hi ? (a01[w] & 0xffff0000u) : (a01[w] << 16));
      const float v10 = __uint_as_float(hi ? (a10[w] & 0xffff0000u) : (a10[w] << 16));
      const float v11 = __uint_as_float(hi ? (a11[w] & 0xffff0000u) : (a11[w] << 16));
      float s = __fmul_rn(v00, w00);
      s = __fadd_rn(s, __fmul_rn(v01, w01));
      s = __fadd_rn(s, __fmul_rn(v10, w10));
      s = __fadd_rn(s, __fmul_rn(v11, w11));
      acc[i] = __fadd_rn(acc[i], __fmul_rn(s, wgt));
    }
  }
  TO* op = out + row * ldo + c;
#pragma unroll
  for (int i = 0; i < 8; ++i) st1(op + i, acc[i]);
}

// ---------------------------------------------------------------------------
// row utilities
// ---------------------------------------------------------------------------
__global__ void rowcopy_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t rows,
                               int n, int64_t ld_in, int64_t ld_out, int act) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * n) return;
  const int64_t r = i / n;
  const int c = int(i - r * n);
  float v = in[r * ld_in + c];
  if (act == 1) v = 1.0f / (1.0f + expf(-v));
  out[r * ld_out + c] = v;
}

__global__ void broadcast_rows_kernel(const float* __restrict__ src, float* __restrict__ out,
                                      __nv_bfloat16* __restrict__ out_bf16, int64_t total,
                                      int64_t per_image) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float v = src[i % per_image];
  if (out) out[i] = v;
  if (out_bf16) out_bf16[i] = __float2bfloat16_rn(v);
}

__global__ void cast_pad_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                int64_t rows, int cols, int64_t ld_src, int64_t ld_dst,
                                int dst_cols, float scale) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * dst_cols) return;
  const int64_t r = i / dst_cols;
  const int c = int(i - r * dst_cols);
  dst[r * ld_dst + c] = __float2bfloat16_rn(c < cols ? scale * src[r * ld_src + c] : 0.f);
}

// x = hi + mid + lo (three bf16 terms carry the full fp32 significand).  The
// six retained partial products of (a_hi+a_mid+a_lo)(w_hi+w_mid+w_lo) are laid
// out as six K segments so that one bf16 GEMM with fp32 accumulation returns
// the fp32 product:   A side  [hi | hi  | mid | hi | lo | mid]
//                     W side  [hi | mid | hi  | lo | hi | mid]
__global__ void split3_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                              int64_t rows, int cols, int64_t ld_src, int64_t ld_dst, int kseg,
                              int w_side) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * kseg) return;
  const int64_t r = i / kseg;
  const int c = int(i - r * kseg);
  float x = c < cols ? src[r * ld_src + c] : 0.f;
  const __nv_bfloat16 hi = __float2bfloat16_rn(x);
  x -= __bfloat162float(hi);
  const __nv_bfloat16 mid = __float2bfloat16_rn(x);
  x -= __bfloat162float(mid);
  const __nv_bfloat16 lo = __float2bfloat16_rn(x);
  __nv_bfloat16* d = dst + r * ld_dst + c;
  if (w_side) {
    d[0] = hi; d[kseg] = mid; d[2 * kseg] = hi; d[3 * kseg] = lo; d[4 * kseg] = hi; d[5 * kseg] = mid;
  } else {
    d[0] = hi; d[kseg] = hi; d[2 * kseg] = mid; d[3 * kseg] = hi; d[4 * kseg] = lo; d[5 * kseg] = mid;
  }
}

}  // namespace
}  // namespace dod

extern "C" int32_t dod_mha_small(const dod_mha_small_args* a, dod_stream_t stream_) {
  using namespace dod;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->q && a->k && a->v && a->out, "dod_mha_small: null pointer");
  DOD_REQUIRE(a->batch > 0 && a->lq > 0 && a->lk > 0 && a->heads > 0 && a->head_dim > 0,
              "dod_mha_small: empty problem");
  DOD_REQUIRE(a->head_dim <= kMhaThreads, "dod_mha_small: head_dim must be <= %d", kMhaThreads);
  DOD_REQUIRE(a->batch <= 65535 && a->heads <= 65535, "dod_mha_small: batch/heads exceed grid limits");
  DOD_REQUIRE(a->dtype == DOD_BF16 || a->dtype == DOD_F32, "dod_mha_small: bad dtype");
  if (a->lq <= 128 && a->lk <= 128) {
    const int pitch = int(a->head_dim) | 1, lkp = int(a->lk) | 1;
    const bool vec = a->dtype == DOD_BF16 && a->head_dim % 8 == 0 && a->ldq % 8 == 0 && a->ldk % 8 == 0 &&
                     a->ldv % 8 == 0 && ((uintptr_t(a->q) | uintptr_t(a->k) | uintptr_t(a->v)) & 15) == 0;
    // vec: K / V are kept as bf16 in shared memory (pitch head_dim + 2 halves), see mha_tiny_kernel
    const size_t kv_words = vec ? 2 * ((size_t(a->lk) * (a->head_dim + 2) + 1) / 2) : size_t(2 * a->lk) * pitch;
    const size_t tiny = sizeof(float) * (size_t(a->lq) * pitch + kv_words + size_t(a->lq) * lkp);
    if (tiny <= 200 * 1024) {
      dim3 grid(unsigned(a->heads), unsigned(a->batch));
#define DOD_MHA_TINY(T, VEC)                                                                                    \
  {                                                                                                             \
    auto kern = mha_tiny_kernel<T, VEC>;                                                                        \
    DOD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));           \
    kern<<<grid, kMhaThreads, tiny, stream>>>((const T*)a->q, (const T*)a->k, (const T*)a->v, (T*)a->out,      \
                                              int(a->lq), int(a->lk), int(a->head_dim), a->ldq, a->ldk, a->ldv, \
                                              a->ldo, a->scale);                                                \
  }
      if (a->dtype == DOD_BF16) {
        if (vec) DOD_MHA_TINY(__nv_bfloat16, true) else DOD_MHA_TINY(__nv_bfloat16, false)
      } else {
        DOD_MHA_TINY(float, false)
      }
#undef DOD_MHA_TINY
      int rc = check_cuda(cudaGetLastError(), "mha_tiny_kernel launch");
      if (rc == 0) count_launch();
      return rc;
    }
  }
  const int lk_pad = int(a->lk) | 1;  // odd stride: conflict-free column writes
  const size_t smem = sizeof(float) * (size_t(kQT) * a->head_dim + size_t(kQT) * lk_pad + kQT);
  DOD_REQUIRE(smem <= 200 * 1024, "dod_mha_small: lk=%lld too long for the shared-memory score tile",
              (long long)a->lk);
  dim3 grid(unsigned((a->lq + kQT - 1) / kQT), unsigned(a->heads), unsigned(a->batch));
  if (a->dtype == DOD_BF16) {
    auto kern = mha_small_kernel<__nv_bfloat16>;
    DOD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    kern<<<grid, kMhaThreads, smem, stream>>>(
        (const __nv_bfloat16*)a->q, (const __nv_bfloat16*)a->k, (const __nv_bfloat16*)a->v,
        (__nv_bfloat16*)a->out, int(a->lq), int(a->lk), int(a->head_dim), a->ldq, a->ldk, a->ldv,
        a->ldo, a->scale, lk_pad);
  } else {
    auto kern = mha_small_kernel<float>;
    DOD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    kern<<<grid, kMhaThreads, smem, stream>>>((const float*)a->q, (const float*)a->k,
                                              (const float*)a->v, (float*)a->out, int(a->lq),
                                              int(a->lk), int(a->head_dim), a->ldq, a->ldk, a->ldv,
                                              a->ldo, a->scale, lk_pad);
  }
  int rc = check_cuda(cudaGetLastError(), "mha_small_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}

extern "C" int32_t dod_deform_sample(const dod_deform_sample_args* a, dod_stream_t stream_) {
  using namespace dod;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->value && a->ref && a->offs && a->logits && a->out,
              "dod_deform_sample: null pointer");
  DOD_REQUIRE(a->batch > 0 && a->queries > 0 && a->heads > 0 && a->points > 0 && a->head_dim > 0 &&
                  a->grid_h > 0 && a->grid_w > 0,
              "dod_deform_sample: bad shape");
  DOD_REQUIRE(a->batch * a->queries < (1ll << 31), "dod_deform_sample: too many rows");
  const int d_model = int(a->heads * a->head_dim);
  const int threads = d_model >= 1024 ? 1024 : ((d_model + 31) / 32) * 32;
  const unsigned grid = unsigned(a->batch * a->queries);
  {
    // vectorised kernel: bf16 values, 8-channel groups inside one head, 16-byte-aligned value rows
    const char* e_vec = getenv("DOD_DEFORM_VEC");  // 0: the thread-per-channel kernel (A/B, bit-equality test)
    const bool no_vec = e_vec != nullptr && e_vec[0] == '0';
    const int groups = d_model / 8;
    if (!no_vec && a->value_dtype == DOD_BF16 && a->head_dim % 8 == 0 && a->ldv % 8 == 0 &&
        (reinterpret_cast<uintptr_t>(a->value) & 15) == 0 && groups >= 1 && groups <= 256) {
      const int rows_per_cta = 256 / groups;
      const int64_t rows = a->batch * a->queries;
      const unsigned vgrid = unsigned((rows + rows_per_cta - 1) / rows_per_cta);
      const int vthreads = rows_per_cta * groups;
#define DOD_LAUNCH_DSV(TO)                                                                                  \
  deform_sample_vec_kernel<TO><<<vgrid, vthreads, 0, stream>>>(                                             \
      (const __nv_bfloat16*)a->value, a->ref, a->offs, a->logits, (TO*)a->out, rows, int(a->queries),       \
      int(a->heads), int(a->points), int(a->head_dim), int(a->grid_h), int(a->grid_w), a->ldv, a->ldref,    \
      a->ldoffs, a->ldlog, a->ldo, a->ref_is_logit)
      if (a->out_dtype == DOD_BF16) DOD_LAUNCH_DSV(__nv_bfloat16);
      else DOD_LAUNCH_DSV(float);
#undef DOD_LAUNCH_DSV
      int rc = check_cuda(cudaGetLastError(), "deform_sample_vec_kernel launch");
      if (rc == 0) count_launch();
      return rc;
    }
  }
#define DOD_LAUNCH_DS(TV, TO)                                                                      \
  deform_sample_kernel<TV, TO><<<grid, threads, 0, stream>>>(                                      \
      (const TV*)a->value, a->ref, a->offs, a->logits, (TO*)a->out, int(a->queries), int(a->heads), \
      int(a->points), int(a->head_dim), int(a->grid_h), int(a->grid_w), a->ldv, a->ldref,          \
      a->ldoffs, a->ldlog, a->ldo, a->ref_is_logit)
  if (a->value_dtype == DOD_BF16 && a->out_dtype == DOD_BF16) DOD_LAUNCH_DS(__nv_bfloat16, __nv_bfloat16);
  else if (a->value_dtype == DOD_F32 && a->out_dtype == DOD_F32) DOD_LAUNCH_DS(float, float);
  else if (a->value_dtype == DOD_BF16 && a->out_dtype == DOD_F32) DOD_LAUNCH_DS(__nv_bfloat16, float);
  else DOD_LAUNCH_DS(float, __nv_bfloat16);
#undef DOD_LAUNCH_DS
  int rc = check_cuda(cudaGetLastError(), "deform_sample_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}

extern "C" int32_t dod_rowcopy(const dod_rowcopy_args* a, dod_stream_t stream_) {
  using namespace dod;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->in && a->out, "dod_rowcopy: null pointer");
  DOD_REQUIRE(a->rows >= 0 && a->n > 0 && a->ld_in >= a->n && a->ld_out >= a->n, "dod_rowcopy: bad shape");
  if (a->rows == 0) return DOD_OK;
  const int64_t total = a->rows * a->n;
  rowcopy_kernel<<<unsigned((total + 255) / 256), 256, 0, stream>>>(a->in, a->out, a->rows, int(a->n),
                                                                   a->ld_in, a->ld_out, a->act);
  int rc = check_cuda(cudaGetLastError(), "rowcopy_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}

extern "C" int32_t dod_broadcast_rows(const dod_broadcast_rows_args* a, dod_stream_t stream_) {
  using namespace dod;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->src && (a->out || a->out_bf16), "dod_broadcast_rows: null pointer");
  DOD_REQUIRE(a->batch > 0 && a->rows > 0 && a->d > 0, "dod_broadcast_rows: bad shape");
  const int64_t per = a->rows * a->d, total = per * a->batch;
  broadcast_rows_kernel<<<unsigned((total + 255) / 256), 256, 0, stream>>>(
      a->src, a->out, reinterpret_cast<__nv_bfloat16*>(a->out_bf16), total, per);
  int rc = check_cuda(cudaGetLastError(), "broadcast_rows_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}

extern "C" int32_t dod_cast_pad_bf16(const dod_cast_pad_args* a, dod_stream_t stream_) {
  using namespace dod;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->src && a->dst, "dod_cast_pad_bf16: null pointer");
  DOD_REQUIRE(a->rows >= 0 && a->cols > 0 && a->dst_cols >= a->cols && a->ld_src >= a->cols &&
                  a->ld_dst >= a->dst_cols,
              "dod_cast_pad_bf16: bad shape");
  if (a->rows == 0) return DOD_OK;
  const int64_t total = a->rows * a->dst_cols;
  cast_pad_kernel<<<unsigned((total + 255) / 256), 256, 0, stream>>>(
      a->src, reinterpret_cast<__nv_bfloat16*>(a->dst), a->rows, int(a->cols), a->ld_src, a->ld_dst,
      int(a->dst_cols), a->scale);
  int rc = check_cuda(cudaGetLastError(), "cast_pad_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}

extern "C" int32_t dod_split3_bf16(const dod_split3_args* a, dod_stream_t stream_) {
  using namespace dod;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->src && a->dst, "dod_split3_bf16: null pointer");
  DOD_REQUIRE(a->rows >= 0 && a->cols > 0 && a->kseg >= a->cols && a->kseg % 8 == 0 &&
                  a->ld_src >= a->cols && a->ld_dst >= 6 * a->kseg,
              "dod_split3_bf16: bad shape (kseg must be >= cols and a multiple of 8)");
  if (a->rows == 0) return DOD_OK;
  const int64_t total = a->rows * a->kseg;
  split3_kernel<<<unsigned((total + 255) / 256), 256, 0, stream>>>(
      a->src, reinterpret_cast<__nv_bfloat16*>(a->dst), a->rows, int(a->cols), a->ld_src, a->ld_dst,
      int(a->kseg), a->w_side);
  int rc = check_cuda(cudaGetLastError(), "split3_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}
