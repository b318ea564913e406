// Hungarian matcher on the GPU (reference dino_detector/matching.py:42-122).
//
//  dod_match_cost  batched cost matrix  C[b, q, j] = (wc*cls + wb*L1) + wg*(-GIoU)
//                  fp32, evaluated in torch's order (separately rounded mul/add,
//                  no FMA contraction) so the matrix tracks the reference to ulps.
//                  HBM/L2-bound: ~15 MB per 256-image batch, launch-latency dominated.
//  dod_lsap_jv     one warp per image; Crouse / Jonker-Volgenant shortest
//                  augmenting path in float64 with scipy's exact scan order
//                  (reverse-filled `remaining`, swap-remove) and tie rule, so
//                  assignments are bit-identical to scipy.optimize.linear_sum_assignment
//                  (call site matching.py:105).  Latency-bound by construction
//                  (<= 50 sequential augmentations x <= 100 pops per image).
#include <math_constants.h>

#include "common.cuh"
#include "../../include/dod.h"

namespace dod {
void count_launch(int n = 1);
namespace {

// ---------------------------------------------------------------------------
// cost matrix
// ---------------------------------------------------------------------------
__device__ __forceinline__ float pow_gamma(float x, float gamma) {
  // torch.pow(x, 2.0) is x*x; other exponents go through powf
  return gamma == 2.0f ? __fmul_rn(x, x) : powf(x, gamma);
}

__global__ void __launch_bounds__(256)
match_cost_kernel(const float* __restrict__ logits, const float* __restrict__ boxes,
                  const int64_t* __restrict__ tgt_labels, const float* __restrict__ tgt_boxes,
                  const int32_t* __restrict__ tgt_offsets, float* __restrict__ cost, int queries,
                  int classes, int max_t, float wc, float wb, float wg, float alpha, float gamma,
                  int use_image0_rows) {
  const int b = blockIdx.y;
  const int t0 = tgt_offsets[b];
  const int n = tgt_offsets[b + 1] - t0;
  const int src_b = use_image0_rows ? 0 : b;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < queries * n;
       idx += gridDim.x * blockDim.x) {
    const int q = idx / n, j = idx - q * n;
    const int64_t label = tgt_labels[t0 + j];
    const float4 tb = *reinterpret_cast<const float4*>(tgt_boxes + int64_t(t0 + j) * 4);
    const float4 pb = *reinterpret_cast<const float4*>(boxes + (int64_t(src_b) * queries + q) * 4);
    const float x = logits[(int64_t(src_b) * queries + q) * classes + label];
    // matching.py:63 sigmoid, :82-86 focal-style class cost
    const float p = 1.0f / (1.0f + expf(-x));
    const float neg = __fmul_rn(__fmul_rn(1.0f - alpha, pow_gamma(p, gamma)),
                                -logf(__fadd_rn(__fsub_rn(1.0f, p), 1e-8f)));
    const float pos = __fmul_rn(__fmul_rn(alpha, pow_gamma(__fsub_rn(1.0f, p), gamma)),
                                -logf(__fadd_rn(p, 1e-8f)));
    const float cc = __fsub_rn(pos, neg);
    // matching.py:89 cdist(p=1) on cxcywh
    float cb = fabsf(__fsub_rn(pb.x, tb.x));
    cb = __fadd_rn(cb, fabsf(__fsub_rn(pb.y, tb.y)));
    cb = __fadd_rn(cb, fabsf(__fsub_rn(pb.z, tb.z)));
    cb = __fadd_rn(cb, fabsf(__fsub_rn(pb.w, tb.w)));
    // utils.py:73-92 cxcywh -> xyxy, :124-164 GIoU (no eps, like the reference)
    const float px0 = __fsub_rn(pb.x, __fmul_rn(0.5f, pb.z)), py0 = __fsub_rn(pb.y, __fmul_rn(0.5f, pb.w));
    const float px1 = __fadd_rn(pb.x, __fmul_rn(0.5f, pb.z)), py1 = __fadd_rn(pb.y, __fmul_rn(0.5f, pb.w));
    const float tx0 = __fsub_rn(tb.x, __fmul_rn(0.5f, tb.z)), ty0 = __fsub_rn(tb.y, __fmul_rn(0.5f, tb.w));
    const float tx1 = __fadd_rn(tb.x, __fmul_rn(0.5f, tb.z)), ty1 = __fadd_rn(tb.y, __fmul_rn(0.5f, tb.w));
    const float area1 = __fmul_rn(__fsub_rn(px1, px0), __fsub_rn(py1, py0));
    const float area2 = __fmul_rn(__fsub_rn(tx1, tx0), __fsub_rn(ty1, ty0));
    const float iw = fmaxf(__fsub_rn(fminf(px1, tx1), fmaxf(px0, tx0)), 0.f);
    const float ih = fmaxf(__fsub_rn(fminf(py1, ty1), fmaxf(py0, ty0)), 0.f);
    const float inter = __fmul_rn(iw, ih);
    const float uni = __fsub_rn(__fadd_rn(area1, area2), inter);
    const float iou = __fdiv_rn(inter, uni);
    const float ew = fmaxf(__fsub_rn(fmaxf(px1, tx1), fminf(px0, tx0)), 0.f);
    const float eh = fmaxf(__fsub_rn(fmaxf(py1, ty1), fminf(py0, ty0)), 0.f);
    const float earea = __fmul_rn(ew, eh);
    const float giou = __fsub_rn(iou, __fdiv_rn(__fsub_rn(earea, uni), earea));
    // matching.py:98   C = wc*cc + wb*cb + wg*(-giou), left to right
    const float c = __fadd_rn(__fadd_rn(__fmul_rn(wc, cc), __fmul_rn(wb, cb)), __fmul_rn(wg, -giou));
    cost[(int64_t(b) * queries + q) * max_t + j] = c;
  }
}

// ---------------------------------------------------------------------------
// LSAP: one warp per image
// ---------------------------------------------------------------------------
struct Best {
  double val;
  int it;      // position in `remaining`
  int unassigned;
};

// total order equivalent to scipy's sequential scan
//   if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; index = it; }
// lower value wins; on equal value an unassigned column wins; among unassigned ties
// the LAST scanned wins, among assigned ties the FIRST scanned wins.
__device__ __forceinline__ bool better(const Best& a, const Best& b) {
  if (a.it < 0) return false;
  if (b.it < 0) return true;
  if (a.val < b.val) return true;
  if (a.val > b.val) return false;
  if (a.unassigned != b.unassigned) return a.unassigned > b.unassigned;
  return a.unassigned ? (a.it > b.it) : (a.it < b.it);
}

__global__ void __launch_bounds__(32)
lsap_kernel(const float* __restrict__ cost, const int32_t* __restrict__ tgt_offsets,
            int32_t* __restrict__ out_q, int32_t* __restrict__ out_t, int32_t* __restrict__ status,
            int queries, int max_t, int max_k, int cost_in_smem) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const int lane = threadIdx.x;
  const int n = tgt_offsets[b + 1] - tgt_offsets[b];
  const float* cb = cost + int64_t(b) * queries * max_t;
  int32_t* oq = out_q + int64_t(b) * max_k;
  int32_t* ot = out_t + int64_t(b) * max_k;

  if (n <= 0 || queries <= 0) {
    if (lane == 0) status[b] = 0;
    return;
  }
  // scipy transposes when nc < nr so that rows are the short side
  const bool transpose = n < queries;
  const int nr = transpose ? n : queries;
  const int nc = transpose ? queries : n;

  double* u = reinterpret_cast<double*>(smem_raw);  // [nr]
  double* v = u + nr;                                // [nc]
  double* spc = v + nc;                              // [nc]
  int* path = reinterpret_cast<int*>(spc + nc);      // [nc]
  int* col4row = path + nc;                          // [nr]
  int* row4col = col4row + nr;                       // [nc]
  int* remaining = row4col + nc;                     // [nc]
  unsigned char* SR = reinterpret_cast<unsigned char*>(remaining + nc);  // [nr]
  unsigned char* SC = SR + nr;                                           // [nc]
  float* cs = reinterpret_cast<float*>(
      (reinterpret_cast<uintptr_t>(SC + nc) + 15) & ~uintptr_t(15));  // [nr][nc] if cost_in_smem

  // validity scan (scipy: NaN or -inf entries -> ValueError) + optional staging
  int bad = 0;
  for (int e = lane; e < nr * nc; e += 32) {
    const int i = e / nc, j = e - i * nc;
    const float c = transpose ? cb[int64_t(j) * max_t + i] : cb[int64_t(i) * max_t + j];
    if (c != c || c == -CUDART_INF_F) bad = 1;
    if (cost_in_smem) cs[e] = c;
  }
  bad = __any_sync(0xffffffffu, bad);
  for (int i = lane; i < nr; i += 32) { u[i] = 0.0; col4row[i] = -1; }
  for (int j = lane; j < nc; j += 32) { v[j] = 0.0; row4col[j] = -1; path[j] = -1; }
  __syncwarp();
  if (bad) {
    if (lane == 0) status[b] = 1;
    return;
  }

  int infeasible = 0;
  for (int cur = 0; cur < nr && !infeasible; ++cur) {
    for (int i = lane; i < nr; i += 32) SR[i] = 0;
    for (int j = lane; j < nc; j += 32) {
      SC[j] = 0;
      spc[j] = CUDART_INF;
      remaining[j] = nc - j - 1;
    }
    __syncwarp();
    int nrem = nc;
    int sink = -1;
    double min_val = 0.0;
    int i = cur;
    while (sink == -1) {
      if (lane == 0) SR[i] = 1;
      const double ui = u[i];
      Best best;
      best.val = CUDART_INF;
      best.it = -1;
      best.unassigned = 0;
      for (int it = lane; it < nrem; it += 32) {
        const int j = remaining[it];
        const float cf = cost_in_smem ? cs[i * nc + j]
                                      : (transpose ? cb[int64_t(j) * max_t + i] : cb[int64_t(i) * max_t + j]);
        const double r = ((min_val + double(cf)) - ui) - v[j];
        double s = spc[j];
        if (r < s) {
          path[j] = i;
          spc[j] = r;
          s = r;
        }
        Best cand;
        cand.val = s;
        cand.it = it;
        cand.unassigned = row4col[j] == -1;
        // scipy only accepts a candidate that is < +inf or (== lowest and unassigned);
        // with lowest starting at +inf an all-inf row leaves index == -1 unless an
        // unassigned column ties at +inf, which `minVal == inf` then rejects anyway.
        if (better(cand, best)) best = cand;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        Best other;
        other.val = __shfl_xor_sync(0xffffffffu, best.val, o);
        other.it = __shfl_xor_sync(0xffffffffu, best.it, o);
        other.unassigned = __shfl_xor_sync(0xffffffffu, best.unassigned, o);
        if (better(other, best)) best = other;
      }
      min_val = best.val;
      if (best.it < 0 || min_val == CUDART_INF) {
        infeasible = 1;
        break;
      }
      __syncwarp();
      const int j = remaining[best.it];
      const int r4c = row4col[j];
      __syncwarp();
      if (r4c == -1) sink = j; else i = r4c;
      if (lane == 0) {
        SC[j] = 1;
        remaining[best.it] = remaining[nrem - 1];
      }
      --nrem;
      __syncwarp();
    }
    if (infeasible) break;
    // dual updates
    if (lane == 0) u[cur] += min_val;
    for (int i2 = lane; i2 < nr; i2 += 32)
      if (SR[i2] && i2 != cur) u[i2] += min_val - spc[col4row[i2]];
    for (int j2 = lane; j2 < nc; j2 += 32)
      if (SC[j2]) v[j2] -= min_val - spc[j2];
    __syncwarp();
    // augment along the alternating path (sequential)
    if (lane == 0) {
      int j = sink;
      while (true) {
        const int i2 = path[j];
        row4col[j] = i2;
        const int tmp = col4row[i2];
        col4row[i2] = j;
        j = tmp;
        if (i2 == cur) break;
      }
    }
    __syncwarp();
  }

  if (infeasible) {
    if (lane == 0) status[b] = 1;
    return;
  }
  // output sorted by query index (scipy: rows ascending; argsort(col4row) if transposed)
  if (!transpose) {
    for (int q = lane; q < nr; q += 32) { oq[q] = q; ot[q] = col4row[q]; }
  } else {
    int base = 0;
    for (int q0 = 0; q0 < nc; q0 += 32) {
      const int q = q0 + lane;
      const int r = q < nc ? row4col[q] : -1;
      const unsigned m = __ballot_sync(0xffffffffu, r >= 0);
      if (r >= 0) {
        const int pos = base + __popc(m & ((1u << lane) - 1));
        oq[pos] = q;
        ot[pos] = r;
      }
      base += __popc(m);
    }
  }
  if (lane == 0) status[b] = 0;
}

size_t lsap_state_bytes(int nr_max, int nc_max) {
  size_t s = sizeof(double) * (size_t(nr_max) + 2 * size_t(nc_max));
  s += sizeof(int) * (3 * size_t(nc_max) + size_t(nr_max));
  s += size_t(nr_max) + size_t(nc_max);
  return s + 16;
}

}  // namespace
}  // namespace dod

extern "C" int32_t dod_match_cost(const dod_match_cost_args* a, dod_stream_t stream_) {
  using namespace dod;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->logits && a->boxes && a->tgt_offsets && a->cost, "dod_match_cost: null pointer");
  DOD_REQUIRE(a->batch > 0 && a->batch <= 65535 && a->queries > 0 && a->classes > 0 && a->max_t >= 0,
              "dod_match_cost: bad shape");
  if (a->max_t == 0) return DOD_OK;
  DOD_REQUIRE(a->tgt_labels && a->tgt_boxes, "dod_match_cost: null targets");
  DOD_REQUIRE((uintptr_t(a->boxes) & 15) == 0 && (uintptr_t(a->tgt_boxes) & 15) == 0,
              "dod_match_cost: boxes must be 16-byte aligned");
  const int64_t work = a->queries * a->max_t;
  unsigned gx = unsigned((work + 255) / 256);
  if (gx > 64) gx = 64;
  match_cost_kernel<<<dim3(gx, unsigned(a->batch)), 256, 0, stream>>>(
      a->logits, a->boxes, a->tgt_labels, a->tgt_boxes, a->tgt_offsets, a->cost, int(a->queries),
      int(a->classes), int(a->max_t), a->w_class, a->w_bbox, a->w_giou, a->alpha, a->gamma,
      a->use_image0_rows);
  int rc = check_cuda(cudaGetLastError(), "match_cost_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}

extern "C" int32_t dod_lsap_jv(const dod_lsap_args* a, dod_stream_t stream_) {
  using namespace dod;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->tgt_offsets && a->status, "dod_lsap_jv: null pointer");
  DOD_REQUIRE(a->batch > 0 && a->queries > 0 && a->max_t >= 0 && a->max_k >= 0, "dod_lsap_jv: bad shape");
  DOD_REQUIRE(a->max_t == 0 || (a->cost && a->out_q && a->out_t), "dod_lsap_jv: null pointer");
  const int64_t kmax = a->queries < a->max_t ? a->queries : a->max_t;
  DOD_REQUIRE(a->max_k >= kmax, "dod_lsap_jv: max_k must be >= min(queries, max_t)");
  DOD_REQUIRE(a->queries <= 4096 && a->max_t <= 4096, "dod_lsap_jv: problem too large");
  // worst-case state: rows = short side, cols = long side
  const int longs = int(a->queries > a->max_t ? a->queries : a->max_t);
  const int shorts = int(kmax > 0 ? kmax : 1);
  size_t state = lsap_state_bytes(shorts, longs);
  size_t with_cost = state + sizeof(float) * size_t(shorts) * longs;
  const int cost_in_smem = with_cost <= 160 * 1024;
  const size_t smem = cost_in_smem ? with_cost : state;
  DOD_REQUIRE(smem <= 200 * 1024, "dod_lsap_jv: problem too large for shared memory");
  DOD_CUDA_OK(cudaFuncSetAttribute(lsap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  lsap_kernel<<<unsigned(a->batch), 32, smem, stream>>>(a->cost, a->tgt_offsets, a->out_q, a->out_t,
                                                       a->status, int(a->queries), int(a->max_t),
                                                       int(a->max_k), cost_in_smem);
  int rc = check_cuda(cudaGetLastError(), "lsap_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}
