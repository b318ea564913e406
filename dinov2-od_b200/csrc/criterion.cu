// Fused SetCriterion forward + backward (reference dino_detector/losses.py:96-241).
//
// One launch computes the three reference losses AND their gradients w.r.t. pred_logits / pred_boxes
// from the device-resident assignment produced by dod_lsap_jv -- no host sync, no one-hot tensor:
//   loss_ce   = sum_{b,q,c} alpha_t (1 - p_t)^gamma BCEwithLogits(x, onehot) / num_boxes   (:96-147)
//   loss_bbox = sum_matched |src - tgt|_1 / num_boxes                                        (:149-175)
//   loss_giou = sum_matched (1 - GIoU(src, tgt)) / num_boxes                                 (:176-187)
// Matched pair (b, i, j): prediction i of IMAGE b against target j of image b (the indices come from
// the matcher, which under reference_compat ranked image 0's predictions, matching.py:102).
// num_boxes is read from device memory (already all-reduced by the caller, losses.py:225-230).
// HBM-bound and tiny: B*Q*C logits read once, gradients written once.
#include "common.cuh"
#include "../../include/dod.h"

namespace dod {
void count_launch(int n = 1);
namespace {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// grid (ceil(Q*C / 256), B): focal classification loss + gradient
__global__ void __launch_bounds__(256)
criterion_labels_kernel(const float* __restrict__ logits, const int32_t* __restrict__ tclass,
                        float* __restrict__ dlogits, float* __restrict__ losses,
                        const float* __restrict__ num_boxes, int queries, int classes, float alpha,
                        float gamma, float w_ce) {
  const int b = blockIdx.y;
  const int e = blockIdx.x * 256 + threadIdx.x;
  float loss = 0.f;
  if (e < queries * classes) {
    const int q = e / classes, c = e - q * classes;
    const int64_t idx = (int64_t(b) * queries + q) * classes + c;
    const float x = logits[idx];
    const float t = tclass[b * queries + q] == c ? 1.0f : 0.0f;
    const float p = 1.0f / (1.0f + expf(-x));
    const float pt = p * t + (1.0f - p) * (1.0f - t);
    const float om = 1.0f - pt;
    const float fw = gamma == 2.0f ? om * om : powf(om, gamma);
    const float aw = alpha * t + (1.0f - alpha) * (1.0f - t);
    const float bce = fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
    loss = aw * fw * bce;
    // d/dx: d pt/dx = p (1-p) (2t - 1);  d bce/dx = p - t
    const float dpt = p * (1.0f - p) * (2.0f * t - 1.0f);
    const float dfw = gamma == 2.0f ? -2.0f * om * dpt : -gamma * powf(om, gamma - 1.0f) * dpt;
    dlogits[idx] = w_ce * aw * (dfw * bce + fw * (p - t)) / num_boxes[0];
  }
  loss = warp_sum_f(loss);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    atomicAdd(&losses[0], w_ce * s / num_boxes[0]);
  }
}

// target class per (b, q): num_classes ("no object") unless matched
__global__ void criterion_tclass_kernel(int32_t* __restrict__ tclass, const int32_t* __restrict__ out_q,
                                        const int32_t* __restrict__ out_t,
                                        const int32_t* __restrict__ tgt_offsets,
                                        const int64_t* __restrict__ tgt_labels, int batch, int queries,
                                        int classes, int max_k) {
  const int b = blockIdx.x;
  for (int q = threadIdx.x; q < queries; q += blockDim.x) tclass[b * queries + q] = classes;
  __syncthreads();
  const int n = tgt_offsets[b + 1] - tgt_offsets[b];
  const int k = n < queries ? n : queries;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const int q = out_q[b * max_k + i], j = out_t[b * max_k + i];
    if (q >= 0 && q < queries && j >= 0 && j < n) tclass[b * queries + q] = int(tgt_labels[tgt_offsets[b] + j]);
  }
}

// one thread per matched pair: L1 + GIoU losses and gradients w.r.t. the predicted cxcywh box
__global__ void __launch_bounds__(128)
criterion_boxes_kernel(const float* __restrict__ boxes, const float* __restrict__ tgt_boxes,
                       const int32_t* __restrict__ out_q, const int32_t* __restrict__ out_t,
                       const int32_t* __restrict__ tgt_offsets, float* __restrict__ dboxes,
                       float* __restrict__ dboxes_giou, float* __restrict__ losses, const float* __restrict__ num_boxes, int queries,
                       int max_k, float w_bbox, float w_giou) {
  const int b = blockIdx.x;
  const int n = tgt_offsets[b + 1] - tgt_offsets[b];
  const int k = n < queries ? n : queries;
  const float inv_nb = 1.0f / num_boxes[0];
  float l1 = 0.f, lg = 0.f;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const int q = out_q[b * max_k + i], j = out_t[b * max_k + i];
    if (q < 0 || q >= queries || j < 0 || j >= n) continue;  // image rejected by the solver (status 1)
    const float4 s = *reinterpret_cast<const float4*>(boxes + (int64_t(b) * queries + q) * 4);
    const float4 t = *reinterpret_cast<const float4*>(tgt_boxes + int64_t(tgt_offsets[b] + j) * 4);
    // L1 (losses.py:171-172)
    l1 += fabsf(s.x - t.x) + fabsf(s.y - t.y) + fabsf(s.z - t.z) + fabsf(s.w - t.w);
    float g[4];
    g[0] = (s.x > t.x) - (s.x < t.x);
    g[1] = (s.y > t.y) - (s.y < t.y);
    g[2] = (s.z > t.z) - (s.z < t.z);
    g[3] = (s.w > t.w) - (s.w < t.w);
    // GIoU on xyxy (utils.py:73-92, 124-164)
    const float ax0 = s.x - 0.5f * s.z, ay0 = s.y - 0.5f * s.w, ax1 = s.x + 0.5f * s.z, ay1 = s.y + 0.5f * s.w;
    const float gx0 = t.x - 0.5f * t.z, gy0 = t.y - 0.5f * t.w, gx1 = t.x + 0.5f * t.z, gy1 = t.y + 0.5f * t.w;
    const float aw_ = ax1 - ax0, ah_ = ay1 - ay0;
    const float area_a = aw_ * ah_, area_g = (gx1 - gx0) * (gy1 - gy0);
    const float iw = fmaxf(fminf(ax1, gx1) - fmaxf(ax0, gx0), 0.f);
    const float ih = fmaxf(fminf(ay1, gy1) - fmaxf(ay0, gy0), 0.f);
    const float inter = iw * ih;
    const float uni = area_a + area_g - inter;
    const float ew = fmaxf(fmaxf(ax1, gx1) - fminf(ax0, gx0), 0.f);
    const float eh = fmaxf(fmaxf(ay1, gy1) - fminf(ay0, gy0), 0.f);
    const float earea = ew * eh;
    const float giou = inter / uni - (earea - uni) / earea;
    lg += 1.0f - giou;
    // gradients of inter / area_a / earea w.r.t. (ax0, ay0, ax1, ay1)
    const float di[4] = {(iw > 0.f && ax0 > gx0) ? -ih : 0.f, (ih > 0.f && ay0 > gy0) ? -iw : 0.f,
                         (iw > 0.f && ax1 < gx1) ? ih : 0.f, (ih > 0.f && ay1 < gy1) ? iw : 0.f};
    const float da[4] = {-ah_, -aw_, ah_, aw_};
    const float de[4] = {(ew > 0.f && ax0 < gx0) ? -eh : 0.f, (eh > 0.f && ay0 < gy0) ? -ew : 0.f,
                         (ew > 0.f && ax1 > gx1) ? eh : 0.f, (eh > 0.f && ay1 > gy1) ? ew : 0.f};
    float dg[4];  // d giou / d xyxy
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float du = da[c] - di[c];
      dg[c] = (di[c] * uni - inter * du) / (uni * uni) + (du * earea - uni * de[c]) / (earea * earea);
    }
    // xyxy -> cxcywh: d/dcx = d/dx0 + d/dx1, d/dw = (d/dx1 - d/dx0) / 2
    const float dgc[4] = {dg[0] + dg[2], dg[1] + dg[3], 0.5f * (dg[2] - dg[0]), 0.5f * (dg[3] - dg[1])};
    const float sb = w_bbox * inv_nb, sgi = -w_giou * inv_nb;
    *reinterpret_cast<float4*>(dboxes + (int64_t(b) * queries + q) * 4) =
        make_float4(sb * g[0], sb * g[1], sb * g[2], sb * g[3]);
    *reinterpret_cast<float4*>(dboxes_giou + (int64_t(b) * queries + q) * 4) =
        make_float4(sgi * dgc[0], sgi * dgc[1], sgi * dgc[2], sgi * dgc[3]);
  }
  l1 = warp_sum_f(l1);
  lg = warp_sum_f(lg);
  if ((threadIdx.x & 31) == 0 && (l1 != 0.f || lg != 0.f)) {
    atomicAdd(&losses[1], w_bbox * l1 * inv_nb);
    atomicAdd(&losses[2], w_giou * lg * inv_nb);
  }
}

}  // namespace
}  // namespace dod

extern "C" int32_t dod_criterion(const dod_criterion_args* a, dod_stream_t stream_) {
  using namespace dod;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->logits && a->boxes && a->tgt_offsets && a->out_q && a->out_t && a->num_boxes &&
                  a->tclass && a->losses && a->dlogits && a->dboxes && a->dboxes_giou,
              "dod_criterion: null pointer");
  DOD_REQUIRE(a->batch > 0 && a->batch <= 65535 && a->queries > 0 && a->classes > 0 && a->max_k >= 1,
              "dod_criterion: bad shape");
  DOD_REQUIRE((uintptr_t(a->boxes) & 15) == 0 && (uintptr_t(a->dboxes) & 15) == 0 &&
                  (!a->tgt_boxes || (uintptr_t(a->tgt_boxes) & 15) == 0),
              "dod_criterion: boxes must be 16-byte aligned");
  // losses[3], dboxes and dboxes_giou must be zeroed by the caller (unmatched queries keep zero)
  criterion_tclass_kernel<<<unsigned(a->batch), 128, 0, stream>>>(a->tclass, a->out_q, a->out_t,
                                                                  a->tgt_offsets, a->tgt_labels,
                                                                  int(a->batch), int(a->queries),
                                                                  int(a->classes), int(a->max_k));
  DOD_CUDA_OK(cudaGetLastError());
  const int64_t per_image = a->queries * a->classes;
  criterion_labels_kernel<<<dim3(unsigned((per_image + 255) / 256), unsigned(a->batch)), 256, 0, stream>>>(
      a->logits, a->tclass, a->dlogits, a->losses, a->num_boxes, int(a->queries), int(a->classes),
      a->focal_alpha, a->focal_gamma, a->w_ce);
  DOD_CUDA_OK(cudaGetLastError());
  if (a->tgt_boxes) {
    criterion_boxes_kernel<<<unsigned(a->batch), 128, 0, stream>>>(
        a->boxes, a->tgt_boxes, a->out_q, a->out_t, a->tgt_offsets, a->dboxes, a->dboxes_giou, a->losses,
        a->num_boxes,
        int(a->queries), int(a->max_k), a->w_bbox, a->w_giou);
    DOD_CUDA_OK(cudaGetLastError());
  }
  count_launch(a->tgt_boxes ? 3 : 2);
  return DOD_OK;
}
