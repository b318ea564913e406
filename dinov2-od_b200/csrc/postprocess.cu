// COCO-format post-processing of the detector outputs (reference utils.py:195-233): sigmoid,
// score > threshold, class 0 skipped, cxcywh -> [x, y, w, h], emitted in the reference's order
// (image, class ascending, query ascending).  The reference does this with a triple python loop and
// a .cpu().numpy() per (image, class); here one CTA per image does an ORDERED compaction (ballot +
// running offset), so validation throughput follows inference throughput.  HBM-bound, tiny.
#include "common.cuh"
#include "../../include/dod.h"

namespace dod {
void count_launch(int n = 1);
namespace {

__global__ void __launch_bounds__(256)
postprocess_kernel(const float* __restrict__ logits, const float* __restrict__ boxes,
                   float* __restrict__ out_score, float* __restrict__ out_box,
                   int32_t* __restrict__ out_class, int32_t* __restrict__ counts, int queries,
                   int classes, int capacity, float threshold) {
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ int warp_cnt[8];
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const int total = (classes - 1) * queries;  // element e -> class 1 + e / Q, query e % Q
  for (int e0 = 0; e0 < total; e0 += 256) {
    const int e = e0 + threadIdx.x;
    bool keep = false;
    float score = 0.f;
    int c = 0, q = 0;
    if (e < total) {
      c = 1 + e / queries;
      q = e - (c - 1) * queries;
      const float x = logits[(int64_t(b) * queries + q) * classes + c];
      score = 1.0f / (1.0f + expf(-x));
      keep = score > threshold;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += warp_cnt[w];
    off += __popc(m & ((1u << lane) - 1));
    if (keep && off < capacity) {
      const float4 bx = *reinterpret_cast<const float4*>(boxes + (int64_t(b) * queries + q) * 4);
      const float x1 = bx.x - 0.5f * bx.z, y1 = bx.y - 0.5f * bx.w;
      const float x2 = bx.x + 0.5f * bx.z, y2 = bx.y + 0.5f * bx.w;
      const int64_t o = int64_t(b) * capacity + off;
      out_score[o] = score;
      out_class[o] = c;
      *reinterpret_cast<float4*>(out_box + o * 4) = make_float4(x1, y1, x2 - x1, y2 - y1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int s = 0;
      for (int w = 0; w < 8; ++w) s += warp_cnt[w];
      base += s;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[b] = base < capacity ? base : capacity;
}

}  // namespace
}  // namespace dod

extern "C" int32_t dod_postprocess(const dod_postprocess_args* a, dod_stream_t stream_) {
  using namespace dod;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->logits && a->boxes && a->out_score && a->out_box && a->out_class && a->counts,
              "dod_postprocess: null pointer");
  DOD_REQUIRE(a->batch > 0 && a->queries > 0 && a->classes > 1 && a->capacity > 0, "dod_postprocess: bad shape");
  DOD_REQUIRE((uintptr_t(a->boxes) & 15) == 0 && (uintptr_t(a->out_box) & 15) == 0,
              "dod_postprocess: boxes must be 16-byte aligned");
  postprocess_kernel<<<unsigned(a->batch), 256, 0, stream>>>(a->logits, a->boxes, a->out_score, a->out_box,
                                                            a->out_class, a->counts, int(a->queries),
                                                            int(a->classes), int(a->capacity), a->threshold);
  int rc = check_cuda(cudaGetLastError(), "postprocess_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}
