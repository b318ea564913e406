// Patch-embedding front end (HBM-bound helpers around the patch GEMM).
//
//  dod_patchify14          im2col of 14x14 / stride-14 patches, fp32 -> bf16, plus
//                          the CLS token rows of the residual stream.
//  dod_pos_resize_bicubic  bicubic resize of the learned position grid (only
//                          needed when the image is not 518x518); computed once
//                          per (H, W) by the host and cached.
//
// Replaces Conv2d's implicit im2col (transformers modeling_dinov2.py:139,148),
// the cls/pos assembly (modeling_dinov2.py:108-112) and
// F.interpolate(mode="bicubic", align_corners=False) (modeling_dinov2.py:84-89).
#include "common.cuh"
#include "../../include/dod.h"

namespace dod {
void count_launch(int n = 1);
namespace {

constexpr int kP = 14;
constexpr int kK = 3 * kP * kP;  // 588

// One thread per 16-byte chunk of an im2col row: eight consecutive k = c*196 + i*14 + j are gathered from the
// image (they sit in at most two image rows, 4 / 1 byte apart) and written with ONE 16-byte store.  A block owns
// one patch row of one image (3 x 14 image rows, ~87 KB as fp32: the scattered 4-byte reads are served from
// L1 / L2 after the first touch of a sector, each input byte leaves DRAM once), so the HBM traffic is the
// algorithmic read + write.  The first version wrote 2-byte elements from a row-major walk over the pixels
// (coalesced reads, 28-byte scattered write pieces): 125 us for 64 images at 518x518 = 2.5 TB/s.
// U8: raw uint8 NHWC images, ToTensor's /255 (reference train.py:584-587, dataset.py:55-66) fused in, so the
// host->device copy is 1 byte per sample instead of 4.
template <bool U8>
__global__ void __launch_bounds__(256)
patchify_kernel(const void* __restrict__ pix_, __nv_bfloat16* __restrict__ out, int H, int W, int gh, int gw,
                int kpad) {
  const int py = blockIdx.x, b = blockIdx.y;
  const int chunks = kpad >> 3;                       // 16-byte chunks per im2col row (74 for kpad = 592)
  const int64_t row0 = (int64_t(b) * gh + py) * gw;
  for (int t = threadIdx.x; t < gw * chunks; t += blockDim.x) {
    const int px = t / chunks, q = t - px * chunks;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = q * 8 + e;
      float x = 0.0f;                                 // K padding: zeros, so that 0-weights never meet garbage
      if (k < kK) {
        const int c = k / (kP * kP), r = k - c * (kP * kP);
        const int i = r / kP, j = r - i * kP;
        const int y = py * kP + i, xx = px * kP + j;
        if constexpr (U8)
          x = float(__ldg(reinterpret_cast<const uint8_t*>(pix_) + ((int64_t(b) * H + y) * W + xx) * 3 + c)) / 255.0f;
        else
          x = __ldg(reinterpret_cast<const float*>(pix_) + ((int64_t(b) * 3 + c) * H + y) * W + xx);
      }
      v[e] = x;
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]);
    o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]);
    o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + (row0 + px) * kpad + q * 8) = o;
  }
}

__global__ void cls_rows_kernel(const float* __restrict__ cls, const float* __restrict__ pos,
                                float* __restrict__ tokens, int64_t n_tokens, int d) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < d; c += blockDim.x)
    tokens[int64_t(b) * n_tokens * d + c] = cls[c] + pos[c];
}

__device__ __forceinline__ void cubic_coeffs(float t, float (&w)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.0f, x3 = 2.0f - t, x2 = 1.0f - t;
  w[0] = ((A * x0 - 5.0f * A) * x0 + 8.0f * A) * x0 - 4.0f * A;
  w[1] = ((A + 2.0f) * t - (A + 3.0f)) * t * t + 1.0f;
  w[2] = ((A + 2.0f) * x2 - (A + 3.0f)) * x2 * x2 + 1.0f;
  w[3] = ((A * x3 - 5.0f * A) * x3 + 8.0f * A) * x3 - 4.0f * A;
}

// one block per output position, threads over channels (channel-last storage)
__global__ void pos_resize_kernel(const float* __restrict__ src, float* __restrict__ dst, int g0,
                                  int gh, int gw, int d) {
  const int o = blockIdx.x;  // 0 = CLS row, 1.. = grid
  if (o == 0) {
    for (int c = threadIdx.x; c < d; c += blockDim.x) dst[c] = src[c];
    return;
  }
  const int oy = (o - 1) / gw, ox = (o - 1) % gw;
  const float sy = float(g0) / float(gh), sx = float(g0) / float(gw);
  const float fy = sy * (oy + 0.5f) - 0.5f, fx = sx * (ox + 0.5f) - 0.5f;
  const int iy = int(floorf(fy)), ix = int(floorf(fx));
  float wy[4], wx[4];
  cubic_coeffs(fy - iy, wy);
  cubic_coeffs(fx - ix, wx);
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float acc = 0.f;
    float rows[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(iy - 1 + a, 0), g0 - 1);
      float r = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int xx = min(max(ix - 1 + e, 0), g0 - 1);
        r += wx[e] * src[(int64_t(1 + yy * g0 + xx)) * d + c];
      }
      rows[a] = r;
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) acc += wy[a] * rows[a];
    dst[int64_t(o) * d + c] = acc;
  }
}

}  // namespace
}  // namespace dod

extern "C" int32_t dod_patchify14(const dod_patchify_args* a, dod_stream_t stream_) {
  using namespace dod;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->pixels && a->patches, "dod_patchify14: null pointer");
  const int gh = int(a->height / kP), gw = int(a->width / kP);
  DOD_REQUIRE(a->batch > 0 && gh > 0 && gw > 0, "dod_patchify14: image smaller than one patch");
  DOD_REQUIRE(a->batch <= 65535, "dod_patchify14: batch too large");
  DOD_REQUIRE(a->kpad >= kK && a->kpad % 8 == 0, "dod_patchify14: kpad must be >= 588 and a multiple of 8");
  DOD_REQUIRE(a->pixel_format == 0 || a->pixel_format == 1, "dod_patchify14: bad pixel_format");
  DOD_REQUIRE((uintptr_t(a->patches) & 15) == 0, "dod_patchify14: patches must be 16-byte aligned");
  if (a->pixel_format == 1)
    patchify_kernel<true><<<dim3(gh, unsigned(a->batch)), 256, 0, stream>>>(
        a->pixels, reinterpret_cast<__nv_bfloat16*>(a->patches), int(a->height), int(a->width), gh, gw,
        int(a->kpad));
  else
    patchify_kernel<false><<<dim3(gh, unsigned(a->batch)), 256, 0, stream>>>(
        a->pixels, reinterpret_cast<__nv_bfloat16*>(a->patches), int(a->height), int(a->width), gh, gw,
        int(a->kpad));
  int rc = check_cuda(cudaGetLastError(), "patchify_kernel launch");
  if (rc) return rc;
  count_launch();
  if (a->tokens) {
    DOD_REQUIRE(a->cls && a->pos && a->d > 0, "dod_patchify14: cls/pos required with tokens");
    cls_rows_kernel<<<unsigned(a->batch), 256, 0, stream>>>(a->cls, a->pos, a->tokens,
                                                            int64_t(gh) * gw + 1, int(a->d));
    rc = check_cuda(cudaGetLastError(), "cls_rows_kernel launch");
    if (rc == 0) count_launch();
  }
  return rc;
}

extern "C" int32_t dod_pos_resize_bicubic(const dod_pos_resize_args* a, dod_stream_t stream_) {
  using namespace dod;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DOD_REQUIRE(a && a->src && a->dst, "dod_pos_resize_bicubic: null pointer");
  DOD_REQUIRE(a->g0 > 0 && a->gh > 0 && a->gw > 0 && a->d > 0, "dod_pos_resize_bicubic: bad shape");
  pos_resize_kernel<<<unsigned(1 + a->gh * a->gw), 256, 0, stream>>>(a->src, a->dst, int(a->g0),
                                                                    int(a->gh), int(a->gw), int(a->d));
  int rc = check_cuda(cudaGetLastError(), "pos_resize_kernel launch");
  if (rc == 0) count_launch();
  return rc;
}
