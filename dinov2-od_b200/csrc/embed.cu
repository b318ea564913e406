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

// grid (patch-row, batch); block 256.  Each image row segment is read coalesced.
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ pix, __nv_bfloat16* __restrict__ out, int H, int W,
                int gh, int gw, int kpad) {
  const int py = blockIdx.x, b = blockIdx.y;
  const int wuse = gw * kP;
  const int64_t row0 = (int64_t(b) * gh + py) * gw;
  for (int ci = 0; ci < 3 * kP; ++ci) {
    const int c = ci / kP, i = ci % kP;
    const float* src = pix + ((int64_t(b) * 3 + c) * H + (py * kP + i)) * W;
    for (int x = threadIdx.x; x < wuse; x += blockDim.x) {
      const int px = x / kP, j = x - px * kP;
      out[(row0 + px) * kpad + c * (kP * kP) + i * kP + j] = __float2bfloat16_rn(src[x]);
    }
  }
  // zero the K padding so that 0-weights never meet NaN garbage
  const int padw = kpad - kK;
  for (int t = threadIdx.x; t < gw * padw; t += blockDim.x) {
    const int px = t / padw, k = kK + t % padw;
    out[(row0 + px) * kpad + k] = __float2bfloat16_rn(0.f);
  }
}

// uint8 NHWC variant (decoder output of PIL / numpy): fuses ToTensor's /255 (reference
// train.py:584-587, dataset.py:55-66) into the im2col, so the host->device copy is 1 byte per
// sample instead of 4.  grid (patch-row, batch); a block walks its 14 image rows, reading the
// interleaved RGB bytes of a row contiguously.
__global__ void __launch_bounds__(256)
patchify_u8_kernel(const uint8_t* __restrict__ pix, __nv_bfloat16* __restrict__ out, int H, int W,
                   int gh, int gw, int kpad) {
  const int py = blockIdx.x, b = blockIdx.y;
  const int wuse = gw * kP * 3;
  const int64_t row0 = (int64_t(b) * gh + py) * gw;
  for (int i = 0; i < kP; ++i) {
    const uint8_t* src = pix + ((int64_t(b) * H + (py * kP + i)) * W) * 3;
    for (int t = threadIdx.x; t < wuse; t += blockDim.x) {
      const int x = t / 3, c = t - x * 3;
      const int px = x / kP, j = x - px * kP;
      out[(row0 + px) * kpad + c * (kP * kP) + i * kP + j] = __float2bfloat16_rn(float(src[t]) / 255.0f);
    }
  }
  const int padw = kpad - kK;
  for (int t = threadIdx.x; t < gw * padw; t += blockDim.x) {
    const int px = t / padw, k = kK + t % padw;
    out[(row0 + px) * kpad + k] = __float2bfloat16_rn(0.f);
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
  if (a->pixel_format == 1)
    patchify_u8_kernel<<<dim3(gh, unsigned(a->batch)), 256, 0, stream>>>(
        reinterpret_cast<const uint8_t*>(a->pixels), reinterpret_cast<__nv_bfloat16*>(a->patches),
        int(a->height), int(a->width), gh, gw, int(a->kpad));
  else
    patchify_kernel<<<dim3(gh, unsigned(a->batch)), 256, 0, stream>>>(
        reinterpret_cast<const float*>(a->pixels), reinterpret_cast<__nv_bfloat16*>(a->patches),
        int(a->height), int(a->width), gh, gw, int(a->kpad));
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
