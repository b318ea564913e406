// Developer microbenchmark: clocks per 128x128 attention tile (per SM) of the softmax inner loop's instruction mix
// on sm_100a, from registers only (no TMEM, no barriers): what the math alone allows.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_mix softmax_mix.cu && ./softmax_mix
// One "tile" = 128 rows x 128 columns of scores = 16384 exp2.  A warp owns 32 rows; COLS columns per thread
// (128: one thread per row, 4 warps per tile; 64: two threads per row, 8 warps per tile).
// Variants:  POLY of every 8 element PAIRS use the FMA-pipe polynomial instead of MUFU.EX2;
//            MAXP: also track the running row maximum (FMNMX3) in the same pass;
//            PACK: convert to bf16x2 (F2FP).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// 2^x for x <= ~8, x >= -126 (clamped): round-to-nearest split + degree-3 minimax on [-0.5, 0.5]
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  const float magic = 12582912.0f;  // 1.5 * 2^23
  x.x = fmaxf(x.x, -126.0f);
  x.y = fmaxf(x.y, -126.0f);
  const float2 t = __fadd2_rn(x, make_float2(magic, magic));
  const float2 xr = __fadd2_rn(t, make_float2(-magic, -magic));
  const float2 f = __fadd2_rn(x, make_float2(-xr.x, -xr.y));
  float2 p = __ffma2_rn(f, make_float2(0.05550357f, 0.05550357f), make_float2(0.24022651f, 0.24022651f));
  p = __ffma2_rn(p, f, make_float2(0.69314718f, 0.69314718f));
  p = __ffma2_rn(p, f, make_float2(1.0f, 1.0f));
  float2 r;
  r.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  r.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return r;
}

template <int COLS, int POLY, bool MAXP, bool PACK>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, int iters, float c) {
  float s[COLS];
#pragma unroll
  for (int i = 0; i < COLS; ++i) s[i] = -0.37f * i - 0.001f * threadIdx.x;
  float m = 0.25f;
  float2 sum = make_float2(0.f, 0.f);
  uint32_t acc = 0;
  float mx0 = -1e30f, mx1 = -1e30f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const float2 sc2 = make_float2(c, c), nm2 = make_float2(-m, -m);
#pragma unroll
    for (int i = 0; i < COLS; i += 2) {
      const float2 x = __ffma2_rn(make_float2(s[i], s[i + 1]), sc2, nm2);
      float2 e;
      if (((i >> 1) & 7) < POLY) e = ex2_poly2(x);
      else e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
      sum = __fadd2_rn(sum, e);
      if (PACK) acc ^= pack_bf16x2(e.x, e.y);
      else acc ^= __float_as_uint(e.x) + __float_as_uint(e.y);
      if (MAXP) {
        if (i & 2) mx0 = fmaxf(mx0, fmaxf(s[i], s[i + 1]));
        else mx1 = fmaxf(mx1, fmaxf(s[i], s[i + 1]));
      }
    }
    m += 0.0009765625f;
    if (MAXP) s[it & (COLS - 1)] += 1e-3f;  // keep the maxima loop-variant
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum.x + sum.y + __uint_as_float(acc) + mx0 + mx1;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int COLS, int POLY, bool MAXP, bool PACK>
void run(int warps_per_smsp, float* out, long long* cyc) {
  const int iters = 400;
  const int threads = warps_per_smsp * 4 * 32;
  k<COLS, POLY, MAXP, PACK><<<148, threads>>>(out, cyc, iters, 0.18f);
  k<COLS, POLY, MAXP, PACK><<<148, threads>>>(out, cyc, iters, 0.18f);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  // per iteration the SM processed threads * COLS elements = that many / 16384 tiles
  const double tiles = double(threads) * COLS / 16384.0;
  printf("cols/thread %3d  warps/SMSP %d  poly %d/8  max %d  pack %d : %7.1f clk per 128x128 tile per SM\n", COLS,
         warps_per_smsp, POLY, int(MAXP), int(PACK), double(h) / iters / tiles);
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 8);
  for (int w : {1, 2, 4}) {
    run<128, 0, false, false>(w, out, cyc);
    run<128, 0, false, true>(w, out, cyc);
    run<128, 0, true, true>(w, out, cyc);
    run<128, 1, true, true>(w, out, cyc);
    run<128, 2, true, true>(w, out, cyc);
    run<128, 3, true, true>(w, out, cyc);
    run<128, 4, true, true>(w, out, cyc);
    run<128, 8, true, true>(w, out, cyc);
    run<64, 0, true, true>(w, out, cyc);
    run<64, 2, true, true>(w, out, cyc);
    run<64, 3, true, true>(w, out, cyc);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
