// Developer microbenchmark: does MUFU.EX2 overlap with FMA-pipe issue on sm_100a?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_overlap mufu_overlap.cu && ./mufu_overlap
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int NM, int NF>
__global__ void k(float* out, long long* cyc, int iters) {
  float m[8], f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { m[i] = -1.0f - threadIdx.x * 1e-3f - i; f[i] = 1.0f + i + threadIdx.x; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < NM; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[i & 7]));
#pragma unroll
      for (int i = 0; i < NF; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i & 7]) : "f"(1.0001f), "f"(0.5f));
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += m[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int NM, int NF>
void run(int warps_per_sm, float* out, long long* cyc) {
  const int iters = 2000;
  k<NM, NF><<<148, warps_per_sm * 32>>>(out, cyc, iters);
  k<NM, NF><<<148, warps_per_sm * 32>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per_iter = double(h) / (iters * 8);
  printf("warps/SM %2d  MUFU %d FFMA %2d per group: %.1f clk/group/warp-set  (per SMSP: %d warps)\n", warps_per_sm, NM, NF, per_iter,
         warps_per_sm / 4);
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 8);
  for (int w : {4, 8, 16}) {
    run<8, 0>(w, out, cyc);
    run<0, 24>(w, out, cyc);
    run<8, 24>(w, out, cyc);
    run<8, 48>(w, out, cyc);
    run<0, 48>(w, out, cyc);
  }
  return 0;
}
