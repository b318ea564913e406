// Developer microbenchmark: issue cost, throughput and commit latency of tcgen05.mma on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../include -o umma_issue umma_issue.cu
#include "../../dinov2-od_b200/csrc/common.cuh"
using namespace dod;

namespace dod { void set_error(const char*, ...) {} int check_cuda(cudaError_t e, const char*) { return int(e); } }

// mode 0: SS M128 N{n} K16 (K-major A,B)   mode 1: TS M128 N64 K16 (B MN-major, A from TMEM)
template <int MODE, int N>
__global__ void __launch_bounds__(128) k(long long* out, int cnt, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<256>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, false, MODE == 1);
    long long issue = 0, total = 0;
    for (int r = 0; r < reps; ++r) {
      long long t0 = clock64();
      for (int i = 0; i < cnt; ++i) {
        const uint32_t off = (i & 7) * 2048;
        if (MODE == 0) {
          const uint64_t da = make_sdesc_sw128(smem_u32(smem) + (i & 3) * 32, 16, 1024);
          const uint64_t db = make_sdesc_sw128(smem_u32(smem + 32768) + (i & 3) * 32, 16, 1024);
          umma_ss(tmem, da, db, idesc, i != 0);
        } else {
          const uint64_t dv = make_sdesc_sw128(smem_u32(smem) + off, 16, 1024);
          umma_ts(tmem + 192, tmem + 128 + 8 * (i & 7), dv, idesc, i != 0);
        }
      }
      long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, r & 1);
      long long t2 = clock64();
      if (r > 0) { issue += t1 - t0; total += t2 - t0; }
    }
    if (blockIdx.x == 0) { out[0] = issue / (reps - 1); out[1] = total / (reps - 1); }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<256>(tmem); }
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// whole warp walks the loop (uniform control flow); only the tcgen05 instructions are elected
template <int N>
__global__ void __launch_bounds__(128) k_conv(long long* out, int cnt, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<256>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, false, false);
    long long issue = 0, total = 0;
    const uint32_t sbase = smem_u32(smem);
    for (int r = 0; r < reps; ++r) {
      long long t0 = clock64();
      for (int i = 0; i < cnt; ++i) {
        const uint64_t da = make_sdesc_sw128(sbase + (i & 3) * 32, 16, 1024);
        const uint64_t db = make_sdesc_sw128(sbase + 32768 + (i & 3) * 32, 16, 1024);
        if (elect_one()) umma_ss(tmem, da, db, idesc, i != 0);
      }
      long long t1 = clock64();
      if (elect_one()) umma_commit(&bar);
      mbar_wait(&bar, r & 1);
      long long t2 = clock64();
      if (r > 0) { issue += t1 - t0; total += t2 - t0; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = issue / (reps - 1); out[1] = total / (reps - 1); }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<256>(tmem); }
}

template <int N>
void run_conv(const char* name, int grid, long long* d) {
  cudaFuncSetAttribute(k_conv<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  for (int cnt : {1, 4, 16, 64}) {
    k_conv<N><<<grid, 128, 66 * 1024>>>(d, cnt, 20);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-22s grid %3d cnt %2d: issue %5lld clk (%.0f/mma)  issue+commit+wait %5lld clk (%.0f/mma) %s\n", name, grid, cnt, h[0],
           double(h[0]) / cnt, h[1], double(h[1]) / cnt, e ? cudaGetErrorString(e) : "");
  }
}

template <int MODE, int N>
void run(const char* name, int grid, long long* d) {
  cudaFuncSetAttribute(k<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  for (int cnt : {1, 2, 4, 8, 16, 64}) {
    k<MODE, N><<<grid, 128, 66 * 1024>>>(d, cnt, 20);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-22s grid %3d cnt %2d: issue %5lld clk (%.0f/mma)  issue+commit+wait %5lld clk (%.0f/mma) %s\n", name, grid, cnt, h[0],
           double(h[0]) / cnt, h[1], double(h[1]) / cnt, e ? cudaGetErrorString(e) : "");
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  run_conv<64>("converged SS N64", 148, d);
  run_conv<128>("converged SS N128", 148, d);
  run_conv<256>("converged SS N256", 148, d);
  for (int grid : {148}) {
    run<0, 64>("SS M128 N64 K16", grid, d);
    run<0, 128>("SS M128 N128 K16", grid, d);
    run<0, 256>("SS M128 N256 K16", grid, d);
    run<1, 64>("TS M128 N64 K16 (PV)", grid, d);
  }
  // two CTAs per SM issuing at once (the FMHA kernel's residency): does the tensor pipe serialise them?
  printf("--- 296 CTAs (2 per SM)\n");
  run<0, 64>("SS M128 N64 K16", 296, d);
  run<0, 128>("SS M128 N128 K16", 296, d);
  run<1, 64>("TS M128 N64 K16 (PV)", 296, d);
  return 0;
}
