// Library plumbing of libdod: error strings, device check, TMA descriptor
// factory (driver entry point resolved at run time so the .so has no link-time
// dependency on libcuda and loads on a CPU-only host), launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "../../include/dod.h"

namespace dod {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error %d (%s) at %s", int(e), cudaGetErrorString(e), what);
  return DOD_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static CUtensorMapDataType tmap_dtype(int elt_bytes) {
  return elt_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
}

int make_tmap_2d(CUtensorMap* out, const void* base, int elt_bytes, uint64_t rows, uint64_t cols,
                 uint64_t ld, uint32_t box_rows, uint32_t box_cols, int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return DOD_ERR_CUDA;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * uint64_t(elt_bytes)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                 : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                       : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = fn(out, tmap_dtype(elt_bytes), 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d rows=%llu cols=%llu ld=%llu box=%ux%u) failed: %d",
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows,
              box_cols, int(r));
    return DOD_ERR_CUDA;
  }
  return 0;
}

int make_tmap_3d(CUtensorMap* out, const void* base, int elt_bytes, uint64_t batch, uint64_t rows,
                 uint64_t cols, uint64_t bs, uint64_t ld, uint32_t box_rows, uint32_t box_cols,
                 int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return DOD_ERR_CUDA;
  }
  cuuint64_t dims[3] = {cols, rows, batch};
  cuuint64_t strides[2] = {ld * uint64_t(elt_bytes), bs * uint64_t(elt_bytes)};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                 : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                       : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = fn(out, tmap_dtype(elt_bytes), 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d batch=%llu rows=%llu cols=%llu) failed: %d",
              (unsigned long long)batch, (unsigned long long)rows, (unsigned long long)cols,
              int(r));
    return DOD_ERR_CUDA;
  }
  return 0;
}

int make_tmap_4d(CUtensorMap* out, const void* base, int elt_bytes, uint64_t outer, uint64_t inner,
                 uint64_t rows, uint64_t cols, uint64_t outer_stride, uint64_t inner_stride, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols, int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return DOD_ERR_CUDA;
  }
  cuuint64_t dims[4] = {cols, rows, inner, outer};
  cuuint64_t strides[3] = {ld * uint64_t(elt_bytes), inner_stride * uint64_t(elt_bytes),
                           outer_stride * uint64_t(elt_bytes)};
  cuuint32_t box[4] = {box_cols, box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                 : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                       : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = fn(out, tmap_dtype(elt_bytes), 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d outer=%llu inner=%llu rows=%llu cols=%llu strides=%llu/%llu/%llu) failed: %d",
              (unsigned long long)outer, (unsigned long long)inner, (unsigned long long)rows,
              (unsigned long long)cols, (unsigned long long)outer_stride, (unsigned long long)inner_stride,
              (unsigned long long)ld, int(r));
    return DOD_ERR_CUDA;
  }
  return 0;
}

}  // namespace dod

extern "C" {

int32_t dod_version(void) { return 100; /* 0.1.0 */ }

const char* dod_last_error(void) { return dod::g_err; }

int32_t dod_device_check(int32_t dev) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return dod::check_cuda(e, "cudaGetDeviceProperties");
  if (prop.major != 10) {
    dod::set_error("device %d is sm_%d%d; libdod is built for sm_100a (B200) only", dev,
                   prop.major, prop.minor);
    return DOD_ERR_DEVICE;
  }
  return DOD_OK;
}

int64_t dod_launch_count(void) { return dod::g_launches.load(std::memory_order_relaxed); }
void dod_launch_count_reset(void) { dod::g_launches.store(0, std::memory_order_relaxed); }

}  // extern "C"
