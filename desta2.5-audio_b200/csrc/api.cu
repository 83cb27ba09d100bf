// Host-side plumbing of libdesta_b200.so: error string, launch checks, TMA descriptor encoding.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "desta_b200.h"

namespace desta {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return DESTA_ERR_CUDA;
  }
  return DESTA_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
      set_error("cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
      return nullptr;
    }
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

int make_tmap(CUtensorMap* out, const void* base, CUtensorMapDataType dt, int elem_bytes, int rank,
              const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = get_encode();
  if (!fn) return DESTA_ERR_CUDA;
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i];
  }
  (void)elem_bytes;
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu] box [%u,%u,%u] base %p", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, base);
    return DESTA_ERR_CUDA;
  }
  return DESTA_OK;
}

bool* device_once_flag(int slot) {
  static bool flags[DESTA_MAX_DEVICES][8] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= DESTA_MAX_DEVICES) dev = 0;
  return &flags[dev][slot & 7];
}

int num_sms() {
  static int sms[DESTA_MAX_DEVICES] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= DESTA_MAX_DEVICES) dev = 0;
  if (sms[dev] == 0) {
    cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (sms[dev] <= 0) sms[dev] = 148;
  }
  return sms[dev];
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows, uint32_t box_cols) {
  uint64_t dims[2] = {cols, rows};
  uint64_t strides[2] = {0, ld_elems * 2};
  uint32_t box[2] = {box_cols, box_rows};
  return make_tmap(out, base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace desta

extern "C" int desta_version(void) { return 100; }
extern "C" const char* desta_last_error(void) { return desta::g_err; }
