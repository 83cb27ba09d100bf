// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16, M = 128) as a function of HOW it is issued.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I desta2.5-audio_b200/csrc tools/umma_bench.cu -o tools/_bin/umma_bench
// STYLE 0: `if (threadIdx.x == 0)` divergent single-thread issue (what the kernels did);
// STYLE 1: the whole warp runs the loop converged, each instruction under elect.sync.
#include <cstdio>
#include "common.cuh"

namespace desta { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }
using namespace desta;

template <int STYLE, int N, bool TS, int AMN = 0, int BMN = 0>
__global__ void __launch_bounds__(128) umma_bench(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t t0 = slot;
  constexpr uint32_t idesc = umma_idesc_bf16(128, N, AMN, BMN);
  // K-major: 16 bf16 of K = 32 B inside the 128-byte row (+2);  MN-major: 16 K rows = two 8-row groups = 2048 B (+128)
  const uint64_t adesc = AMN ? umma_desc_sw128(smem_u32(smem), 16384, 1024) : umma_desc_sw128(smem_u32(smem), 16, 1024);
  const uint64_t bdesc = BMN ? umma_desc_sw128(smem_u32(smem) + 32768, 1024, 1024) : umma_desc_sw128(smem_u32(smem) + 32768, 16, 1024);
  constexpr int AS = AMN ? 128 : 2, BS = BMN ? 128 : 2;
  if (STYLE == 0) {
    if (threadIdx.x == 0) {
      for (int i = 0; i < 8; ++i) umma_bf16_ss(t0, adesc, bdesc, idesc, 1);
      umma_commit(&bar); mbar_wait(&bar, 0);
      const long long c0 = clock64();
      for (int i = 0; i < iters; i += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (TS) umma_bf16_ts(t0, t0 + 256 + 8 * k, bdesc + BS * k, idesc, 1);
          else umma_bf16_ss(t0, adesc + AS * k, bdesc + BS * k, idesc, 1);
        }
      }
      umma_commit(&bar); mbar_wait(&bar, 1);
      if (blockIdx.x == 0) out[0] = clock64() - c0;
    }
  } else {
    if (warp == 0) {
      if (elect_one()) { for (int i = 0; i < 8; ++i) umma_bf16_ss(t0, adesc, bdesc, idesc, 1); umma_commit(&bar); }
      mbar_wait(&bar, 0);
      __syncwarp();
      const long long c0 = clock64();
      for (int i = 0; i < iters; i += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (elect_one()) {
            if (TS) umma_bf16_ts(t0, t0 + 256 + 8 * k, bdesc + BS * k, idesc, 1);
            else umma_bf16_ss(t0, adesc + AS * k, bdesc + BS * k, idesc, 1);
          }
        }
      }
      if (elect_one()) umma_commit(&bar);
      mbar_wait(&bar, 1);
      if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = clock64() - c0;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<512>(t0); }
}

template <int STYLE, int N, bool TS, int AMN = 0, int BMN = 0>
void run(long long* d, const char* name) {
  const int iters = 1024;
  cudaFuncSetAttribute(umma_bench<STYLE, N, TS, AMN, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  umma_bench<STYLE, N, TS, AMN, BMN><<<148, 128, 65536>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("style %d  128x%3dx16 %s  %-28s %7.1f cyc/mma  (math floor %d)%s\n", STYLE, N, TS ? "TS" : "SS", name, (double)h / iters,
         128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  run<0, 64, false>(d, "thread-0 branch");   run<1, 64, false>(d, "converged warp + elect");
  run<0, 64, true>(d, "thread-0 branch");    run<1, 64, true>(d, "converged warp + elect");
  run<0, 128, false>(d, "thread-0 branch");  run<1, 128, false>(d, "converged warp + elect");
  run<0, 256, false>(d, "thread-0 branch");  run<1, 256, false>(d, "converged warp + elect");
  run<1, 64, false, 0, 1>(d, "B MN-major");
  run<1, 64, false, 1, 0>(d, "A MN-major");
  run<1, 64, false, 1, 1>(d, "A and B MN-major");
  run<1, 64, true, 0, 1>(d, "A TMEM, B MN-major");
  return 0;
}
