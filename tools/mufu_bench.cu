// Micro-benchmark: MUFU ex2 throughput per SM for f32 vs packed f16x2 / bf16x2 inputs (is the packed form 2 results per op?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/mufu_bench.cu -o tools/_bin/mufu_bench
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, long long* cyc) {
  float x[8];
  unsigned h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = -0.001f * (threadIdx.x + i); h[i] = 0xbc00b800u + i; }   // (-1.0, -0.5) as f16x2
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) asm volatile("ex2.approx.f16 %0, %0;" : "+h"(*reinterpret_cast<unsigned short*>(&h[i])));
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  float* o; long long* c;
  cudaMalloc(&o, 148 * 256 * 4); cudaMalloc(&c, 8);
  const int iters = 4096;
  const char* names[] = {"ex2.approx.ftz.f32", "ex2.approx.f16x2", "ex2.approx.ftz.bf16x2", "ex2.approx.f16"};
  for (int m = 0; m < 4; ++m) {
    if (m == 0) k<0><<<148, 256>>>(o, iters, c);
    if (m == 1) k<1><<<148, 256>>>(o, iters, c);
    if (m == 2) k<2><<<148, 256>>>(o, iters, c);
    if (m == 3) k<3><<<148, 256>>>(o, iters, c);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    // 8 warps per SM (2 per SMSP), 8 independent ops per thread per iteration
    const double per_warp_instr = (double)h / (iters * 8.0 * 2.0);   // cycles per warp-instruction per SMSP
    printf("%-24s %8.2f cycles per warp instruction per SMSP  -> %.1f results / clk / SM %s\n", names[m], per_warp_instr,
           (m == 1 || m == 2 ? 64.0 : 32.0) * 4 / per_warp_instr, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
