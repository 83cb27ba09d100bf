// Micro-benchmark: which instructions of the attention softmax share the MUFU (XU) pipe with ex2?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/xu_mix_bench.cu -o tools/_bin/xu_mix_bench
// Modes: 0 ex2 only | 1 cvt.rn.bf16x2.f32 only | 2 ex2 x2 + cvt x1 (the softmax mix) | 3 ex2 x2 + prmt x1 (truncating pack)
//        4 ex2 x2 + ffma x2 (control: FMA pipe work does not slow ex2 down)
#include <cstdio>
#include <cuda_bf16.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, long long* cyc) {
  float x[8];
  unsigned h[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = i;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      if (MODE == 0 || MODE >= 2) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i + 1]));
      }
      if (MODE == 1 || MODE == 2) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i >> 1]) : "f"(x[i + 1]), "f"(x[i]));
      if (MODE == 3) asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(h[i >> 1]) : "r"(__float_as_uint(x[i])), "r"(__float_as_uint(x[i + 1])));
      if (MODE == 4) {
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[i]) : "f"(x[i + 1]));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[i + 1]) : "f"(x[i]));
      }
      if (MODE == 1) { x[i] += __uint_as_float(h[i >> 1] & 1u); }
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) s += __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  float* o; long long* c;
  cudaMalloc(&o, 148 * 256 * 4); cudaMalloc(&c, 8);
  const int iters = 4096;
  const char* names[] = {"ex2 only (8 per iter)", "cvt.rn.bf16x2.f32 only (4 per iter)", "ex2 x8 + cvt.bf16x2 x4", "ex2 x8 + prmt x4", "ex2 x8 + ffma x8"};
  for (int m = 0; m < 5; ++m) {
    if (m == 0) k<0><<<148, 256>>>(o, iters, c);
    if (m == 1) k<1><<<148, 256>>>(o, iters, c);
    if (m == 2) k<2><<<148, 256>>>(o, iters, c);
    if (m == 3) k<3><<<148, 256>>>(o, iters, c);
    if (m == 4) k<4><<<148, 256>>>(o, iters, c);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    // 8 warps per SM = 2 per SMSP; cycles per iteration per SMSP (both warps)
    printf("%-40s %8.2f cycles per iteration per warp pair (SMSP) %s\n", names[m], (double)h / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
