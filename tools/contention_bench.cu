// Micro-benchmark: do the tensor pipe (tcgen05.mma reading / writing TMEM) and the softmax warps (tcgen05.ld / MUFU /
// tcgen05.st) of the attention kernel overlap, or do they serialise on a shared resource?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I desta2.5-audio_b200/csrc -I include tools/contention_bench.cu -o tools/_bin/contention_bench
// One or two CTAs per SM; per CTA: warp 0 issues the attention kernel's MMA mix per "block" (4 SS 128x64x16 + 4 TS
// 128x64x16, independent of the softmax data), warps 1-4 run the softmax loop of tools/tmem_bench.cu on other TMEM columns.
// Each role runs a fixed number of blocks; the roles are timed alone and together.
#include <cstdio>

#include "common.cuh"

namespace desta { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }
using namespace desta;

__device__ __forceinline__ void fence32(uint32_t (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(v[i]));
}

// SM: 0 = ld + st only (values passed through), 1 = lean (ex2, pack, row sum), 2 = today's full body
template <int SM>
__device__ __forceinline__ void chunk(const uint32_t (&v)[32], uint32_t (&pk)[16], float& m0, float& m1, uint64_t& lacc) {
  const uint64_t c2 = pack_f32x2(1.0f, 1.0f), nmc2 = pack_f32x2(-0.5f, -0.5f);
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    float x0 = __uint_as_float(v[i]), x1 = __uint_as_float(v[i + 1]), x2 = __uint_as_float(v[i + 2]), x3 = __uint_as_float(v[i + 3]);
    if (SM == 2) {
      m0 = max3(m0, x0, x1);
      m1 = max3(m1, x2, x3);
      unpack_f32x2(ffma2(pack_f32x2(x0, x1), c2, nmc2), x0, x1);
      unpack_f32x2(ffma2(pack_f32x2(x2, x3), c2, nmc2), x2, x3);
    }
    if (SM >= 1) {
      x0 = fast_exp2(x0); x1 = fast_exp2(x1); x2 = fast_exp2(x2); x3 = fast_exp2(x3);
      lacc = fadd2(lacc, fadd2(pack_f32x2(x0, x1), pack_f32x2(x2, x3)));
    }
    pk[i >> 1] = pack_bf16(x0, x1);
    pk[(i >> 1) + 1] = pack_bf16(x2, x3);
  }
}

// ROLES: bit 0 = MMA warp active, bit 1 = softmax warps active
template <int ROLES, int SM>
__global__ void __launch_bounds__(160, 2) k(int blocks, long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<256>(&slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t t0 = slot;
  if (warp == 0) {
    if (ROLES & 1) {
      constexpr uint32_t idesc_ss = umma_idesc_bf16(128, 64, 0, 0), idesc_ts = umma_idesc_bf16(128, 64, 0, 1);
      const uint64_t adesc = umma_desc_sw128(smem_u32(smem), 16, 1024);
      const uint64_t bdesc = umma_desc_sw128(smem_u32(smem) + 16384, 16, 1024);
      const uint64_t vdesc = umma_desc_sw128(smem_u32(smem) + 32768, 1024, 1024);
      const long long c0 = clock64();
      for (int j = 0; j < blocks; ++j) {
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16_ts(t0 + 192, t0 + 160 + 8 * kk, vdesc + (uint64_t)(kk * 128), idesc_ts, 1);   // O += P V
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(t0 + (j % 2) * 64 + 0, adesc + 2 * kk, bdesc + 2 * kk, idesc_ss, kk != 0);   // S = Q K^T
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bar);
      mbar_wait(&bar, 0);
      if (threadIdx.x == 0) atomicMax((unsigned long long*)&out[0], (unsigned long long)(clock64() - c0));
    }
  } else if (ROLES & 2) {
    const uint32_t base = t0 + ((uint32_t)((warp - 1) * 32) << 16) + 128;     // columns [128, 192): not touched by the MMAs above... (P operand at 160 is only read)
    uint32_t a[32], b[32], pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = 0xbc00bc00u + i;
    tmem_st_32x16(base, pk); tmem_st_32x16(base + 16, pk); tmem_st_32x16(base + 32, pk); tmem_st_32x16(base + 48, pk);
    tmem_wait_st();
    float m0 = -1e30f, m1 = -1e30f;
    uint64_t lacc = 0;
    const long long c0 = clock64();
    tmem_ld_32x32(base, a);
    for (int it = 0; it < blocks; ++it) {
      tmem_wait_ld(); fence32(a);
      tmem_ld_32x32(base + 32, b);
      chunk<SM>(a, pk, m0, m1, lacc);
      if (it > 0) tmem_wait_st();
      tmem_st_32x16(base, pk);
      tmem_wait_ld(); fence32(b);
      tmem_ld_32x32(base, a);
      chunk<SM>(b, pk, m0, m1, lacc);
      tmem_st_32x16(base + 16, pk);
    }
    tmem_wait_ld(); tmem_wait_st();
    const long long c1 = clock64();
    float l0, l1; unpack_f32x2(lacc, l0, l1);
    float s = m0 + m1 + l0 + l1 + __uint_as_float(pk[5]);
#pragma unroll
    for (int i = 0; i < 32; ++i) s += __uint_as_float(a[i]) + __uint_as_float(b[i]);
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)&out[1], (unsigned long long)(c1 - c0));
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<256>(t0); }
}

template <int ROLES, int SM>
static void run(const char* name, int ctas_per_sm, long long* d, float* sink) {
  const int blocks = 2048;
  cudaMemset(d, 0, 16);
  cudaFuncSetAttribute(k<ROLES, SM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
  k<ROLES, SM><<<148 * ctas_per_sm, 160, 49152>>>(blocks, d, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-52s %d CTA/SM: MMA warp %7.1f cycles / block (8 UMMAs), softmax warps %7.1f cycles / block%s\n", name, ctas_per_sm,
         (double)h[0] / blocks, (double)h[1] / blocks, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d; float* sink;
  cudaMalloc(&d, 16); cudaMalloc(&sink, 148 * 2 * 160 * 4);
  for (int n = 1; n <= 2; ++n) {
    run<1, 0>("MMA stream alone", n, d, sink);
    run<2, 0>("softmax alone: ld + pack + st (no MUFU)", n, d, sink);
    run<3, 0>("both:          ld + pack + st (no MUFU)", n, d, sink);
    run<2, 1>("softmax alone: lean (ex2, pack, row sum)", n, d, sink);
    run<3, 1>("both:          lean", n, d, sink);
    run<2, 2>("softmax alone: today's body", n, d, sink);
    run<3, 2>("both:          today's body", n, d, sink);
  }
  return 0;
}
