// Micro-benchmark: what bounds the softmax warps of the attention kernel when the tensor pipe is out of the picture?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I desta2.5-audio_b200/csrc -I include tools/tmem_bench.cu -o tools/_bin/tmem_bench
// Modes (4 softmax-like warps per CTA, each on its own TMEM lane quarter; 1 or 2 CTAs per SM):
//   0  tcgen05.ld 32x32b.x32 back to back                      -> TMEM read bandwidth
//   1  tcgen05.st 32x32b.x16 back to back                      -> TMEM write bandwidth
//   2  ld.x32 + 32 ex2 (+ FFMA2 / max3 / add2 / pack as the kernel does) + st.x16, software pipelined like attention.cu
//   3  as 2 without the exponentials (the "no exp" experiment of profiles/r1d_ncu_summary.md in isolation)
//   4  as 2 without any TMEM traffic (registers only)          -> pure XU + issue floor
// Output: cycles per 32-column chunk per warp, and per "64-key block pair" (8 warps x 2 chunks) for comparison with the
// 1540 cycles the kernel takes and the 1024-cycle XU floor.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

using namespace desta;

__device__ __forceinline__ void fence32(uint32_t (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(v[i]));
}

template <int MODE>
__device__ __forceinline__ void chunk(const uint32_t (&v)[32], uint32_t (&pk)[16], float c, float mc, float& m0, float& m1, uint64_t& lacc) {
  const uint64_t c2 = pack_f32x2(c, c), nmc2 = pack_f32x2(-mc, -mc);
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float s0 = __uint_as_float(v[i]), s1 = __uint_as_float(v[i + 1]), s2 = __uint_as_float(v[i + 2]), s3 = __uint_as_float(v[i + 3]);
    m0 = max3(m0, s0, s1);
    m1 = max3(m1, s2, s3);
    float x0, x1, x2, x3;
    unpack_f32x2(ffma2(pack_f32x2(s0, s1), c2, nmc2), x0, x1);
    unpack_f32x2(ffma2(pack_f32x2(s2, s3), c2, nmc2), x2, x3);
    float p0 = x0, p1 = x1, p2 = x2, p3 = x3;
    if (MODE != 3) { p0 = fast_exp2(x0); p1 = fast_exp2(x1); p2 = fast_exp2(x2); p3 = fast_exp2(x3); }
    lacc = fadd2(lacc, fadd2(pack_f32x2(p0, p1), pack_f32x2(p2, p3)));
    pk[i >> 1] = pack_bf16(p0, p1);
    pk[(i >> 1) + 1] = pack_bf16(p2, p3);
  }
}

__device__ __forceinline__ void chunk_lean(const uint32_t (&v)[32], uint32_t (&pk)[16], uint64_t& lacc) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float p0 = fast_exp2(__uint_as_float(v[i])), p1 = fast_exp2(__uint_as_float(v[i + 1]));
    const float p2 = fast_exp2(__uint_as_float(v[i + 2])), p3 = fast_exp2(__uint_as_float(v[i + 3]));
    lacc = fadd2(lacc, fadd2(pack_f32x2(p0, p1), pack_f32x2(p2, p3)));
    pk[i >> 1] = pack_bf16(p0, p1);
    pk[(i >> 1) + 1] = pack_bf16(p2, p3);
  }
}

// software-pipelined bodies: the pack / row-sum of a group of 4 exponentials trails their MUFU issue by LAG groups, so that the
// consumers never wait on the XU latency (ptxas places them right behind their producers otherwise)
template <int LAG, bool FULL>
__device__ __forceinline__ void chunk_swp(const uint32_t (&v)[32], uint32_t (&pk)[16], float c, float mc, float& m0, float& m1, uint64_t& lacc) {
  float p[32];
  const uint64_t c2 = pack_f32x2(c, c), nmc2 = pack_f32x2(-mc, -mc);
#pragma unroll
  for (int g = 0; g < 8 + LAG; ++g) {
    if (g < 8) {
      const int i = 4 * g;
      float x0 = __uint_as_float(v[i]), x1 = __uint_as_float(v[i + 1]), x2 = __uint_as_float(v[i + 2]), x3 = __uint_as_float(v[i + 3]);
      if (FULL) {
        m0 = max3(m0, x0, x1);
        m1 = max3(m1, x2, x3);
        unpack_f32x2(ffma2(pack_f32x2(x0, x1), c2, nmc2), x0, x1);
        unpack_f32x2(ffma2(pack_f32x2(x2, x3), c2, nmc2), x2, x3);
      }
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p[i]) : "f"(x0));
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p[i + 1]) : "f"(x1));
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p[i + 2]) : "f"(x2));
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p[i + 3]) : "f"(x3));
    }
    if (g >= LAG) {
      const int i = 4 * (g - LAG);
      uint64_t t;
      asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(pack_f32x2(p[i], p[i + 1])), "l"(pack_f32x2(p[i + 2], p[i + 3])));
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(lacc) : "l"(t));
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[i >> 1]) : "f"(p[i + 1]), "f"(p[i]));
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[(i >> 1) + 1]) : "f"(p[i + 3]), "f"(p[i + 2]));
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(128, 2) k(float* out, int iters, long long* cyc) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<256>(&slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t base = slot + ((uint32_t)(warp * 32) << 16);
  uint32_t a[32], b[32], pk[16];
#pragma unroll
  for (int i = 0; i < 32; ++i) { a[i] = __float_as_uint(-0.01f * (threadIdx.x + i)); b[i] = a[i] ^ 0x100; }
#pragma unroll
  for (int i = 0; i < 16; ++i) pk[i] = i;
  // initialise the columns that are read (finite values) so that the MUFU inputs are well defined
  for (int cidx = 0; cidx < 256; cidx += 16) tmem_st_32x16(base + cidx, pk);
  tmem_wait_st();
  float m0 = -1e30f, m1 = -1e30f;
  uint64_t lacc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (MODE == 0) {
    for (int it = 0; it < iters; ++it) {
      tmem_ld_32x32(base + (it & 3) * 64, a);
      tmem_ld_32x32(base + (it & 3) * 64 + 32, b);
      tmem_wait_ld();
      fence32(a); fence32(b);
    }
  } else if (MODE == 1) {
    for (int it = 0; it < iters; ++it) {
      tmem_st_32x16(base + (it & 3) * 64, pk);
      tmem_st_32x16(base + (it & 3) * 64 + 16, pk);
      tmem_wait_st();
    }
  } else if (MODE == 4) {
    for (int it = 0; it < iters; ++it) {
      chunk<2>(a, pk, 1.0f, 0.5f, m0, m1, lacc);
      fence32(a);
      chunk<2>(b, pk, 1.0f, 0.5f, m0, m1, lacc);
      fence32(b);
    }
  } else {
    // block = two chunks; the next block's first chunk is requested while the second chunk is processed (as attention.cu)
    tmem_ld_32x32(base, a);
    for (int it = 0; it < iters; ++it) {
      const uint32_t tS = base + (it % 3) * 64, tN = base + ((it + 1) % 3) * 64;
      tmem_wait_ld(); fence32(a);
      tmem_ld_32x32(tS + 32, b);
      if (MODE == 5) chunk_lean(a, pk, lacc);
      else if (MODE == 6) chunk_swp<2, false>(a, pk, 1.0f, 0.5f, m0, m1, lacc);
      else if (MODE == 7) chunk_swp<2, true>(a, pk, 1.0f, 0.5f, m0, m1, lacc);
      else if (MODE == 8) chunk_swp<3, true>(a, pk, 1.0f, 0.5f, m0, m1, lacc);
      else if (MODE == 9) chunk_swp<1, true>(a, pk, 1.0f, 0.5f, m0, m1, lacc);
      else chunk<MODE>(a, pk, 1.0f, 0.5f, m0, m1, lacc);
      if (it > 0) tmem_wait_st();
      tmem_st_32x16(tS, pk);
      tmem_wait_ld(); fence32(b);
      tmem_ld_32x32(tN, a);
      if (MODE == 5) chunk_lean(b, pk, lacc);
      else if (MODE == 6) chunk_swp<2, false>(b, pk, 1.0f, 0.5f, m0, m1, lacc);
      else if (MODE == 7) chunk_swp<2, true>(b, pk, 1.0f, 0.5f, m0, m1, lacc);
      else if (MODE == 8) chunk_swp<3, true>(b, pk, 1.0f, 0.5f, m0, m1, lacc);
      else if (MODE == 9) chunk_swp<1, true>(b, pk, 1.0f, 0.5f, m0, m1, lacc);
      else chunk<MODE>(b, pk, 1.0f, 0.5f, m0, m1, lacc);
      tmem_st_32x16(tS + 16, pk);
    }
    tmem_wait_ld(); tmem_wait_st();
  }
  const long long t1 = clock64();
  float s = m0 + m1;
  float l0, l1; unpack_f32x2(lacc, l0, l1);
#pragma unroll
  for (int i = 0; i < 32; ++i) s += __uint_as_float(a[i]) + __uint_as_float(b[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + l0 + l1 + __uint_as_float(pk[3]);
  if (threadIdx.x == 0) atomicMax((unsigned long long*)cyc, (unsigned long long)(t1 - t0));
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<256>(slot); }
}

template <int MODE>
static void run(const char* name, int ctas_per_sm, float* o, long long* c) {
  const int iters = 2048;
  cudaMemset(c, 0, 8);
  k<MODE><<<148 * ctas_per_sm, 128>>>(o, iters, c);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  const double per_block = (double)h / iters;      // per warp: one "block" = two 32-column chunks (= 64 keys of one row set)
  printf("%-44s %d CTA/SM: %8.1f cycles per 64-column block per warp  (x32 ld pair / x16 st pair)%s\n", name, ctas_per_sm, per_block,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  float* o; long long* c;
  cudaMalloc(&o, 148 * 2 * 128 * 4); cudaMalloc(&c, 8);
  for (int n = 1; n <= 2; ++n) {
    run<1>("1 tcgen05.st x16 x2 only", n, o, c);
    run<2>("2 ld + softmax chunk (ex2) + st, pipelined", n, o, c);
    run<3>("3 same without ex2", n, o, c);
    run<5>("5 ld + LEAN chunk (ex2, pack, row sum) + st", n, o, c);
    run<6>("6 LEAN, software pipelined, lag 2 groups", n, o, c);
    run<7>("7 today's body, software pipelined, lag 2", n, o, c);
    run<8>("8 today's body, software pipelined, lag 3", n, o, c);
    run<9>("9 today's body, software pipelined, lag 1", n, o, c);
  }
  printf("reference points: kernel softmax loop 1540 cycles per block (2 CTAs/SM), XU floor 1024, bytes: ld 8 KB + st 4 KB per warp per block\n");
  return 0;
}
