// Micro-benchmark: do MUFU.EX2 and the other instructions of the softmax loop overlap on one SM sub-partition, or do they
// serialise on the dispatch port?  (profiles/r2_issue_bench.md)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I desta2.5-audio_b200/csrc -I include tools/issue_bench.cu -o tools/_bin/issue_bench
// Every stream is 8 independent loop-carried chains per thread; W warps per SM sub-partition (W = 1, 2, 4).
#include <cstdio>

#include "common.cuh"

using namespace desta;

enum { S_MUFU, S_FFMA, S_FFMA2, S_FADD2, S_FMNMX3, S_F2FP, S_HMNMX2, S_LEA, S_MUFU_FFMA, S_MUFU_2FFMA, S_MUFU_FFMA2, S_MUFU_F2FP,
       S_MUFU_FMNMX3, S_MUFU_4FFMA, S_SOFTMAX_NOW, S_SOFTMAX_NEW, S_EX2_H2, S_EX2_BF2, S_CVT_H2, S_HADD2, S_EX2H2_BODY, S_COUNT };

template <int S>
__global__ void __launch_bounds__(512) k(float* out, int iters, long long* cyc) {
  float x[8], y[8], w[8];
  uint64_t q[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = -0.001f * (threadIdx.x + i);
    y[i] = 0.5f + 0.001f * i;
    w[i] = 0.25f + 0.002f * i;
    q[i] = pack_f32x2(y[i], x[i]);
    h[i] = 0x3f803f00u + i;
  }
  const float a = 0.999f, b = 0.001f;
  const uint64_t a2 = pack_f32x2(a, a), b2 = pack_f32x2(b, b);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (S == S_MUFU || S >= S_MUFU_FFMA && S <= S_MUFU_4FFMA) x[i] = fast_exp2(x[i]);
      if (S == S_FFMA || S == S_MUFU_FFMA || S == S_MUFU_2FFMA || S == S_MUFU_4FFMA) y[i] = fmaf(y[i], a, b);
      if (S == S_MUFU_2FFMA || S == S_MUFU_4FFMA) w[i] = fmaf(w[i], a, b);
      if (S == S_MUFU_4FFMA) { y[i] = fmaf(y[i], a, b); w[i] = fmaf(w[i], a, b); }
      if (S == S_FFMA2 || S == S_MUFU_FFMA2) q[i] = ffma2(q[i], a2, b2);
      if (S == S_FADD2) q[i] = fadd2(q[i], b2);
      if (S == S_FMNMX3 || S == S_MUFU_FMNMX3) y[i] = max3(y[i], w[i], w[(i + 1) & 7]) * 1.0f, w[i] = y[(i + 3) & 7];
      if (S == S_F2FP || S == S_MUFU_F2FP) { h[i] = pack_bf16(__uint_as_float(h[i]), __uint_as_float(h[(i + 1) & 7])); }
      if (S == S_HMNMX2) asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(h[i]) : "r"(h[(i + 1) & 7]));
      if (S == S_LEA) h[i] = (h[i] << 23) + h[(i + 1) & 7];
      if (S == S_EX2_H2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (S == S_EX2_BF2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (S == S_CVT_H2) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(__uint_as_float(h[i])), "f"(__uint_as_float(h[(i + 1) & 7])));
      if (S == S_HADD2) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(h[i]) : "r"(h[(i + 1) & 7]));
    }
    if (S == S_SOFTMAX_NOW) {
      // per 4 elements as attention.cu does today: 2 max3, 2 FFMA2, 4 ex2, 2 FADD2, 2 F2FP  (x2 to cover 8 elements)
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int i = 4 * g;
        y[0] = max3(y[0], x[i], x[i + 1]);
        y[1] = max3(y[1], x[i + 2], x[i + 3]);
        float x0, x1, x2, x3;
        unpack_f32x2(ffma2(pack_f32x2(x[i], x[i + 1]), a2, b2), x0, x1);
        unpack_f32x2(ffma2(pack_f32x2(x[i + 2], x[i + 3]), a2, b2), x2, x3);
        const float p0 = fast_exp2(x0), p1 = fast_exp2(x1), p2 = fast_exp2(x2), p3 = fast_exp2(x3);
        q[0] = fadd2(q[0], fadd2(pack_f32x2(p0, p1), pack_f32x2(p2, p3)));
        h[2 * g] = pack_bf16(p0, p1);
        h[2 * g + 1] = pack_bf16(p2, p3);
        x[i] = p0 - 1.5f; x[i + 1] = p1 - 1.5f; x[i + 2] = p2 - 1.5f; x[i + 3] = p3 - 1.5f;   // loop-carried (4 extra FADD, see NEW too)
      }
    }
    if (S == S_EX2H2_BODY) {
      // 8 elements: 4 cvt f32x2->f16x2, 4 ex2.f16x2, 3 add.f16x2 (sum of 8 in two halves), row sum to fp32 (2 cvt + 1 ffma2)
      uint32_t e[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(e[g]) : "f"(x[2 * g + 1]), "f"(x[2 * g]));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(e[g]));
        h[g] = e[g];
      }
      uint32_t s01, s23, s4;
      asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(s01) : "r"(e[0]), "r"(e[1]));
      asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(s23) : "r"(e[2]), "r"(e[3]));
      asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(s4) : "r"(s01), "r"(s23));
      float f0, f1;
      asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %2; cvt.f32.f16 %0, lo; cvt.f32.f16 %1, hi;}" : "=f"(f0), "=f"(f1) : "r"(s4));
      q[0] = ffma2(pack_f32x2(f0, f1), a2, q[0]);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = x[i] * 0.999f - 0.01f;   // loop-carried (8 FFMA)
    }
    if (S == S_SOFTMAX_NEW) {
      // proposed: 4 ex2, 2 F2FP, 1 max.bf16x2 per 4 elements (shift folded into the MMA, row sum from a ones column)
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int i = 4 * g;
        const float p0 = fast_exp2(x[i]), p1 = fast_exp2(x[i + 1]), p2 = fast_exp2(x[i + 2]), p3 = fast_exp2(x[i + 3]);
        h[2 * g] = pack_bf16(p0, p1);
        h[2 * g + 1] = pack_bf16(p2, p3);
        asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(h[4]) : "r"(h[2 * g]));
        asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(h[5]) : "r"(h[2 * g + 1]));
        x[i] = p0 - 1.5f; x[i + 1] = p1 - 1.5f; x[i + 2] = p2 - 1.5f; x[i + 3] = p3 - 1.5f;
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
  float l0, l1;
#pragma unroll
  for (int i = 0; i < 8; ++i) { unpack_f32x2(q[i], l0, l1); s += x[i] + y[i] + w[i] + l0 + l1 + __uint_as_float(h[i]); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) atomicMax((unsigned long long*)cyc, (unsigned long long)(t1 - t0));
}

static const char* NAMES[S_COUNT] = {"8 MUFU.EX2", "8 FFMA", "8 FFMA2", "8 FADD2", "8 FMNMX3 (+8 FMUL)", "8 F2FP.BF16", "8 HMNMX2.BF16",
                                     "8 SHL+IADD (LEA)", "8 MUFU + 8 FFMA", "8 MUFU + 16 FFMA", "8 MUFU + 8 FFMA2", "8 MUFU + 8 F2FP",
                                     "8 MUFU + 8 FMNMX3 (+8 FMUL)", "8 MUFU + 32 FFMA", "softmax body today (8 elems, +8 FADD)",
                                     "softmax body proposed (8 elems, +8 FADD)", "8 MUFU.EX2.F16x2 (16 exps)", "8 MUFU.EX2.BF16x2 (16 exps)", "8 cvt.rn.f16x2.f32", "8 HADD2.F16", "f16x2 softmax body (8 elems, +8 FFMA)"};

template <int S>
static void run(float* o, long long* c) {
  const int iters = 4096;
  for (int warps_per_smsp = 1; warps_per_smsp <= 4; warps_per_smsp *= 2) {
    cudaMemset(c, 0, 8);
    k<S><<<148, 128 * warps_per_smsp>>>(o, iters, c);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %d warp(s)/SMSP: %7.2f cycles per iteration per SMSP (= %6.2f per warp)%s\n", NAMES[S], warps_per_smsp,
           (double)h / iters, (double)h / iters / warps_per_smsp, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
}

int main() {
  float* o; long long* c;
  cudaMalloc(&o, 148 * 512 * 4); cudaMalloc(&c, 8);
  run<S_MUFU>(o, c); run<S_FFMA>(o, c); run<S_FFMA2>(o, c); run<S_FADD2>(o, c); run<S_FMNMX3>(o, c); run<S_F2FP>(o, c);
  run<S_HMNMX2>(o, c); run<S_LEA>(o, c); run<S_MUFU_FFMA>(o, c); run<S_MUFU_2FFMA>(o, c); run<S_MUFU_4FFMA>(o, c);
  run<S_MUFU_FFMA2>(o, c); run<S_MUFU_F2FP>(o, c); run<S_MUFU_FMNMX3>(o, c); run<S_SOFTMAX_NOW>(o, c); run<S_SOFTMAX_NEW>(o, c);
  run<S_EX2_H2>(o, c); run<S_EX2_BF2>(o, c); run<S_CVT_H2>(o, c); run<S_HADD2>(o, c); run<S_EX2H2_BODY>(o, c);
  return 0;
}
