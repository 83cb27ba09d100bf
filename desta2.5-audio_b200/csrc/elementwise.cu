// Memory-bound helpers around the tensor-core kernels: LayerNorm, conv-stem im2col, Q-Former layer mix.
// One warp per row, 128-bit coalesced loads/stores; LayerNorm runs persistent warps with a one-row prefetch.
#include "common.cuh"
#include "desta_b200.h"

namespace desta {

// ---- LayerNorm: fp32 row in registers (D <= 32*4*MAXV), two-pass mean / variance like ATen's layer_norm. ----
template <int MAXV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float eps, int M, int D, __nv_bfloat16* __restrict__ out_bf16, int ld_bf16, float* __restrict__ out_f32,
                 int ld_f32) {
  // Persistent warps: one row at a time per warp, held in registers; the loads of the warp's NEXT row are issued before the
  // statistics of the current one, so every warp always has a full row of HBM reads in flight.
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= M) return;
  const int nv = D >> 2;                     // float4 per row
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  float4 v[MAXV], nx[MAXV];
  {
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * ldx);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv) nx[i] = xr[idx];
    }
  }
  for (; row < M; row += nwarps) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) v[i] = nx[i];
    const int next = row + nwarps;
    if (next < M) {
      const float4* xr = reinterpret_cast<const float4*>(x + (size_t)next * ldx);
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nv) nx[i] = xr[idx];
      }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)D;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        ss += (a * a + b * b) + (c * c + d * d);
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss / (float)D + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv) {
        const float4 g = __ldg(g4 + idx), b = __ldg(b4 + idx);
        float4 y;
        y.x = (v[i].x - mean) * rstd * g.x + b.x;
        y.y = (v[i].y - mean) * rstd * g.y + b.y;
        y.z = (v[i].z - mean) * rstd * g.z + b.z;
        y.w = (v[i].w - mean) * rstd * g.w + b.w;
        if (out_f32) reinterpret_cast<float4*>(out_f32 + (size_t)row * ld_f32)[idx] = y;
        if (out_bf16) {
          uint2 u;
          u.x = pack_bf16(y.x, y.y);
          u.y = pack_bf16(y.z, y.w);
          reinterpret_cast<uint2*>(out_bf16 + (size_t)row * ld_bf16)[idx] = u;
        }
      }
    }
  }
}

// ---- im2col for Conv1d(k=3, pad=1, stride s) on time-major bf16 input. One thread moves 8 channels (16 B). ----
__global__ void __launch_bounds__(256)
im2col_k3_kernel(const uint4* __restrict__ x, int N, int T_in, int C8, int stride, int T_out, uint4* __restrict__ cols,
                 const unsigned* __restrict__ clip_max) {
  const size_t total = (size_t)N * T_out * 3 * C8;
  int floor_n = -1;              // clip whose floor is held in `fl` (recomputed only when the grid-stride loop crosses a clip)
  uint32_t fl = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    const size_t r = i / C8;
    const int tap = (int)(r % 3);
    const size_t row = r / 3;
    const int t_out = (int)(row % T_out);
    const int n = (int)(row / T_out);
    const int t_in = t_out * stride + tap - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (t_in >= 0 && t_in < T_in) {
      v = x[((size_t)n * T_in + t_in) * C8 + c];
      if (clip_max) {
        // deferred log-mel floor (desta_logmel with out_f32 == NULL): max(L, Lmax - 8) in the affine / bf16 domain
        if (n != floor_n) {
          const unsigned u = __ldg(clip_max + n);
          const float lmax = __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
          fl = pack_bf16(((lmax - 8.0f) + 4.0f) * 0.25f, ((lmax - 8.0f) + 4.0f) * 0.25f);
          floor_n = n;
        }
        asm("max.bf16x2 %0, %0, %1;" : "+r"(v.x) : "r"(fl));
        asm("max.bf16x2 %0, %0, %1;" : "+r"(v.y) : "r"(fl));
        asm("max.bf16x2 %0, %0, %1;" : "+r"(v.z) : "r"(fl));
        asm("max.bf16x2 %0, %0, %1;" : "+r"(v.w) : "r"(fl));
      }
    }
    cols[i] = v;
  }
}

// ---- layer mix: softmax over taps of layer_weights[q,:], weighted sum of the taps' Q-Former outputs. ----
__global__ void __launch_bounds__(256)
layer_mix_kernel(const float* __restrict__ x, const float* __restrict__ lw, int taps, int rows, int Q, int D4,
                 float* __restrict__ out) {
  const size_t total = (size_t)rows * D4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / D4);
    const int q = row % Q;
    float mx = -INFINITY;
    for (int k = 0; k < taps; ++k) mx = fmaxf(mx, __ldg(lw + q * taps + k));
    float den = 0.f;
    for (int k = 0; k < taps; ++k) den += expf(__ldg(lw + q * taps + k) - mx);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < taps; ++k) {
      const float w = expf(__ldg(lw + q * taps + k) - mx) / den;
      const float4 v = reinterpret_cast<const float4*>(x)[(size_t)k * total + i];
      acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
    }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
}


}  // namespace desta

using namespace desta;

extern "C" int desta_layernorm(const float* x, int ldx, const float* gamma, const float* beta, float eps, int M, int D,
                               void* out_bf16, int ld_bf16, float* out_f32, int ld_f32, void* stream) {
  if (!x || !gamma || !beta || (!out_bf16 && !out_f32)) { set_error("desta_layernorm: null pointer"); return DESTA_ERR_ARG; }
  if (M <= 0) return DESTA_OK;
  if (D <= 0 || (D % 4) || D > 32 * 4 * 16 || (ldx % 4) || (out_bf16 && (ld_bf16 % 4)) || (out_f32 && (ld_f32 % 4))) {
    set_error("desta_layernorm: D must be a multiple of 4 and <= 2048, leading dims multiples of 4 (D=%d)", D);
    return DESTA_ERR_ARG;
  }
  const int threads = 256, rows_per_block = threads / 32;
  int grid = (M + rows_per_block - 1) / rows_per_block;
  const int cap = num_sms() * 2;             // persistent: two 8-warp CTAs per SM (128 registers), each warp walks rows with a one-row prefetch
  if (grid > cap) grid = cap;
  cudaStream_t s = (cudaStream_t)stream;
  if (D <= 512)
    layernorm_kernel<4><<<grid, threads, 0, s>>>(x, ldx, gamma, beta, eps, M, D, (__nv_bfloat16*)out_bf16, ld_bf16, out_f32, ld_f32);
  else if (D <= 1280)
    layernorm_kernel<10><<<grid, threads, 0, s>>>(x, ldx, gamma, beta, eps, M, D, (__nv_bfloat16*)out_bf16, ld_bf16, out_f32, ld_f32);
  else
    layernorm_kernel<16><<<grid, threads, 0, s>>>(x, ldx, gamma, beta, eps, M, D, (__nv_bfloat16*)out_bf16, ld_bf16, out_f32, ld_f32);
  return check_launch("layernorm_kernel");
}

extern "C" int desta_im2col_k3(const void* x, int N, int T_in, int C, int stride, void* cols, const float* mel_clip_max, void* stream) {
  if (!x || !cols) { set_error("desta_im2col_k3: null pointer"); return DESTA_ERR_ARG; }
  if (N <= 0) return DESTA_OK;
  if ((C % 8) || (stride != 1 && stride != 2) || (T_in % stride)) {
    set_error("desta_im2col_k3: C must be a multiple of 8, stride 1 or 2 (C=%d stride=%d T_in=%d)", C, stride, T_in);
    return DESTA_ERR_ARG;
  }
  const int T_out = T_in / stride;
  const size_t total = (size_t)N * T_out * 3 * (C / 8);
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  im2col_k3_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, N, T_in, C / 8, stride, T_out, (uint4*)cols,
                                                                        (const unsigned*)mel_clip_max);
  return check_launch("im2col_k3_kernel");
}

extern "C" int desta_layer_mix(const float* x, const float* layer_weights, int taps, int N, int Q, int D, float* out,
                               void* stream) {
  if (!x || !layer_weights || !out) { set_error("desta_layer_mix: null pointer"); return DESTA_ERR_ARG; }
  if (N <= 0) return DESTA_OK;
  if (D % 4) { set_error("desta_layer_mix: D must be a multiple of 4"); return DESTA_ERR_ARG; }
  const size_t total = (size_t)N * Q * (D / 4);
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  layer_mix_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, layer_weights, taps, N * Q, Q, D / 4, out);
  return check_launch("layer_mix_kernel");
}

// ---- per-row content hash (scope row n2, see desta_b200.h) ----
namespace desta {
__device__ __forceinline__ uint64_t mix64(uint32_t w, uint64_t i) {
  uint64_t z = ((uint64_t)w | (i << 32)) + 0x9E3779B97F4A7C15ull * (i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
constexpr int RH_CHUNK = 8192;      // 32-bit words per block
__global__ void __launch_bounds__(256)
row_hash_kernel(const uint8_t* __restrict__ rows, int64_t words, int64_t stride_bytes, unsigned long long* __restrict__ out) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(rows + (size_t)blockIdx.y * stride_bytes);
  const int64_t i0 = (int64_t)blockIdx.x * RH_CHUNK;
  const int64_t i1 = i0 + RH_CHUNK < words ? i0 + RH_CHUNK : words;
  uint64_t acc = 0;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) acc += mix64(__ldg(r + i), (uint64_t)i);
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ uint64_t part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t t = 0;
    for (int w = 0; w < 8; ++w) t += part[w];
    atomicAdd(out + blockIdx.y, (unsigned long long)t);
  }
}
}  // namespace desta

extern "C" int desta_row_hash(const void* rows, int64_t n_rows, int64_t row_bytes, int64_t row_stride_bytes, uint64_t* out,
                              void* stream) {
  using namespace desta;
  if (n_rows <= 0) return DESTA_OK;
  if (!rows || !out || row_bytes <= 0 || (row_bytes & 3) || (row_stride_bytes & 3) || ((uintptr_t)rows & 3) || n_rows > 65535) {
    set_error("desta_row_hash: rows must be 4-byte aligned multiples of 4 bytes, n_rows <= 65535");
    return DESTA_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(out, 0, (size_t)n_rows * 8, s) != cudaSuccess) return check_launch("desta_row_hash memset");
  const int64_t words = row_bytes / 4;
  const dim3 grid((unsigned)((words + RH_CHUNK - 1) / RH_CHUNK), (unsigned)n_rows);
  row_hash_kernel<<<grid, 256, 0, s>>>((const uint8_t*)rows, words, row_stride_bytes, (unsigned long long*)out);
  return check_launch("row_hash_kernel");
}
