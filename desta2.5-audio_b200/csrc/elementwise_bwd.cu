// Memory-bound pieces of the connector backward (autograd of LayerNorm / GELU / bias / layer mix in
// $TF/models/bert/modeling_bert.py:294-298, 339-356 and desta/models/modeling_desta25.py:600-606).
#include "common.cuh"
#include "desta_b200.h"

namespace desta {


// ---- LayerNorm backward.  One warp per row (row kept in registers), statistics recomputed from the saved input.
//      dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat));  dgamma += sum_rows dy*xhat;  dbeta += sum_rows dy.
//      Per-warp partial dgamma/dbeta live in registers across the rows a warp owns, then shared -> global atomics. ----
template <int MAXV>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ x, int ldx,
                     const float* __restrict__ gamma, float eps, int M, int D, float* __restrict__ dx_f32, int ld_f32,
                     __nv_bfloat16* __restrict__ dx_bf16, int ld_bf16, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ float s_red[];   // [2][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int nv = D >> 2;
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  float4 ag[MAXV], ab[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) { ag[i] = make_float4(0, 0, 0, 0); ab[i] = make_float4(0, 0, 0, 0); }
  for (int row = blockIdx.x * nwarp + warp; row < M; row += gridDim.x * nwarp) {
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * ldx);
    const float4* dr = reinterpret_cast<const float4*>(dy + (size_t)row * lddy);
    float4 xv[MAXV], dv[MAXV];
    // all 2 MAXV 16-byte loads of the row are issued before the first use (one memory round trip per row; with the sum in
    // the same loop the compiler kept load -> use pairs and every pair paid the full DRAM latency: ncu long_scoreboard 6.8)
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + 32 * i;
      xv[i] = idx < nv ? __ldcs(xr + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + 32 * i;
      dv[i] = idx < nv ? __ldcs(dr + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)D;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv) {
        const float a = xv[i].x - mean, b = xv[i].y - mean, c = xv[i].z - mean, d = xv[i].w - mean;
        ss += (a * a + b * b) + (c * c + d * d);
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss / (float)D + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + idx);
        // xhat into xv, g*dy into a temp; accumulate dgamma / dbeta partials
        xv[i].x = (xv[i].x - mean) * rstd; xv[i].y = (xv[i].y - mean) * rstd;
        xv[i].z = (xv[i].z - mean) * rstd; xv[i].w = (xv[i].w - mean) * rstd;
        ag[i].x += dv[i].x * xv[i].x; ag[i].y += dv[i].y * xv[i].y; ag[i].z += dv[i].z * xv[i].z; ag[i].w += dv[i].w * xv[i].w;
        ab[i].x += dv[i].x; ab[i].y += dv[i].y; ab[i].z += dv[i].z; ab[i].w += dv[i].w;
        dv[i].x *= g.x; dv[i].y *= g.y; dv[i].z *= g.z; dv[i].w *= g.w;
        s1 += (dv[i].x + dv[i].y) + (dv[i].z + dv[i].w);
        s2 += (dv[i].x * xv[i].x + dv[i].y * xv[i].y) + (dv[i].z * xv[i].z + dv[i].w * xv[i].w);
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    const float m1 = s1 / (float)D, m2 = s2 / (float)D;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv) {
        float4 r;
        r.x = rstd * (dv[i].x - m1 - xv[i].x * m2); r.y = rstd * (dv[i].y - m1 - xv[i].y * m2);
        r.z = rstd * (dv[i].z - m1 - xv[i].z * m2); r.w = rstd * (dv[i].w - m1 - xv[i].w * m2);
        if (dx_f32) reinterpret_cast<float4*>(dx_f32 + (size_t)row * ld_f32)[idx] = r;
        if (dx_bf16) {
          uint2 u; u.x = pack_bf16(r.x, r.y); u.y = pack_bf16(r.z, r.w);
          reinterpret_cast<uint2*>(dx_bf16 + (size_t)row * ld_bf16)[idx] = u;
        }
      }
    }
  }
  // warps add their partials to the block's copy one after the other (plain 16-byte read-modify-writes: shared-memory
  // atomics on the same 2 D addresses from 8 warps cost more than the whole row loop)
  for (int w = 0; w < nwarp; ++w) {
    if (warp == w) {
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nv) {
          float4* pg = reinterpret_cast<float4*>(s_red) + idx;
          float4* pb = reinterpret_cast<float4*>(s_red + D) + idx;
          float4 a = *pg, b = *pb;
          a.x += ag[i].x; a.y += ag[i].y; a.z += ag[i].z; a.w += ag[i].w;
          b.x += ab[i].x; b.y += ab[i].y; b.z += ab[i].z; b.w += ab[i].w;
          *pg = a; *pb = b;
        }
      }
    }
    __syncthreads();
  }
  // one 16-byte reduction per four columns (red.global.add.v4.f32): with one CTA per SM every address sees <= 148 of them
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(s_red)[i], b = reinterpret_cast<const float4*>(s_red + D)[i];
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dgamma + 4 * i), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dbeta + 4 * i), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
  }
}

// ---- erf-GELU forward on a saved pre-activation, and its backward:  d/dh [0.5 h (1 + erf(h/sqrt2))] ----
__global__ void __launch_bounds__(256)
gelu_fwd_kernel(const uint4* __restrict__ h, uint4* __restrict__ y, size_t n8) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v = h[i];
    __nv_bfloat162* b = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = __bfloat1622float2(b[t]);
      b[t] = __floats2bfloat162_rn(gelu_erf(f.x), gelu_erf(f.y));
    }
    y[i] = v;
  }
}
__device__ __forceinline__ float gelu_grad(float h) {
  const float cdf = 0.5f * (1.0f + erff(h * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * h * h);
  return cdf + h * pdf;
}
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ h, uint4* __restrict__ dh, size_t n8) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    uint4 g = dy[i];
    const uint4 hv = h[i];
    __nv_bfloat162* gb = reinterpret_cast<__nv_bfloat162*>(&g);
    const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&hv);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 gf = __bfloat1622float2(gb[t]), hf = __bfloat1622float2(hb[t]);
      gb[t] = __floats2bfloat162_rn(gf.x * gelu_grad(hf.x), gf.y * gelu_grad(hf.y));
    }
    dh[i] = g;
  }
}

// ---- column sums (bias gradients): out[j] += sum_m x[m, j].  Block = 32 columns x 8 row-lanes, grid-stride over rows. ----
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, int ld, int M, int N, float* __restrict__ out, int rows_per_block) {
  __shared__ float s[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float acc = 0.f;
  if (col < N)
    for (int r = r0 + ty; r < r1; r += 8) acc += (float)x[(size_t)r * ld + col];
  s[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && col < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s[i][tx];
    atomicAdd(out + col, t);
  }
}

// Vectorised variant (16-byte loads, 4 rows in flight per thread): a warp covers 32 x VEC consecutive columns of a row,
// the 8 warps of a block stride over the rows; per-thread partial sums, one shared-memory reduction, fp32 atomics.
// Used whenever the rows are 16-byte aligned (the 5.9 GB dK/dV tensor of the K/V-projection bias gradient goes from
// 1.8 to > 5 TB/s).
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const T* __restrict__ x, int ld, int M, int N, float* __restrict__ out, int rows_per_block) {
  __shared__ float s[8][32 * VEC + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + lane) * VEC;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
  if (col < N) {
    const T* p = x + col;
    int r = r0 + warp;
    for (; r + 24 < r1; r += 32) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(p + (size_t)(r + 8 * u) * ld));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (sizeof(T) == 4) {
          const float* f = reinterpret_cast<const float*>(&v[u]);
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] += f[i];
        } else {
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[u]);
#pragma unroll
          for (int i = 0; i < VEC / 2; ++i) { const float2 f = __bfloat1622float2(h[i]); acc[2 * i] += f.x; acc[2 * i + 1] += f.y; }
        }
      }
    }
    for (; r < r1; r += 8) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(p + (size_t)r * ld));
      if (sizeof(T) == 4) {
        const float* f = reinterpret_cast<const float*>(&v);
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] += f[i];
      } else {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < VEC / 2; ++i) { const float2 f = __bfloat1622float2(h[i]); acc[2 * i] += f.x; acc[2 * i + 1] += f.y; }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) s[warp][lane * VEC + i] = acc[i];
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * VEC; c += 256) {
    const int gc = blockIdx.x * 32 * VEC + c;
    if (gc < N) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += s[w][c];
      atomicAdd(out + gc, t);
    }
  }
}

// ---- layer mix backward: dx[k,r,:] = w[q,k] * dm[r,:];  dlw[q,k] = w[q,k] * (g[q,k] - sum_k' w[q,k'] g[q,k']),
//      g[q,k] = sum_{n,d} dm[n,q,d] * x[k,n,q,d].  Grid (Q, clip chunks): the softmax backward is linear in g, so every CTA
//      turns its partial g into a partial dlw and adds it atomically.  16-byte loads / stores. ----
template <int TAPS>
__global__ void __launch_bounds__(256)
layer_mix_bwd_kernel(const float* __restrict__ dm, const float* __restrict__ x, const float* __restrict__ lw, int taps_rt, int N,
                     int Q, int D, int clips_per_cta, float* __restrict__ dx, float* __restrict__ dlw) {
  __shared__ float s_w[16], s_g[16][8];
  const int taps = TAPS ? TAPS : taps_rt;
  const int q = blockIdx.x;
  const int n0 = blockIdx.y * clips_per_cta, n1 = min(N, n0 + clips_per_cta);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    float mx = -INFINITY, den = 0.f;
    for (int k = 0; k < taps; ++k) mx = fmaxf(mx, lw[q * taps + k]);
    for (int k = 0; k < taps; ++k) den += expf(lw[q * taps + k] - mx);
    for (int k = 0; k < taps; ++k) s_w[k] = expf(lw[q * taps + k] - mx) / den;
  }
  __syncthreads();
  float g[TAPS ? TAPS : 16], wk[TAPS ? TAPS : 16];
#pragma unroll
  for (int k = 0; k < (TAPS ? TAPS : 16); ++k) { g[k] = 0.f; wk[k] = k < taps ? s_w[k] : 0.f; }
  const size_t tap_stride4 = (size_t)N * Q * D / 4;
  const int D4 = D >> 2, total = (n1 - n0) * D4;
  const float4* dm4 = reinterpret_cast<const float4*>(dm);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float4* dx4 = reinterpret_cast<float4*>(dx);
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int n = n0 + i / D4, d = i - (i / D4) * D4;
    const size_t at = ((size_t)n * Q + q) * D4 + d;
    const float4 v = dm4[at];
#pragma unroll
    for (int k = 0; k < (TAPS ? TAPS : 16); ++k) {
      if (k < taps) {
        const float4 xv = x4[k * tap_stride4 + at];
        g[k] += (v.x * xv.x + v.y * xv.y) + (v.z * xv.z + v.w * xv.w);
        dx4[k * tap_stride4 + at] = make_float4(wk[k] * v.x, wk[k] * v.y, wk[k] * v.z, wk[k] * v.w);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < (TAPS ? TAPS : 16); ++k) {
    if (k < taps) {
      float t = g[k];
#pragma unroll
      for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) s_g[k][warp] = t;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float gk[16], dot = 0.f;
    for (int k = 0; k < taps; ++k) {
      float t = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_g[k][w];
      gk[k] = t;
      dot += s_w[k] * t;
    }
    for (int k = 0; k < taps; ++k) atomicAdd(dlw + q * taps + k, s_w[k] * (gk[k] - dot));
  }
}

__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float4* __restrict__ x, uint2* __restrict__ y, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    uint2 u; u.x = pack_bf16(v.x, v.y); u.y = pack_bf16(v.z, v.w);
    y[i] = u;
  }
}

static int ew_grid(size_t n) {
  size_t b = (n + 255) / 256;
  const size_t cap = (size_t)num_sms() * 16;
  return (int)(b > cap ? cap : (b ? b : 1));
}

}  // namespace desta

using namespace desta;

extern "C" int desta_layernorm_bwd(const float* dy, int lddy, const float* x, int ldx, const float* gamma, float eps, int M,
                                   int D, float* dx_f32, int ld_f32, void* dx_bf16, int ld_bf16, float* dgamma, float* dbeta,
                                   void* stream) {
  if (!dy || !x || !gamma || !dgamma || !dbeta || (!dx_f32 && !dx_bf16)) { set_error("desta_layernorm_bwd: null pointer"); return DESTA_ERR_ARG; }
  if (M <= 0) return DESTA_OK;
  if (D <= 0 || (D % 4) || D > 2048 || (lddy % 4) || (ldx % 4) || (dx_f32 && (ld_f32 % 4)) || (dx_bf16 && (ld_bf16 % 4))) {
    set_error("desta_layernorm_bwd: D must be a multiple of 4 and <= 2048 (D=%d)", D);
    return DESTA_ERR_ARG;
  }
  int grid = (M + 7) / 8;
  // one CTA per SM: the per-CTA dgamma / dbeta partials end in global reductions onto the same 2 D addresses, and
  // same-address reductions serialise in L2 (592 CTAs: 0.11 ms per call for 147 MB; one per SM: see profiles/r2_train_kernels.md)
  const int cap = num_sms();
  if (grid > cap) grid = cap;
  const size_t sm = 2 * (size_t)D * 4;
  cudaStream_t s = (cudaStream_t)stream;
  if (D <= 512)
    layernorm_bwd_kernel<4><<<grid, 256, sm, s>>>(dy, lddy, x, ldx, gamma, eps, M, D, dx_f32, ld_f32, (__nv_bfloat16*)dx_bf16, ld_bf16, dgamma, dbeta);
  else if (D <= 1280)
    layernorm_bwd_kernel<10><<<grid, 256, sm, s>>>(dy, lddy, x, ldx, gamma, eps, M, D, dx_f32, ld_f32, (__nv_bfloat16*)dx_bf16, ld_bf16, dgamma, dbeta);
  else
    layernorm_bwd_kernel<16><<<grid, 256, sm, s>>>(dy, lddy, x, ldx, gamma, eps, M, D, dx_f32, ld_f32, (__nv_bfloat16*)dx_bf16, ld_bf16, dgamma, dbeta);
  return check_launch("layernorm_bwd_kernel");
}

extern "C" int desta_gelu_fwd(const void* h, void* y, size_t n, void* stream) {
  if (!h || !y || (n % 8)) { set_error("desta_gelu_fwd: null pointer or n %% 8 != 0"); return DESTA_ERR_ARG; }
  if (!n) return DESTA_OK;
  gelu_fwd_kernel<<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>((const uint4*)h, (uint4*)y, n / 8);
  return check_launch("gelu_fwd_kernel");
}

extern "C" int desta_gelu_bwd(const void* dy, const void* h, void* dh, size_t n, void* stream) {
  if (!dy || !h || !dh || (n % 8)) { set_error("desta_gelu_bwd: null pointer or n %% 8 != 0"); return DESTA_ERR_ARG; }
  if (!n) return DESTA_OK;
  gelu_bwd_kernel<<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>((const uint4*)dy, (const uint4*)h, (uint4*)dh, n / 8);
  return check_launch("gelu_bwd_kernel");
}

extern "C" int desta_colsum(const void* x, int dtype, int ld, int M, int N, float* out, void* stream) {
  if (!x || !out) { set_error("desta_colsum: null pointer"); return DESTA_ERR_ARG; }
  if (M <= 0 || N <= 0) return DESTA_OK;
  const int vec = dtype == 0 ? 4 : 8, esz = dtype == 0 ? 4 : 2;
  if ((N % vec) == 0 && ((size_t)ld * esz) % 16 == 0 && ((uintptr_t)x & 15) == 0) {
    const int col_blocks = (N + 32 * vec - 1) / (32 * vec);
    int row_blocks = (num_sms() * 8 + col_blocks - 1) / col_blocks;
    int rows_per_block = (M + row_blocks - 1) / row_blocks;
    if (rows_per_block < 256) rows_per_block = 256;
    row_blocks = (M + rows_per_block - 1) / rows_per_block;
    if (row_blocks > 65535) { rows_per_block = (M + 65534) / 65535; row_blocks = (M + rows_per_block - 1) / rows_per_block; }
    const dim3 grid(col_blocks, row_blocks);
    if (dtype == 0) colsum_vec_kernel<float, 4><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, ld, M, N, out, rows_per_block);
    else colsum_vec_kernel<__nv_bfloat16, 8><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ld, M, N, out, rows_per_block);
    return check_launch("colsum_vec_kernel");
  }
  int row_blocks = (num_sms() * 8) / ((N + 31) / 32);
  if (row_blocks < 1) row_blocks = 1;
  int rows_per_block = (M + row_blocks - 1) / row_blocks;
  if (rows_per_block < 64) rows_per_block = 64;
  row_blocks = (M + rows_per_block - 1) / rows_per_block;
  if (row_blocks > 65535) { rows_per_block = (M + 65534) / 65535; row_blocks = (M + rows_per_block - 1) / rows_per_block; }
  const dim3 grid((N + 31) / 32, row_blocks);
  if (dtype == 0) colsum_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, ld, M, N, out, rows_per_block);
  else colsum_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ld, M, N, out, rows_per_block);
  return check_launch("colsum_kernel");
}

// out[h*64 + c] += sum over (b, t) of w[(b*heads + h)*Tq + t] * x[(b*Tq + t)*ld + h*64 + c]:  the value-bias gradient of an
// attention from dO (see desta_colsum_heads in the header).  grid (heads, row chunks), 256 threads = 8 column octets x 32
// row lanes; one 16-byte load per thread and row, octet partials folded through shared memory, 64 atomics per block.
__global__ void __launch_bounds__(256) colsum_heads_kernel(const __nv_bfloat16* __restrict__ x, int ld, const float* __restrict__ w,
                                                           int rows, int heads, int Tq, float* __restrict__ out, int rows_per_block) {
  __shared__ float part[32][65];
  const int h = blockIdx.x, oct = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int r = r0 + rl; r < r1; r += 32) {
    const uint4 u = *reinterpret_cast<const uint4*>(x + (size_t)r * ld + h * 64 + oct * 8);
    const int b = r / Tq, t = r - b * Tq;
    const float wt = w ? __ldg(w + ((size_t)b * heads + h) * Tq + t) : 1.0f;
    const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(v2[i]);
      acc[2 * i] = fmaf(wt, f.x, acc[2 * i]);
      acc[2 * i + 1] = fmaf(wt, f.y, acc[2 * i + 1]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) part[rl][oct * 8 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 64) {
    float sum = 0.f;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) sum += part[i][threadIdx.x];
    atomicAdd(out + h * 64 + threadIdx.x, sum);
  }
}

extern "C" int desta_colsum_heads(const void* x, int ld, const float* w, int batch, int heads, int Tq, float* out, void* stream) {
  if (!x || !out) { set_error("desta_colsum_heads: null pointer"); return DESTA_ERR_ARG; }
  if (batch <= 0 || heads <= 0 || Tq <= 0) return DESTA_OK;
  if ((ld % 8) || ld < heads * 64 || ((uintptr_t)x & 15)) { set_error("desta_colsum_heads: ld must be a multiple of 8 and >= heads*64, x 16-byte aligned"); return DESTA_ERR_ARG; }
  if (heads > 65535) { set_error("desta_colsum_heads: too many heads"); return DESTA_ERR_ARG; }
  const long long rows = (long long)batch * Tq;
  if (rows > 0x7fffffffLL) { set_error("desta_colsum_heads: too many rows"); return DESTA_ERR_ARG; }
  int row_blocks = (num_sms() * 4 + heads - 1) / heads;
  int rows_per_block = (int)((rows + row_blocks - 1) / row_blocks);
  if (rows_per_block < 128) rows_per_block = 128;
  row_blocks = (int)((rows + rows_per_block - 1) / rows_per_block);
  if (row_blocks > 65535) { rows_per_block = (int)((rows + 65534) / 65535); row_blocks = (int)((rows + rows_per_block - 1) / rows_per_block); }
  colsum_heads_kernel<<<dim3(heads, row_blocks), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ld, w, (int)rows, heads, Tq, out, rows_per_block);
  return check_launch("colsum_heads_kernel");
}

extern "C" int desta_layer_mix_bwd(const float* dm, const float* x, const float* layer_weights, int taps, int N, int Q, int D,
                                   float* dx, float* dlw, void* stream) {
  if (!dm || !x || !layer_weights || !dx || !dlw) { set_error("desta_layer_mix_bwd: null pointer"); return DESTA_ERR_ARG; }
  if (taps > 16) { set_error("desta_layer_mix_bwd: at most 16 taps"); return DESTA_ERR_ARG; }
  if (N <= 0) return DESTA_OK;
  if (D % 4) { set_error("desta_layer_mix_bwd: D must be a multiple of 4"); return DESTA_ERR_ARG; }
  // enough (query, clip chunk) CTAs for ~4 per SM
  int chunks = (4 * num_sms() + Q - 1) / Q;
  if (chunks > N) chunks = N;
  const int per = (N + chunks - 1) / chunks;
  const dim3 grid(Q, (N + per - 1) / per);
  if (taps == 4) layer_mix_bwd_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(dm, x, layer_weights, taps, N, Q, D, per, dx, dlw);
  else layer_mix_bwd_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(dm, x, layer_weights, taps, N, Q, D, per, dx, dlw);
  return check_launch("layer_mix_bwd_kernel");
}

extern "C" int desta_cast_f32_bf16(const float* x, void* y, size_t n, void* stream) {
  if (!x || !y || (n % 4)) { set_error("desta_cast_f32_bf16: null pointer or n %% 4 != 0"); return DESTA_ERR_ARG; }
  if (!n) return DESTA_OK;
  cast_f32_bf16_kernel<<<ew_grid(n / 4), 256, 0, (cudaStream_t)stream>>>((const float4*)x, (uint2*)y, n / 4);
  return check_launch("cast_f32_bf16_kernel");
}

// ---- train-mode dropout of the Q-Former's dense outputs (BertSelfOutput / BertOutput: LayerNorm(dropout(dense) + x),
//      $TF/models/bert/modeling_bert.py:294-298, 352-356).  One thread per aligned group of 8 elements = one Philox call;
//      the same kernel serves the forward (resid != NULL) and the backward of the dropped branch (resid == NULL). ----
namespace desta {

__global__ void __launch_bounds__(256)
dropout_add_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ resid, int ldr, int M, int ngrp,
                   DropoutCfg dc, float* __restrict__ out_f32, int ld_f32, __nv_bfloat16* __restrict__ out_bf16, int ld_bf16) {
  const size_t total = (size_t)M * ngrp;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(g / ngrp), col = (int)(g % ngrp) * 8;
    const float4* xp = reinterpret_cast<const float4*>(x + (size_t)row * ldx + col);
    float4 a = xp[0], b = xp[1];
    const uint32_t keep = dc.thr16 ? dropout_keep8(dc, g) : 0xFFu;
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = ((keep >> i) & 1) ? v[i] * dc.scale : 0.f;
    if (resid) {
      const float4* rp = reinterpret_cast<const float4*>(resid + (size_t)row * ldr + col);
      const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
      v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
    }
    if (out_f32) {
      float4* op = reinterpret_cast<float4*>(out_f32 + (size_t)row * ld_f32 + col);
      op[0] = make_float4(v[0], v[1], v[2], v[3]);
      op[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (out_bf16)
      *reinterpret_cast<uint4*>(out_bf16 + (size_t)row * ld_bf16 + col) =
          make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
}

__global__ void __launch_bounds__(256)
dropout_mask_kernel(DropoutCfg dc, int rows, int cols, int ngrp, uint8_t* __restrict__ mask) {
  const size_t total = (size_t)rows * ngrp;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(g / ngrp), col = (int)(g % ngrp) * 8;
    const uint32_t keep = dc.thr16 ? dropout_keep8(dc, g) : 0xFFu;
    for (int i = 0; i < 8 && col + i < cols; ++i) mask[(size_t)row * cols + col + i] = (keep >> i) & 1;
  }
}

}  // namespace desta

extern "C" int desta_dropout_add(const float* x, int ldx, const float* resid, int ld_resid, int M, int N,
                                 const desta_dropout* drop, float* out_f32, int ld_f32, void* out_bf16, int ld_bf16,
                                 void* stream) {
  if (!x || (!out_f32 && !out_bf16)) { set_error("desta_dropout_add: null pointer"); return DESTA_ERR_ARG; }
  if (M <= 0 || N <= 0) return DESTA_OK;
  if ((N % 8) || (ldx % 4) || (resid && (ld_resid % 4)) || (out_f32 && (ld_f32 % 4)) || (out_bf16 && (ld_bf16 % 8))) {
    set_error("desta_dropout_add: N must be a multiple of 8 and leading dims 16-byte aligned (N=%d)", N);
    return DESTA_ERR_ARG;
  }
  const size_t total = (size_t)M * (N / 8);
  dropout_add_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(x, ldx, resid, ld_resid, M, N / 8, make_dropout_cfg(drop),
                                                                      out_f32, ld_f32, (__nv_bfloat16*)out_bf16, ld_bf16);
  return check_launch("dropout_add_kernel");
}

extern "C" int desta_dropout_mask(const desta_dropout* drop, int rows, int cols, uint8_t* mask, void* stream) {
  if (!mask) { set_error("desta_dropout_mask: null pointer"); return DESTA_ERR_ARG; }
  if (rows <= 0 || cols <= 0) return DESTA_OK;
  const int ngrp = (cols + 7) / 8;
  dropout_mask_kernel<<<ew_grid((size_t)rows * ngrp), 256, 0, (cudaStream_t)stream>>>(make_dropout_cfg(drop), rows, cols, ngrp, mask);
  return check_launch("dropout_mask_kernel");
}
