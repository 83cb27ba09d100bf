// Whisper log-mel front-end on the GPU: framing (reflect pad, Hann), 400-point real FFT, power, slaney mel projection,
// log10 / clamp, per-clip max floor and affine normalisation.
// Replaces WhisperFeatureExtractor._torch_extract_fbank_features ($TF/models/whisper/feature_extraction_whisper.py:135-164)
// as called at desta/models/modeling_desta25.py:1570 and desta/trainer/data/simple_dataset.py:239-243.
//
// Kernel 1 (logmel_fft_kernel): one CTA = 32 consecutive frames of one clip (8 warps x 4 frames).  The 5360 samples the
//   tile touches are staged once in shared memory with coalesced loads (frames overlap 60 %), each warp runs four
//   400-point real FFTs at a time in fp32 as a 200-point complex FFT (radix 8 x 5 x 5, in-place in a per-warp scratch)
//   plus the real-input post-pass, applies the sparse triangular mel filters, takes log10 and stores the raw tile
//   [mel][32 frames] with 128-byte coalesced rows; the clip maximum is reduced warp -> block -> one atomicMax.
// Kernel 2 (logmel_finish_kernel): max(L, clip_max - 8), (x + 4) / 4 in place (reads that hit L2 right after kernel 1),
//   optional bf16 time-major copy [N, 3000, n_mels] that feeds the conv-stem GEMM.
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "desta_b200.h"

namespace desta {

constexpr int MEL_NFFT = 400, MEL_HOP = 160, MEL_NFREQ = 201;
constexpr int MEL_TILE = 32;                 // frames per CTA
constexpr int MEL_WARPS = 8, MEL_FPW = 4;    // frames per warp pass
constexpr int MEL_WAVE = MEL_TILE * MEL_HOP + (MEL_NFFT - MEL_HOP);  // 5360 staged samples
constexpr int PLAN_HDR = 16;

struct MelPlanView {
  int n_mels, max_w;
  const float* hann;    // [400]
  const float2* tw200;  // [200] e^{-2 pi i j / 200}
  const float2* tw400;  // [201] e^{-2 pi i k / 400}
  const int* lut;       // [201] float index of P[k] inside the per-frame scratch
  const int* lo;        // [n_mels]
  const int* cnt;       // [n_mels]
  const float* w;       // [n_mels * max_w]
};

__host__ __device__ inline size_t plan_words(int n_mels, int max_w) {
  return PLAN_HDR + 400 + 400 + 402 + 201 + 2 * (size_t)n_mels + (size_t)n_mels * max_w;
}

// ---- host: slaney mel filter bank in fp64 ($TF/audio_utils.py:263-333, 356-376, 453-546), cast to fp32 ----
static double hz_to_mel(double f) { return f >= 1000.0 ? 15.0 + log(f / 1000.0) * (27.0 / log(6.4)) : 3.0 * f / 200.0; }
static double mel_to_hz(double m) { return m >= 15.0 ? 1000.0 * exp((log(6.4) / 27.0) * (m - 15.0)) : 200.0 * m / 3.0; }

static void build_filters(int n_mels, double* fb /*[201][n_mels]*/) {
  const double mel_min = hz_to_mel(0.0), mel_max = hz_to_mel(8000.0);
  double* hz = new double[n_mels + 2];
  for (int i = 0; i < n_mels + 2; ++i) hz[i] = mel_to_hz(mel_min + (mel_max - mel_min) * i / (n_mels + 1));
  for (int k = 0; k < MEL_NFREQ; ++k) {
    const double f = 8000.0 * k / (MEL_NFREQ - 1);
    for (int m = 0; m < n_mels; ++m) {
      const double down = -(hz[m] - f) / (hz[m + 1] - hz[m]);
      const double up = (hz[m + 2] - f) / (hz[m + 2] - hz[m + 1]);
      double v = down < up ? down : up;
      if (v < 0.0) v = 0.0;
      fb[k * n_mels + m] = v * (2.0 / (hz[m + 2] - hz[m]));
    }
  }
  delete[] hz;
}

static int filters_max_width(int n_mels, const double* fb) {
  int mw = 1;
  for (int m = 0; m < n_mels; ++m) {
    int lo = -1, hi = -1;
    for (int k = 0; k < MEL_NFREQ; ++k)
      if ((float)fb[k * n_mels + m] != 0.0f) { if (lo < 0) lo = k; hi = k; }
    if (lo >= 0 && hi - lo + 1 > mw) mw = hi - lo + 1;
  }
  return mw;
}

#ifdef __CUDACC__
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_negi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)
__device__ __forceinline__ float2 mul_posi(float2 a) { return make_float2(-a.y, a.x); }   // a * (+i)

__device__ __forceinline__ void dft4(float2 c0, float2 c1, float2 c2, float2 c3, float2& x0, float2& x1, float2& x2, float2& x3) {
  const float2 s02 = cadd(c0, c2), d02 = csub(c0, c2), s13 = cadd(c1, c3), d13 = csub(c1, c3);
  x0 = cadd(s02, s13);
  x2 = csub(s02, s13);
  x1 = cadd(d02, mul_negi(d13));
  x3 = cadd(d02, mul_posi(d13));
}
__device__ __forceinline__ void dft8(float2 (&a)[8]) {
  const float r = 0.70710678118654752f;
  const float2 b0 = cadd(a[0], a[4]), b1 = cadd(a[1], a[5]), b2 = cadd(a[2], a[6]), b3 = cadd(a[3], a[7]);
  float2 b4 = csub(a[0], a[4]), b5 = csub(a[1], a[5]), b6 = csub(a[2], a[6]), b7 = csub(a[3], a[7]);
  b5 = make_float2(r * (b5.x + b5.y), r * (b5.y - b5.x));      // * (1 - i)/sqrt2
  b6 = mul_negi(b6);                                           // * (-i)
  b7 = make_float2(r * (b7.y - b7.x), -r * (b7.x + b7.y));     // * (-1 - i)/sqrt2
  dft4(b0, b1, b2, b3, a[0], a[2], a[4], a[6]);
  dft4(b4, b5, b6, b7, a[1], a[3], a[5], a[7]);
}
__device__ __forceinline__ void dft5(float2 (&x)[5]) {
  const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;
  const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
  const float2 p1 = cadd(x[1], x[4]), p2 = cadd(x[2], x[3]), d1 = csub(x[1], x[4]), d2 = csub(x[2], x[3]);
  const float2 a1 = make_float2(x[0].x + c1 * p1.x + c2 * p2.x, x[0].y + c1 * p1.y + c2 * p2.y);
  const float2 a2 = make_float2(x[0].x + c2 * p1.x + c1 * p2.x, x[0].y + c2 * p1.y + c1 * p2.y);
  const float2 b1 = make_float2(s1 * d1.x + s2 * d2.x, s1 * d1.y + s2 * d2.y);
  const float2 b2 = make_float2(s2 * d1.x - s1 * d2.x, s2 * d1.y - s1 * d2.y);
  x[0] = cadd(x[0], cadd(p1, p2));
  x[1] = make_float2(a1.x + b1.y, a1.y - b1.x);   // a1 - i b1
  x[4] = make_float2(a1.x - b1.y, a1.y + b1.x);   // a1 + i b1
  x[2] = make_float2(a2.x + b2.y, a2.y - b2.x);
  x[3] = make_float2(a2.x - b2.y, a2.y + b2.x);
}

__device__ __forceinline__ unsigned order_f32(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unorder_f32(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

__global__ void __launch_bounds__(MEL_WARPS * 32)
logmel_fft_kernel(const float* __restrict__ wave, const int32_t* __restrict__ n_valid, const float* __restrict__ plan,
                  float* __restrict__ out, unsigned* __restrict__ clip_max) {
  extern __shared__ __align__(16) float sm[];
  const int* hdr = reinterpret_cast<const int*>(plan);
  const int n_mels = hdr[0], max_w = hdr[1];
  const int tab_words = (int)plan_words(n_mels, max_w) - PLAN_HDR;
  float* s_tab = sm;                                 // plan tables
  float* s_wave = s_tab + ((tab_words + 3) & ~3);    // [MEL_WAVE]
  float* s_scr = s_wave + MEL_WAVE;                  // [warps][4 frames][400]
  float* s_tile = s_scr + MEL_WARPS * MEL_FPW * 400; // [n_mels][33]
  __shared__ float s_wmax[MEL_WARPS];

  const float* s_hann = s_tab;
  const float2* s_tw200 = reinterpret_cast<const float2*>(s_tab + 400);
  const float2* s_tw400 = reinterpret_cast<const float2*>(s_tab + 800);
  const int* s_lut = reinterpret_cast<const int*>(s_tab + 1202);
  const int* s_lo = s_lut + 201;
  const int* s_cnt = s_lo + n_mels;
  const float* s_w = reinterpret_cast<const float*>(s_cnt + n_mels);

  const int clip = blockIdx.y, t0 = blockIdx.x * MEL_TILE;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < tab_words; i += blockDim.x) s_tab[i] = plan[PLAN_HDR + i];
  {
    const float* w = wave + (size_t)clip * DESTA_MEL_SAMPLES;
    const int nv = n_valid ? n_valid[clip] : DESTA_MEL_SAMPLES;
    const int s0 = t0 * MEL_HOP - MEL_NFFT / 2;
    for (int j = tid; j < MEL_WAVE; j += blockDim.x) {
      int s = s0 + j;
      if (s < 0) s = -s;                                              // reflect (torch.stft center=True)
      if (s >= DESTA_MEL_SAMPLES) s = 2 * (DESTA_MEL_SAMPLES - 1) - s;
      s_wave[j] = (s < nv) ? __ldg(w + s) : 0.0f;
    }
  }
  __syncthreads();

  float2* scr = reinterpret_cast<float2*>(s_scr + warp * MEL_FPW * 400);   // 4 frames x 200 complex
  const int fbase = warp * MEL_FPW;                                         // first local frame of this warp
  float lmax = -INFINITY;

  // ---- step 1: radix-8 over n1 (inputs z[25 n1 + n2]), twiddle W200^(n2 k1), store Y[k1][n2] ----
  for (int it = lane; it < MEL_FPW * 25; it += 32) {
    const int j = it / 25, n2 = it - j * 25;
    const float* x = s_wave + (fbase + j) * MEL_HOP;
    float2 a[8];
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) {
      const int idx = 50 * n1 + 2 * n2;
      const float2 xv = *reinterpret_cast<const float2*>(x + idx);
      const float2 hv = *reinterpret_cast<const float2*>(s_hann + idx);
      a[n1] = make_float2(xv.x * hv.x, xv.y * hv.y);
    }
    dft8(a);
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) scr[j * 200 + k1 * 25 + n2] = cmul(a[k1], s_tw200[n2 * k1]);
  }
  __syncwarp();
  // ---- step 2a: radix-5 over a (inputs [k1][5a+b]), twiddle W25^(b c), store [k1][5c+b] ----
  for (int it = lane; it < MEL_FPW * 40; it += 32) {
    const int j = it / 40, r = it - j * 40, k1 = r / 5, b = r - k1 * 5;
    float2* p = scr + j * 200 + k1 * 25 + b;
    float2 x[5];
#pragma unroll
    for (int a = 0; a < 5; ++a) x[a] = p[5 * a];
    dft5(x);
#pragma unroll
    for (int c = 0; c < 5; ++c) p[5 * c] = cmul(x[c], s_tw200[8 * b * c]);
  }
  __syncwarp();
  // ---- step 2b: radix-5 over b (inputs [k1][5c+b]) -> Z[k1 + 8(c + 5e)] stored at [k1][5c+e] ----
  for (int it = lane; it < MEL_FPW * 40; it += 32) {
    const int j = it / 40, r = it - j * 40, k1 = r / 5, c = r - k1 * 5;
    float2* p = scr + j * 200 + k1 * 25 + 5 * c;
    float2 x[5];
#pragma unroll
    for (int b = 0; b < 5; ++b) x[b] = p[b];
    dft5(x);
#pragma unroll
    for (int e = 0; e < 5; ++e) p[e] = x[e];
  }
  __syncwarp();
  // ---- real-input post-pass: X[k] = E + T, X[200-k] = conj(E - T); power written back in place via the LUT ----
  {
    float* fs = reinterpret_cast<float*>(scr);
    for (int it = lane; it < MEL_FPW * 101; it += 32) {
      const int j = it / 101, k = it - j * 101;
      float* f = fs + j * 400;
      const int ia = s_lut[k], ib = (k == 0) ? s_lut[0] : s_lut[200 - k];
      const float2 A = *reinterpret_cast<const float2*>(f + ia);
      const float2 B = *reinterpret_cast<const float2*>(f + ib);
      const float2 E = make_float2(0.5f * (A.x + B.x), 0.5f * (A.y - B.y));
      const float2 O = make_float2(0.5f * (A.x - B.x), 0.5f * (A.y + B.y));
      const float2 T = cmul(s_tw400[k], mul_negi(O));
      const float2 Xk = cadd(E, T), Xm = csub(E, T);
      const float pk = Xk.x * Xk.x + Xk.y * Xk.y, pm = Xm.x * Xm.x + Xm.y * Xm.y;
      // every item reads and then overwrites only its own two complex slots: no cross-lane hazard
      f[ia] = pk;
      if (k == 0) f[ia + 1] = pm;           // P[200]
      else if (k != 100) f[ib] = pm;        // P[200-k]
    }
  }
  __syncwarp();
  // ---- sparse mel projection + log10 ----
  {
    const float* fs = reinterpret_cast<const float*>(scr);
    for (int it = lane; it < MEL_FPW * n_mels; it += 32) {
      const int j = it / n_mels, m = it - j * n_mels;
      const float* f = fs + j * 400;
      const int lo = s_lo[m], cnt = s_cnt[m];
      const float* wv = s_w + m * max_w;
      float acc = 0.f;
      for (int i = 0; i < cnt; ++i) acc = fmaf(wv[i], f[s_lut[lo + i]], acc);
      const float L = 0.30102999566398120f * __log2f(fmaxf(acc, 1e-10f));
      const int tl = fbase + j;
      s_tile[m * (MEL_TILE + 1) + tl] = L;
      if (t0 + tl < DESTA_MEL_FRAMES) lmax = fmaxf(lmax, L);
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if (lane == 0) s_wmax[warp] = lmax;
  __syncthreads();
  if (tid == 0) {
    float m = s_wmax[0];
    for (int i = 1; i < MEL_WARPS; ++i) m = fmaxf(m, s_wmax[i]);
    atomicMax(clip_max + clip, order_f32(m));
  }
  // raw tile -> out[clip][m][t0 .. t0+31]: each warp writes whole 128-byte rows
  float* o = out + (size_t)clip * n_mels * DESTA_MEL_FRAMES;
  for (int m = warp; m < n_mels; m += MEL_WARPS) {
    const int t = t0 + lane;
    if (t < DESTA_MEL_FRAMES) o[(size_t)m * DESTA_MEL_FRAMES + t] = s_tile[m * (MEL_TILE + 1) + lane];
  }
}

// in-place floor + affine; optional bf16 time-major copy (transposed through shared memory).
__global__ void __launch_bounds__(256)
logmel_finish_kernel(float* __restrict__ out, const unsigned* __restrict__ clip_max, int n_mels,
                     __nv_bfloat16* __restrict__ out_tm) {
  extern __shared__ __align__(16) float s_t[];   // [n_mels][33]
  const int clip = blockIdx.y, t0 = blockIdx.x * MEL_TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float floor_v = unorder_f32(clip_max[clip]) - 8.0f;
  float* o = out + (size_t)clip * n_mels * DESTA_MEL_FRAMES;
  const int t = t0 + lane;
  for (int m = warp; m < n_mels; m += 8) {
    float v = 0.f;
    if (t < DESTA_MEL_FRAMES) {
      float* p = o + (size_t)m * DESTA_MEL_FRAMES + t;
      v = (fmaxf(*p, floor_v) + 4.0f) * 0.25f;
      *p = v;
    }
    s_t[m * (MEL_TILE + 1) + lane] = v;
  }
  if (!out_tm) return;
  __syncthreads();
  __nv_bfloat16* om = out_tm + (size_t)clip * DESTA_MEL_FRAMES * n_mels;
  for (int i = threadIdx.x; i < MEL_TILE * n_mels; i += blockDim.x) {
    const int tl = i / n_mels, m = i - tl * n_mels;
    if (t0 + tl < DESTA_MEL_FRAMES) om[(size_t)(t0 + tl) * n_mels + m] = __float2bfloat16_rn(s_t[m * (MEL_TILE + 1) + tl]);
  }
}
#endif

}  // namespace desta

using namespace desta;

extern "C" size_t desta_logmel_plan_bytes(int n_mels) {
  if (n_mels <= 0 || n_mels > 256) return 0;
  double* fb = new double[MEL_NFREQ * n_mels];
  build_filters(n_mels, fb);
  const int mw = filters_max_width(n_mels, fb);
  delete[] fb;
  return plan_words(n_mels, mw) * 4;
}

extern "C" int desta_logmel_plan_build(int n_mels, void* host_buf, size_t bytes) {
  const size_t need = desta_logmel_plan_bytes(n_mels);
  if (!need || !host_buf || bytes < need) { set_error("desta_logmel_plan_build: bad n_mels %d or buffer too small", n_mels); return DESTA_ERR_ARG; }
  double* fb = new double[MEL_NFREQ * n_mels];
  build_filters(n_mels, fb);
  const int mw = filters_max_width(n_mels, fb);
  memset(host_buf, 0, need);
  int* hdr = (int*)host_buf;
  hdr[0] = n_mels; hdr[1] = mw; hdr[2] = (int)(need / 4);
  float* p = (float*)host_buf + PLAN_HDR;
  const double PI = 3.14159265358979323846;
  for (int n = 0; n < 400; ++n) p[n] = (float)(0.5 - 0.5 * cos(2.0 * PI * n / 400.0));   // periodic Hann
  p += 400;
  for (int j = 0; j < 200; ++j) { p[2 * j] = (float)cos(2.0 * PI * j / 200.0); p[2 * j + 1] = (float)(-sin(2.0 * PI * j / 200.0)); }
  p += 400;
  for (int k = 0; k <= 200; ++k) { p[2 * k] = (float)cos(2.0 * PI * k / 400.0); p[2 * k + 1] = (float)(-sin(2.0 * PI * k / 400.0)); }
  p += 402;
  int* lut = (int*)p;
  for (int k = 0; k < 200; ++k) {
    const int k1 = k % 8, k2 = k / 8, c = k2 % 5, e = k2 / 5;
    lut[k] = 2 * (k1 * 25 + 5 * c + e);
  }
  lut[200] = lut[0] + 1;
  int* lo = lut + 201;
  int* cnt = lo + n_mels;
  float* w = (float*)(cnt + n_mels);
  for (int m = 0; m < n_mels; ++m) {
    int first = -1, last = -1;
    for (int k = 0; k < MEL_NFREQ; ++k)
      if ((float)fb[k * n_mels + m] != 0.0f) { if (first < 0) first = k; last = k; }
    lo[m] = first < 0 ? 0 : first;
    cnt[m] = first < 0 ? 0 : last - first + 1;
    for (int i = 0; i < cnt[m]; ++i) w[m * mw + i] = (float)fb[(first + i) * n_mels + m];
  }
  delete[] fb;
  return DESTA_OK;
}

extern "C" int desta_logmel(const float* wave, const int32_t* n_valid, int N, int n_mels, const void* plan, float* out_f32,
                            void* out_bf16_tm, float* clip_max, void* stream) {
  if (!wave || !plan || !out_f32 || !clip_max) { set_error("desta_logmel: null pointer"); return DESTA_ERR_ARG; }
  if (N <= 0) return DESTA_OK;
  const size_t pb = desta_logmel_plan_bytes(n_mels);
  if (!pb) { set_error("desta_logmel: bad n_mels %d", n_mels); return DESTA_ERR_ARG; }
  if (N > 65535) { set_error("desta_logmel: at most 65535 clips per call"); return DESTA_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  const size_t tab = ((pb / 4 - PLAN_HDR) + 3) & ~(size_t)3;
  const size_t smem1 = (tab + MEL_WAVE + MEL_WARPS * MEL_FPW * 400 + (size_t)n_mels * (MEL_TILE + 1)) * 4;
  static size_t attr1 = 0;
  if (smem1 > attr1) {
    cudaError_t e = cudaFuncSetAttribute(logmel_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
    if (e != cudaSuccess) { set_error("logmel smem attribute: %s", cudaGetErrorString(e)); return DESTA_ERR_CUDA; }
    attr1 = smem1;
  }
  cudaError_t e = cudaMemsetAsync(clip_max, 0, (size_t)N * 4, s);
  if (e != cudaSuccess) { set_error("desta_logmel memset: %s", cudaGetErrorString(e)); return DESTA_ERR_CUDA; }
  const dim3 grid((DESTA_MEL_FRAMES + MEL_TILE - 1) / MEL_TILE, N);
  logmel_fft_kernel<<<grid, MEL_WARPS * 32, smem1, s>>>(wave, n_valid, (const float*)plan, out_f32, (unsigned*)clip_max);
  int rc = check_launch("logmel_fft_kernel");
  if (rc) return rc;
  logmel_finish_kernel<<<grid, 256, (size_t)n_mels * (MEL_TILE + 1) * 4, s>>>(out_f32, (const unsigned*)clip_max, n_mels,
                                                                              (__nv_bfloat16*)out_bf16_tm);
  return check_launch("logmel_finish_kernel");
}
