// Whisper log-mel front-end on the GPU: framing (reflect pad, Hann), 400-point real FFT, power, slaney mel projection,
// log10 / clamp, per-clip max floor and affine normalisation.
// Replaces WhisperFeatureExtractor._torch_extract_fbank_features ($TF/models/whisper/feature_extraction_whisper.py:135-164)
// as called at desta/models/modeling_desta25.py:1570 and desta/trainer/data/simple_dataset.py:239-243.
//
// Kernel 1 (logmel_fft_kernel): one CTA = 32 consecutive frames of one clip (8 warps x 4 frames).  The 5360 samples the
//   tile touches are staged once in shared memory with coalesced loads (frames overlap 60 %; 16 floats of padding per
//   hop keep the four frames of a warp on different banks).  Each 400-point real FFT runs as a 200-point complex FFT
//   held ENTIRELY IN REGISTERS by 8 threads: thread t owns z[8m + t] (25 values), does the 25-point DFT (radix 5 x 5,
//   compile-time twiddles) in registers, and the radix-8 step across the 8 threads is three shuffle butterfly stages
//   (decimation in frequency, outputs land bit-reversed over the lanes).  The real-input post-pass fetches the mirror
//   bin with one more shuffle, the power spectrum goes through a small conflict-free shared buffer to the sparse
//   triangular mel filters (one lane = one mel bin for the warp's 4 frames -> 16-byte stores), log10, and the clip
//   maximum is reduced warp -> block -> one atomicMax.
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

constexpr int MEL_HOP_PAD = MEL_HOP + 16;                       // padded hop in the staged wave (bank spreading)
constexpr int MEL_WAVE_PAD = MEL_WAVE + 16 * (MEL_WAVE / MEL_HOP + 1);
constexpr int MEL_PSTRIDE = 216;                                // per-frame stride of the power buffer (conflict-free)

// W25^(b c) = exp(-2 pi i b c / 25): compile-time twiddles of the in-register 25-point DFT
__device__ constexpr float W25C[5][5] = {
    {1.000000000f, 1.000000000f, 1.000000000f, 1.000000000f, 1.000000000f},
    {1.000000000f, 0.968583161f, 0.876306680f, 0.728968627f, 0.535826795f},
    {1.000000000f, 0.876306680f, 0.535826795f, 0.062790520f, -0.425779292f},
    {1.000000000f, 0.728968627f, 0.062790520f, -0.637423990f, -0.992114701f},
    {1.000000000f, 0.535826795f, -0.425779292f, -0.992114701f, -0.637423990f}};
__device__ constexpr float W25S[5][5] = {
    {0.f, 0.f, 0.f, 0.f, 0.f},
    {0.f, -0.248689887f, -0.481753674f, -0.684547106f, -0.844327926f},
    {0.f, -0.481753674f, -0.844327926f, -0.998026728f, -0.904827052f},
    {0.f, -0.684547106f, -0.998026728f, -0.770513243f, -0.125333234f},
    {0.f, -0.844327926f, -0.904827052f, -0.125333234f, 0.770513243f}};

__global__ void __launch_bounds__(MEL_WARPS * 32, 2)
logmel_fft_kernel(const float* __restrict__ wave, const int32_t* __restrict__ n_valid, const float* __restrict__ plan,
                  float* __restrict__ out, unsigned* __restrict__ clip_max) {
  extern __shared__ __align__(16) float sm[];
  const int* hdr = reinterpret_cast<const int*>(plan);
  const int n_mels = hdr[0], max_w = hdr[1];
  const int tab_words = (int)plan_words(n_mels, max_w) - PLAN_HDR;
  float* s_tab = sm;                                            // plan tables
  float* s_wave = s_tab + ((tab_words + 3) & ~3);               // [MEL_WAVE_PAD] padded staged samples
  float* s_pow = s_wave + MEL_WAVE_PAD;                         // [warps][4 frames][MEL_PSTRIDE] power spectra
  __shared__ float s_wmax[MEL_WARPS];

  const float* s_hann = s_tab;
  const float2* s_tw200 = reinterpret_cast<const float2*>(s_tab + 400);
  const float2* s_tw400 = reinterpret_cast<const float2*>(s_tab + 800);
  const int* s_lo = reinterpret_cast<const int*>(s_tab + 1202) + 201;
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
      s_wave[j + 16 * (j / MEL_HOP)] = (s < nv) ? __ldg(w + s) : 0.0f;
    }
  }
  __syncthreads();

  const int f = lane >> 3, t = lane & 7;                 // frame within the warp, thread within the frame's octet
  const int fl = warp * MEL_FPW + f;                     // local frame
  const int kap = ((t & 1) << 2) | (t & 2) | (t >> 2);   // bit-reversed octet index: this thread ends up with bins [25 kap, 25 kap + 25)
  float2 z[25];
  // ---- load x_t[m] = z[8m + t] = (x[16m + 2t], x[16m + 2t + 1]) * hann ----
  {
    const float* xb = s_wave + fl * MEL_HOP_PAD + 2 * t;
    const float* hb = s_hann + 2 * t;
#pragma unroll
    for (int m = 0; m < 25; ++m) {
      const float2 xv = *reinterpret_cast<const float2*>(xb + 16 * m + 16 * (m / 10));
      const float2 hv = *reinterpret_cast<const float2*>(hb + 16 * m);
      z[m] = make_float2(xv.x * hv.x, xv.y * hv.y);
    }
  }
  // ---- 25-point DFT in registers: m = 5a + b, k1 = c + 5e; result X[k1] in register 5c + e ----
#pragma unroll
  for (int b = 0; b < 5; ++b) {
    float2 v[5] = {z[b], z[5 + b], z[10 + b], z[15 + b], z[20 + b]};
    dft5(v);
#pragma unroll
    for (int c = 0; c < 5; ++c) z[5 * c + b] = (b == 0 || c == 0) ? v[c] : cmul(v[c], make_float2(W25C[b][c], W25S[b][c]));
  }
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    float2 v[5] = {z[5 * c], z[5 * c + 1], z[5 * c + 2], z[5 * c + 3], z[5 * c + 4]};
    dft5(v);
#pragma unroll
    for (int e = 0; e < 5; ++e) z[5 * c + e] = v[e];
  }
  // ---- twiddle W200^(t k1), then radix-8 across the octet: 3 shuffle butterfly stages (DIF) ----
  {
    const float r2 = 0.70710678118654752f;
    // stage-1 twiddle of the upper lanes: W8^(t & 3); stage-2: W4^(t & 1)
    const int q = t & 3;
    const float2 w1 = (t & 4) ? (q == 0 ? make_float2(1.f, 0.f) : q == 1 ? make_float2(r2, -r2) : q == 2 ? make_float2(0.f, -1.f) : make_float2(-r2, -r2))
                              : make_float2(1.f, 0.f);
    const float2 w2 = ((t & 2) && (t & 1)) ? make_float2(0.f, -1.f) : make_float2(1.f, 0.f);
    // butterfly as  d = o + sgn * v  (lower lane: v + o, upper lane: o - v), then * w (w = 1 on lower lanes)
    const float sg4 = (t & 4) ? -1.f : 1.f, sg2 = (t & 2) ? -1.f : 1.f, sg1 = (t & 1) ? -1.f : 1.f;
#pragma unroll
    for (int r = 0; r < 25; ++r) {
      const int k1 = (r / 5) + 5 * (r % 5);
      float2 v = (k1 == 0) ? z[r] : cmul(z[r], s_tw200[t * k1]);
      float2 o, d;
      o.x = __shfl_xor_sync(0xffffffffu, v.x, 4); o.y = __shfl_xor_sync(0xffffffffu, v.y, 4);
      d = make_float2(fmaf(sg4, v.x, o.x), fmaf(sg4, v.y, o.y));
      v = cmul(d, w1);
      o.x = __shfl_xor_sync(0xffffffffu, v.x, 2); o.y = __shfl_xor_sync(0xffffffffu, v.y, 2);
      d = make_float2(fmaf(sg2, v.x, o.x), fmaf(sg2, v.y, o.y));
      v = cmul(d, w2);
      o.x = __shfl_xor_sync(0xffffffffu, v.x, 1); o.y = __shfl_xor_sync(0xffffffffu, v.y, 1);
      z[r] = make_float2(fmaf(sg1, v.x, o.x), fmaf(sg1, v.y, o.y));
    }
  }
  // now z[5c + e] = Z[25 kap + c + 5e]
  // ---- real-input post-pass: X[k] = E + W400^k (-i O) with the mirror bin Z[200 - k] fetched from lane t ^ 7 ----
  float* pw = s_pow + (warp * MEL_FPW + f) * MEL_PSTRIDE;
  {
    const int kap0 = (8 - kap) & 7;                                       // holder of Z[200 - 25 kap] (k1 = 0 case)
    const int src0 = (lane & ~7) | (((kap0 & 1) << 2) | (kap0 & 2) | (kap0 >> 2));
    const float2* tw = s_tw400 + 25 * kap;
#pragma unroll
    for (int r = 0; r < 25; ++r) {
      const int c = r / 5, e = r % 5, k1 = c + 5 * e;
      float2 B;
      if (k1 == 0) {
        B.x = __shfl_sync(0xffffffffu, z[0].x, src0); B.y = __shfl_sync(0xffffffffu, z[0].y, src0);
      } else {
        const int rp = (c >= 1) ? 5 * (5 - c) + (4 - e) : 5 - e;          // register of bin 25 - k1 in the mirror lane
        B.x = __shfl_xor_sync(0xffffffffu, z[rp].x, 7); B.y = __shfl_xor_sync(0xffffffffu, z[rp].y, 7);
      }
      const float2 A = z[r];
      const float2 E = make_float2(0.5f * (A.x + B.x), 0.5f * (A.y - B.y));
      const float2 O = make_float2(0.5f * (A.x - B.x), 0.5f * (A.y + B.y));
      const float2 T = cmul(tw[k1], mul_negi(O));
      const float2 X = cadd(E, T);
      pw[25 * kap + k1] = X.x * X.x + X.y * X.y;
    }
    if (t == 0) { const float d = z[0].x - z[0].y; pw[200] = d * d; }     // bin 200 (kap == 0 holds Z[0])
  }
  __syncwarp();
  // ---- sparse mel projection + log10: one lane = one mel bin for the warp's 4 frames -> one 16-byte store ----
  float lmax = -INFINITY;
  const int tw0 = t0 + warp * MEL_FPW;                                   // first frame of this warp (multiple of 4)
  const bool frames_ok = tw0 < DESTA_MEL_FRAMES;                         // 3000 % 4 == 0: all four valid or none
  const float* p0 = s_pow + warp * MEL_FPW * MEL_PSTRIDE;
  float* o = out + (size_t)clip * n_mels * DESTA_MEL_FRAMES;
  for (int m = lane; m < n_mels; m += 32) {
    const int lo = s_lo[m], cnt = s_cnt[m];
    const float* wv = s_w + m * max_w;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int i = 0; i < cnt; ++i) {
      const float wi = wv[i];
      const float* pp = p0 + lo + i;
      a0 = fmaf(wi, pp[0], a0);
      a1 = fmaf(wi, pp[MEL_PSTRIDE], a1);
      a2 = fmaf(wi, pp[2 * MEL_PSTRIDE], a2);
      a3 = fmaf(wi, pp[3 * MEL_PSTRIDE], a3);
    }
    const float k = 0.30102999566398120f;
    float4 L;
    L.x = k * __log2f(fmaxf(a0, 1e-10f)); L.y = k * __log2f(fmaxf(a1, 1e-10f));
    L.z = k * __log2f(fmaxf(a2, 1e-10f)); L.w = k * __log2f(fmaxf(a3, 1e-10f));
    if (frames_ok) {
      lmax = fmaxf(lmax, fmaxf(fmaxf(L.x, L.y), fmaxf(L.z, L.w)));
      *reinterpret_cast<float4*>(o + (size_t)m * DESTA_MEL_FRAMES + tw0) = L;
    }
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, off));
  if (lane == 0) s_wmax[warp] = lmax;
  __syncthreads();
  if (tid == 0) {
    float mx = s_wmax[0];
    for (int i = 1; i < MEL_WARPS; ++i) mx = fmaxf(mx, s_wmax[i]);
    if (mx > -INFINITY) atomicMax(clip_max + clip, order_f32(mx));
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
  for (int tl = warp; tl < MEL_TILE; tl += 8) {          // one warp per frame row, lanes over mel pairs: 4-byte stores
    if (t0 + tl >= DESTA_MEL_FRAMES) break;
    __nv_bfloat162* row = reinterpret_cast<__nv_bfloat162*>(om + (size_t)(t0 + tl) * n_mels);
    for (int m2 = lane; 2 * m2 < n_mels; m2 += 32)
      row[m2] = __floats2bfloat162_rn(s_t[(2 * m2) * (MEL_TILE + 1) + tl], s_t[(2 * m2 + 1) * (MEL_TILE + 1) + tl]);
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
  const size_t smem1 = (tab + MEL_WAVE_PAD + MEL_WARPS * MEL_FPW * MEL_PSTRIDE) * 4;
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
