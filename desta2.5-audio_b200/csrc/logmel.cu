// Whisper log-mel front-end on the GPU: framing (reflect pad, Hann), 400-point real FFT, power, slaney mel projection,
// log10 / clamp, per-clip max floor and affine normalisation.
// Replaces WhisperFeatureExtractor._torch_extract_fbank_features ($TF/models/whisper/feature_extraction_whisper.py:135-164)
// as called at desta/models/modeling_desta25.py:1570 and desta/trainer/data/simple_dataset.py:239-243.
//
// Kernel 1 (logmel_tile_kernel, persistent, 2 CTAs / SM): one tile = 32 consecutive frames of one clip, 320 threads = 10 per
//   frame.  The 5360 samples of a tile are staged once (16-byte loads; frames overlap 60 %), with 20 pad words per hop so
//   that every shared-memory access below is conflict-free.  Each 400-point real FFT is a 200-point complex FFT split
//   10 x 20 (four-step):
//     stage 1  thread t of the frame owns z[10 m + t], m < 20: window, 20-point DFT (4 x 5, compile-time twiddles) entirely in
//              registers, twiddle W200^(t k1), one 8-byte store per value into the exchange buffer [frame][k1][t];
//     stage 2  thread u owns the MIRROR PAIR of 10-point DFTs k1 = u and 20 - u (u = 0: k1 = 0 and 10): five 16-byte loads
//              each, 2 x 5 butterflies in registers; because bin k and bin 200 - k then sit in the same thread, the
//              real-input untangling X[k] = Xe + W400^k Xo, X[200-k] = conj(Xe - W400^k Xo) and the power spectrum need
//              no further exchange (the Hann table carries the factor 1/2 of Xe / Xo).
//   Compared with round 1 (200-point FFT held by 8 threads, radix-8 across lanes by 6 shuffles + 3 twiddle products per
//   value) the exchange costs one store + 0.5 loads per value and no shuffles.
//   Mel projection: lane = frame, warp = a cost-balanced range of mel bins; the sparse slaney weights are warp-uniform
//   (broadcast loads), the power rows have an odd stride.  log10, per-clip maximum (block reduce + one ordered atomicMax).
//   Outputs: fp32 [N, n_mels, 3000] (128-byte coalesced stores; floor / affine by kernel 2), or - product path - ONLY the
//   bf16 time-major tensor [N, 3000, n_mels] (staged through shared memory, a tile is one contiguous 8 KB run) holding
//   (L + 4) / 4 WITHOUT the per-clip floor: max(L, Lmax - 8) commutes with the monotone affine map and with the bf16
//   rounding, so the consumer (im2col of conv1, desta_im2col_k3) applies max(v, bf16((Lmax - 8 + 4) / 4)) on read and the
//   fp32 [N, 128, 3000] round trip of round 1 (write, re-read, re-write: 1.6 x the algorithmic traffic) disappears.
// Kernel 2 (logmel_finish_kernel, fp32 output only): max(L, clip_max - 8), (x + 4) / 4 in place, optional bf16 time-major copy.
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "desta_b200.h"

namespace desta {

constexpr int MEL_NFFT = 400, MEL_HOP = 160, MEL_NFREQ = 201;
constexpr int MEL_TILE = 32;                  // frames per tile (also the tile of the finish kernel)
constexpr int MT_TPF = 10;                    // threads per frame
constexpr int MT_THREADS = MEL_TILE * MT_TPF; // 320
constexpr int MT_WARPS = MT_THREADS / 32;     // 10
constexpr int MT_WAVE = (MEL_TILE - 1) * MEL_HOP + MEL_NFFT;          // 5360 staged samples
constexpr int MT_HOP_PAD = MEL_HOP + 20;      // staged hop: 180 = 20 (mod 32) -> lane address = 2 tid (mod 32)
constexpr int MT_WAVE_PAD = MT_WAVE + 20 * (MT_WAVE / MEL_HOP + 1);   // 6040
constexpr int MT_EX = 404;                    // floats per frame of the exchange buffer (200 complex + 4; 404 = 20 mod 32)
constexpr int MT_PS = 205;                    // floats per frame of the power buffer (odd stride; [201, 205) stay zero)
constexpr int MT_OUT_HW = 130;                // halfwords per frame of the bf16 output staging (65 words: odd)
constexpr int PLAN_HDR = 32;
constexpr int TILES_PER_CLIP = (DESTA_MEL_FRAMES + MEL_TILE - 1) / MEL_TILE;   // 94

// plan (device words after the header): hann/2 [400] | tw1 [20][16] float2 | tw2 [11][16] float2 | meta [n_mels] int2 |
// w4 [total4] float4      meta[m] = {first bin | groups << 16, index of the bin's first float4}: the sparse slaney taps of
// a mel bin, zero-padded to groups of four.  header: [0] n_mels, [1] total4, [2] words, [3..13] mel range of each of the 10 warps
__host__ __device__ inline size_t plan_words(int n_mels, int total4) {
  return PLAN_HDR + 400 + 640 + 352 + 2 * (size_t)n_mels + 4 * (size_t)total4;
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

#ifdef __CUDACC__
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_negi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)
__device__ __forceinline__ float2 mul_posi(float2 a) { return make_float2(-a.y, a.x); }   // a * (+i)

__device__ __forceinline__ void dft4(float2& c0, float2& c1, float2& c2, float2& c3) {
  const float2 s02 = cadd(c0, c2), d02 = csub(c0, c2), s13 = cadd(c1, c3), d13 = csub(c1, c3);
  c0 = cadd(s02, s13);
  c2 = csub(s02, s13);
  c1 = cadd(d02, mul_negi(d13));
  c3 = cadd(d02, mul_posi(d13));
}
__device__ __forceinline__ void dft5(float2& x0, float2& x1, float2& x2, float2& x3, float2& x4) {
  const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;
  const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
  const float2 p1 = cadd(x1, x4), p2 = cadd(x2, x3), d1 = csub(x1, x4), d2 = csub(x2, x3);
  const float2 a1 = make_float2(x0.x + c1 * p1.x + c2 * p2.x, x0.y + c1 * p1.y + c2 * p2.y);
  const float2 a2 = make_float2(x0.x + c2 * p1.x + c1 * p2.x, x0.y + c2 * p1.y + c1 * p2.y);
  const float2 b1 = make_float2(s1 * d1.x + s2 * d2.x, s1 * d1.y + s2 * d2.y);
  const float2 b2 = make_float2(s2 * d1.x - s1 * d2.x, s2 * d1.y - s1 * d2.y);
  x0 = cadd(x0, cadd(p1, p2));
  x1 = make_float2(a1.x + b1.y, a1.y - b1.x);   // a1 - i b1
  x4 = make_float2(a1.x - b1.y, a1.y + b1.x);   // a1 + i b1
  x2 = make_float2(a2.x + b2.y, a2.y - b2.x);
  x3 = make_float2(a2.x - b2.y, a2.y + b2.x);
}

__device__ __forceinline__ unsigned order_f32(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unorder_f32(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// W20^(b c) = exp(-2 pi i b c / 20) for b = 1..4, c = 1..3 (20-point DFT = 4 x 5) and W10^c, c = 1..4 (10-point = 2 x 5)
__device__ constexpr float W20C[5][4] = {{1.f, 1.f, 1.f, 1.f},
                                         {1.f, 0.951056516f, 0.809016994f, 0.587785252f},
                                         {1.f, 0.809016994f, 0.309016994f, -0.309016994f},
                                         {1.f, 0.587785252f, -0.309016994f, -0.951056516f},
                                         {1.f, 0.309016994f, -0.809016994f, -0.809016994f}};
__device__ constexpr float W20S[5][4] = {{0.f, 0.f, 0.f, 0.f},
                                         {0.f, -0.309016994f, -0.587785252f, -0.809016994f},
                                         {0.f, -0.587785252f, -0.951056516f, -0.951056516f},
                                         {0.f, -0.809016994f, -0.951056516f, -0.309016994f},
                                         {0.f, -0.951056516f, -0.587785252f, 0.587785252f}};
__device__ constexpr float W10C[5] = {1.f, 0.809016994f, 0.309016994f, -0.309016994f, -0.809016994f};
__device__ constexpr float W10S[5] = {0.f, -0.587785252f, -0.951056516f, -0.951056516f, -0.587785252f};

// 10-point DFT of v[0..9] (2 x 5): out[k2], k2 = c + 5 e
__device__ __forceinline__ void dft10(const float4 (&q)[5], float2 (&z)[10]) {
  // q[i] = (v[2i], v[2i+1])
  float2 e0 = make_float2(q[0].x, q[0].y), e1 = make_float2(q[1].x, q[1].y), e2 = make_float2(q[2].x, q[2].y),
         e3 = make_float2(q[3].x, q[3].y), e4 = make_float2(q[4].x, q[4].y);
  float2 o0 = make_float2(q[0].z, q[0].w), o1 = make_float2(q[1].z, q[1].w), o2 = make_float2(q[2].z, q[2].w),
         o3 = make_float2(q[3].z, q[3].w), o4 = make_float2(q[4].z, q[4].w);
  dft5(e0, e1, e2, e3, e4);
  dft5(o0, o1, o2, o3, o4);
  o1 = cmul(o1, make_float2(W10C[1], W10S[1]));
  o2 = cmul(o2, make_float2(W10C[2], W10S[2]));
  o3 = cmul(o3, make_float2(W10C[3], W10S[3]));
  o4 = cmul(o4, make_float2(W10C[4], W10S[4]));
  z[0] = cadd(e0, o0); z[5] = csub(e0, o0);
  z[1] = cadd(e1, o1); z[6] = csub(e1, o1);
  z[2] = cadd(e2, o2); z[7] = csub(e2, o2);
  z[3] = cadd(e3, o3); z[8] = csub(e3, o3);
  z[4] = cadd(e4, o4); z[9] = csub(e4, o4);
}

// real-input untangling of the bin pair (k, 200 - k): A = Z[k], B = Z[200 - k] (of the half-scaled input), w = W400^k
__device__ __forceinline__ void untangle(float2 A, float2 B, float2 w, float& pk, float& pmk) {
  const float2 xe = make_float2(A.x + B.x, A.y - B.y);
  const float2 xo = make_float2(A.y + B.y, B.x - A.x);
  const float2 t = cmul(w, xo);
  const float ax = xe.x + t.x, ay = xe.y + t.y, bx = xe.x - t.x, by = xe.y - t.y;
  pk = ax * ax + ay * ay;
  pmk = bx * bx + by * by;
}

__global__ void __launch_bounds__(MT_THREADS, 2)
logmel_tile_kernel(const float* __restrict__ wave, const int32_t* __restrict__ n_valid, const float* __restrict__ plan, int n_tiles,
                   float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_tm, unsigned* __restrict__ clip_max) {
  extern __shared__ __align__(16) float sm[];
  const int* hdr = reinterpret_cast<const int*>(plan);
  const int n_mels = hdr[0], total_w = hdr[1];
  const int tab_words = (int)plan_words(n_mels, total_w) - PLAN_HDR;
  float* s_tab = sm;
  float* s_wave = s_tab + ((tab_words + 3) & ~3);               // [MT_WAVE_PAD]
  float* s_ex = s_wave + MT_WAVE_PAD;                           // [32][MT_EX]; reused as bf16 output staging [32][MT_OUT_HW]
  float* s_p = s_ex + MEL_TILE * MT_EX;                         // [32][MT_PS]
  __shared__ float s_wmax[MT_WARPS];

  const float* s_hann = s_tab;
  const float2* s_tw1 = reinterpret_cast<const float2*>(s_tab + 400);
  const float2* s_tw2 = reinterpret_cast<const float2*>(s_tab + 1040);
  const int2* s_meta = reinterpret_cast<const int2*>(s_tab + 1392);
  const float4* s_w4 = reinterpret_cast<const float4*>(s_tab + 1392 + 2 * n_mels);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // tells the compiler that the warp index is warp-uniform
  for (int i = tid; i < tab_words; i += MT_THREADS) s_tab[i] = plan[PLAN_HDR + i];
  for (int i = tid; i < MEL_TILE * 4; i += MT_THREADS) s_p[(i >> 2) * MT_PS + MEL_NFREQ + (i & 3)] = 0.f;   // zero pads of the power rows
  const int m_begin = hdr[3 + warp], m_end = hdr[4 + warp];
  const int f = tid / MT_TPF, t = tid - MT_TPF * f;             // stage 1: frame of the tile, thread of the frame

  // Interior tiles are fetched with 16-byte cp.async into the padded layout; the fetch of the NEXT tile is issued as soon
  // as stage 1 has consumed the buffer and lands behind stage 2 and the mel projection.
  auto interior = [&](int tile) {
    const int clip = tile / TILES_PER_CLIP, t0 = (tile - clip * TILES_PER_CLIP) * MEL_TILE;
    const int nv = n_valid ? min(n_valid[clip], DESTA_MEL_SAMPLES) : DESTA_MEL_SAMPLES;
    const int s0 = t0 * MEL_HOP - MEL_NFFT / 2;
    return s0 >= 0 && s0 + MT_WAVE <= nv;
  };
  auto prefetch = [&](int tile) {
    const int clip = tile / TILES_PER_CLIP, t0 = (tile - clip * TILES_PER_CLIP) * MEL_TILE;
    const float* src = wave + (size_t)clip * DESTA_MEL_SAMPLES + (t0 * MEL_HOP - MEL_NFFT / 2);
    for (int i = tid; i < MT_WAVE / 4; i += MT_THREADS) {
      const int j = 4 * i;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s_wave + j + 20 * (j / MEL_HOP))), "l"(src + j) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  bool fetched = false;
  if ((int)blockIdx.x < n_tiles && interior(blockIdx.x)) { prefetch(blockIdx.x); fetched = true; }

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int clip = tile / TILES_PER_CLIP, t0 = (tile - clip * TILES_PER_CLIP) * MEL_TILE;
    // ---- the tile's samples (reflect padding of torch.stft(center=True); samples >= n_valid read as 0) ----
    if (fetched) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
      const float* w = wave + (size_t)clip * DESTA_MEL_SAMPLES;
      const int nv = n_valid ? min(n_valid[clip], DESTA_MEL_SAMPLES) : DESTA_MEL_SAMPLES;
      const int s0 = t0 * MEL_HOP - MEL_NFFT / 2;
      for (int j = tid; j < MT_WAVE; j += MT_THREADS) {
        int sidx = s0 + j;
        if (sidx < 0) sidx = -sidx;
        if (sidx >= DESTA_MEL_SAMPLES) sidx = 2 * (DESTA_MEL_SAMPLES - 1) - sidx;
        s_wave[j + 20 * (j / MEL_HOP)] = (sidx < nv) ? __ldg(w + sidx) : 0.0f;
      }
    }
    __syncthreads();
    // ---- stage 1: z[m] = x_t[m] = (x[20 m + 2 t], x[20 m + 2 t + 1]) * hann / 2;  20-point DFT over m = 5 a + b ----
    {
      float2 z[20];
      const float* xb = s_wave + MT_HOP_PAD * f + 2 * t;
      const float* hb = s_hann + 2 * t;
#pragma unroll
      for (int m = 0; m < 20; ++m) {
        const float2 xv = *reinterpret_cast<const float2*>(xb + 20 * m + 20 * (m / 8));
        const float2 hv = *reinterpret_cast<const float2*>(hb + 20 * m);
        z[m] = make_float2(xv.x * hv.x, xv.y * hv.y);
      }
#pragma unroll
      for (int b = 0; b < 5; ++b) {           // 4-point DFTs over a (stride 5), twiddle W20^(b c); U_b[c] stays at z[5 c + b]
        dft4(z[b], z[5 + b], z[10 + b], z[15 + b]);
        if (b) {
#pragma unroll
          for (int c = 1; c < 4; ++c) z[5 * c + b] = cmul(z[5 * c + b], make_float2(W20C[b][c], W20S[b][c]));
        }
      }
      float2* ex = reinterpret_cast<float2*>(s_ex) + f * (MT_EX / 2) + t;
#pragma unroll
      for (int c = 0; c < 4; ++c) {           // 5-point DFTs over b: Y[c + 4 e] lands in z[5 c + e]
        dft5(z[5 * c], z[5 * c + 1], z[5 * c + 2], z[5 * c + 3], z[5 * c + 4]);
#pragma unroll
        for (int e = 0; e < 5; ++e) {
          const int k1 = c + 4 * e;
          const float2 v = (k1 == 0) ? z[5 * c + e] : cmul(z[5 * c + e], s_tw1[k1 * 16 + t]);   // W200^(t k1)
          ex[k1 * MT_TPF] = v;
        }
      }
    }
    __syncthreads();
    {
      const int next = tile + gridDim.x;
      fetched = false;
      if (next < n_tiles && interior(next)) { prefetch(next); fetched = true; }
    }
    // ---- stage 2 (lane = frame, warp = u): 10-point DFTs over t for the mirror pair k1 = u, 20 - u (u = 0: 0 and 10);
    //      real-input untangling and power.  F[k2] = Z[k1a + 20 k2], S[k2] = Z[k1b + 20 k2]. ----
    {
      const int u = warp;
      const int k1a = u, k1b = u ? 20 - u : 10;
      float2 F[10], S[10];
      {
        float4 q[5];
        const float4* src = reinterpret_cast<const float4*>(s_ex + lane * MT_EX + 20 * k1a);
#pragma unroll
        for (int i = 0; i < 5; ++i) q[i] = src[i];
        dft10(q, F);
        src = reinterpret_cast<const float4*>(s_ex + lane * MT_EX + 20 * k1b);
#pragma unroll
        for (int i = 0; i < 5; ++i) q[i] = src[i];
        dft10(q, S);
      }
      float* pw = s_p + lane * MT_PS;
      if (u) {                                 // bin k = u + 20 i pairs F[i] with S[9 - i]  (200 - k = (20 - u) + 20 (9 - i))
#pragma unroll
        for (int i = 0; i < 10; ++i) {
          float pk, pmk;
          untangle(F[i], S[9 - i], s_tw2[i * 16 + u], pk, pmk);
          pw[u + 20 * i] = pk;
          pw[200 - u - 20 * i] = pmk;
        }
      } else {                                 // bins 20 i pair F[i] with F[(10 - i) % 10]; bins 10 + 20 i pair S[i] with S[9 - i]
#pragma unroll
        for (int i = 0; i <= 5; ++i) {
          float pk, pmk;
          untangle(F[i], F[(10 - i) % 10], s_tw2[i * 16], pk, pmk);
          pw[20 * i] = pk;
          pw[200 - 20 * i] = pmk;
        }
#pragma unroll
        for (int i = 0; i <= 4; ++i) {
          float pk, pmk;
          untangle(S[i], S[9 - i], s_tw2[(6 + i) * 16], pk, pmk);
          pw[10 + 20 * i] = pk;
          pw[190 - 20 * i] = pmk;
        }
      }
    }
    __syncthreads();
    // ---- sparse mel projection + log10: lane = frame, this warp's mel bins (taps zero-padded to groups of four) ----
    {
      const float* prow = s_p + lane * MT_PS;
      const int tf = t0 + lane;
      const bool fvalid = tf < DESTA_MEL_FRAMES;
      float lmax = -INFINITY;
      float* o32 = out_f32 ? out_f32 + (size_t)clip * n_mels * DESTA_MEL_FRAMES + tf : nullptr;
      uint32_t* so = reinterpret_cast<uint32_t*>(s_ex) + lane * (MT_OUT_HW / 2);
      // meta and taps are warp-uniform (broadcast loads); the group count is 1 for most bins (<= 4 taps), so the extra
      // groups hang off uniform branches instead of a loop
      auto mel_bin = [&](int m) {
        const int2 mt = s_meta[m];
        const float* pp = prow + (mt.x & 0xffff);
        const float4* w4 = s_w4 + mt.y;
        const int n4 = mt.x >> 16;
        float4 w = w4[0];
        float a0 = w.x * pp[0], a1 = w.y * pp[1], a2 = w.z * pp[2], a3 = w.w * pp[3];
        if (n4 > 1) {
          w = w4[1];
          a0 = fmaf(w.x, pp[4], a0); a1 = fmaf(w.y, pp[5], a1); a2 = fmaf(w.z, pp[6], a2); a3 = fmaf(w.w, pp[7], a3);
          if (n4 > 2) {
            w = w4[2];
            a0 = fmaf(w.x, pp[8], a0); a1 = fmaf(w.y, pp[9], a1); a2 = fmaf(w.z, pp[10], a2); a3 = fmaf(w.w, pp[11], a3);
            for (int g = 3; g < n4; ++g) {
              w = w4[g];
              a0 = fmaf(w.x, pp[4 * g], a0); a1 = fmaf(w.y, pp[4 * g + 1], a1);
              a2 = fmaf(w.z, pp[4 * g + 2], a2); a3 = fmaf(w.w, pp[4 * g + 3], a3);
            }
          }
        }
        return 0.30102999566398120f * __log2f(fmaxf((a0 + a1) + (a2 + a3), 1e-10f));
      };
      for (int m = m_begin; m < m_end; m += 2) {      // the ranges start and end on even bins
        const float L0 = mel_bin(m), L1 = mel_bin(m + 1);
        lmax = fmaxf(lmax, fmaxf(L0, L1));
        if (o32) {
          if (fvalid) { o32[(size_t)m * DESTA_MEL_FRAMES] = L0; o32[(size_t)(m + 1) * DESTA_MEL_FRAMES] = L1; }
        } else {
          so[m >> 1] = pack_bf16((L0 + 4.0f) * 0.25f, (L1 + 4.0f) * 0.25f);
        }
      }
      if (!fvalid) lmax = -INFINITY;
#pragma unroll
      for (int off = 16; off; off >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, off));
      if (lane == 0) s_wmax[warp] = lmax;
    }
    __syncthreads();
    if (tid == 0) {
      float mx = s_wmax[0];
#pragma unroll
      for (int i = 1; i < MT_WARPS; ++i) mx = fmaxf(mx, s_wmax[i]);
      if (mx > -INFINITY) atomicMax(clip_max + clip, order_f32(mx));
    }
    if (!out_f32) {
      // the tile's bf16 rows are one contiguous run of the time-major tensor: 4-byte words, 128 B per warp instruction
      const uint32_t* sw = reinterpret_cast<const uint32_t*>(s_ex);
      const int words = n_mels >> 1;
      uint32_t* ow = reinterpret_cast<uint32_t*>(out_tm + ((size_t)clip * DESTA_MEL_FRAMES + t0) * n_mels);
      const int valid = min(MEL_TILE, DESTA_MEL_FRAMES - t0) * words;
      for (int i = tid; i < valid; i += MT_THREADS) {
        const int fr = i / words, wd = i - fr * words;
        ow[i] = sw[fr * (MT_OUT_HW / 2) + wd];
      }
    }
    // the next tile's first barrier orders these reads before anything overwrites s_ex / s_wmax
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

// mel bins of a warp: prefix split of the per-bin cost (tap groups + fixed overhead) into MT_WARPS ranges with even ends
static void split_mels(int n_mels, const int* n4, int* split /*[MT_WARPS + 1]*/) {
  long total = 0;
  for (int m = 0; m < n_mels; ++m) total += 11 * n4[m] + 11;
  long acc = 0;
  int w = 1;
  split[0] = 0;
  for (int m = 0; m + 1 < n_mels && w < MT_WARPS; m += 2) {
    acc += 11 * (n4[m] + n4[m + 1]) + 22;
    while (w < MT_WARPS && acc * MT_WARPS >= total * w) split[w++] = m + 2;
  }
  for (; w <= MT_WARPS; ++w) split[w] = n_mels;
}

static void filter_support(int n_mels, const double* fb, int m, int* first, int* last) {
  *first = -1; *last = -1;
  for (int k = 0; k < MEL_NFREQ; ++k)
    if ((float)fb[k * n_mels + m] != 0.0f) { if (*first < 0) *first = k; *last = k; }
}

static int filters_total_groups(int n_mels, const double* fb) {
  int total = 0;
  for (int m = 0; m < n_mels; ++m) {
    int lo, hi;
    filter_support(n_mels, fb, m, &lo, &hi);
    if (lo >= 0) total += (hi - lo + 1 + 3) / 4;
  }
  return total;
}

extern "C" size_t desta_logmel_plan_bytes(int n_mels) {
  if (n_mels <= 0 || n_mels > 256 || (n_mels & 1)) return 0;
  double* fb = new double[MEL_NFREQ * n_mels];
  build_filters(n_mels, fb);
  const int t4 = filters_total_groups(n_mels, fb);
  delete[] fb;
  return plan_words(n_mels, t4) * 4;
}

extern "C" int desta_logmel_plan_build(int n_mels, void* host_buf, size_t bytes) {
  const size_t need = desta_logmel_plan_bytes(n_mels);
  if (!need || !host_buf || bytes < need) { set_error("desta_logmel_plan_build: bad n_mels %d or buffer too small", n_mels); return DESTA_ERR_ARG; }
  double* fb = new double[MEL_NFREQ * n_mels];
  build_filters(n_mels, fb);
  const int total4 = filters_total_groups(n_mels, fb);
  memset(host_buf, 0, need);
  int* hdr = (int*)host_buf;
  hdr[0] = n_mels; hdr[1] = total4; hdr[2] = (int)(need / 4);
  float* p = (float*)host_buf + PLAN_HDR;
  const double PI = 3.14159265358979323846;
  // periodic Hann, halved: the untangling of the real FFT then needs no 1/2 (exact scaling)
  for (int n = 0; n < 400; ++n) p[n] = (float)(0.5 * (0.5 - 0.5 * cos(2.0 * PI * n / 400.0)));
  p += 400;
  for (int k1 = 0; k1 < 20; ++k1)              // tw1[k1][t] = W200^(t k1)
    for (int t = 0; t < 10; ++t) {
      p[2 * (k1 * 16 + t)] = (float)cos(2.0 * PI * (t * k1) / 200.0);
      p[2 * (k1 * 16 + t) + 1] = (float)(-sin(2.0 * PI * (t * k1) / 200.0));
    }
  p += 640;
  for (int i = 0; i < 11; ++i)                 // tw2[slot][u] = W400^k of the slot's first bin (see stage 2 of the kernel)
    for (int u = 0; u < 10; ++u) {
      int k;
      if (u == 0) k = i <= 5 ? 20 * i : 10 + 20 * (i - 6);
      else k = i < 10 ? u + 20 * i : 0;
      p[2 * (i * 16 + u)] = (float)cos(2.0 * PI * k / 400.0);
      p[2 * (i * 16 + u) + 1] = (float)(-sin(2.0 * PI * k / 400.0));
    }
  p += 352;
  int* meta = (int*)p;
  float* w = (float*)(meta + 2 * n_mels);
  int* n4 = new int[n_mels];
  int pos = 0;                                 // in float4 units
  for (int m = 0; m < n_mels; ++m) {
    int first, last;
    filter_support(n_mels, fb, m, &first, &last);
    const int cnt = first < 0 ? 0 : last - first + 1;
    n4[m] = (cnt + 3) / 4;
    meta[2 * m] = (first < 0 ? 0 : first) | (n4[m] << 16);
    meta[2 * m + 1] = pos;
    for (int i = 0; i < cnt; ++i) w[4 * pos + i] = (float)fb[(first + i) * n_mels + m];
    pos += n4[m];
  }
  split_mels(n_mels, n4, hdr + 3);
  delete[] n4;
  delete[] fb;
  return DESTA_OK;
}

extern "C" int desta_logmel(const float* wave, const int32_t* n_valid, int N, int n_mels, const void* plan, float* out_f32,
                            void* out_bf16_tm, float* clip_max, void* stream) {
  if (!wave || !plan || !clip_max || (!out_f32 && !out_bf16_tm)) { set_error("desta_logmel: null pointer"); return DESTA_ERR_ARG; }
  if (N <= 0) return DESTA_OK;
  const size_t pb = desta_logmel_plan_bytes(n_mels);
  if (!pb) { set_error("desta_logmel: bad n_mels %d (even, <= 256)", n_mels); return DESTA_ERR_ARG; }
  if (N > 65535) { set_error("desta_logmel: at most 65535 clips per call"); return DESTA_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  const size_t tab = ((pb / 4 - PLAN_HDR) + 3) & ~(size_t)3;
  const size_t smem1 = (tab + MT_WAVE_PAD + MEL_TILE * MT_EX + MEL_TILE * MT_PS) * 4;
  bool& attr_set = *device_once_flag(5);
  static size_t attr_bytes[DESTA_MAX_DEVICES] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= DESTA_MAX_DEVICES) dev = 0;
  if (!attr_set || smem1 > attr_bytes[dev]) {
    cudaError_t e = cudaFuncSetAttribute(logmel_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
    if (e != cudaSuccess) { set_error("logmel smem attribute: %s", cudaGetErrorString(e)); return DESTA_ERR_CUDA; }
    attr_set = true;
    attr_bytes[dev] = smem1;
  }
  cudaError_t e = cudaMemsetAsync(clip_max, 0, (size_t)N * 4, s);
  if (e != cudaSuccess) { set_error("desta_logmel memset: %s", cudaGetErrorString(e)); return DESTA_ERR_CUDA; }
  const int n_tiles = N * TILES_PER_CLIP;
  int grid = 2 * num_sms();
  if (grid > n_tiles) grid = n_tiles;
  logmel_tile_kernel<<<grid, MT_THREADS, smem1, s>>>(wave, n_valid, (const float*)plan, n_tiles, out_f32,
                                                     out_f32 ? nullptr : (__nv_bfloat16*)out_bf16_tm, (unsigned*)clip_max);
  int rc = check_launch("logmel_tile_kernel");
  if (rc || !out_f32) return rc;
  const dim3 fgrid(TILES_PER_CLIP, N);
  logmel_finish_kernel<<<fgrid, 256, (size_t)n_mels * (MEL_TILE + 1) * 4, s>>>(out_f32, (const unsigned*)clip_max, n_mels,
                                                                               (__nv_bfloat16*)out_bf16_tm);
  return check_launch("logmel_finish_kernel");
}
