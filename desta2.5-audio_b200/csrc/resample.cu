// PCM decode + channel average + polyphase resample to the model rate + right-pad / truncate, one launch per group of
// clips that share a container format (scope row n3; reference: desta/utils/audio.py:147-166, 221-243, 305-315).
//
// Output n of a clip sits at input time n*M/L.  With n = q*L + r the fractional position depends on r only, so the filter
// is a table of L rows (one per output residue r) of P taps, and output (q, r) reads the input from
// k0 = q*M + floor(r*M/L) - (P/2 - 1).  A CTA owns R residues x Q periods of one clip: it stages the R coefficient rows and
// the (Q-1)*M + P (+ R*M/L) input samples it needs in shared memory - the sample-format conversion and the channel average
// happen in that staging pass, so the raw bytes are read once, coalesced - and every thread then produces U = 4 outputs
// of one residue (one coefficient read feeds four FMAs).  The arithmetic is P fp32 FMAs per output (P = 248 for
// 22.05 kHz, 494 for 44.1 kHz): the kernel is bound by the FP32 / shared-memory pipes, not by HBM.
//
// That kernel (decode_resample_kernel) works for every ratio but pays 1.25 shared-memory loads per FMA, with bank
// conflicts when neighbouring residues start 2-3 samples apart.  The common ratios run on decode_resample_tiled_kernel
// instead: lanes run along the PERIOD axis (input stride M, made odd by one pad word per period -> conflict-free), a warp
// owns four consecutive residues whose coefficient rows are staged interleaved and shifted onto a common input index
// ([i][4], zero-padded), so the inner step is one broadcast LDS.128 of coefficients + QU input loads for 4*QU FMAs
// (0.375-0.5 loads per FMA).  Small L (32 kHz: L = 1) is handled by treating s consecutive periods as one period of
// s*L residues (the tables are indexed r mod L), which gives every warp four distinct output phases to share an input.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "desta_b200.h"

namespace desta {

constexpr int RS_U = 4;          // outputs (periods) per thread
constexpr int RS_THREADS = 256;
constexpr int RS_MAX_R = 32;

struct ResampleArgs {
  const desta_pcm_clip* clips;
  const float* taps;
  float* out;
  long long out_stride, out_len;
  int fmt, channels;
  int L, M, P, Ps;     // Ps: shared-memory row stride of the coefficient rows (odd -> conflict-free across residues)
  int R, Qt;           // residues per CTA, threads along the period axis (CTA covers Q = Qt * U periods)
  int U;               // periods per thread: RS_U, or 1 when one period of input (M samples) is already a large window
  int win_cap;         // floats reserved for the input window
  float inv_ch;
};

// Samples are read through the read-only global path (__ldg): the clip pointers come out of a descriptor array, so the
// compiler could not otherwise tell the address space and would emit generic loads.
template <int FMT, typename Index>
__device__ __forceinline__ float pcm_sample(const uint8_t* p, Index idx) {
  if (FMT == DESTA_PCM_U8) return ((float)__ldg(p + idx) - 128.0f) * (1.0f / 128.0f);
  if (FMT == DESTA_PCM_S16) return (float)__ldg(reinterpret_cast<const short*>(p) + idx) * (1.0f / 32768.0f);
  if (FMT == DESTA_PCM_S24) {
    const uint8_t* b = p + idx * 3;
    const int v = (int)((uint32_t)__ldg(b) << 8 | (uint32_t)__ldg(b + 1) << 16 | (uint32_t)__ldg(b + 2) << 24) >> 8;
    return (float)v * (1.0f / 8388608.0f);
  }
  if (FMT == DESTA_PCM_S32) return (float)__ldg(reinterpret_cast<const int*>(p) + idx) * (1.0f / 2147483648.0f);
  if (FMT == DESTA_PCM_F32) return __ldg(reinterpret_cast<const float*>(p) + idx);
  return (float)__ldg(reinterpret_cast<const double*>(p) + idx);
}

// One frame -> mono float32 exactly as the reference: convert every channel to float32, sum left to right in float32,
// divide by the channel count (np.mean over the last axis, audio.py:98).  CH = 1 / 2: compile-time channel count (straight
// code, so that the staging loops keep their loads batched); CH = 0: any count.  inv_ch = 1 / channels when that is exact
// (power of two: multiplying is then bit-identical to dividing), else 0.
template <int FMT, int CH, typename Index>
__device__ __forceinline__ float pcm_frame(const uint8_t* p, Index frame, int channels, float inv_ch) {
  if (CH == 1) return pcm_sample<FMT, Index>(p, frame);
  if (CH == 2) {
    if (FMT == DESTA_PCM_S16) {
      const uint32_t v = __ldg(reinterpret_cast<const unsigned int*>(p) + frame);   // both channels in one aligned load
      return ((float)(int16_t)(v & 0xffffu) * (1.0f / 32768.0f) + (float)(int16_t)(v >> 16) * (1.0f / 32768.0f)) * 0.5f;
    }
    return (pcm_sample<FMT, Index>(p, frame * 2) + pcm_sample<FMT, Index>(p, frame * 2 + 1)) * 0.5f;
  }
  float s = 0.f;
  for (int c = 0; c < channels; ++c) s += pcm_sample<FMT, Index>(p, frame * channels + c);
  return inv_ch != 0.f ? s * inv_ch : __fdiv_rn(s, (float)channels);
}

__host__ __device__ inline float exact_inverse(int channels) {
  return (channels & (channels - 1)) == 0 ? 1.0f / (float)channels : 0.f;
}

template <int FMT, int CH>
__global__ void __launch_bounds__(RS_THREADS)
decode_resample_kernel(const ResampleArgs a) {
  extern __shared__ __align__(16) float rs_smem[];
  float* s_coef = rs_smem;                       // [R][Ps]
  float* s_win = rs_smem + a.R * a.Ps;           // input window
  const desta_pcm_clip clip = a.clips[blockIdx.z];
  const long long n_out = clip.n_out < a.out_len ? clip.n_out : a.out_len;
  const int r0 = blockIdx.y * a.R;
  const long long q0 = (long long)blockIdx.x * (a.Qt * a.U);
  const int tid = threadIdx.x;
  const int r_l = tid % a.R, qt = tid / a.R;
  const bool live = qt < a.Qt && r0 + r_l < a.L;
  float* orow = a.out + (size_t)clip.out_row * a.out_stride;

  if (q0 * a.L + r0 >= n_out) {
    // whole CTA beyond the end of the clip: zero padding only
    if (live) {
#pragma unroll
      for (int u = 0; u < RS_U; ++u) {
        const long long n = (q0 + qt + u * a.Qt) * a.L + r0 + r_l;
        if (u < a.U && n < a.out_len) orow[n] = 0.f;
      }
    }
    return;
  }

  const int H = a.P / 2 - (a.P > 1 ? 1 : 0);
  const long long off0 = (long long)r0 * a.M / a.L;
  const long long k_min = q0 * a.M + off0 - H;
  const int r_last = min(r0 + a.R, a.L) - 1;
  const int win_len = (a.Qt * a.U - 1) * a.M + (int)((long long)r_last * a.M / a.L - off0) + a.P;
  const uint8_t* pcm = reinterpret_cast<const uint8_t*>(clip.pcm);
  for (int i = tid; i < win_len; i += RS_THREADS) {
    const long long k = k_min + i;
    s_win[i] = (k >= 0 && k < clip.n_frames) ? pcm_frame<FMT, CH>(pcm, k, a.channels, a.inv_ch) : 0.f;
  }
  for (int i = tid; i < a.R * a.P; i += RS_THREADS) {
    const int rr = i / a.P, j = i - rr * a.P;
    s_coef[rr * a.Ps + j] = (r0 + rr < a.L) ? a.taps[(size_t)(r0 + rr) * a.P + j] : 0.f;
  }
  __syncthreads();
  if (!live) return;

  const float* coef = s_coef + r_l * a.Ps;
  const int base = qt * a.M + (int)((long long)(r0 + r_l) * a.M / a.L - off0);
  const int step = a.U > 1 ? a.Qt * a.M : 0;      // U == 1: the unused accumulators re-read this thread's own window
  const float* w0 = s_win + base;
  float acc[RS_U] = {0.f, 0.f, 0.f, 0.f};
  int j = 0;
  for (; j + 4 <= a.P; j += 4) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const float c = coef[j + jj];
#pragma unroll
      for (int u = 0; u < RS_U; ++u) acc[u] = fmaf(c, w0[u * step + j + jj], acc[u]);
    }
  }
  for (; j < a.P; ++j) {
    const float c = coef[j];
#pragma unroll
    for (int u = 0; u < RS_U; ++u) acc[u] = fmaf(c, w0[u * step + j], acc[u]);
  }
#pragma unroll
  for (int u = 0; u < RS_U; ++u) {
    const long long n = (q0 + qt + u * a.Qt) * a.L + r0 + r_l;
    if (u < a.U && n < a.out_len) orow[n] = n < n_out ? acc[u] : 0.f;
  }
}

// ---- host side: filter design (fp64) ----
static long long gcd_ll(long long a, long long b) { while (b) { const long long t = a % b; a = b; b = t; } return a; }

// sr_in == sr_out: conversion + channel average + zero padding only - a streaming kernel, four outputs per thread.
template <int FMT, int CH>
__global__ void __launch_bounds__(256)
decode_copy_kernel(const desta_pcm_clip* clips, float* out, long long out_stride, long long out_len, int channels, float inv_ch,
                   int vec_store) {
  const desta_pcm_clip clip = clips[blockIdx.y];
  const long long n_out = clip.n_frames < out_len ? clip.n_frames : out_len;
  const uint8_t* pcm = reinterpret_cast<const uint8_t*>(clip.pcm);
  const bool aligned16 = (reinterpret_cast<uintptr_t>(pcm) & 15) == 0;
  float* orow = out + (size_t)clip.out_row * out_stride;
  for (long long n = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; n < out_len; n += (long long)gridDim.x * blockDim.x * 4) {
    float v[4];
    if (FMT == DESTA_PCM_S16 && (CH == 1 || CH == 2) && aligned16 && n + 3 < n_out) {
      // 16-bit PCM, the common case: four frames in one 8- or 16-byte load
      constexpr float k = 1.0f / 32768.0f;
      if (CH == 1) {
        const uint2 w = __ldg(reinterpret_cast<const uint2*>(pcm) + (n >> 2));
        v[0] = (float)(int16_t)(w.x & 0xffffu) * k; v[1] = (float)(int16_t)(w.x >> 16) * k;
        v[2] = (float)(int16_t)(w.y & 0xffffu) * k; v[3] = (float)(int16_t)(w.y >> 16) * k;
      } else {
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(pcm) + (n >> 2));
        const uint32_t f[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = ((float)(int16_t)(f[e] & 0xffffu) * k + (float)(int16_t)(f[e] >> 16) * k) * 0.5f;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = n + e < n_out ? pcm_frame<FMT, CH>(pcm, n + e, channels, inv_ch) : 0.f;
    }
    if (vec_store && n + 3 < out_len) {
      *reinterpret_cast<float4*>(orow + n) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (n + e < out_len) orow[n + e] = v[e];
    }
  }
}

struct TiledArgs {
  const desta_pcm_clip* clips;
  const float* taps;
  float* out;
  long long out_stride, out_len;
  int channels;
  int L, M, P;         // unscaled ratio and taps per output
  int Lp, Mp;          // scaled period: Lp = s L residues, Mp = s M input samples
  int pad, Sp;         // one pad word per period when Mp is even; Sp = Mp + pad (odd -> conflict-free lane stride)
  int Kw;              // staged taps per warp: P + shift of the warp's last residue, rounded up to a multiple of 4
  int warps;           // warps per CTA; the CTA owns R = 4 * warps residues x 32 * QU periods
  int win_len;         // input samples staged per block of periods
  int nb;              // blocks of 32 * QU periods a CTA walks with one coefficient tile
  int vec_store;       // 1 -> the four residues of a thread are stored as one float4
  float inv_ch;
};

// Window staging for the tiled kernel: eight independent loads in flight per thread (the loads, not the arithmetic, are
// what a block of periods waits for), conversion + channel average on the way into shared memory.  Windows that lie wholly
// inside the clip - all but the first and last of a clip - take the path without range checks and 64-bit index arithmetic.
template <int FMT, int CH>
__device__ __forceinline__ void stage_window(float* s_win, int win_len, int pad, int Mp, const uint8_t* pcm, long long k_min,
                                             long long n_frames, int channels, float inv_ch) {
  const int nthr = blockDim.x;
  const bool interior = k_min >= 0 && k_min + win_len <= n_frames;
  constexpr int fb = FMT == DESTA_PCM_U8 ? 1 : FMT == DESTA_PCM_S16 ? 2 : FMT == DESTA_PCM_S24 ? 3 : FMT == DESTA_PCM_F64 ? 8 : 4;
  const uint8_t* base = pcm + (interior ? k_min * channels * fb : 0);      // frame 0 of the window
  for (int b = threadIdx.x; b < win_len; b += nthr * 8) {
    float v[8];
    if (interior) {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int idx = min(b + t * nthr, win_len - 1);
        v[t] = pcm_frame<FMT, CH>(base, idx, channels, inv_ch);
      }
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int idx = b + t * nthr;
        const long long k = k_min + idx;
        v[t] = (idx < win_len && k >= 0 && k < n_frames) ? pcm_frame<FMT, CH>(pcm, k, channels, inv_ch) : 0.f;
      }
    }
    if (pad) {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int idx = b + t * nthr;
        if (idx < win_len) s_win[idx + idx / Mp] = v[t];
      }
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int idx = b + t * nthr;
        if (idx < win_len) s_win[idx] = v[t];
      }
    }
  }
}

template <int FMT, int CH, int QU>
__global__ void __launch_bounds__(256)
decode_resample_tiled_kernel(const TiledArgs a) {
  extern __shared__ __align__(16) float rs_smem[];
  float* s_coef = rs_smem;                                   // [warps][Kw][4]
  float* s_win = rs_smem + a.warps * a.Kw * 4;               // padded input window: rows of Mp samples, stride Sp
  const desta_pcm_clip clip = a.clips[blockIdx.z];
  const long long n_out = clip.n_out < a.out_len ? clip.n_out : a.out_len;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = 4 * a.warps;
  const int r0c = blockIdx.y * R;
  const int r0w = r0c + 4 * warp;
  float* orow = a.out + (size_t)clip.out_row * a.out_stride;
  const uint8_t* pcm = reinterpret_cast<const uint8_t*>(clip.pcm);
  const int H = a.P / 2 - 1;
  const long long off0 = (long long)r0c * a.M / a.L;
  const long long offw = (long long)r0w * a.M / a.L;
  const int wbase = (int)(offw - off0);
  const int im0 = wbase % a.Mp;
  const int pos0 = wbase + a.pad * (wbase / a.Mp);
  bool coef_ready = false;

  // The CTA walks nb blocks of 32*QU periods with one coefficient tile.
  for (int blk = 0; blk < a.nb; ++blk) {
    const long long q0 = ((long long)blockIdx.x * a.nb + blk) * (32 * QU);
    if (q0 * a.Lp + r0c >= a.out_len) break;
    if (q0 * a.Lp + r0c >= n_out) {
      // beyond the end of the clip: zero padding only
#pragma unroll
      for (int u = 0; u < QU; ++u) {
        const long long n = (q0 + lane + 32 * u) * a.Lp + r0w;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (r0w + e < a.Lp && n + e < a.out_len) orow[n + e] = 0.f;
      }
      continue;
    }
    __syncthreads();                                         // the previous block's reads of the window are done
    if (!coef_ready) {
      // coefficient tile of this warp: residue e is shifted by the distance of its first input sample from the first
      // input sample of residue 0, so that all four residues read the same window index i
      int sh[4], row[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        sh[e] = (int)((long long)(r0w + e) * a.M / a.L - offw);
        row[e] = (r0w + e < a.Lp) ? (r0w + e) % a.L : -1;
      }
      float4* cwr = reinterpret_cast<float4*>(s_coef) + (size_t)warp * a.Kw;
      for (int i = lane; i < a.Kw; i += 32) {
        float4 c;
        float* cp = reinterpret_cast<float*>(&c);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = i - sh[e];
          cp[e] = (row[e] >= 0 && j >= 0 && j < a.P) ? __ldg(a.taps + (size_t)row[e] * a.P + j) : 0.f;
        }
        cwr[i] = c;
      }
      coef_ready = true;
    }
    stage_window<FMT, CH>(s_win, a.win_len, a.pad, a.Mp, pcm, q0 * a.Mp + off0 - H, clip.n_frames, a.channels, a.inv_ch);
    __syncthreads();

    int im = im0, pos = pos0;
    const float* xw = s_win + lane * a.Sp;
    const float4* cw = reinterpret_cast<const float4*>(s_coef) + (size_t)warp * a.Kw;
    float acc[QU][4];
#pragma unroll
    for (int u = 0; u < QU; ++u) { acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.f; }
    const int ustep = 32 * a.Sp;
    if (a.pad == 0) {
      const float* x0 = xw + pos;
#pragma unroll 4
      for (int i = 0; i < a.Kw; ++i) {
        const float4 c = cw[i];
#pragma unroll
        for (int u = 0; u < QU; ++u) {
          const float x = x0[i + u * ustep];
          acc[u][0] = fmaf(c.x, x, acc[u][0]); acc[u][1] = fmaf(c.y, x, acc[u][1]);
          acc[u][2] = fmaf(c.z, x, acc[u][2]); acc[u][3] = fmaf(c.w, x, acc[u][3]);
        }
      }
    } else {
      // runs of taps up to the next period boundary (warp-uniform), then skip the pad word
      int i = 0;
      while (i < a.Kw) {
        const int run = min(a.Kw - i, a.Mp - im);
        const float* x0 = xw + pos;
#pragma unroll 4
        for (int t = 0; t < run; ++t) {
          const float4 c = cw[i + t];
#pragma unroll
          for (int u = 0; u < QU; ++u) {
            const float x = x0[t + u * ustep];
            acc[u][0] = fmaf(c.x, x, acc[u][0]); acc[u][1] = fmaf(c.y, x, acc[u][1]);
            acc[u][2] = fmaf(c.z, x, acc[u][2]); acc[u][3] = fmaf(c.w, x, acc[u][3]);
          }
        }
        i += run;
        pos += run + 1;
        im = 0;
      }
    }
#pragma unroll
    for (int u = 0; u < QU; ++u) {
      const long long n = (q0 + lane + 32 * u) * a.Lp + r0w;
      if (a.vec_store && r0w + 3 < a.Lp && n + 3 < a.out_len) {
        float4 v;
        v.x = n < n_out ? acc[u][0] : 0.f; v.y = n + 1 < n_out ? acc[u][1] : 0.f;
        v.z = n + 2 < n_out ? acc[u][2] : 0.f; v.w = n + 3 < n_out ? acc[u][3] : 0.f;
        *reinterpret_cast<float4*>(orow + n) = v;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (r0w + e < a.Lp && n + e < a.out_len) orow[n + e] = n + e < n_out ? acc[u][e] : 0.f;
      }
    }
  }
}

// Chooses the tiling of decode_resample_tiled_kernel; false -> the ratio runs on the general kernel.
static bool plan_tiled(const desta_resample_plan& p, TiledArgs& t, int& QU, size_t& smem) {
  if (p.P < 2) return false;
  long long s = 1;
  if (p.L < 32) { s = 32 / gcd_ll(p.L, 32); }
  const long long Lp = s * p.L, Mp = s * p.M;
  if (Mp > 8192) return false;
  t.L = p.L; t.M = p.M; t.P = p.P; t.Lp = (int)Lp; t.Mp = (int)Mp;
  t.pad = (Mp & 1) ? 0 : 1;
  t.Sp = t.Mp + t.pad;
  t.Kw = (p.P + (int)((3LL * p.M) / p.L) + 1 + 3) & ~3;
  // (warps, QU, shared-memory budget in KB), first fit: eight warps first - also as the ONLY CTA of an SM (44.1 kHz:
  // 178 KB, QU = 2, 2.5 ms per 64 clips against 3.5 ms for two 4-warp CTAs of 89 KB) - then narrower CTAs
  static const int cand[10][3] = {{8, 2, 100}, {8, 1, 100}, {8, 2, 200}, {8, 1, 200}, {4, 2, 100}, {4, 1, 100},
                                  {4, 2, 200}, {4, 1, 200}, {2, 2, 200}, {2, 1, 200}};
  for (int c = 0; c < 10; ++c) {
    const int warps = cand[c][0], qu = cand[c][1];
    const size_t limit = (size_t)cand[c][2] * 1024;
    const int R = 4 * warps;
    const long long win = (32LL * qu - 1) * Mp + ((long long)(R - 4) * p.M) / p.L + 1 + t.Kw;
    const size_t bytes = ((size_t)warps * t.Kw * 4 + (size_t)win + (size_t)t.pad * (win / Mp + 1)) * sizeof(float);
    if (bytes <= limit) {
      t.warps = warps; t.win_len = (int)win; QU = qu; smem = bytes;
      return true;
    }
  }
  return false;
}

static double bessel_i0(double x) {
  double sum = 1.0, term = 1.0;
  const double q = x * x / 4.0;
  for (int k = 1; k < 500; ++k) {
    term *= q / ((double)k * (double)k);
    sum += term;
    if (term < 1e-20 * sum) break;
  }
  return sum;
}

constexpr double RS_PASS = 0.913, RS_STOP = 1.0;     // soxr "HQ" band edges, fractions of the lower Nyquist frequency
constexpr double RS_ATTEN_DB = 120.0;                // ~20-bit precision

static bool plan_of(int sr_in, int sr_out, desta_resample_plan* p) {
  if (sr_in <= 0 || sr_out <= 0) return false;
  const long long g = gcd_ll(sr_in, sr_out);
  p->L = (int32_t)(sr_out / g);
  p->M = (int32_t)(sr_in / g);
  p->reserved = 0;
  if (p->L == 1 && p->M == 1) { p->P = 1; return true; }
  const double low = sr_in < sr_out ? sr_in : sr_out;
  const double delta = 0.5 * (RS_STOP - RS_PASS) * low / (double)sr_in;       // transition width, cycles per input sample
  int taps = (int)ceil((RS_ATTEN_DB - 7.95) / (14.357 * delta));
  taps += taps & 1;
  p->P = taps;
  return true;
}

}  // namespace desta

using namespace desta;

extern "C" size_t desta_resample_plan_bytes(int sr_in, int sr_out, desta_resample_plan* plan) {
  desta_resample_plan p;
  if (!plan_of(sr_in, sr_out, &p)) { set_error("desta_resample_plan: bad rates %d -> %d", sr_in, sr_out); return 0; }
  if (plan) *plan = p;
  return (size_t)p.L * p.P * sizeof(float);
}

extern "C" int desta_resample_plan_build(int sr_in, int sr_out, float* host_taps, size_t bytes) {
  desta_resample_plan p;
  if (!host_taps || !plan_of(sr_in, sr_out, &p)) { set_error("desta_resample_plan_build: bad arguments"); return DESTA_ERR_ARG; }
  if (bytes < (size_t)p.L * p.P * sizeof(float)) { set_error("desta_resample_plan_build: buffer too small"); return DESTA_ERR_ARG; }
  if (p.P == 1) { host_taps[0] = 1.0f; return DESTA_OK; }
  const double low = sr_in < sr_out ? sr_in : sr_out;
  const double fc = 0.25 * (RS_PASS + RS_STOP) * low / (double)sr_in;         // cut-off, cycles per input sample
  const double beta = 0.1102 * (RS_ATTEN_DB - 8.7);
  const double inv_i0 = 1.0 / bessel_i0(beta);
  const double W = 0.5 * p.P;
  const int H = p.P / 2 - 1;
  for (int r = 0; r < p.L; ++r) {
    const double frac = (double)(((long long)r * p.M) % p.L) / (double)p.L;
    for (int j = 0; j < p.P; ++j) {
      const double tau = frac + (double)(H - j);
      const double u = tau / W;
      double g = 0.0;
      if (fabs(u) < 1.0) {
        const double x = 2.0 * fc * tau;
        const double sinc = fabs(x) < 1e-12 ? 1.0 : sin(M_PI * x) / (M_PI * x);
        g = 2.0 * fc * sinc * bessel_i0(beta * sqrt(1.0 - u * u)) * inv_i0;
      }
      host_taps[(size_t)r * p.P + j] = (float)g;
    }
  }
  return DESTA_OK;
}

extern "C" int desta_decode_resample(const desta_pcm_clip* clips_dev, int n_clips, int fmt, int channels,
                                     const desta_resample_plan* plan, const float* taps_dev, float* out, int64_t out_stride,
                                     int64_t out_len, void* stream) {
  if (n_clips <= 0 || out_len <= 0) return DESTA_OK;
  if (!clips_dev || !plan || !taps_dev || !out) { set_error("desta_decode_resample: null pointer"); return DESTA_ERR_ARG; }
  if (fmt < DESTA_PCM_U8 || fmt > DESTA_PCM_F64 || channels < 1 || channels > 64) {
    set_error("desta_decode_resample: unsupported sample format %d / channel count %d", fmt, channels);
    return DESTA_ERR_ARG;
  }
  if (plan->L < 1 || plan->M < 1 || plan->P < 1 || out_stride < out_len || n_clips > 65535) {
    set_error("desta_decode_resample: bad plan or shape (L=%d M=%d P=%d, stride %lld < len %lld, %d clips)", plan->L, plan->M,
              plan->P, (long long)out_stride, (long long)out_len, n_clips);
    return DESTA_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (plan->L == 1 && plan->M == 1 && !(plan->reserved & 1)) {
    const int vec = (out_stride % 4 == 0) && ((uintptr_t)out % 16 == 0);
    const float inv = exact_inverse(channels);
    long long bx = (out_len + 1023) / 1024;
    if (bx > 4096) bx = 4096;
    const dim3 grid((unsigned)bx, (unsigned)n_clips);
#define RS_COPY_CH(F, C) decode_copy_kernel<F, C><<<grid, 256, 0, s>>>(clips_dev, out, out_stride, out_len, channels, inv, vec);
#define RS_COPY(F) case F: if (channels == 1) RS_COPY_CH(F, 1) else if (channels == 2) RS_COPY_CH(F, 2) else RS_COPY_CH(F, 0) break;
    switch (fmt) {
      RS_COPY(DESTA_PCM_U8)
      RS_COPY(DESTA_PCM_S16)
      RS_COPY(DESTA_PCM_S24)
      RS_COPY(DESTA_PCM_S32)
      RS_COPY(DESTA_PCM_F32)
      RS_COPY(DESTA_PCM_F64)
    }
#undef RS_COPY
#undef RS_COPY_CH
    return check_launch("decode_copy_kernel");
  }
  {
    TiledArgs t;
    int QU = 0;
    size_t smem_t = 0;
    if (!(plan->reserved & 1) && plan_tiled(*plan, t, QU, smem_t)) {
      t.clips = clips_dev; t.taps = taps_dev; t.out = out; t.out_stride = out_stride; t.out_len = out_len; t.channels = channels;
      t.vec_store = (t.Lp % 4 == 0) && (out_stride % 4 == 0) && ((uintptr_t)out % 16 == 0);
      t.inv_ch = exact_inverse(channels);
      const int R = 4 * t.warps, Qc = 32 * QU;
      const long long periods = (out_len + t.Lp - 1) / t.Lp;
      const long long blocks = (periods + Qc - 1) / Qc;
      const long long groups = (t.Lp + R - 1) / R;
      // one coefficient tile serves nb period blocks; keep >= ~4 CTAs per SM in flight when the batch allows it
      long long nb = blocks * groups * n_clips / (4LL * num_sms());
      nb = nb < 1 ? 1 : (nb > 16 ? 16 : nb);
      t.nb = (int)nb;
      const dim3 grid((unsigned)((blocks + nb - 1) / nb), (unsigned)groups, (unsigned)n_clips);
      if (grid.y <= 65535) {
#define RS_TILED(F, C, Q)                                                                                                   \
  {                                                                                                                         \
    if (smem_t > 48 * 1024) {                                                                                               \
      cudaError_t e = cudaFuncSetAttribute(decode_resample_tiled_kernel<F, C, Q>,                                           \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);                        \
      if (e != cudaSuccess) { set_error("decode_resample smem attribute: %s", cudaGetErrorString(e)); return DESTA_ERR_CUDA; } \
    }                                                                                                                       \
    decode_resample_tiled_kernel<F, C, Q><<<grid, 32 * t.warps, smem_t, s>>>(t);                                            \
  }
#define RS_TILED_CH(F, C) { if (QU == 2) RS_TILED(F, C, 2) else RS_TILED(F, C, 1) }
#define RS_TILED_FMT(F) case F: if (channels == 1) RS_TILED_CH(F, 1) else if (channels == 2) RS_TILED_CH(F, 2) else RS_TILED_CH(F, 0) break;
        switch (fmt) {
          RS_TILED_FMT(DESTA_PCM_U8)
          RS_TILED_FMT(DESTA_PCM_S16)
          RS_TILED_FMT(DESTA_PCM_S24)
          RS_TILED_FMT(DESTA_PCM_S32)
          RS_TILED_FMT(DESTA_PCM_F32)
          RS_TILED_FMT(DESTA_PCM_F64)
        }
#undef RS_TILED_FMT
#undef RS_TILED_CH
#undef RS_TILED
        return check_launch("decode_resample_tiled_kernel");
      }
    }
  }
  ResampleArgs a;
  a.clips = clips_dev; a.taps = taps_dev; a.out = out; a.out_stride = out_stride; a.out_len = out_len;
  a.fmt = fmt; a.channels = channels; a.L = plan->L; a.M = plan->M; a.P = plan->P; a.Ps = plan->P | 1;
  a.R = plan->L < RS_MAX_R ? plan->L : RS_MAX_R;
  a.Qt = RS_THREADS / a.R;
  a.U = RS_U;
  a.inv_ch = exact_inverse(channels);
  // keep the input window under ~96 KB: fewer periods per CTA for large decimation factors / long periods
  while (a.Qt > 1 && (long long)a.Qt * a.U * a.M > 24576) a.Qt >>= 1;
  if ((long long)a.Qt * a.U * a.M > 24576) a.U = 1;
  const int Q = a.Qt * a.U;
  a.win_cap = (Q - 1) * a.M + (int)(((long long)a.R * a.M + a.L - 1) / a.L) + a.P + 1;
  const size_t smem = ((size_t)a.R * a.Ps + a.win_cap) * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("desta_decode_resample: rate ratio %d/%d needs %zu bytes of shared memory", plan->L, plan->M, smem);
    return DESTA_ERR_SHAPE;
  }
  const long long periods = (out_len + a.L - 1) / a.L;
  const dim3 grid((unsigned)((periods + Q - 1) / Q), (unsigned)((a.L + a.R - 1) / a.R), (unsigned)n_clips);
  if (grid.y > 65535) { set_error("desta_decode_resample: rate ratio %d/%d has too many phases", plan->L, plan->M); return DESTA_ERR_SHAPE; }
#define RS_LAUNCH_CH(F, C)                                                                                                  \
  {                                                                                                                         \
    if (smem > 48 * 1024) {                                                                                                 \
      cudaError_t e = cudaFuncSetAttribute(decode_resample_kernel<F, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                           200 * 1024);                                                                     \
      if (e != cudaSuccess) { set_error("decode_resample smem attribute: %s", cudaGetErrorString(e)); return DESTA_ERR_CUDA; } \
    }                                                                                                                       \
    decode_resample_kernel<F, C><<<grid, RS_THREADS, smem, s>>>(a);                                                         \
  }
#define RS_LAUNCH(F) case F: if (channels == 1) RS_LAUNCH_CH(F, 1) else if (channels == 2) RS_LAUNCH_CH(F, 2) else RS_LAUNCH_CH(F, 0) break;
  switch (fmt) {
    RS_LAUNCH(DESTA_PCM_U8)
    RS_LAUNCH(DESTA_PCM_S16)
    RS_LAUNCH(DESTA_PCM_S24)
    RS_LAUNCH(DESTA_PCM_S32)
    RS_LAUNCH(DESTA_PCM_F32)
    RS_LAUNCH(DESTA_PCM_F64)
  }
#undef RS_LAUNCH
#undef RS_LAUNCH_CH
  return check_launch("decode_resample_kernel");
}
