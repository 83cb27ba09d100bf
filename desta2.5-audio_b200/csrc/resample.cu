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
#include <math.h>

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
  int R, Qt;           // residues per CTA, threads along the period axis (CTA covers Q = Qt * RS_U periods)
  int win_cap;         // floats reserved for the input window
};

template <int FMT>
__device__ __forceinline__ float pcm_sample(const uint8_t* p, long long idx) {
  if (FMT == DESTA_PCM_U8) return ((float)p[idx] - 128.0f) * (1.0f / 128.0f);
  if (FMT == DESTA_PCM_S16) return (float)reinterpret_cast<const int16_t*>(p)[idx] * (1.0f / 32768.0f);
  if (FMT == DESTA_PCM_S24) {
    const uint8_t* b = p + idx * 3;
    const int v = (int)((uint32_t)b[0] << 8 | (uint32_t)b[1] << 16 | (uint32_t)b[2] << 24) >> 8;
    return (float)v * (1.0f / 8388608.0f);
  }
  if (FMT == DESTA_PCM_S32) return (float)reinterpret_cast<const int32_t*>(p)[idx] * (1.0f / 2147483648.0f);
  if (FMT == DESTA_PCM_F32) return reinterpret_cast<const float*>(p)[idx];
  return (float)reinterpret_cast<const double*>(p)[idx];
}

// One frame -> mono float32 exactly as the reference: convert every channel to float32, sum left to right in float32,
// divide by the channel count (np.mean over the last axis, audio.py:98).
template <int FMT>
__device__ __forceinline__ float pcm_frame(const uint8_t* p, long long frame, int channels) {
  if (channels == 1) return pcm_sample<FMT>(p, frame);
  float s = 0.f;
  for (int c = 0; c < channels; ++c) s += pcm_sample<FMT>(p, frame * channels + c);
  return __fdiv_rn(s, (float)channels);
}

template <int FMT>
__global__ void __launch_bounds__(RS_THREADS)
decode_resample_kernel(const ResampleArgs a) {
  extern __shared__ __align__(16) float rs_smem[];
  float* s_coef = rs_smem;                       // [R][Ps]
  float* s_win = rs_smem + a.R * a.Ps;           // input window
  const desta_pcm_clip clip = a.clips[blockIdx.z];
  const long long n_out = clip.n_out < a.out_len ? clip.n_out : a.out_len;
  const int r0 = blockIdx.y * a.R;
  const long long q0 = (long long)blockIdx.x * (a.Qt * RS_U);
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
        if (n < a.out_len) orow[n] = 0.f;
      }
    }
    return;
  }

  const int H = a.P / 2 - (a.P > 1 ? 1 : 0);
  const long long off0 = (long long)r0 * a.M / a.L;
  const long long k_min = q0 * a.M + off0 - H;
  const int r_last = min(r0 + a.R, a.L) - 1;
  const int win_len = (a.Qt * RS_U - 1) * a.M + (int)((long long)r_last * a.M / a.L - off0) + a.P;
  const uint8_t* pcm = reinterpret_cast<const uint8_t*>(clip.pcm);
  for (int i = tid; i < win_len; i += RS_THREADS) {
    const long long k = k_min + i;
    s_win[i] = (k >= 0 && k < clip.n_frames) ? pcm_frame<FMT>(pcm, k, a.channels) : 0.f;
  }
  for (int i = tid; i < a.R * a.P; i += RS_THREADS) {
    const int rr = i / a.P, j = i - rr * a.P;
    s_coef[rr * a.Ps + j] = (r0 + rr < a.L) ? a.taps[(size_t)(r0 + rr) * a.P + j] : 0.f;
  }
  __syncthreads();
  if (!live) return;

  const float* coef = s_coef + r_l * a.Ps;
  const int base = qt * a.M + (int)((long long)(r0 + r_l) * a.M / a.L - off0);
  const int step = a.Qt * a.M;
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
    if (n < a.out_len) orow[n] = n < n_out ? acc[u] : 0.f;
  }
}

// ---- host side: filter design (fp64) ----
static long long gcd_ll(long long a, long long b) { while (b) { const long long t = a % b; a = b; b = t; } return a; }

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
  ResampleArgs a;
  a.clips = clips_dev; a.taps = taps_dev; a.out = out; a.out_stride = out_stride; a.out_len = out_len;
  a.fmt = fmt; a.channels = channels; a.L = plan->L; a.M = plan->M; a.P = plan->P; a.Ps = plan->P | 1;
  a.R = plan->L < RS_MAX_R ? plan->L : RS_MAX_R;
  a.Qt = RS_THREADS / a.R;
  // keep the input window under ~96 KB: fewer periods per CTA for large decimation factors
  while (a.Qt > 1 && (long long)a.Qt * RS_U * a.M > 24576) a.Qt >>= 1;
  const int Q = a.Qt * RS_U;
  a.win_cap = (Q - 1) * a.M + (int)(((long long)a.R * a.M + a.L - 1) / a.L) + a.P + 1;
  const size_t smem = ((size_t)a.R * a.Ps + a.win_cap) * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("desta_decode_resample: rate ratio %d/%d needs %zu bytes of shared memory", plan->L, plan->M, smem);
    return DESTA_ERR_SHAPE;
  }
  const long long periods = (out_len + a.L - 1) / a.L;
  const dim3 grid((unsigned)((periods + Q - 1) / Q), (unsigned)((a.L + a.R - 1) / a.R), (unsigned)n_clips);
  if (grid.y > 65535) { set_error("desta_decode_resample: rate ratio %d/%d has too many phases", plan->L, plan->M); return DESTA_ERR_SHAPE; }
  cudaStream_t s = (cudaStream_t)stream;
#define RS_LAUNCH(F)                                                                                                          \
  case F: {                                                                                                                   \
    if (smem > 48 * 1024) {                                                                                                   \
      cudaError_t e = cudaFuncSetAttribute(decode_resample_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
      if (e != cudaSuccess) { set_error("decode_resample smem attribute: %s", cudaGetErrorString(e)); return DESTA_ERR_CUDA; } \
    }                                                                                                                         \
    decode_resample_kernel<F><<<grid, RS_THREADS, smem, s>>>(a);                                                              \
    break;                                                                                                                    \
  }
  switch (fmt) {
    RS_LAUNCH(DESTA_PCM_U8)
    RS_LAUNCH(DESTA_PCM_S16)
    RS_LAUNCH(DESTA_PCM_S24)
    RS_LAUNCH(DESTA_PCM_S32)
    RS_LAUNCH(DESTA_PCM_F32)
    RS_LAUNCH(DESTA_PCM_F64)
  }
#undef RS_LAUNCH
  return check_launch("decode_resample_kernel");
}
