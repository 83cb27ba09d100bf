// Shared device-side PTX wrappers for sm_100a (mbarrier, TMA, tcgen05/TMEM) and host-side helpers.
// Hand-written inline PTX; descriptor bit layouts follow the PTX ISA "tcgen05" matrix/instruction descriptors.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "desta_b200.h"

namespace desta {

// ----------------------------------------------------------------------------------------------
// error plumbing (host)
// ----------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

// TMA descriptor encode through the driver entry point (no link-time libcuda dependency).
// dims/strides: innermost first; strides in BYTES for dims 1..rank-1. Returns 0 on success.
int make_tmap(CUtensorMap* out, const void* base, CUtensorMapDataType dt, int elem_bytes, int rank,
              const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz);
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows, uint32_t box_cols);
// Per-device one-time state (kernel attributes such as the dynamic shared-memory limit are per device): returns a
// pointer to a flag for (current device, slot) that is false until the caller sets it.  The library keeps no other
// mutable state, so one process may drive several GPUs.
constexpr int DESTA_MAX_DEVICES = 64;
bool* device_once_flag(int slot);
int num_sms();

// ---- dropout configuration as the kernels take it (by value); built from the ABI's desta_dropout (NULL -> off) ----
struct DropoutCfg {
  uint32_t thr16;      // keep iff rnd16 >= thr16; 0 disables dropout
  float scale;         // 1 / (1 - p)
  uint32_t seed_lo, seed_hi;
  uint32_t site;       // which dropout call of the step (layer / position), part of the counter
  float* row_ratio;    // attention forward only: optional per-row (dropped sum / plain sum) of the probabilities
};
inline DropoutCfg make_dropout_cfg(const desta_dropout* d) {
  DropoutCfg c = {0u, 1.0f, 0u, 0u, 0u, nullptr};
  if (d && d->thr16) {
    c.thr16 = d->thr16 > 65535u ? 65535u : d->thr16;
    c.scale = d->scale;
    c.seed_lo = (uint32_t)d->seed;
    c.seed_hi = (uint32_t)(d->seed >> 32);
    c.site = d->site;
    c.row_ratio = d->row_keep_ratio;
  }
  return c;
}

#ifdef __CUDACC__
// ----------------------------------------------------------------------------------------------
// small device helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// variant with a suspend-time hint: the thread sleeps in hardware instead of spinning.  Wake-up is slower, so it is
// only used for long, latency-tolerant waits (GEMM producer waiting for a free stage, epilogue waiting for a tile).
__device__ __forceinline__ bool mbar_try_wait_sleep(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// Spin on try_wait (hardware-assisted sleep).  A wait that lasts > ~2 s can only be a protocol bug: report and trap
// instead of hanging the GPU.
static __device__ __noinline__ void mbar_timeout(uint64_t* bar, uint32_t parity) {
  printf("desta_b200: mbarrier timeout block (%d,%d,%d) thread %d bar 0x%x parity %u\n", blockIdx.x, blockIdx.y,
         blockIdx.z, threadIdx.x, smem_u32(bar), parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) mbar_timeout(bar, parity);
  }
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_sleep(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) mbar_timeout(bar, parity);
  }
}

// ---- TMA --------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ---- tcgen05 / TMEM ---------------------------------------------------------------------------
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 inputs, fp32 accumulate. Issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the A operand in TENSOR MEMORY (M x K, K-major: lane = row, two bf16 per 32-bit column), B from smem.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives TMEM lane (base_lane + t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- descriptors ------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_128B (layout_type 2 at bits [61,64)), descriptor version 1 (bits [46,48)).
//   start address [0,14) (>>4), leading byte offset [16,30) (>>4), stride byte offset [32,46) (>>4).
// K-major tile written by TMA with a 128-byte inner box: rows are 128 B apart, 8-row groups 1024 B apart (SBO).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Same with no swizzle ("interleaved" canonical K-major layout, layout_type 0): core matrix = 8 rows x 16 bytes stored
// contiguously (128 B); LBO = byte distance between the two 16-byte K chunks of a K = 16 step, SBO = byte distance
// between consecutive 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32.
//   c_format [4,6)=1 (F32), a_format [7,10)=1 (BF16), b_format [10,13)=1, a_major bit 15, b_major bit 16
//   (0 = K-major, 1 = MN-major), n_dim [17,23) = N>>3, m_dim [24,29) = M>>4.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- math --------------------------------------------------------------------------------------
// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2: two lanes per issued instruction) ----
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}


// erf-GELU (the reference uses the exact erf form: nn.functional.gelu / HF "gelu").
// erf via Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7) with fast exp / reciprocal.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.0f - p * t * __expf(-z * z);   // erf(|x|/sqrt2)
  const float erfv = copysignf(e, x);
  return 0.5f * x * (1.0f + erfv);
}

// erf-GELU for 32 accumulator values at once, MUFU-free and written stage-by-stage so that the 32 independent
// Horner chains interleave (the per-element form above serialises on the MUFU scoreboard inside a GEMM epilogue).
// erf(z) ~ z * P(z^2) on |z| <= 3, z = x / sqrt(2) clamped to [-3, 3].  P is a minimax fit (|err| < 8.9e-5 in fp32) under
// the constraint 3 P(9) = 1: at the clamp the approximation is +-1 up to fp32 rounding (1 - 1.3e-6), so the tails are
// right instead of merely small - GELU(x) -> 0 for x << 0 (|err| < 1e-6 |x|: 4e-5 at x = -60; real Whisper checkpoints
// have large fc1 outliers) and GELU(x) -> x for x >> 0.  |GELU error| <= 1.9e-4 absolute (worst near |x| = 4.2), below
// the bf16 rounding of the stored value; tests/test_kernels_gpu.py::test_gemm_gelu_tails checks |x| up to 60.
__device__ __forceinline__ void gelu_erf_x32(float (&f)[32]) {
  // packed fp32x2 arithmetic: 16 register pairs, every polynomial stage is 16 independent FFMA2 (half the issued
  // instructions of the scalar form; the fc1 epilogue was at 48 % issue utilisation and held the tensor pipe at 85 %)
  uint64_t z[16], z2[16], p[16];
  const uint64_t k = pack_f32x2(0.70710678118654752f, 0.70710678118654752f);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float a, b;
    unpack_f32x2(fmul2(pack_f32x2(f[2 * j], f[2 * j + 1]), k), a, b);
    a = fminf(fmaxf(a, -3.0f), 3.0f);
    b = fminf(fmaxf(b, -3.0f), 3.0f);
    z[j] = pack_f32x2(a, b);
    z2[j] = fmul2(z[j], z[j]);
  }
  const uint64_t c0 = pack_f32x2(-3.901861874e-07f, -3.901861874e-07f), c1 = pack_f32x2(1.668067671e-05f, 1.668067671e-05f);
#pragma unroll
  for (int j = 0; j < 16; ++j) p[j] = ffma2(c0, z2[j], c1);
  const float cs[6] = {-3.086585493e-04f, 3.281595651e-03f, -2.256293967e-02f, 1.075120047e-01f, -3.730821311e-01f, 1.127865195e+00f};
#pragma unroll
  for (int s = 0; s < 6; ++s) {
    const uint64_t cc = pack_f32x2(cs[s], cs[s]);
#pragma unroll
    for (int j = 0; j < 16; ++j) p[j] = ffma2(p[j], z2[j], cc);
  }
  const uint64_t half = pack_f32x2(0.5f, 0.5f);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const uint64_t hx = fmul2(pack_f32x2(f[2 * j], f[2 * j + 1]), half);
    unpack_f32x2(ffma2(hx, fmul2(p[j], z[j]), hx), f[2 * j], f[2 * j + 1]);
  }
}

// ---- dropout RNG: Philox-4x32-7, counter-based (the same (seed, site, element) always yields the same bit, so the
//      backward regenerates the forward's mask instead of storing it).  One call = 8 elements x 16 random bits.
//      Q-Former dropout of the reference: BertConfig hidden_dropout_prob = attention_probs_dropout_prob = 0.1
//      ($TF/models/bert/modeling_bert.py:198-201 attention probs, :296,354 dense outputs), active in train mode. ----
__device__ __forceinline__ uint4 philox4x32_7(uint32_t k0, uint32_t k1, uint4 c) {
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c;
}
// 8 keep-bits (bit i = element 8*group + i is kept) for the aligned group of 8 elements `group`
__device__ __forceinline__ uint32_t dropout_keep8(const DropoutCfg& d, uint64_t group) {
  const uint4 r = philox4x32_7(d.seed_lo, d.seed_hi, make_uint4((uint32_t)group, (uint32_t)(group >> 32), d.site, 0x0DE57Au));
  uint32_t m = 0;
  m |= ((r.x & 0xFFFFu) >= d.thr16) << 0; m |= ((r.x >> 16) >= d.thr16) << 1;
  m |= ((r.y & 0xFFFFu) >= d.thr16) << 2; m |= ((r.y >> 16) >= d.thr16) << 3;
  m |= ((r.z & 0xFFFFu) >= d.thr16) << 4; m |= ((r.z >> 16) >= d.thr16) << 5;
  m |= ((r.w & 0xFFFFu) >= d.thr16) << 6; m |= ((r.w >> 16) >= d.thr16) << 7;
  return m;
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
#endif  // __CUDACC__

}  // namespace desta
