// Fused non-causal multi-head attention for head_dim 64 on tcgen05/TMEM:  O = softmax(scale * Q K^T) V.
// Replaces WhisperAttention's SDPA ($TF/models/whisper/modeling_whisper.py:284-358; T = 1500, scale 1 because q is
// pre-scaled) and the eager Bert self / cross attention of the Q-Former ($TF/models/bert/modeling_bert.py:168-207,
// 234-284; 64 queries x 64 or 1500 keys, scale 1/8).
//
// One CTA = 128 query rows of one (batch, head); two CTAs are co-resident per SM (96 KB smem, 256 TMEM columns each).
// Keys are processed in blocks of 64.  TMEM holds THREE score buffers of 64 fp32 columns and the 64-column output
// accumulator.  P never touches shared memory: each softmax thread reads its row of S into registers, and writes
// the probabilities back IN PLACE as packed bf16 (32 columns) with tcgen05.st; the P V product takes its A operand
// straight from TMEM.  (At head_dim 64 every UMMA 128x64x16 re-reads 4 KB of A and 2 KB of B from shared memory, so a
// P round trip through shared memory was 40 % of all shared-memory traffic, the busiest unit of the first version.)
// With three buffers the MMA warp runs S = Q K^T two full blocks ahead of the softmax, so neither the wake-up of the
// MMA warp nor the tensor-pipe latency of P V -> next S is ever on the softmax warps' critical path.
//   warp 0   : TMA producer  (Q once; K/V blocks of 64 keys through a 5-stage mbarrier ring, SWIZZLE_128B)
//   warp 1   : TMEM allocator + MMA issuer: S_j = Q K_j^T (UMMA 128x64x16, both operands K-major in smem) into buffer
//              j % 3; O += P_j V_j (UMMA 128x64x16, A = P_j from TMEM, B = V MN-major as it lies in memory); tcgen05
//              MMAs execute in issue order, so S_{j+3} may overwrite buffer j % 3 as soon as P_j V_j has been issued
//   warps 2-5: softmax, one query row per thread (TMEM lane == row).  ONE pass per block: exponentials are taken
//              against the running max of the previous blocks while the block max is gathered in the same pass; only
//              when a row's max grows by more than 2^20 is the reference raised (O and l rescaled in TMEM) and the block
//              redone from the scores still held in registers.  Final 1/l and 128-byte row stores.
//
// UNIT mode (scale * log2(e) == 1: the Whisper encoder, whose q is pre-scaled by the QKV GEMM epilogue - the log2(e) factor
// is folded into that scale): the softmax shift is folded into the tensor-core product.  tools/issue_bench.cu shows that the
// softmax loop is not bound by the MUFU pipe alone (8 cycles per warp instruction) but by the scale-subtract / max / row-sum
// instructions around it, whose dependent issue leaves the XU idle a quarter of the time (2 warps per sub-partition: 73 %
// XU utilisation with today's instruction mix, 92 % with exp + pack + row-sum only).  So S' = Q K^T - m 1^T is computed BY
// THE MMA: a fifth K = 16 step multiplies a [128 x 16] operand holding -m_row (bf16) in its first column with a [64 x 16]
// operand of ones (both un-swizzled K-major tiles in shared memory), and the softmax threads take p = ex2(S') straight from
// the TMEM registers - no FFMA, no running max (a row sum above 2^20 triggers the same lazy rescale).  The shift is a
// per-row constant, so its bf16 rounding cancels exactly in the softmax; it only changes at the first block and at
// rescale events, and every thread remembers which shift each in-flight S block was computed with.
// Start of a tile: S_0 is issued with a zero shift; the softmax threads publish the row maxima of block 0 as the shift
// BEFORE they start its exponentials (mbarrier shift_ready), and only then are S_1 and S_2 issued - their latency hides
// behind block 0's exponentials, and every block but the first runs the lean step (round 2: 1.059 -> 0.997 ms per encoder
// launch together with the items below).  The ragged last key block (1500 = 23 * 64 + 28) runs the lean step with selects
// and skips the exponentials of a fully padded 32-key half.  One of every 8 (or 4) exponentials is evaluated on the FMA pipe
// (poly_exp2): the XU pipe is the floor at head_dim 64, and profiles/r2_poly_attention.md shows where the issue slots run
// out (12.5-25 %: 0.995 ms; 37.5 %: 1.07 ms; 50 %: 1.16 ms).  ex2.approx.f16x2 is no shortcut: MUFU.EX2.F16x2 issues
// every 16 cycles per warp, i.e. the same 8 cycles per exponential (tools/issue_bench.cu).
//
// File map: attn_tile_body<DROP, UNIT, FALLBACK> is the kernel described above (attn_fwd_kernel<DROP, UNIT> wraps it; it is the
// product path of every call).  attn_fwd_unit8_kernel further down is a second design for the encoder's calls that is NOT the
// default (desta_attention_variant(): eight softmax warps, Q in tensor memory; it reuses attn_tile_body<.., FALLBACK = true> to
// recompute a tile whose fixed reference overflowed) - measured slower, kept for A/B runs: profiles/r2_attention_unit8.md.
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "desta_b200.h"

namespace desta {

constexpr int AT_BQ = 128, AT_BKV = 64, AT_HD = 64, AT_STAGES = 5, AT_SBUF = 3;
constexpr int AT_Q_BYTES = AT_BQ * AT_HD * 2;      // 16 KB
constexpr int AT_KV_BYTES = AT_BKV * AT_HD * 2;    // 8 KB (K) + 8 KB (V) per stage
constexpr int AT_AX_BYTES = AT_BQ * 32;            // shift operand  [128 rows x 16 bf16], un-swizzled K-major (UNIT mode)
constexpr int AT_BX_BYTES = AT_BKV * 32;           // ones operand   [64 keys x 16 bf16]
constexpr int AT_SMEM = AT_Q_BYTES + AT_STAGES * 2 * AT_KV_BYTES + AT_AX_BYTES + AT_BX_BYTES + 256;   // 104,704 B -> 2 CTAs / SM
constexpr int AT_THREADS = 192;
constexpr int AT_TMEM_COLS = 256;                  // S/P buffers: [0,64) [64,128) [128,192)   O: [192,256)
constexpr int AT_TMEM_O = AT_SBUF * AT_BKV;

struct AttnArgs {
  __nv_bfloat16* out; int ldo;
  int Tq, Tk;
  float scale_log2;    // scale * log2(e)
  int q_rows, kv_rows; // TMA box rows actually loaded (min(tile, T))
  float* lse;          // optional [batch, heads, Tq]: log2-domain log-sum-exp of the scaled scores (for the backward)
  DropoutCfg drop;     // train-mode dropout on the probabilities (modeling_bert.py:198-201); thr16 == 0 -> off
  float one;           // 1.0f (an opaque multiplier for the packed row-sum FMA)
  int heads, nq, n_tiles;   // tile index = q_tile + nq * (head + heads * batch)
  int stagger, n_sm;   // first wave: the second CTA of every SM starts `stagger` cycles late (see stagger_first_wave)
};

// All tiles of a launch take the same time, so the two CTAs that share an SM start, reach their epilogues and are replaced
// in lockstep: both sit in the latency-bound start of a tile (Q / K / V loads, S_0, shift hand-over) and in its end (last
// P V, O read-out, stores, CTA turnover) at the same moments, and the MUFU pipe idles.  Delaying the second CTA of every
// SM once, in the first wave, by about half a tile keeps the pair out of phase for the whole launch.
__device__ __forceinline__ void stagger_first_wave(const AttnArgs& a) {
  const int lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  if (a.stagger > 0 && lin >= a.n_sm && lin < 2 * a.n_sm) {
    const long long t0 = clock64();
    while (clock64() - t0 < a.stagger) { }
  }
}

// Compiler-level ordering helper: the 32 registers written by an (asynchronous) tcgen05.ld may only be consumed after
// tcgen05.wait::ld; routing them through an empty asm after the wait makes that dependency explicit.
__device__ __forceinline__ void reg_fence32(uint32_t (&v)[32]) {
  asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                    "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                    "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                    "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31]));
}

struct SoftmaxAcc { float l0, l1, m0, m1, t; };     // t (DROP only): sum of the dropped, rescaled probabilities



#ifdef DESTA_ATTN_TRACE
// Debug build only (tools/attn_trace.py): per-CTA cycle accounting of the MMA warp and of softmax warp 2.
__device__ long long g_attn_trace[8];
#define TRACE_T0() const long long _t0 = clock64()
#define TRACE_ADD(slot) do { trace_acc[slot] += clock64() - _t0; } while (0)
#define TRACE_FLUSH() do { if (lane == 0) for (int _i = 0; _i < 8; ++_i) if (trace_acc[_i]) atomicAdd((unsigned long long*)&g_attn_trace[_i], (unsigned long long)trace_acc[_i]); } while (0)
// timeline of ONE CTA (blockIdx 5,7,30): [event][block] clock64 stamps.  events: 0 MMA saw p_ready[j], 1 PV_j issued,
// 2 QK_{j+3} issued (slot j+3), 3..6 softmax warps 2..5 step j start, 7..10 warps 2..5 got s_full[j+1], 11 warp 2 step end
__device__ long long g_attn_tl[16][32];   // rows 12-15: attn_fwd_unit8_kernel only
#define TL(ev, j) do { if (tl_on && lane == 0 && (j) < 32) g_attn_tl[ev][j] = clock64(); } while (0)
#else
#define TL(ev, j)
#define TRACE_T0()
#define TRACE_ADD(slot)
#define TRACE_FLUSH()
#endif

// 2^x on the FMA pipe (no MUFU): round-to-nearest split x = n + f with the 1.5 * 2^23 magic add (n sits in the low
// mantissa bits of t), cubic minimax polynomial for 2^f on [-0.5, 0.5] (max relative error 7.5e-5, a tenth of the bf16
// rounding of P), exponent insertion by an integer shift-add.  x <= -126 flushes to ~0 like ex2.approx.ftz; x up to +127 is
// exact in range (the lazy rescale keeps x below 2^40's exponent).
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 5.517132208e-02f, 2.426105440e-01f);
  p = fmaf(p, f, 6.932609677e-01f);
  p = fmaf(p, f, 9.999281168e-01f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// Which of every 8 consecutive exponentials are evaluated by poly_exp2 instead of MUFU.EX2 (bit e = element e of the octet).
// The XU pipe (16 ex2 / clk / SM) is the floor of this kernel at head_dim 64; every exponential moved to the FMA pipe
// shortens that floor as long as the issue slots last (profiles/r2_issue_bench.txt: 32 FFMA hide beside 8 MUFU).
#ifndef DESTA_ATTN_POLY_MASK
#define DESTA_ATTN_POLY_MASK 0x08
#endif
template <int E>
__device__ __forceinline__ float exp2_sel(float x) {
  return ((DESTA_ATTN_POLY_MASK >> (E & 7)) & 1) ? poly_exp2(x) : fast_exp2(x);
}

// UNIT-mode hot path: the scores arrive shifted (S' = s - m, log2 domain) -> p = 2^S', bf16 pack, row sum.  Nothing else.
// The row sum runs as fma.rn.f32x2 with a unit multiplier on two accumulator pairs (profiles/r2_issue_bench.txt: FFMA2
// issues every 2.25 cycles, FADD2 every 3.4).  MASKED (ragged last key block): columns >= nvalid contribute p = 0.
template <bool MASKED>
__device__ __forceinline__ void softmax_chunk_unit(const uint32_t (&v)[32], uint32_t (&pk)[16], uint64_t& lacc0, uint64_t& lacc1,
                                                   int nvalid) {
  const uint64_t one2 = pack_f32x2(1.0f, 1.0f);
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    float p0 = exp2_sel<0>(__uint_as_float(v[i])), p1 = exp2_sel<1>(__uint_as_float(v[i + 1]));
    float p2 = exp2_sel<2>(__uint_as_float(v[i + 2])), p3 = exp2_sel<3>(__uint_as_float(v[i + 3]));
    float p4 = exp2_sel<4>(__uint_as_float(v[i + 4])), p5 = exp2_sel<5>(__uint_as_float(v[i + 5]));
    float p6 = exp2_sel<6>(__uint_as_float(v[i + 6])), p7 = exp2_sel<7>(__uint_as_float(v[i + 7]));
    if (MASKED) {
      p0 = i < nvalid ? p0 : 0.f; p1 = i + 1 < nvalid ? p1 : 0.f; p2 = i + 2 < nvalid ? p2 : 0.f; p3 = i + 3 < nvalid ? p3 : 0.f;
      p4 = i + 4 < nvalid ? p4 : 0.f; p5 = i + 5 < nvalid ? p5 : 0.f; p6 = i + 6 < nvalid ? p6 : 0.f; p7 = i + 7 < nvalid ? p7 : 0.f;
    }
    lacc0 = ffma2(pack_f32x2(p0, p1), one2, lacc0);
    lacc1 = ffma2(pack_f32x2(p2, p3), one2, lacc1);
    lacc0 = ffma2(pack_f32x2(p4, p5), one2, lacc0);
    lacc1 = ffma2(pack_f32x2(p6, p7), one2, lacc1);
    pk[i >> 1] = pack_bf16(p0, p1);
    pk[(i >> 1) + 1] = pack_bf16(p2, p3);
    pk[(i >> 1) + 2] = pack_bf16(p4, p5);
    pk[(i >> 1) + 3] = pack_bf16(p6, p7);
  }
}

// 32 score columns (one TMEM chunk held in registers) of one query row: p = exp2(c*s - mc) packed as bf16 pairs into
// pk; running row sum / row max accumulated in `acc`.  MASKED (ragged last key block only) selects -inf for padded keys
// with predicated selects - no per-element branches.
// DROP: the row sum l is taken over the undropped probabilities (softmax normaliser); the P tile fed to P V holds
// keep ? p / (1 - p_drop) : 0, one Philox call per aligned group of 8 keys (grp0 = group index of this block's key 0).
template <bool MASKED, bool DROP>
__device__ __forceinline__ void softmax_chunk(const uint32_t (&v)[32], int ch, uint32_t (&pk)[16], float c, float mc,
                                              int kv_valid, SoftmaxAcc& acc, const DropoutCfg& dc, uint64_t grp0) {
  if (!MASKED && !DROP) {
    // hot path: packed scale-subtract and row sums (one FFMA2 / FADD2 per two elements: the step is power-capped, so
    // issued instructions - not only stalls - are what the kernel pays for)
    const uint64_t c2 = pack_f32x2(c, c), nmc2 = pack_f32x2(-mc, -mc);
    uint64_t lacc = pack_f32x2(acc.l0, acc.l1);
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float s0 = __uint_as_float(v[i]), s1 = __uint_as_float(v[i + 1]), s2 = __uint_as_float(v[i + 2]), s3 = __uint_as_float(v[i + 3]);
      acc.m0 = max3(acc.m0, s0, s1);
      acc.m1 = max3(acc.m1, s2, s3);
      float x0, x1, x2, x3;
      unpack_f32x2(ffma2(pack_f32x2(s0, s1), c2, nmc2), x0, x1);
      unpack_f32x2(ffma2(pack_f32x2(s2, s3), c2, nmc2), x2, x3);
      const float p0 = fast_exp2(x0), p1 = fast_exp2(x1), p2 = fast_exp2(x2), p3 = fast_exp2(x3);
      lacc = fadd2(lacc, fadd2(pack_f32x2(p0, p1), pack_f32x2(p2, p3)));
      pk[i >> 1] = pack_bf16(p0, p1);
      pk[(i >> 1) + 1] = pack_bf16(p2, p3);
    }
    unpack_f32x2(lacc, acc.l0, acc.l1);
    return;
  }
  uint32_t keep = 0;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    float s0 = __uint_as_float(v[i]), s1 = __uint_as_float(v[i + 1]), s2 = __uint_as_float(v[i + 2]), s3 = __uint_as_float(v[i + 3]);
    if (MASKED) {
      const int col = ch * 32 + i;
      s0 = col < kv_valid ? s0 : -INFINITY;
      s1 = col + 1 < kv_valid ? s1 : -INFINITY;
      s2 = col + 2 < kv_valid ? s2 : -INFINITY;
      s3 = col + 3 < kv_valid ? s3 : -INFINITY;
    }
    acc.m0 = max3(acc.m0, s0, s1);
    acc.m1 = max3(acc.m1, s2, s3);
    float p0 = fast_exp2(fmaf(s0, c, -mc)), p1 = fast_exp2(fmaf(s1, c, -mc));
    float p2 = fast_exp2(fmaf(s2, c, -mc)), p3 = fast_exp2(fmaf(s3, c, -mc));
    acc.l0 += p0 + p1;
    acc.l1 += p2 + p3;
    if (DROP) {
      if ((i & 7) == 0) keep = dropout_keep8(dc, grp0 + (uint64_t)((ch * 32 + i) >> 3));
      const uint32_t kb = keep >> (i & 4);
      p0 = (kb & 1) ? p0 * dc.scale : 0.f;
      p1 = (kb & 2) ? p1 * dc.scale : 0.f;
      p2 = (kb & 4) ? p2 * dc.scale : 0.f;
      p3 = (kb & 8) ? p3 * dc.scale : 0.f;
      acc.t += (p0 + p1) + (p2 + p3);
    }
    pk[i >> 1] = pack_bf16(p0, p1);
    pk[(i >> 1) + 1] = pack_bf16(p2, p3);
  }
}

// Both chunks of a 64-key block from the scores in registers; P written in place over the block's S columns [0, 32).
template <bool MASKED, bool DROP>
__device__ __forceinline__ void softmax_block(const uint32_t (&va)[32], const uint32_t (&vb)[32], uint32_t tS, float c, float mc,
                                              int kv_valid, float& l_blk, float& m_blk, float& t_blk, const DropoutCfg& dc, uint64_t grp0) {
  SoftmaxAcc acc = {0.f, 0.f, -INFINITY, -INFINITY, 0.f};
  uint32_t pk[16];
  softmax_chunk<MASKED, DROP>(va, 0, pk, c, mc, kv_valid, acc, dc, grp0);
  tmem_st_32x16(tS, pk);            // keys [0, 32): A operand of P V, K-major, two bf16 per column
  softmax_chunk<MASKED, DROP>(vb, 1, pk, c, mc, kv_valid, acc, dc, grp0);
  tmem_st_32x16(tS + 16, pk);       // keys [32, 64)
  l_blk = acc.l0 + acc.l1;
  m_blk = fmaxf(acc.m0, acc.m1);
  t_blk = acc.t;
}

// bf16 round-to-nearest-even of a float, returned as float (the shift operand holds bf16 values)
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// mbarrier wait without the out-of-line report of mbar_wait: an ABI call inside the softmax loop makes ptxas keep a spill
// frame for the registers that are live across it (96 registers per thread in this kernel).  Same 2 s guard, plain trap.
__device__ __forceinline__ void mbar_wait_lean(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// Waits of the two issuing warps.  They wait most of every step (S_{j+2} and P_j V_j are short bursts), so a plain spin on
// try_wait costs their sub-partitions' softmax warps issue slots, and the hardware sleep (mbar_wait_sleep) puts its wake-up
// latency on the hand-over chain.  Default: poll with a short nanosleep in between (DESTA_ATTN_MMA_NAP ns).
// -DDESTA_ATTN_MMA_SLEEP / -DDESTA_ATTN_MMA_SPIN select the other two for A/B runs.
#ifndef DESTA_ATTN_MMA_NAP
#define DESTA_ATTN_MMA_NAP 32
#endif
__device__ __forceinline__ void mbar_wait_nap(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  int spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(DESTA_ATTN_MMA_NAP);
    if (++spins > (1 << 25)) __trap();
  }
}
#if defined(DESTA_ATTN_MMA_SLEEP)
#define MMA_WAIT(bar, par) mbar_wait_sleep(bar, par)
#elif defined(DESTA_ATTN_MMA_SPIN)
#define MMA_WAIT(bar, par) mbar_wait_lean(bar, par)
#else
#define MMA_WAIT(bar, par) mbar_wait_nap(bar, par)
#endif

// waits of the MMA warp of the general / four-warp kernels: -DDESTA_ATTN_MMA4=0 hardware sleep (0.993 ms per encoder launch), =1 spin (0.986),
// =2 (default) poll with a short nanosleep (0.984)
#ifndef DESTA_ATTN_MMA4
#define DESTA_ATTN_MMA4 2
#endif
#if DESTA_ATTN_MMA4 == 1
#define AT_MMA_WAIT(bar, par) mbar_wait_lean(bar, par)
#elif DESTA_ATTN_MMA4 == 2
#define AT_MMA_WAIT(bar, par) mbar_wait_nap(bar, par)
#else
#define AT_MMA_WAIT(bar, par) mbar_wait_sleep(bar, par)
#endif

// One tile (128 query rows of one (batch, head)) by warps 0-5 of the CTA; every thread of the CTA must call it (it
// contains __syncthreads).  FALLBACK: called by attn_fwd_unit8_kernel for a tile whose fast pass overflowed - TMEM is
// already allocated (`tmem_given`), the barriers live in a second block (`bar_off`), warps >= 6 only join the barriers.
template <bool DROP, bool UNIT, bool FALLBACK>
__device__ __forceinline__ void attn_tile_body(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                                               const AttnArgs& a, uint8_t* smem, int bar_off, uint32_t tmem_given,
                                               int q0, int head, int batch) {
  static_assert(!(DROP && UNIT), "the dropout variant keeps the general path");
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + AT_Q_BYTES;                       // stage s: K at s*16K, V at s*16K + 8K
  uint8_t* sAx = sKV + AT_STAGES * 2 * AT_KV_BYTES;     // UNIT: [128 x 16] bf16, column 0 = -shift of the row, rest 0
  uint8_t* sBx = sAx + AT_AX_BYTES;                     // UNIT: [64 x 16] bf16, column 0 = 1, rest 0
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBx + AT_BX_BYTES + bar_off);
  uint64_t* q_full = bars;                              // 1
  uint64_t* kv_full = bars + 1;                         // [AT_STAGES]
  uint64_t* kv_empty = kv_full + AT_STAGES;             // [AT_STAGES]
  uint64_t* s_full = kv_empty + AT_STAGES;              // [3]  S_j landed in buffer j % 3
  uint64_t* p_ready = s_full + AT_SBUF;                 // [3]  P_j written over S_j (128 arrivals), O rescaled if needed
  uint64_t* pv_done = p_ready + AT_SBUF;                // [3]  P_j V_j accumulated into O
  uint64_t* shift_ready = pv_done + AT_SBUF;             // 1   UNIT: every row's first reference sits in the shift operand
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(shift_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef DESTA_ATTN_TRACE
  long long trace_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
  const int n_kv = (a.Tk + AT_BKV - 1) / AT_BKV;
#ifdef DESTA_ATTN_TRACE
  const bool tl_on = !FALLBACK && q0 == 5 * AT_BQ && head == 7 && batch == 30;
#endif

  const bool early_loads = !(a.kv_rows < AT_BKV || a.q_rows < AT_BQ);
  if (!early_loads) {
    // short sequences: rows the TMA boxes never write must read as zeros (keys: P = 0 there; queries: finite scores, so
    // that the unused rows of a 64-query Q-Former tile can never send their warps down the rescale path)
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < (AT_Q_BYTES + AT_STAGES * 2 * AT_KV_BYTES) / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  if (UNIT) {
    // un-swizzled K-major operand tiles ("interleaved" canonical layout): core matrix = 8 rows x 16 B stored contiguously,
    // 8-row groups SBO = 128 B apart, the two 16-byte K chunks LBO = rows*16 B apart.  Shift operand: all zero for now
    // (block 0 is issued before any row max is known); ones operand: element [key][0] = 1.
    uint4* z = reinterpret_cast<uint4*>(sAx);
    for (int i = threadIdx.x; i < (AT_AX_BYTES + AT_BX_BYTES) / 16; i += blockDim.x) {
      const bool ones_chunk0 = i >= AT_AX_BYTES / 16 && i < AT_AX_BYTES / 16 + AT_BKV;      // K chunk 0 of the ones tile
      z[i] = make_uint4(ones_chunk0 ? 0x00003F80u : 0u, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < AT_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < AT_SBUF; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 128); mbar_init(&pv_done[i], 1); }
    mbar_init(shift_ready, 128);
    fence_barrier_init();
    if (early_loads) {
      // full tiles: Q and the first ring of K / V leave before the TMEM allocation and the CTA barrier - the start of a tile is
      // pure latency (tools/attn_trace.py: ~1 000 cycles from the first TMA instruction to the first scores)
      mbar_arrive_expect_tx(q_full, a.q_rows * AT_HD * 2);
      tma_load_3d(sQ, &tmQ, q_full, head * AT_HD, q0, batch);
      for (int j = 0; j < AT_STAGES && j < n_kv; ++j) {
        mbar_arrive_expect_tx(&kv_full[j], 2 * a.kv_rows * AT_HD * 2);
        tma_load_3d(sKV + j * 2 * AT_KV_BYTES, &tmK, &kv_full[j], head * AT_HD, j * AT_BKV, batch);
        tma_load_3d(sKV + j * 2 * AT_KV_BYTES + AT_KV_BYTES, &tmV, &kv_full[j], head * AT_HD, j * AT_BKV, batch);
      }
    }
  }
  if (!FALLBACK && warp == 1) tmem_alloc<AT_TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = FALLBACK ? tmem_given : *tmem_slot;
  const uint32_t tmem_O = tmem_base + AT_TMEM_O;

  // Warps 0 and 1 run their loops CONVERGED (all 32 lanes wait on the mbarriers) and hand each TMA / tcgen05
  // instruction group to one lane with elect.sync: ptxas then keeps descriptors in uniform registers and issues
  // UTMALDG / UTCHMMA directly, instead of the per-instruction waterfall loop it emits inside an `if (lane == 0)` branch
  // (tools/umma_bench.cu: 48 -> 32 cycles per 128x64x16 TS instruction; the issue loop, not the tensor pipe, was the
  // bottleneck of this kernel).
  if (warp == 0) {
    if (!early_loads && elect_one()) {
      mbar_arrive_expect_tx(q_full, a.q_rows * AT_HD * 2);
      tma_load_3d(sQ, &tmQ, q_full, head * AT_HD, q0, batch);
    }
    for (int j = early_loads ? (n_kv < AT_STAGES ? n_kv : AT_STAGES) : 0; j < n_kv; ++j) {
      const int st = j % AT_STAGES;
      mbar_wait_sleep(&kv_empty[st], ((j / AT_STAGES) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&kv_full[st], 2 * a.kv_rows * AT_HD * 2);
        tma_load_3d(sKV + st * 2 * AT_KV_BYTES, &tmK, &kv_full[st], head * AT_HD, j * AT_BKV, batch);
        tma_load_3d(sKV + st * 2 * AT_KV_BYTES + AT_KV_BYTES, &tmV, &kv_full[st], head * AT_HD, j * AT_BKV, batch);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_qk = umma_idesc_bf16(AT_BQ, AT_BKV, 0, 0);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(AT_BQ, AT_HD, 0, 1);   // B (= V) is MN-major
    const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ), 16, 1024);
    const uint64_t axdesc = umma_desc_nosw(smem_u32(sAx), AT_BQ * 16, 128);
    const uint64_t bxdesc = umma_desc_nosw(smem_u32(sBx), AT_BKV * 16, 128);
    // S_j = Q K_j^T (- shift 1^T in UNIT mode) into buffer j % 3; issued three blocks ahead of the P V product of the buffer
    auto issue_qk = [&](int j) {
      const int st = j % AT_STAGES;
      { TRACE_T0(); AT_MMA_WAIT(&kv_full[st], (j / AT_STAGES) & 1); TRACE_ADD(0); }
      tcgen05_fence_after();
      const uint64_t kdesc = umma_desc_sw128(smem_u32(sKV + st * 2 * AT_KV_BYTES), 16, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < AT_HD / 16; ++k)
          umma_bf16_ss(tmem_base + (j % AT_SBUF) * AT_BKV, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
        if (UNIT) umma_bf16_ss(tmem_base + (j % AT_SBUF) * AT_BKV, axdesc, bxdesc, idesc_qk, 1);
        umma_commit(&s_full[j % AT_SBUF]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    TRACE_T0();
    if (UNIT) {
      // S_0 with a zero shift; S_1 and S_2 only once the softmax threads have put the row maxima of block 0 into the shift
      // operand (a few hundred cycles after S_0 lands, hidden behind the exponentials of block 0), so that every block but
      // the first runs the lean step.
      issue_qk(0);
      if (n_kv > 1) {
        mbar_wait(shift_ready, 0);
        tcgen05_fence_after();
        for (int j = 1; j < AT_SBUF && j < n_kv; ++j) issue_qk(j);
      }
    } else {
      for (int j = 0; j < AT_SBUF && j < n_kv; ++j) issue_qk(j);
    }
    for (int j = 0; j < n_kv; ++j) {
      const int st = j % AT_STAGES, b = j % AT_SBUF;
      { TRACE_T0(); AT_MMA_WAIT(&p_ready[b], (j / AT_SBUF) & 1); TRACE_ADD(1); }     // P_j in TMEM (over S_j), O rescaled if needed
      TL(0, j);
      tcgen05_fence_after();
      // V stage: 64 key rows x 128 B (64 head-dim values) -> MN-major B operand, 8-row groups 1024 B apart
      const uint64_t vdesc = umma_desc_sw128(smem_u32(sKV + st * 2 * AT_KV_BYTES + AT_KV_BYTES), 1024, 1024);
      const uint32_t tP = tmem_base + b * AT_BKV;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < AT_BKV / 16; ++k)     // 16 keys = 8 packed columns of P per UMMA
          umma_bf16_ts(tmem_O, tP + k * 8, vdesc + (uint64_t)(k * 128), idesc_pv, (j | k) != 0);
        umma_commit(&kv_empty[st]);
        umma_commit(&pv_done[b]);
      }
      __syncwarp();
      TL(1, j);
      if (j + AT_SBUF < n_kv) { issue_qk(j + AT_SBUF); TL(2, j + AT_SBUF); }   // in-order tensor pipe: overwrites buffer b after P_j V_j
    }
    TRACE_ADD(3);
  } else if (warp < 6 && q0 + (warp & 3) * 32 >= a.Tq) {
    // ---- a softmax warp whose 32 rows all lie past Tq (the Q-Former's 64-query tiles: two of four warps; the encoder's last
    //      q-tile, 1500 = 11 * 128 + 92: one of four).  Rows are independent (O[r] = sum_k P[r,k] V[k]) and these are never
    //      stored, so whatever the P columns of these lanes hold is harmless: the warp only keeps the hand-over protocol in
    //      step (an arrival per block, paced by the scores it would have read) and spends no MUFU / issue slots on padding. ----
    if (UNIT) mbar_arrive(shift_ready);
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait_sleep(&s_full[j % AT_SBUF], (j / AT_SBUF) & 1);
      mbar_arrive(&p_ready[j % AT_SBUF]);
    }
  } else if (warp < 6) {
    // ---- softmax warps: TMEM lane quarter = warp % 4, one query row per thread ----
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                 // row inside the tile
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float c = UNIT ? 1.0f : a.scale_log2;
    // everything below lives in the log2 domain: m_ref = reference exponent of the row (P, l, O are relative to 2^m_ref)
    float m_ref = -INFINITY, l_run = 0.f;
    float t_run = 0.f;                                  // DROP: running sum of the dropped, rescaled probabilities (same reference)
    // UNIT: shift folded into S by the MMA.  sh_cur = shift this thread last wrote to the operand tile; bk0..2 = the shift
    // block j % 3 was (or will be) computed with - a QK^T issued after this thread's arrival on p_ready[j-3] sees the
    // operand as it was then.
    float sh_cur = 0.f, bk0 = 0.f, bk1 = 0.f, bk2 = 0.f;
    __nv_bfloat16* ax_elem = reinterpret_cast<__nv_bfloat16*>(sAx + (r >> 3) * 128 + (r & 7) * 16);
    // dropout group index of key 0 for this thread's row: row = (batch*heads + head)*Tq + query
    const uint64_t grp_row = DROP ? (((uint64_t)batch * a.heads + head) * a.Tq + (q0 + r)) * (uint64_t)((a.Tk + 7) >> 3) : 0;
    // New shift for the row: only legal while no S block that reads the operand is in flight, i.e. once S_{j+1} and S_{j+2}
    // (issued earlier, with the old shift) have landed; S_{j+3} is issued after this thread's next p_ready arrival.
    auto set_shift = [&](int j, float sh) {
      if (j + 1 < n_kv) mbar_wait(&s_full[(j + 1) % AT_SBUF], ((j + 1) / AT_SBUF) & 1);
      if (j + 2 < n_kv) mbar_wait(&s_full[(j + 2) % AT_SBUF], ((j + 2) / AT_SBUF) & 1);
      *ax_elem = __float2bfloat16_rn(-sh);
      fence_proxy_async_smem();
      sh_cur = sh;
    };
    // Software pipeline over 32-column chunks with three register sets: while the exponentials of one chunk occupy the
    // MUFU pipe, the tcgen05.ld of the next chunk (the second half of this block, then the first half of the NEXT block,
    // which the three S buffers make available early) is already in flight, and the tcgen05.st of the previous block's P
    // is only waited for - and announced to the MMA warp - one chunk later.  `ca` = this block's keys [0,32) (already
    // requested), `cb` = keys [32,64), `na` = the next block's keys [0,32).

    // announce P_{j-1} (its stores were issued 32 exponentials ago); S_{j+2} will be computed with the current shift
    auto announce_prev = [&](int j) {
      tmem_wait_st();
      tcgen05_fence_before();
      mbar_arrive(&p_ready[(j - 1) % AT_SBUF]);
      if (UNIT) {
        const int pb = (j - 1) % AT_SBUF;
        bk0 = pb == 0 ? sh_cur : bk0; bk1 = pb == 1 ? sh_cur : bk1; bk2 = pb == 2 ? sh_cur : bk2;
      }
    };
    auto request_next = [&](int j, uint32_t (&na)[32]) {
      if (j + 1 < n_kv) {
        { TRACE_T0(); mbar_wait(&s_full[(j + 1) % AT_SBUF], ((j + 1) / AT_SBUF) & 1); if (warp == 2) TRACE_ADD(2); if (warp == 4) TRACE_ADD(5); if (warp == 5) TRACE_ADD(6); if (warp == 3) TRACE_ADD(7); }
        TL(5 + warp, j + 1);
        tcgen05_fence_after();
        tmem_ld_32x32(tmem_base + ((j + 1) % AT_SBUF) * AT_BKV + lane_off, na);
      }
    };
    // rare: raise the reference to m_blk, rescale O and l, redo the block's exponentials from the registers
    auto rescale = [&](int j, uint32_t (&ca)[32], uint32_t (&cb)[32], uint32_t tS, float baked, int kv_valid, uint64_t grp0,
                       float m_blk, float& l_blk, float& t_blk, bool from_fast_loop) {
      float m_new = fmaxf(m_ref, m_blk);
      if (UNIT) m_new = bf16_round(m_new);
      if (j > 0) {
        mbar_wait(&pv_done[(j - 1) % AT_SBUF], ((j - 1) / AT_SBUF) & 1);   // P_{j-1} V_{j-1} must have landed in O
        tcgen05_fence_after();
        const float f = exp2f(m_ref - m_new);
        l_run *= f;
        t_run *= f;
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t o[16];
          tmem_ld_32x16(tmem_O + lane_off + ch * 16, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
          tmem_st_32x16(tmem_O + lane_off + ch * 16, o);
        }
      }
      m_ref = m_new;
      float m_unused;
      softmax_block<true, DROP>(ca, cb, tS, c, m_ref - baked, kv_valid, l_blk, m_unused, t_blk, a.drop, grp0);
      if (UNIT && m_ref != sh_cur) {
        if (from_fast_loop) bk0 = bk1 = bk2 = sh_cur;   // the fast loop does not track the slots: S_{j+1}, S_{j+2} carry the old shift
        set_shift(j, m_ref);
      }
    };
    // general step: any block (first block, ragged last block, dropout, blocks whose S carries a stale shift)
    auto step = [&](int j, uint32_t (&ca)[32], uint32_t (&cb)[32], uint32_t (&na)[32]) {
      const int b = j % AT_SBUF;
      const uint64_t grp0 = grp_row + (uint64_t)(j * (AT_BKV / 8));
      const uint32_t tS = tmem_base + b * AT_BKV + lane_off;
      const int kv_valid = a.Tk - j * AT_BKV;
      const float baked = UNIT ? (b == 0 ? bk0 : (b == 1 ? bk1 : bk2)) : 0.f;
      TL(1 + warp, j);
      tmem_wait_ld();                                   // ca
      reg_fence32(ca);
      tmem_ld_32x32(tS + 32, cb);
      if (j == 0) {
        // first block: no reference yet -> its own row max
        tmem_wait_ld();
        reg_fence32(cb);
        float m0 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          m0 = max3(m0, (i < kv_valid) ? __uint_as_float(ca[i]) : -INFINITY, (i + 1 < kv_valid) ? __uint_as_float(ca[i + 1]) : -INFINITY);
          m0 = max3(m0, (32 + i < kv_valid) ? __uint_as_float(cb[i]) : -INFINITY, (33 + i < kv_valid) ? __uint_as_float(cb[i + 1]) : -INFINITY);
        }
        m_ref = UNIT ? bf16_round(m0) : m0 * c;
        if (UNIT) {
          // nothing reads the shift operand yet (S_1, S_2 are issued behind shift_ready): publish the row's reference now
          *ax_elem = __float2bfloat16_rn(-m_ref);
          fence_proxy_async_smem();
          sh_cur = bk1 = bk2 = m_ref;
          mbar_arrive(shift_ready);
        }
      }
      // exp pass against the reference of the PREVIOUS blocks; the block max comes out of the same pass
      const float off = m_ref - baked;                  // p = exp2(c * S - off)
      SoftmaxAcc acc = {0.f, 0.f, -INFINITY, -INFINITY, 0.f};
      uint32_t pk[16];
      if (kv_valid >= AT_BKV) softmax_chunk<false, DROP>(ca, 0, pk, c, off, kv_valid, acc, a.drop, grp0);
      else softmax_chunk<true, DROP>(ca, 0, pk, c, off, kv_valid, acc, a.drop, grp0);
      if (j > 0) announce_prev(j);
      tmem_st_32x16(tS, pk);                            // P keys [0,32) in place over S columns [0,16)
      tmem_wait_ld();                                   // cb
      reg_fence32(cb);
      request_next(j, na);
      if (kv_valid >= AT_BKV) softmax_chunk<false, DROP>(cb, 1, pk, c, off, kv_valid, acc, a.drop, grp0);
      else softmax_chunk<true, DROP>(cb, 1, pk, c, off, kv_valid, acc, a.drop, grp0);
      tmem_st_32x16(tS + 16, pk);                       // P keys [32,64) over S columns [16,32)
      float l_blk = acc.l0 + acc.l1, t_blk = acc.t;
      const float m_blk = fmaxf(acc.m0, acc.m1) * c + baked;      // block max in the log2 domain
      // P is bf16 and O / l are fp32, so exponents up to 2^20 above the reference are harmless; only beyond that is the
      // reference raised (FA-style lazy rescale; the common case never takes the slow path)
      const bool need = (m_blk - m_ref) > 20.0f;
      if (__builtin_expect(__any_sync(0xffffffffu, need), 0)) rescale(j, ca, cb, tS, baked, kv_valid, grp0, m_blk, l_blk, t_blk, false);
      l_run += l_blk;
      if (DROP) t_run += t_blk;
    };
    // UNIT fast step: a full block whose S already carries this row's current reference as shift (S' = s - m_ref):
    // p = ex2(S'), bf16 pack, row sum - nothing else.  A row sum above 2^40 (some p far above the reference, inf included)
    // sends the warp through the same rescale as the general step; below that bf16 P / fp32 l, O lose nothing.
    // `tracked` = false inside the fast loop (all three slots equal the current shift there, nothing to maintain).
    // Returns true (warp-uniform) if the block had to raise the reference.
    auto step_unit = [&](int j, uint32_t (&ca)[32], uint32_t (&cb)[32], uint32_t (&na)[32], bool tracked, int kv_valid) -> bool {
      const uint32_t tS = tmem_base + (j % AT_SBUF) * AT_BKV + lane_off;
      TL(1 + warp, j);
      tmem_wait_ld();                                   // ca
      reg_fence32(ca);
      tmem_ld_32x32(tS + 32, cb);
      uint64_t lacc0 = 0, lacc1 = 0;                    // (0.f, 0.f)
      uint32_t pk[16];
      if (kv_valid >= 32) softmax_chunk_unit<false>(ca, pk, lacc0, lacc1, 32);
      else softmax_chunk_unit<true>(ca, pk, lacc0, lacc1, kv_valid);
      if (tracked) announce_prev(j); else { tmem_wait_st(); tcgen05_fence_before(); mbar_arrive(&p_ready[(j - 1) % AT_SBUF]); }
      tmem_st_32x16(tS, pk);
      tmem_wait_ld();                                   // cb
      reg_fence32(cb);
      request_next(j, na);
      if (kv_valid >= AT_BKV) softmax_chunk_unit<false>(cb, pk, lacc0, lacc1, 32);
      else if (kv_valid > 32) softmax_chunk_unit<true>(cb, pk, lacc0, lacc1, kv_valid - 32);
      else {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;        // ragged block with <= 32 keys: no exponentials for the padded half
      }
      tmem_st_32x16(tS + 16, pk);
      float la, lb, lc, ld;
      unpack_f32x2(lacc0, la, lb);
      unpack_f32x2(lacc1, lc, ld);
      float l_blk = (la + lb) + (lc + ld);
      const bool need = __any_sync(0xffffffffu, !(l_blk <= 1.0995116e12f));
      if (__builtin_expect(need, 0)) {
        float mx = -INFINITY;                           // the fast pass keeps no maximum: take it from the registers
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          mx = max3(mx, (i < kv_valid) ? __uint_as_float(ca[i]) : -INFINITY, (i + 1 < kv_valid) ? __uint_as_float(ca[i + 1]) : -INFINITY);
          mx = max3(mx, (32 + i < kv_valid) ? __uint_as_float(cb[i]) : -INFINITY, (33 + i < kv_valid) ? __uint_as_float(cb[i + 1]) : -INFINITY);
        }
        float t_unused = 0.f;
        rescale(j, ca, cb, tS, m_ref, kv_valid, 0, mx + m_ref, l_blk, t_unused, !tracked);
      }
      l_run += l_blk;
      return need;
    };
    // the fast step needs, for EVERY row of the warp, a full block computed with the row's current reference as shift
    auto unit_ready = [&](int j) {
      if (!UNIT) return false;
      const int b = j % AT_SBUF;
      const float baked = b == 0 ? bk0 : (b == 1 ? bk1 : bk2);
      return __all_sync(0xffffffffu, baked == m_ref) != 0;
    };
    uint32_t ra[32], rb[32], rc[32];
    TRACE_T0();
    mbar_wait(&s_full[0], 0);
    tcgen05_fence_after();
    tmem_ld_32x32(tmem_base + lane_off, ra);
    int j = 0;
    if (UNIT) {
      // block 0 was issued before any row max was known: general step (shift 0 baked in); it publishes the shift that
      // S_1, S_2, ... are computed with
      step(0, ra, rb, rc);
      if (n_kv > 1) { if (unit_ready(1)) step_unit(1, rc, rb, ra, true, a.Tk - AT_BKV); else step(1, rc, rb, ra); }
      if (n_kv > 2) { if (unit_ready(2)) step_unit(2, ra, rb, rc, true, a.Tk - 2 * AT_BKV); else step(2, ra, rb, rc); }
      j = n_kv < 3 ? n_kv : 3;
      // fast loop over the full blocks: two steps per iteration (register sets alternate), no per-step bookkeeping.
      // Entered only if every row of the warp has all in-flight blocks on its current reference; left on a rescale.
      const int n_full = a.Tk / AT_BKV;                 // blocks without padded keys
      bool fast_ok = __all_sync(0xffffffffu, bk0 == m_ref && bk1 == m_ref && sh_cur == m_ref);
      if (fast_ok) bk2 = sh_cur;                        // S_5 is issued behind this thread's next arrival (p_ready[2])
      while (fast_ok && j + 1 < n_full) {               // j is odd here
        if (step_unit(j, rc, rb, ra, false, AT_BKV)) { j += 1; break; }
        if (step_unit(j + 1, ra, rb, rc, false, AT_BKV)) { j += 2; break; }
        j += 2;
      }
    }
    if (UNIT) {
      for (; j < n_kv; ++j) {                           // blocks after a rescale inside the fast loop, ragged last block
        const int kvv = a.Tk - j * AT_BKV;
        if (j & 1) { if (unit_ready(j)) step_unit(j, rc, rb, ra, true, kvv); else step(j, rc, rb, ra); }
        else { if (unit_ready(j)) step_unit(j, ra, rb, rc, true, kvv); else step(j, ra, rb, rc); }
      }
    } else {
      for (; j < n_kv; j += 2) {
        step(j, ra, rb, rc);
        if (j + 1 < n_kv) step(j + 1, rc, rb, ra);
      }
    }
    tmem_wait_st();                 // the last P (and a rescaled O) have landed in TMEM
    tcgen05_fence_before();         // orders this thread's TMEM accesses before the MMA warp's next MMAs
    mbar_arrive(&p_ready[(n_kv - 1) % AT_SBUF]);
    if (warp == 2) TRACE_ADD(4);
    // ---- epilogue: O / l -> bf16, one 128-byte row per thread ----
    mbar_wait(&pv_done[(n_kv - 1) % AT_SBUF], ((n_kv - 1) / AT_SBUF) & 1);
    tcgen05_fence_after();
    const float inv = 1.0f / l_run;
    const int row = q0 + r;
    if (a.lse && row < a.Tq) a.lse[((size_t)batch * a.heads + head) * a.Tq + row] = m_ref + log2f(l_run);
    if (DROP && a.drop.row_ratio && row < a.Tq) a.drop.row_ratio[((size_t)batch * a.heads + head) * a.Tq + row] = t_run * inv;
    __nv_bfloat16* orow = a.out + ((size_t)batch * a.Tq + row) * a.ldo + head * AT_HD;
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t o[32];
      tmem_ld_32x32(tmem_O + lane_off + ch * 32, o);
      tmem_wait_ld();
      if (row < a.Tq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
          u.y = pack_bf16(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
          u.z = pack_bf16(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
          u.w = pack_bf16(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + ch * 32 + i) = u;
        }
      }
    }
  }
  TRACE_FLUSH();
  tcgen05_fence_before();
  __syncthreads();
  if (!FALLBACK && warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<AT_TMEM_COLS>(tmem_base);
  }
  if (FALLBACK) {
    // a persistent CTA may recompute several tiles: hand the barrier block back as raw memory (every phase has completed:
    // the epilogue waited for the last P V, whose commits are the last arrivals) so that the next call may initialise it
    if (warp == 0 && lane == 0)
      for (int i = 0; i < 2 + 2 * AT_STAGES + 3 * AT_SBUF; ++i)
        asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&bars[i])) : "memory");
    __syncthreads();
  }
}

template <bool DROP, bool UNIT>
__global__ void __launch_bounds__(AT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];   // SWIZZLE_128B tiles need 1024-byte alignment
  if (smem_u32(smem) & 1023) __trap();
  stagger_first_wave(a);
  attn_tile_body<DROP, UNIT, false>(tmQ, tmK, tmV, a, smem, 0, 0u, blockIdx.x * AT_BQ, blockIdx.y, blockIdx.z);
}

// ================================================================================================================
// The encoder's kernel (scale * log2(e) == 1: q arrives pre-scaled, scores are base-2 exponents): EIGHT softmax warps.
// tools/attn_trace.py and tools/contention_bench.cu on the four-warp kernel above show two things: (1) the UMMA stream is
// paced by SHARED MEMORY, not by the tensor pipe - an SS-form 128x64x16 step reads 4 KB of Q and 2 KB of K (48 cycles at
// 128 B / clk against 32 cycles of tensor time), and with the shift step of the UNIT variant two co-resident CTAs need
// 2 x (5 x 48 + 4 x 32) = 736 cycles per 64-key block against a MUFU floor of 896 (1024 x 7/8); (2) with two softmax warps
// per sub-partition the exp / pack / sum loop is latency-bound (1 150 - 1 250 cycles per block even without any MMA), and
// the hand-over chain  P_j announced -> MMA warp -> P_j V_j -> S_{j+3} -> tcgen05.ld  was as long as the look-ahead.
// This kernel changes all three:
//   * Q lives in TENSOR MEMORY (32 packed columns, written once per tile by the threads that own the rows), so S = Q K^T is
//     a TS-form product like P V: 2 KB of shared-memory reads per step instead of 6 KB, and the shift is subtracted by the
//     softmax threads (one packed add.f32x2 per two scores) instead of a fifth, shared-memory-heavy MMA step:
//     2 x 8 x 32 = 512 cycles of UMMA stream per block pair.
//   * warps 2-5 take score columns [0, 32) of every 64-key block, warps 6-9 columns [32, 64) (a warp pair w, w + 4 shares a
//     TMEM lane quarter): FOUR softmax warps per sub-partition; the waits around a block (tcgen05.wait, mbarrier, TMEM
//     store) of one warp are covered by three others, and there are issue slots to spare for the packed subtract.
//   * TMEM: two score buffers, ONE packed-probability buffer, Q, O  (64 + 64 + 32 + 32 + 64 = 256 columns).  A score buffer
//     is handed back (s_free, 256 arrivals) as soon as every thread holds its columns in registers - at the START of step j -
//     so S_{j+2} is issued a whole step before P_j V_j and the softmax warps never wait for scores; P_j V_j is issued a
//     quarter step after P_j is stored and has long finished when P_{j+1} is stored over it.
//   * the row reference is the maximum of block 0 (the two half-row maxima are exchanged through shared memory, named
//     barrier of the warp pair) and stays fixed: P is bf16 and l / O are fp32, so a later score above the reference costs no
//     precision until a sum overflows.  Every thread checks its final row sum (<= 2^100, finite) and its O values; a
//     violation sets a CTA flag and the WHOLE tile is recomputed by the general path (attn_tile_body<false, false, true>:
//     running max, lazy rescale) right here in the same CTA.  Scores that exceed the first 64 keys' maximum by 100 octaves
//     do not occur in a trained encoder; tests force the path with ramped keys.
//   * row sums accumulate over the whole tile in two packed fp32x2 registers per thread and are combined by the warp pair
//     once, in the epilogue; each thread normalises and stores its 32 output columns.
// ================================================================================================================
constexpr int AT8_THREADS = 320;
constexpr int AT8_MAX_IT = 128;                      // tiles per CTA in the persistent launch (host: otherwise one tile per CTA)
constexpr int AT8_SMEM = AT_SMEM + 256 + 2048 + AT8_MAX_IT;   // + fallback barrier block + max / sum exchange (2 x 256 floats) + tile flags
constexpr int AT8_TMEM_P = 2 * AT_BKV;               // TMEM: S buffers [0,64) [64,128), packed P [128,160), packed Q [160,192), O [192,256)
constexpr int AT8_TMEM_Q = AT8_TMEM_P + 32;
constexpr int AT8_KST = 4, AT8_VST = 6;              // K / V rings of 8 KB stages (together the 80 KB of the general kernel's ring)
static_assert((AT8_KST + AT8_VST) * AT_KV_BYTES == AT_STAGES * 2 * AT_KV_BYTES, "the fallback path reuses the same shared memory");
static_assert((1 + 2 * AT8_KST + 2 * AT8_VST + 7) * 8 + 16 <= 256, "barrier block");

// named barrier of one warp pair (ids 1-4, 64 threads); immediate ids keep the CTA's barrier count at five
__device__ __forceinline__ void pair_sync(int quarter) {
  switch (quarter) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}

// columns [I0, I1) of the thread's 32-column chunk: p = 2^(s - m) (nm2 = (-m, -m)), bf16 pack, row sum.
// MASKED: columns >= nvalid are padded keys and contribute p = 0.
template <int I0, int I1, bool MASKED>
__device__ __forceinline__ void softmax8_cols(const uint32_t (&v)[32], uint32_t (&pk)[16], uint64_t& lacc0, uint64_t& lacc1,
                                              int nvalid, uint64_t nm2, uint64_t one2) {
#pragma unroll
  for (int i = I0; i < I1; i += 8) {
    float x0, x1, x2, x3, x4, x5, x6, x7;
    unpack_f32x2(fadd2(pack_f32x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), nm2), x0, x1);
    unpack_f32x2(fadd2(pack_f32x2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])), nm2), x2, x3);
    unpack_f32x2(fadd2(pack_f32x2(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])), nm2), x4, x5);
    unpack_f32x2(fadd2(pack_f32x2(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])), nm2), x6, x7);
    float p0 = exp2_sel<0>(x0), p1 = exp2_sel<1>(x1), p2 = exp2_sel<2>(x2), p3 = exp2_sel<3>(x3);
    float p4 = exp2_sel<4>(x4), p5 = exp2_sel<5>(x5), p6 = exp2_sel<6>(x6), p7 = exp2_sel<7>(x7);
    if (MASKED) {
      p0 = i < nvalid ? p0 : 0.f; p1 = i + 1 < nvalid ? p1 : 0.f; p2 = i + 2 < nvalid ? p2 : 0.f; p3 = i + 3 < nvalid ? p3 : 0.f;
      p4 = i + 4 < nvalid ? p4 : 0.f; p5 = i + 5 < nvalid ? p5 : 0.f; p6 = i + 6 < nvalid ? p6 : 0.f; p7 = i + 7 < nvalid ? p7 : 0.f;
    }
    lacc0 = ffma2(pack_f32x2(p0, p1), one2, lacc0);
    lacc1 = ffma2(pack_f32x2(p2, p3), one2, lacc1);
    lacc0 = ffma2(pack_f32x2(p4, p5), one2, lacc0);
    lacc1 = ffma2(pack_f32x2(p6, p7), one2, lacc1);
    pk[i >> 1] = pack_bf16(p0, p1);
    pk[(i >> 1) + 1] = pack_bf16(p2, p3);
    pk[(i >> 1) + 2] = pack_bf16(p4, p5);
    pk[(i >> 1) + 3] = pack_bf16(p6, p7);
  }
}

// the rare path, kept out of line so that its registers (three 32-column sets) do not shape the fast kernel's allocation
static __device__ __noinline__ void attn_tile_fallback(const CUtensorMap* tmQ, const CUtensorMap* tmK, const CUtensorMap* tmV,
                                                       const AttnArgs* a, uint8_t* smem, uint32_t tmem_base, int q0, int head, int batch) {
  attn_tile_body<false, false, true>(*tmQ, *tmK, *tmV, *a, smem, 256, tmem_base, q0, head, batch);
}

struct TileCoord { int q0, head, batch; };
__device__ __forceinline__ TileCoord tile_coord(const AttnArgs& a, int t) {
  TileCoord c;
  const int qt = t % a.nq, hb = t / a.nq;
  c.q0 = qt * AT_BQ; c.head = hb % a.heads; c.batch = hb / a.heads;
  return c;
}

// PERSISTENT when the host launches fewer CTAs than tiles (two per SM; only for n_kv % 24 == 0, which makes every mbarrier
// parity but those of the once-per-tile Q hand-over repeat from tile to tile): a CTA walks the tiles t = blockIdx.x,
// blockIdx.x + gridDim.x, ... without giving up its TMEM columns, barriers or warps, the K / V rings run on across the tile
// boundary (the next tile's first blocks and its Q are in shared memory before the current tile's epilogue starts), and a
// tile that overflowed is only noted and recomputed after the loop.  The timeline (tools/attn_trace.py) of the
// one-tile-per-CTA version showed ~3 000 cycles from CTA start to the first scores and a 1 300-cycle stall of the neighbour
// CTA at every turnover, a quarter of a 34 000-cycle tile together with the epilogue.
__global__ void __launch_bounds__(AT8_THREADS, 2)
attn_fwd_unit8_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023) __trap();
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AT_Q_BYTES;                        // K ring: AT8_KST stages of 8 KB (a stage is free once S_j = Q K_j^T has run)
  uint8_t* sV = sK + AT8_KST * AT_KV_BYTES;             // V ring: AT8_VST stages of 8 KB (free once P_j V_j has run, a step and a half later)
  uint8_t* sTail = sQ + AT_Q_BYTES + AT_STAGES * 2 * AT_KV_BYTES + AT_AX_BYTES + AT_BX_BYTES;   // same offsets as the general kernel
  uint64_t* bars = reinterpret_cast<uint64_t*>(sTail);
  uint64_t* q_full = bars;                              // Q tile landed in shared memory (once per tile)
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + AT8_KST;
  uint64_t* v_full = k_empty + AT8_KST;
  uint64_t* v_empty = v_full + AT8_VST;
  uint64_t* s_full = v_empty + AT8_VST;                 // [2]  S_j landed in score buffer j % 2
  uint64_t* s_free = s_full + 2;                        // [2]  256 arrivals: every thread holds its part of S_j in registers
  uint64_t* p_ready = s_free + 2;                       // 256 arrivals: P_j stored in the probability buffer
  uint64_t* pv_done = p_ready + 1;                      // P_j V_j accumulated: the probability buffer is free again
  uint64_t* q_ready = pv_done + 1;                      // 128 arrivals: Q copied to tensor memory (once per tile)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_ready + 1);
  float* xm = reinterpret_cast<float*>(sTail + 512);    // [2][128] half-row maxima of block 0
  float* xl = xm + 256;                                 // [2][128] half-row sums of the tile
  uint8_t* bad_tiles = sTail + 512 + 2048;              // [AT8_MAX_IT] tile `it` of this CTA overflowed: recompute after the loop

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv = (a.Tk + AT_BKV - 1) / AT_BKV;
  const int n_it = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // tiles of this CTA
#ifdef DESTA_ATTN_TRACE
  const int tl_it = n_it > 20 ? 20 : 0;
  bool tl_on = false;
#endif

  const bool short_seq = a.kv_rows < AT_BKV || a.q_rows < AT_BQ;     // rows the TMA boxes never write must read as zeros
  if (short_seq) {
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < (AT_Q_BYTES + AT_STAGES * 2 * AT_KV_BYTES) / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  for (int i = threadIdx.x; i < AT8_MAX_IT; i += blockDim.x) bad_tiles[i] = 0;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < AT8_KST; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < AT8_VST; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 256); }
    mbar_init(p_ready, 256);
    mbar_init(pv_done, 1);
    mbar_init(q_ready, 128);
    fence_barrier_init();
    if (!short_seq) {
      // full tiles: the first load leaves before the TMEM allocation and the CTA barrier (a CTA's start is pure latency)
      const TileCoord c = tile_coord(a, blockIdx.x);
      mbar_arrive_expect_tx(q_full, a.q_rows * AT_HD * 2);
      tma_load_3d(sQ, &tmQ, q_full, c.head * AT_HD, c.q0, c.batch);
    }
  }
  if (warp == 1) tmem_alloc<AT_TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + AT_TMEM_O;

  // Warps 0 and 1 share the tensor-core issue: warp 1 issues S_j = Q K_j^T and refills the K ring, warp 0 issues O += P_j V_j
  // and refills the V ring (and Q).  The two products are independent (S_{j+2} only needs the score buffer of block j back,
  // P_j V_j only P_j), and each warp's own stream stays ordered.  Both loops are unrolled over 12 blocks (K ring 4, V ring 6,
  // score buffers 2) so that stages, buffers and all mbarrier parities but one are compile-time constants.
  if (warp == 0) {
    constexpr uint32_t idesc_pv = umma_idesc_bf16(AT_BQ, AT_HD, 0, 1);   // B (= V) is MN-major
    const uint32_t tP = tmem_base + AT8_TMEM_P;
    // V stage: 64 key rows x 128 B (64 head-dim values) -> MN-major B operand, 8-row groups 1024 B apart
    const uint64_t vdesc0 = umma_desc_sw128(smem_u32(sV), 1024, 1024);
    {
      const TileCoord c = tile_coord(a, blockIdx.x);
      if (elect_one()) {
        if (short_seq) {
          mbar_arrive_expect_tx(q_full, a.q_rows * AT_HD * 2);
          tma_load_3d(sQ, &tmQ, q_full, c.head * AT_HD, c.q0, c.batch);
        }
        for (int jj = 0; jj < AT8_VST && jj < n_kv; ++jj) {
          mbar_arrive_expect_tx(&v_full[jj], a.kv_rows * AT_HD * 2);
          tma_load_3d(sV + jj * AT_KV_BYTES, &tmV, &v_full[jj], c.head * AT_HD, jj * AT_BKV, c.batch);
        }
      }
      __syncwarp();
    }
    for (int it = 0; it < n_it; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const TileCoord c = tile_coord(a, tile);
      const bool has_next = it + 1 < n_it;
      const TileCoord cn = tile_coord(a, has_next ? tile + (int)gridDim.x : tile);
#ifdef DESTA_ATTN_TRACE
      tl_on = blockIdx.x == 37 && it == tl_it;
#endif
      if (has_next) {
        // the Q tile in shared memory is free once every row has been copied to tensor memory: fetch the next tile's Q now
        MMA_WAIT(q_ready, it & 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(q_full, a.q_rows * AT_HD * 2);
          tma_load_3d(sQ, &tmQ, q_full, cn.head * AT_HD, cn.q0, cn.batch);
        }
        __syncwarp();
      }
      for (int jb = 0; jb < n_kv; jb += 12) {
#pragma unroll
        for (int u = 0; u < 12; ++u) {
          const int j = jb + u;
          if (j >= n_kv) break;
          MMA_WAIT(&v_full[u % AT8_VST], (u / AT8_VST) & 1);
          MMA_WAIT(p_ready, u & 1);
          TL(0, j);
          tcgen05_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < AT_BKV / 16; ++k)     // 16 keys = 8 packed columns of P per UMMA
              umma_bf16_ts(tmem_O, tP + k * 8, vdesc0 + (uint64_t)((u % AT8_VST) * (AT_KV_BYTES >> 4) + k * 128), idesc_pv, (j | k) != 0);
            umma_commit(&v_empty[u % AT8_VST]);
            umma_commit(pv_done);
          }
          __syncwarp();
          TL(1, j);
          // V of block j + 5 (possibly of the next tile) into the stage P_{j-1} V_{j-1} released a step ago
          int nj = j + AT8_VST - 1;
          const bool wrap = nj >= n_kv;
          if (wrap) nj -= n_kv;
          if ((j >= 1 || it > 0) && (!wrap || has_next)) {
            const int vs = (u + AT8_VST - 1) % AT8_VST;
            MMA_WAIT(&v_empty[vs], u >= 1 ? ((u - 1) / AT8_VST) & 1 : 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&v_full[vs], a.kv_rows * AT_HD * 2);
              tma_load_3d(sV + vs * AT_KV_BYTES, &tmV, &v_full[vs], (wrap ? cn.head : c.head) * AT_HD, nj * AT_BKV, wrap ? cn.batch : c.batch);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_qk = umma_idesc_bf16(AT_BQ, AT_BKV, 0, 0);
    const uint32_t tQ = tmem_base + AT8_TMEM_Q;
    const uint64_t kdesc0 = umma_desc_sw128(smem_u32(sK), 16, 1024);
    {
      const TileCoord c = tile_coord(a, blockIdx.x);
      if (elect_one()) {
        for (int jj = 0; jj < AT8_KST && jj < n_kv; ++jj) {
          mbar_arrive_expect_tx(&k_full[jj], a.kv_rows * AT_HD * 2);
          tma_load_3d(sK + jj * AT_KV_BYTES, &tmK, &k_full[jj], c.head * AT_HD, jj * AT_BKV, c.batch);
        }
      }
      __syncwarp();
    }
    for (int it = 0; it < n_it; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const TileCoord c = tile_coord(a, tile);
      const bool has_next = it + 1 < n_it;
      const TileCoord cn = tile_coord(a, has_next ? tile + (int)gridDim.x : tile);
#ifdef DESTA_ATTN_TRACE
      tl_on = blockIdx.x == 37 && it == tl_it;
#endif
      MMA_WAIT(q_ready, it & 1);
      TL(13, 2);
      tcgen05_fence_after();
      // S_j = Q K_j^T: A = Q from tensor memory (8 packed columns per K = 16 step), B = K_j, K-major in shared memory.
      // Issued as soon as the score buffer of block j - 2 is back (every softmax thread holds its columns: start of step j - 2).
      int ob = 0;                                        // (jb / 12) & 1: the only run-time part of the K-ring parities
      for (int jb = 0; jb < n_kv; jb += 12, ob ^= 1) {
#pragma unroll
        for (int u = 0; u < 12; ++u) {
          const int j = jb + u;
          if (j >= n_kv) break;
          if (j >= 2 || it > 0) MMA_WAIT(&s_free[u & 1], ((u >> 1) + 1) & 1);
          TL(14, j);
          MMA_WAIT(&k_full[u % AT8_KST], (ob + u / AT8_KST) & 1);
          TL(15, j);
          tcgen05_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < AT_HD / 16; ++k)
              umma_bf16_ts(tmem_base + (u & 1) * AT_BKV, tQ + k * 8, kdesc0 + (uint64_t)((u % AT8_KST) * (AT_KV_BYTES >> 4) + 2 * k), idesc_qk, k != 0);
            umma_commit(&s_full[u & 1]);
            umma_commit(&k_empty[u % AT8_KST]);
          }
          __syncwarp();
          TL(2, j);
          // K of block j + 3 (possibly of the next tile) into the stage S_{j-1} released a step ago
          int nj = j + AT8_KST - 1;
          const bool wrap = nj >= n_kv;
          if (wrap) nj -= n_kv;
          if ((j >= 1 || it > 0) && (!wrap || has_next)) {
            const int ks = (u + AT8_KST - 1) % AT8_KST;
            MMA_WAIT(&k_empty[ks], u >= 1 ? (ob + (u - 1) / AT8_KST) & 1 : ob ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&k_full[ks], a.kv_rows * AT_HD * 2);
              tma_load_3d(sK + ks * AT_KV_BYTES, &tmK, &k_full[ks], (wrap ? cn.head : c.head) * AT_HD, nj * AT_BKV, wrap ? cn.batch : c.batch);
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    const int half = (warp - 2) >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t coff = (uint32_t)half * 32;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t tbase = tmem_base + lane_off + coff;                            // this thread's score columns (buffer 0)
    const uint32_t pbase = tmem_base + lane_off + AT8_TMEM_P + (uint32_t)half * 16;   // its packed probability columns
    const uint64_t one2 = pack_f32x2(a.one, a.one);     // from the argument block: ptxas must not fold the row-sum FFMA2 into FADD2
    const int t_rem = a.Tk - (int)coff - 32;
    const int jfull = t_rem >= 0 ? t_rem / AT_BKV + 1 : 0;     // blocks [0, jfull) have no padded key in this thread's columns
    const int nv0 = (a.Tk < AT_BKV ? a.Tk : AT_BKV) - (int)coff;

    for (int it = 0; it < n_it; ++it) {
      const TileCoord c = tile_coord(a, blockIdx.x + it * gridDim.x);
#ifdef DESTA_ATTN_TRACE
      tl_on = blockIdx.x == 37 && it == tl_it;
#endif
      if (warp == 2) TL(13, 0);
      uint64_t lacc0 = 0, lacc1 = 0;
      uint32_t ra[32], rb[32], pk[16];

      // ---- Q row -> tensor memory (lower half of the pair; the 128-byte row is 8 chunks of 16 B, SWIZZLE_128B).  Every S of
      //      the previous tile has been read (this thread waited for each), so no product still reads the old Q. ----
      if (half == 0) {
        mbar_wait_lean(q_full, it & 1);
        if (warp == 2) TL(13, 1);
        const uint8_t* qrow = sQ + r * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint4 u4 = *reinterpret_cast<const uint4*>(qrow + ((ch ^ (r & 7)) << 4));
          ra[4 * ch] = u4.x; ra[4 * ch + 1] = u4.y; ra[4 * ch + 2] = u4.z; ra[4 * ch + 3] = u4.w;
        }
        uint32_t lo[16], hi[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { lo[i] = ra[i]; hi[i] = ra[16 + i]; }
        tmem_st_32x16(tmem_base + lane_off + AT8_TMEM_Q, lo);
        tmem_st_32x16(tmem_base + lane_off + AT8_TMEM_Q + 16, hi);
        tmem_wait_st();
        tcgen05_fence_before();
        mbar_arrive(q_ready);
      }

      auto announce = [&](int jp) {                     // P_jp: its stores were issued a quarter step ago
        tmem_wait_st();
        tcgen05_fence_before();
        mbar_arrive(p_ready);
        if (warp == 2) TL(12, jp);
      };
      auto request = [&](int jn, uint32_t (&nx)[32]) {
        if (jn < n_kv) {
          mbar_wait_lean(&s_full[jn & 1], (jn >> 1) & 1);
          if (warp == 2) TL(11, jn);
          tcgen05_fence_after();
          tmem_ld_32x32(tbase + (jn & 1) * AT_BKV, nx);
        }
      };

      // ---- block 0: its row maximum is the reference of the whole tile ----
      mbar_wait_lean(&s_full[0], 0);
      if (warp == 2) TL(13, 3);
      TL(1 + warp, 0);
      tcgen05_fence_after();
      tmem_ld_32x32(tbase, ra);
      tmem_wait_ld();
      reg_fence32(ra);
      tcgen05_fence_before();
      mbar_arrive(&s_free[0]);
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 2)
        mx = max3(mx, (i < nv0) ? __uint_as_float(ra[i]) : -INFINITY, (i + 1 < nv0) ? __uint_as_float(ra[i + 1]) : -INFINITY);
      xm[half * 128 + r] = mx;
      pair_sync(quarter);
      const float m_ref = fmaxf(xm[r], xm[128 + r]);
      const uint64_t nm2 = pack_f32x2(-m_ref, -m_ref);
      if (warp == 2) TL(13, 4);
      if (nv0 >= 32) softmax8_cols<0, 32, false>(ra, pk, lacc0, lacc1, 32, nm2, one2);
      else if (nv0 > 0) softmax8_cols<0, 32, true>(ra, pk, lacc0, lacc1, nv0, nm2, one2);
      else {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
      }
      tmem_st_32x16(pbase, pk);                          // (the previous tile's last P V was waited for in its epilogue)
      request(1, rb);

      // ---- blocks 1 .. : p = ex2(s - m), pack, sum ----
      auto step = [&](int j, uint32_t (&cu)[32], uint32_t (&nx)[32], auto masked_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        tmem_wait_ld();
        TL(1 + warp, j);
        reg_fence32(cu);
        tcgen05_fence_before();
        mbar_arrive(&s_free[j & 1]);                    // S_{j+2} may overwrite the buffer
        const int nv = MASKED ? a.Tk - j * AT_BKV - (int)coff : 32;
        if (!MASKED || nv > 0) {
          softmax8_cols<0, 8, MASKED>(cu, pk, lacc0, lacc1, nv, nm2, one2);
          announce(j - 1);
          softmax8_cols<8, 24, MASKED>(cu, pk, lacc0, lacc1, nv, nm2, one2);
          request(j + 1, nx);
          softmax8_cols<24, 32, MASKED>(cu, pk, lacc0, lacc1, nv, nm2, one2);
        } else {
          announce(j - 1);
          request(j + 1, nx);
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = 0u;       // fully padded half: no exponentials
        }
        mbar_wait_lean(pv_done, (j - 1) & 1);           // P_{j-1} V_{j-1} has read the probability buffer (half a step ago)
        tcgen05_fence_after();
        tmem_st_32x16(pbase, pk);
      };
      int j = 1;
      for (; j + 1 < jfull; j += 2) {
        step(j, rb, ra, std::false_type{});
        step(j + 1, ra, rb, std::false_type{});
      }
      for (; j < n_kv; ++j) {
        if (j & 1) step(j, rb, ra, std::true_type{});
        else step(j, ra, rb, std::true_type{});
      }
      announce(n_kv - 1);
      if (warp == 2) TL(13, 5);

      // ---- epilogue: combine the two half-row sums, normalise and store this thread's 32 output columns ----
      float la, lb, lc, ld;
      unpack_f32x2(lacc0, la, lb);
      unpack_f32x2(lacc1, lc, ld);
      xl[half * 128 + r] = (la + lb) + (lc + ld);
      pair_sync(quarter);
      const float l_run = xl[r] + xl[128 + r];
      bool bad = !(l_run <= 1.2676506e30f);             // 2^100; also catches inf / NaN
      mbar_wait_lean(pv_done, (n_kv - 1) & 1);
      if (warp == 2) TL(13, 6);
      tcgen05_fence_after();
      const float inv = 1.0f / l_run;
      const int row = c.q0 + r;
      if (half == 0 && a.lse && row < a.Tq) a.lse[((size_t)c.batch * a.heads + c.head) * a.Tq + row] = m_ref + log2f(l_run);
      __nv_bfloat16* orow = a.out + ((size_t)c.batch * a.Tq + row) * a.ldo + c.head * AT_HD + coff;
      tmem_ld_32x32(tmem_O + lane_off + coff, ra);
      tmem_wait_ld();
      reg_fence32(ra);
      float chk = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) chk = fmaf(__uint_as_float(ra[i]), 0.f, chk);    // NaN iff some O value is inf / NaN
      bad = bad || (chk != 0.f);
      if (row < a.Tq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 u4;
          u4.x = pack_bf16(__uint_as_float(ra[i]) * inv, __uint_as_float(ra[i + 1]) * inv);
          u4.y = pack_bf16(__uint_as_float(ra[i + 2]) * inv, __uint_as_float(ra[i + 3]) * inv);
          u4.z = pack_bf16(__uint_as_float(ra[i + 4]) * inv, __uint_as_float(ra[i + 5]) * inv);
          u4.w = pack_bf16(__uint_as_float(ra[i + 6]) * inv, __uint_as_float(ra[i + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + i) = u4;
        }
        if (bad) bad_tiles[it] = 1;                     // rows past Tq are padding: their sums never matter
      }
      if (warp == 2) TL(13, 7);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  for (int it = 0; it < n_it; ++it) {
    if (bad_tiles[it]) {                                // CTA-uniform
      const TileCoord c = tile_coord(a, blockIdx.x + it * gridDim.x);
      attn_tile_fallback(&tmQ, &tmK, &tmV, &a, smem, tmem_base, c.q0, c.head, c.batch);
    }
  }
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<AT_TMEM_COLS>(tmem_base);
  }
}

static int make_qkv_tmap(CUtensorMap* m, const void* base, int heads, int T, int ld, int batch, int box_rows) {
  uint64_t dims[3] = {(uint64_t)heads * AT_HD, (uint64_t)T, (uint64_t)batch};
  uint64_t strides[3] = {0, (uint64_t)ld * 2, (uint64_t)T * ld * 2};
  uint32_t box[3] = {AT_HD, (uint32_t)box_rows, 1};
  return make_tmap(m, base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace desta

using namespace desta;

#ifdef DESTA_ATTN_TRACE
extern "C" int desta_attention_trace(long long* host8, int reset) {
  long long z[8] = {0};
  if (reset) return cudaMemcpyToSymbol(g_attn_trace, z, sizeof(z)) == cudaSuccess ? 0 : -2;
  return cudaMemcpyFromSymbol(host8, g_attn_trace, sizeof(z)) == cudaSuccess ? 0 : -2;
}
extern "C" int desta_attention_timeline(long long* host_12x32 /* [16][32] */) {
  return cudaMemcpyFromSymbol(host_12x32, g_attn_tl, sizeof(long long) * 16 * 32) == cudaSuccess ? 0 : -2;
}
#endif

// Which kernel serves UNIT-mode calls (the encoder): 0 = four softmax warps, shift applied by the tensor cores (default: the
// faster one in every measurement of round 2, profiles/r2_attention_unit8.md); 1 = attn_fwd_unit8_kernel, one tile per CTA;
// 2 = attn_fwd_unit8_kernel, persistent where it applies.  Set by DESTA_ATTN_UNIT8 at load time or desta_attention_variant().
static int g_unit_variant = [] { const char* e = getenv("DESTA_ATTN_UNIT8"); return e ? atoi(e) : 0; }();
static bool unit8_enabled() { return g_unit_variant >= 1; }
static bool unit8_persistent() { return g_unit_variant >= 2; }
extern "C" int desta_attention_variant(int v) {
  const int old = g_unit_variant;
  if (v >= 0) g_unit_variant = v;
  return old;
}

extern "C" int desta_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                               int batch, int heads, int Tq, int Tk, float scale, float* lse, const desta_dropout* drop,
                               void* stream) {
  if (!q || !k || !v || !out) { set_error("desta_attention: null pointer"); return DESTA_ERR_ARG; }
  if (batch <= 0 || heads <= 0 || Tq <= 0 || Tk <= 0) { set_error("desta_attention: bad shape"); return DESTA_ERR_ARG; }
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8) || ldq < heads * 64 || ldk < heads * 64 || ldv < heads * 64 || ldo < heads * 64) {
    set_error("desta_attention: leading dims must be multiples of 8 and >= heads*64");
    return DESTA_ERR_ARG;
  }
  if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15) { set_error("desta_attention: pointers must be 16-byte aligned"); return DESTA_ERR_ARG; }
  if (heads > 65535 || batch > 65535) { set_error("desta_attention: heads/batch too large for one launch"); return DESTA_ERR_ARG; }
  AttnArgs a;
  a.out = (__nv_bfloat16*)out; a.ldo = ldo; a.Tq = Tq; a.Tk = Tk;
  const double scale_log2 = (double)scale * 1.4426950408889634;
  a.scale_log2 = (float)scale_log2;
  // UNIT mode: the scores are already log2-domain exponents (the encoder folds head_dim^-1/2 * log2(e) into q)
  const bool unit = !(drop && drop->thr16) && fabs(scale_log2 - 1.0) < 1e-6;
  a.lse = lse;
  a.drop = make_dropout_cfg(drop);
  {
    static const int stagger_env = [] { const char* e = getenv("DESTA_ATTN_STAGGER"); return e ? atoi(e) : 0; }();
    a.stagger = stagger_env;
    a.n_sm = num_sms();
  }
  a.one = 1.0f;
  a.q_rows = Tq < AT_BQ ? Tq : AT_BQ;
  a.kv_rows = Tk < AT_BKV ? Tk : AT_BKV;
  CUtensorMap tmQ, tmK, tmV;
  if (make_qkv_tmap(&tmQ, q, heads, Tq, ldq, batch, a.q_rows)) return DESTA_ERR_CUDA;
  if (make_qkv_tmap(&tmK, k, heads, Tk, ldk, batch, a.kv_rows)) return DESTA_ERR_CUDA;
  if (make_qkv_tmap(&tmV, v, heads, Tk, ldv, batch, a.kv_rows)) return DESTA_ERR_CUDA;
  bool& attr_set = *device_once_flag(2);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_unit8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT8_SMEM);
    if (e != cudaSuccess) { set_error("attention smem attribute: %s", cudaGetErrorString(e)); return DESTA_ERR_CUDA; }
    attr_set = true;
  }
  const dim3 grid((Tq + AT_BQ - 1) / AT_BQ, heads, batch);
  a.heads = heads; a.nq = (int)grid.x;
  const long long n_tiles = (long long)grid.x * heads * batch;
  if (unit && unit8_enabled() && n_tiles < (1ll << 30)) {
    // persistent (two CTAs per SM walking the tile list) when the barrier parities repeat from tile to tile; else one tile per CTA
    a.n_tiles = (int)n_tiles;
    const int n_kv = (Tk + AT_BKV - 1) / AT_BKV, resident = 2 * a.n_sm;
    const bool persistent = unit8_persistent() && n_kv % 24 == 0 && a.q_rows == AT_BQ && a.kv_rows == AT_BKV && n_tiles > resident &&
                            (n_tiles + resident - 1) / resident <= AT8_MAX_IT;
    attn_fwd_unit8_kernel<<<persistent ? resident : (int)n_tiles, AT8_THREADS, AT8_SMEM, (cudaStream_t)stream>>>(tmQ, tmK, tmV, a);
    return check_launch("attn_fwd_unit8_kernel");
  }
  if (a.drop.thr16) attn_fwd_kernel<true, false><<<grid, AT_THREADS, AT_SMEM, (cudaStream_t)stream>>>(tmQ, tmK, tmV, a);
  else if (unit) attn_fwd_kernel<false, true><<<grid, AT_THREADS, AT_SMEM, (cudaStream_t)stream>>>(tmQ, tmK, tmV, a);
  else attn_fwd_kernel<false, false><<<grid, AT_THREADS, AT_SMEM, (cudaStream_t)stream>>>(tmQ, tmK, tmV, a);
  return check_launch("attn_fwd_kernel");
}
