// Persistent warp-specialised tcgen05 GEMM for sm_100a:   C[M,N] = epilogue( A[M,K] · W[N,K]^T )
//   A, W bf16 K-major (nn.Linear layout), fp32 accumulation in TMEM (2 x 256-column accumulator buffers so the
//   epilogue of tile i overlaps the MMAs of tile i+1), operands staged by TMA (SWIZZLE_128B, 4-stage mbarrier ring).
// Tile 128 x 256 x 64, UMMA 128x256x16 (cta_group::1), 384 threads:
//   warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warp 2 = TMEM allocator, warps 4..11 = epilogue.
// Epilogue (fused, per the reference ops it replaces):
//   +bias, column-range scale (Whisper q·hd^-1/2 after bias, modeling_whisper.py:310-332), erf-GELU
//   (modeling_whisper.py:404-406, modeling_bert.py:339-342, modeling_desta25.py:563-564), + row-periodic fp32 table
//   (embed_positions add, modeling_desta25.py:566-570), + fp32 residual (modeling_whisper.py:400,408),
//   fp32 and/or bf16 stores (bf16 copy = the tap hand-off, modeling_desta25.py:585-595).
//
// MODE_WGRAD (connector backward, autograd of $TF/models/bert/modeling_bert.py Linear layers):
//   C[Nout,Kin] += dY[Mred,Nout]^T · X[Mred,Kin] — both operands are read as they lie in memory (reduction dim = rows),
//   i.e. MN-major UMMA operands: TMA boxes of 64 rows x 64 columns land as 128-byte-swizzled atoms, LBO = 8 KB between
//   the 64-column atoms, SBO = 1 KB between 8-row groups.  Optional split-K over the reduction with fp32 atomics.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "desta_b200.h"

namespace desta {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4, UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;   // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int GEMM_THREADS = 384;
constexpr int EPI_WARP0 = 4, EPI_WARPS = 8;
constexpr int GEMM_SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct GemmArgs {
  int M, N, K;
  float* out_f32; int ld_f32;
  __nv_bfloat16* out_bf16; int ld_bf16;
  const float* bias;
  const float* resid; int ld_resid;
  const float* pos; int pos_rows;
  float scale; int scale_cols;
  int gelu;
  // work items (MODE_WGRAD): the first n_full tiles are one item each (whole reduction; dW += acc by plain
  // read-modify-write through `resid`); the remaining tiles - the partial last wave of the persistent grid, or every tile
  // of a small output - are split `splits` ways along the reduction and accumulate with fp32 atomics
  int n_full;
  int splits;
};

constexpr int MODE_TN = 0, MODE_WGRAD = 1;
constexpr int ATOM_BYTES = 64 * 128;   // one 64-row x 64-column bf16 box

template <int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                 // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;      // [2] MMA -> epilogue
  uint64_t* acc_empty = bars + 2 * STAGES + 2; // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (g.M + BM - 1) / BM, n_tiles = (g.N + BN - 1) / BN;
  const int num_tiles = g.n_full + (m_tiles * n_tiles - g.n_full) * g.splits;          // work items
  const int k_blocks_total = (g.K + BK - 1) / BK;
  const int kb_per = (k_blocks_total + g.splits - 1) / g.splits;
  // item -> (tile, k-block range, split?)
  auto decode = [&](int item, int& tile, int& kb0, int& kb1) {
    if (item < g.n_full) { tile = item; kb0 = 0; kb1 = k_blocks_total; return false; }
    const int j = item - g.n_full;
    tile = g.n_full + j / g.splits;
    const int split = j - (j / g.splits) * g.splits;
    kb0 = split * kb_per;
    kb1 = min(k_blocks_total, kb0 + kb_per);
    return g.splits > 1;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Warps 0 and 1 run their loops converged and hand each TMA / tcgen05 group to one lane with elect.sync (ptxas then
  // issues UTMALDG / UTCHMMA from uniform registers instead of a per-instruction waterfall loop; tools/umma_bench.cu).
  if (warp == 0) {
    {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_tiles; item += gridDim.x) {
        int tile, kb0, kb1;
        decode(item, tile, kb0, kb1);
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_sleep(&empty[stage], phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
            uint8_t* sa = smem + stage * STAGE_BYTES;
            if (MODE == MODE_TN) {
              tma_load_2d(sa, &tmA, &full[stage], kb * BK, m_blk * BM);
              tma_load_2d(sa + A_STAGE_BYTES, &tmB, &full[stage], kb * BK, n_blk * BN);
            } else {
#pragma unroll
              for (int a = 0; a < BM / 64; ++a) tma_load_2d(sa + a * ATOM_BYTES, &tmA, &full[stage], m_blk * BM + a * 64, kb * BK);
#pragma unroll
              for (int a = 0; a < BN / 64; ++a)
                tma_load_2d(sa + A_STAGE_BYTES + a * ATOM_BYTES, &tmB, &full[stage], n_blk * BN + a * 64, kb * BK);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, MODE == MODE_WGRAD, MODE == MODE_WGRAD);
      int stage = 0; uint32_t phase = 0;
      int local = 0;
      for (int item = blockIdx.x; item < num_tiles; item += gridDim.x, ++local) {
        int tile_unused, kb0, kb1;
        decode(item, tile_unused, kb0, kb1);
        const int as = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait(&acc_empty[as], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          if (elect_one()) {
          if (MODE == MODE_TN) {
            const uint64_t adesc = umma_desc_sw128(sa, 16, 1024);
            const uint64_t bdesc = umma_desc_sw128(sa + A_STAGE_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 16 bf16 = 32 B along K inside the 128-byte swizzle atom: +2 in the (addr>>4) field
              umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb0 || k) ? 1u : 0u);
            }
          } else {
            const uint64_t adesc = umma_desc_sw128(sa, ATOM_BYTES, 1024);
            const uint64_t bdesc = umma_desc_sw128(sa + A_STAGE_BYTES, ATOM_BYTES, 1024);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // 16 reduction rows = two 8-row groups = 2048 B
              umma_bf16_ss(d_tmem, adesc + 128 * k, bdesc + 128 * k, idesc, (kb > kb0 || k) ? 1u : 0u);
            }
          }
          umma_commit(&empty[stage]);
          if (kb == kb1 - 1) umma_commit(&acc_full[as]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - EPI_WARP0) >> 2;  // column half handled by this warp
    int local = 0;
    for (int item = blockIdx.x; item < num_tiles; item += gridDim.x, ++local) {
      int tile, kb0_unused, kb1_unused;
      const bool atomic = decode(item, tile, kb0_unused, kb1_unused);
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int as = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait_sleep(&acc_full[as], acc_phase);
      tcgen05_fence_after();
      const int row = m_blk * BM + q * 32 + lane;
      const bool row_ok = row < g.M;
      const uint32_t taddr0 = tmem_base + as * BN + half * 128 + ((uint32_t)(q * 32) << 16);
      const float* pos_row = g.pos ? g.pos + (size_t)(row % g.pos_rows) * g.N : nullptr;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(taddr0 + c * 32, v);
        tmem_wait_ld();
        if (c == 3) {   // accumulator fully drained into registers: hand the TMEM buffer back to the MMA warp
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[as]);
        }
        const int col0 = n_blk * BN + half * 128 + c * 32;
        if (col0 >= g.N) continue;
        const bool full_cols = (col0 + 32 <= g.N);
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (g.bias) {
          if (full_cols) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + col0 + j));
              f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < g.N) f[j] += __ldg(g.bias + col0 + j);
          }
        }
        if (col0 < g.scale_cols) {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (col0 + j < g.scale_cols) f[j] *= g.scale;
        }
        if (g.gelu) gelu_erf_x32(f);
        if (row_ok) {
        if (pos_row) {
          if (full_cols) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(pos_row + col0 + j));
              f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < g.N) f[j] += __ldg(pos_row + col0 + j);
          }
        }
        if (g.resid && !atomic) {
          const float* r = g.resid + (size_t)row * g.ld_resid + col0;
          if (full_cols) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = *reinterpret_cast<const float4*>(r + j);
              f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < g.N) f[j] += r[j];
          }
        }
        if (g.out_f32 && atomic) {
          float* o = g.out_f32 + (size_t)row * g.ld_f32 + col0;
#pragma unroll
          for (int j = 0; j < 32; ++j) if (col0 + j < g.N) atomicAdd(o + j, f[j]);
        } else if (g.out_f32) {
          float* o = g.out_f32 + (size_t)row * g.ld_f32 + col0;
          if (full_cols) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < g.N) o[j] = f[j];
          }
        }
        if (g.out_bf16) {
          __nv_bfloat16* o = g.out_bf16 + (size_t)row * g.ld_bf16 + col0;
          if (full_cols) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 u;
              u.x = pack_bf16(f[j], f[j + 1]); u.y = pack_bf16(f[j + 2], f[j + 3]);
              u.z = pack_bf16(f[j + 4], f[j + 5]); u.w = pack_bf16(f[j + 6], f[j + 7]);
              *reinterpret_cast<uint4*>(o + j) = u;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < g.N) o[j] = __float2bfloat16_rn(f[j]);
          }
        }
        }  // row_ok
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

int launch_gemm_2cta(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const desta_gemm_epilogue* ep,
                     cudaStream_t stream);


}  // namespace desta

using namespace desta;

static int launch_gemm(int mode, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& g, cudaStream_t stream) {
  bool& attr_set = *device_once_flag(0);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<MODE_TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_bf16_kernel<MODE_WGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
    if (e != cudaSuccess) { set_error("gemm smem attribute: %s", cudaGetErrorString(e)); return DESTA_ERR_CUDA; }
    attr_set = true;
  }
  const int m_tiles = (g.M + BM - 1) / BM, n_tiles = (g.N + BN - 1) / BN;
  const int items = g.n_full + (m_tiles * n_tiles - g.n_full) * g.splits;
  const int grid = items < num_sms() ? items : num_sms();
  if (mode == MODE_TN) gemm_bf16_kernel<MODE_TN><<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(tmA, tmB, g);
  else gemm_bf16_kernel<MODE_WGRAD><<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(tmA, tmB, g);
  return check_launch("gemm_bf16_kernel");
}

extern "C" int desta_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                               const desta_gemm_epilogue* ep, void* stream) {
  if (!A || !W || !ep) { set_error("desta_gemm_bf16: null pointer"); return DESTA_ERR_ARG; }
  if (M <= 0 || N <= 0 || K <= 0) { set_error("desta_gemm_bf16: bad shape M=%d N=%d K=%d", M, N, K); return DESTA_ERR_ARG; }
  if ((lda % 8) || (ldw % 8) || lda < K || ldw < K) {
    set_error("desta_gemm_bf16: lda/ldw must be >= K and multiples of 8 (got %d, %d, K=%d)", lda, ldw, K);
    return DESTA_ERR_ARG;
  }
  if (!ep->out_f32 && !ep->out_bf16) { set_error("desta_gemm_bf16: no output"); return DESTA_ERR_ARG; }
  if ((ep->out_f32 && (ep->ld_f32 % 4)) || (ep->out_bf16 && (ep->ld_bf16 % 8)) || (ep->resid && (ep->ld_resid % 4))) {
    set_error("desta_gemm_bf16: output/residual leading dims must keep 16-byte row alignment");
    return DESTA_ERR_ARG;
  }
  static int force_1cta = -1;
  if (force_1cta < 0) { const char* e = getenv("DESTA_GEMM_1CTA"); force_1cta = (e && e[0] == '1') ? 1 : 0; }
  if (!force_1cta && M > 128) return launch_gemm_2cta(A, lda, W, ldw, M, N, K, ep, (cudaStream_t)stream);
  if (ep->ln_stats_out || ep->ln_stats_in) { set_error("desta_gemm_bf16: the LayerNorm hand-over needs M > 128 (CTA-pair kernel)"); return DESTA_ERR_ARG; }
  CUtensorMap tmA, tmB;
  if (make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK)) return DESTA_ERR_CUDA;
  if (make_tmap_2d_bf16(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, BN, BK)) return DESTA_ERR_CUDA;
  GemmArgs g;
  g.M = M; g.N = N; g.K = K;
  g.out_f32 = (float*)ep->out_f32; g.ld_f32 = ep->ld_f32;
  g.out_bf16 = (__nv_bfloat16*)ep->out_bf16; g.ld_bf16 = ep->ld_bf16;
  g.bias = ep->bias; g.resid = ep->resid; g.ld_resid = ep->ld_resid;
  g.pos = ep->pos; g.pos_rows = ep->pos_rows > 0 ? ep->pos_rows : 1;
  g.scale = ep->scale; g.scale_cols = ep->scale_cols; g.gelu = ep->gelu;
  g.splits = 1; g.n_full = 0;
  return launch_gemm(MODE_TN, tmA, tmB, g, (cudaStream_t)stream);
}

extern "C" int desta_gemm_bf16_wgrad(const void* dY, int ldy, const void* X, int ldx, int Mred, int Nout, int Kin,
                                     float* dW, int ldw, void* stream) {
  if (!dY || !X || !dW) { set_error("desta_gemm_bf16_wgrad: null pointer"); return DESTA_ERR_ARG; }
  if (Mred <= 0 || Nout <= 0 || Kin <= 0) { set_error("desta_gemm_bf16_wgrad: bad shape"); return DESTA_ERR_ARG; }
  if ((ldy % 8) || (ldx % 8) || ldy < Nout || ldx < Kin || (ldw % 4)) {
    set_error("desta_gemm_bf16_wgrad: leading dims must cover the row and keep 16-byte alignment");
    return DESTA_ERR_ARG;
  }
  CUtensorMap tmA, tmB;
  // operands are loaded as 64-row x 64-column boxes, exactly as they lie in memory (reduction dim = rows)
  if (make_tmap_2d_bf16(&tmA, dY, (uint64_t)Mred, (uint64_t)Nout, (uint64_t)ldy, 64, 64)) return DESTA_ERR_CUDA;
  if (make_tmap_2d_bf16(&tmB, X, (uint64_t)Mred, (uint64_t)Kin, (uint64_t)ldx, 64, 64)) return DESTA_ERR_CUDA;
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = Nout; g.N = Kin; g.K = Mred;
  g.out_f32 = dW; g.ld_f32 = ldw;
  g.pos_rows = 1; g.scale = 1.f;
  const int tiles = ((Nout + BM - 1) / BM) * ((Kin + BN - 1) / BN);
  const int kblocks = (Mred + BK - 1) / BK;
  // The persistent grid runs one CTA per SM.  Full waves of tiles are processed whole; the tiles of the partial last wave
  // (all tiles of a small output) are split along the reduction so that every SM gets a share of them: the K/V-projection
  // weight gradient has 600 tiles = 4.05 waves on 148 SMs, and its 8 tail tiles would otherwise cost a whole fifth round.
  const int sms = num_sms();
  const int n_full = (tiles / sms) * sms;
  const int tail = tiles - n_full;
  int splits = 1;
  if (tail > 0) {
    splits = sms / tail;
    const int max_splits = kblocks / 8 > 0 ? kblocks / 8 : 1;    // keep >= 8 k-blocks per split
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    const int kb_per = (kblocks + splits - 1) / splits;
    splits = (kblocks + kb_per - 1) / kb_per;                    // no empty split
  }
  g.n_full = splits > 1 ? n_full : tiles;
  g.splits = splits;
  g.resid = dW; g.ld_resid = ldw;                                // whole-reduction items: dW += acc (one writer per element)
  return launch_gemm(MODE_WGRAD, tmA, tmB, g, (cudaStream_t)stream);
}
