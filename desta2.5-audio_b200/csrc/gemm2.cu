// 2-CTA (cta_group::2) persistent tcgen05 GEMM:  C[M,N] = epilogue( A[M,K] · W[N,K]^T ), bf16 in, fp32 accumulate.
//
// A CTA pair (thread-block cluster of 2, one TPC) owns a 256 x 256 output tile: each CTA stages its own 128 rows of A
// and HALF of the W tile (128 of the 256 N-rows) per k-block, and the leader CTA issues UMMA 256x256x16 — the tensor
// cores of both SMs read the two B halves from both shared memories, which halves the shared-memory traffic per FLOP
// compared with the 1-CTA kernel in gemm.cu (ncu showed that one pinned at ~75 % tensor-pipe by smem bandwidth).
//   per CTA: 5-stage ring of (16 KB A + 16 KB B), 2 x 256 TMEM columns (epilogue of tile i overlaps MMAs of tile i+1),
//   warp 0 TMA producer (2SM loads signal the leader's mbarrier), warp 1 MMA issuer (leader only), warp 2 TMEM alloc,
//   warps 4..11 epilogue.
// Epilogue: same fused ops as gemm.cu; bf16 outputs leave through shared-memory staging + TMA bulk stores
// (32x32 boxes per warp, SWIZZLE_64B) instead of scattered 16-byte row stores through the LSU.
#include "common.cuh"
#include "desta_b200.h"

namespace desta {

namespace g2 {
constexpr int BM = 128, BN = 256, BNH = 128, BK = 64, UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BNH * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;   // 32 KB
constexpr int EPI_WARP0 = 4, EPI_WARPS = 8, THREADS = 384;
// two epilogue flavours: bf16 outputs (32 x 32 bf16 boxes) and fp32 outputs with optional fp32 residual (32 x 16 fp32
// boxes, loaded AND stored by TMA); both stage through two 2 KB boxes per warp (fp32 mode: a third one for the optional
// bf16 copy of the same rows, 32 x 32 bf16, also TMA-stored), which leaves room for a 5-stage operand ring in either mode
// (the K = 5120 fc2 GEMM lost 10 % going from 4 to 3 stages)
constexpr int EPI_BF16 = 0, EPI_F32 = 1;
__host__ __device__ constexpr int stages_of(int mode) { return 5; }
__host__ __device__ constexpr int staging_of(int mode) { return mode == EPI_F32 ? 3 * 2048 : 2 * 2048; }   // EPI_F32: + one bf16 box
constexpr int BAR_BYTES = 512;
__host__ __device__ constexpr int smem_of(int mode) { return stages_of(mode) * STAGE_BYTES + EPI_WARPS * staging_of(mode) + 1024 + BAR_BYTES + EPI_WARPS * 1024; }
}  // namespace g2

struct Gemm2Args {
  int M, N, K;
  float* out_f32; int ld_f32;
  __nv_bfloat16* out_bf16; int ld_bf16;   // EPI_F32 only: optional direct bf16 copy (tap hand-off)
  int has_bf16;            // EPI_BF16: bf16 output through tmC
  const float* bias;
  const float* resid; int ld_resid;
  const float* pos; int pos_rows;
  float scale; int scale_cols;
  int gelu;
  // fused LayerNorm hand-over (see desta_gemm_epilogue): EPI_F32 emits per-row partial (sum, sum of squares) of its final
  // output per 128-column group; EPI_BF16 applies  rstd * (acc - mean * colsum[n]) + bias[n]  with the row statistics of
  // the tensor A was copied from (A = bf16 of the un-normalised rows, W carries the LayerNorm gain)
  float2* ln_stats_out; int ln_groups_out;
  const float2* ln_stats_in; int ln_groups_in; float ln_inv_cols, ln_eps;
  const float* ln_colsum;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
// 2SM TMA load: data lands in THIS CTA's smem, transaction bytes are counted on the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  const uint32_t leader_bar = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(m), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all previously issued MMAs completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_plain(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(g2::THREADS, 1)
gemm_bf16_tn_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                         const __grid_constant__ CUtensorMap tmD, const Gemm2Args g) {
  using namespace g2;
  constexpr int STAGES = stages_of(MODE);
  constexpr int STAGING_PER_WARP = staging_of(MODE);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + EPI_WARPS * STAGING_PER_WARP);
  uint64_t* full = bars;                        // [STAGES] used in the leader: 2 producer arrivals + tx bytes of both CTAs
  uint64_t* empty = bars + STAGES;              // [STAGES] per CTA, arrived by the leader's multicast commit
  uint64_t* acc_full = bars + 2 * STAGES;       // [2] per CTA (multicast commit)
  uint64_t* acc_empty = bars + 2 * STAGES + 2;  // [2] used in the leader: epilogue warps of both CTAs
  uint64_t* rbar = bars + 2 * STAGES + 4;       // [EPI_WARPS][2] residual boxes landed (EPI_F32)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rbar + 2 * EPI_WARPS);
  float* bias_all = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + BAR_BYTES);   // [EPI_WARPS][128]
  float* colsum_all = bias_all + EPI_WARPS * 128;                                              // [EPI_WARPS][128] (LayerNorm mode)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int m_tiles = (g.M + 2 * BM - 1) / (2 * BM), n_tiles = (g.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = (g.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if (MODE == EPI_F32 && g.resid) tma_prefetch_desc(&tmR);
    if (MODE == EPI_F32 && g.out_bf16) tma_prefetch_desc(&tmD);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 2); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 2 * EPI_WARPS); }
    for (int i = 0; i < 2 * EPI_WARPS; ++i) mbar_init(&rbar[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();          // peer barriers initialised / TMEM allocated in both CTAs
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Warps 0 and 1 run their loops converged and hand each TMA / tcgen05 group to one lane with elect.sync (ptxas then
  // issues UTMALDG / UTCHMMA from uniform registers instead of a per-instruction waterfall loop; tools/umma_bench.cu).
  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int tile = pair; tile < num_tiles; tile += n_pairs) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int row_a = m_blk * 2 * BM + (int)rank * BM;
      const int row_b = n_blk * BN + (int)rank * BNH;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait_sleep(&empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * STAGE_BYTES;
          tma_load_2d_2sm(sa, &tmA, &full[stage], kb * BK, row_a);
          tma_load_2d_2sm(sa + A_BYTES, &tmB, &full[stage], kb * BK, row_b);
          if (leader) mbar_arrive_expect_tx(&full[stage], 2 * STAGE_BYTES);
          else mbar_arrive_cluster(&full[stage], 0);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int local = 0;
      for (int tile = pair; tile < num_tiles; tile += n_pairs, ++local) {
        const int as = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait(&acc_empty[as], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t adesc = umma_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = umma_desc_sw128(sa + A_BYTES, 16, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) umma_bf16_ss_2cta(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            umma_commit_2cta(&empty[stage]);
            if (kb == k_blocks - 1) umma_commit_2cta(&acc_full[as]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    const int ew = warp - EPI_WARP0;
    const int q = warp & 3;
    const int half = ew >> 2;
    uint8_t* stg = staging + ew * STAGING_PER_WARP;
    uint64_t* my_rbar = rbar + 2 * ew;
    int local = 0;
    uint32_t nchunk = 0;           // staged boxes issued so far by this warp (buffer = nchunk & 1)
    const bool use_resid = (MODE == EPI_F32) && g.resid != nullptr;
    for (int tile = pair; tile < num_tiles; tile += n_pairs, ++local) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int as = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int row0 = m_blk * 2 * BM + (int)rank * BM + q * 32;
      const int row = row0 + lane;
      const bool row_ok = row < g.M;
      const int colbase = n_blk * BN + half * 128;
      float* bias_s = bias_all + ew * 128;
      if (g.bias) {       // this warp's 128 bias values, fetched while the MMAs of the tile are still running
        const int cb = colbase + lane * 4;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cb + 3 < g.N) b4 = __ldg(reinterpret_cast<const float4*>(g.bias + cb));
        else {
          if (cb < g.N) b4.x = __ldg(g.bias + cb);
          if (cb + 1 < g.N) b4.y = __ldg(g.bias + cb + 1);
          if (cb + 2 < g.N) b4.z = __ldg(g.bias + cb + 2);
        }
        __syncwarp();
        reinterpret_cast<float4*>(bias_s)[lane] = b4;
        __syncwarp();
      }
      const bool ln_in = (MODE == EPI_BF16) && g.ln_stats_in != nullptr;
      float* colsum_s = colsum_all + ew * 128;
      float ln_rstd = 1.f, ln_nmu = 0.f;
      if (ln_in) {
        const int cb = colbase + lane * 4;
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cb + 3 < g.N) s4 = __ldg(reinterpret_cast<const float4*>(g.ln_colsum + cb));
        else {
          if (cb < g.N) s4.x = __ldg(g.ln_colsum + cb);
          if (cb + 1 < g.N) s4.y = __ldg(g.ln_colsum + cb + 1);
          if (cb + 2 < g.N) s4.z = __ldg(g.ln_colsum + cb + 2);
        }
        __syncwarp();
        reinterpret_cast<float4*>(colsum_s)[lane] = s4;
        __syncwarp();
        if (row_ok) {      // mean / rstd of this thread's row from the producer's per-group partial sums
          const float2* st = g.ln_stats_in + (size_t)row * g.ln_groups_in;
          float s1 = 0.f, s2 = 0.f;
          for (int i = 0; i < g.ln_groups_in; ++i) { const float2 p = __ldg(st + i); s1 += p.x; s2 += p.y; }
          const float mu = s1 * g.ln_inv_cols;
          const float var = fmaxf(fmaf(s2, g.ln_inv_cols, -mu * mu), 0.f);
          ln_rstd = rsqrtf(var + g.ln_eps);
          ln_nmu = -mu;
        }
      }
      uint64_t st_sum = 0, st_sq = 0;     // EPI_F32 + ln_stats_out: (even, odd) column partial sums of this row
      if (use_resid && lane == 0 && colbase < g.N && row0 < g.M) {
        // residual box of the first half-chunk (the buffer's previous store was issued >= 2 boxes ago)
        bulk_wait_read<1>();
        mbar_arrive_expect_tx(&my_rbar[nchunk & 1], 2048);
        tma_load_2d_plain(stg + (nchunk & 1) * 2048, &tmR, &my_rbar[nchunk & 1], colbase, row0);
      }
      mbar_wait_sleep(&acc_full[as], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr0 = tmem_base + as * BN + half * 128 + ((uint32_t)(q * 32) << 16);
      const float* pos_row = g.pos ? g.pos + (size_t)(row % g.pos_rows) * g.N : nullptr;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(taddr0 + c * 32, v);
        tmem_wait_ld();
        if (c == 3) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&acc_empty[as], 0);
        }
        const int col0 = colbase + c * 32;
        if (col0 >= g.N || row0 >= g.M) continue;     // warp-uniform
        const bool full_cols = (col0 + 32 <= g.N);
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (ln_in) {
          // LayerNorm folded into this GEMM: acc = sum_k bf16(x_k) (gain_k W_nk)  ->  rstd (acc - mean colsum_n) + bias_n
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 sn = reinterpret_cast<const float4*>(colsum_s + c * 32)[j >> 2];
            const float4 b = reinterpret_cast<const float4*>(bias_s + c * 32)[j >> 2];
            f[j] = fmaf(ln_rstd, fmaf(ln_nmu, sn.x, f[j]), b.x);
            f[j + 1] = fmaf(ln_rstd, fmaf(ln_nmu, sn.y, f[j + 1]), b.y);
            f[j + 2] = fmaf(ln_rstd, fmaf(ln_nmu, sn.z, f[j + 2]), b.z);
            f[j + 3] = fmaf(ln_rstd, fmaf(ln_nmu, sn.w, f[j + 3]), b.w);
          }
        } else if (g.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = reinterpret_cast<const float4*>(bias_s + c * 32)[j >> 2];   // broadcast read
            f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
          }
        }
        if (col0 < g.scale_cols) {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (col0 + j < g.scale_cols) f[j] *= g.scale;
        }
        if (g.gelu) gelu_erf_x32(f);
        if (pos_row && row_ok) {
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
        if (MODE == EPI_BF16) {
          // stage 32 rows x 32 bf16 (64-byte rows, SWIZZLE_64B) and hand the box to the TMA (clips at M / N itself)
          uint8_t* buf = stg + (nchunk & 1) * 2048;
          if (lane == 0) bulk_wait_read<1>();          // the store that used this buffer two chunks ago has read it
          __syncwarp();
          uint8_t* rowp = buf + lane * 64;
          const int sw = (lane >> 1) & 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            u.x = pack_bf16(f[8 * j], f[8 * j + 1]); u.y = pack_bf16(f[8 * j + 2], f[8 * j + 3]);
            u.z = pack_bf16(f[8 * j + 4], f[8 * j + 5]); u.w = pack_bf16(f[8 * j + 6], f[8 * j + 7]);
            *reinterpret_cast<uint4*>(rowp + ((j ^ sw) << 4)) = u;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmC, buf, col0, row0);
            bulk_commit();
          }
          ++nchunk;
        } else {
          // fp32: two boxes of 32 rows x 16 floats (64-byte rows, SWIZZLE_64B) per 32-column chunk.  The residual box was
          // TMA-loaded into the very buffer; add in place, then TMA-store.  The next box's residual is requested first.
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int colh = col0 + 16 * h;
            if (colh >= g.N) break;                     // warp-uniform
            uint8_t* buf = stg + (nchunk & 1) * 2048;
            const bool more = h == 0 ? (colh + 16 < g.N) : ((c < 3) && (col0 + 32 < g.N));
            if (use_resid) {
              if (lane == 0 && more) {
                bulk_wait_read<0>();                    // the other buffer's store (previous box) has been read
                mbar_arrive_expect_tx(&my_rbar[(nchunk + 1) & 1], 2048);
                tma_load_2d_plain(stg + ((nchunk + 1) & 1) * 2048, &tmR, &my_rbar[(nchunk + 1) & 1], colh + 16, row0);
              }
              mbar_wait(&my_rbar[nchunk & 1], (nchunk >> 1) & 1);
            } else {
              if (lane == 0) bulk_wait_read<1>();
              __syncwarp();
            }
            uint8_t* rowp = buf + lane * 64;
            const int sw = (lane >> 1) & 3;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4* q4 = reinterpret_cast<float4*>(rowp + ((j ^ sw) << 4));
              float4 o4 = make_float4(f[16 * h + 4 * j], f[16 * h + 4 * j + 1], f[16 * h + 4 * j + 2], f[16 * h + 4 * j + 3]);
              if (use_resid) {
                const float4 r4 = *q4;
                o4.x += r4.x; o4.y += r4.y; o4.z += r4.z; o4.w += r4.w;
              }
              *q4 = o4;
              f[16 * h + 4 * j] = o4.x; f[16 * h + 4 * j + 1] = o4.y; f[16 * h + 4 * j + 2] = o4.z; f[16 * h + 4 * j + 3] = o4.w;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, buf, colh, row0);
              bulk_commit();
            }
            ++nchunk;
          }
          if (g.ln_stats_out) {
            const uint64_t one2 = pack_f32x2(1.f, 1.f);
            if (full_cols) {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const uint64_t v2 = pack_f32x2(f[j], f[j + 1]);
                st_sum = ffma2(v2, one2, st_sum);
                st_sq = ffma2(v2, v2, st_sq);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const uint64_t v2 = pack_f32x2(col0 + j < g.N ? f[j] : 0.f, col0 + j + 1 < g.N ? f[j + 1] : 0.f);
                st_sum = ffma2(v2, one2, st_sum);
                st_sq = ffma2(v2, v2, st_sq);
              }
            }
          }
          if (g.out_bf16) {
            // bf16 copy of the final rows (tap hand-off / A operand of the next GEMM): third staging box, 32 rows x 32 bf16
            // (64-byte rows, SWIZZLE_64B), TMA-stored (clips at M / N itself).  Its previous store lies >= 3 bulk groups back.
            uint8_t* buf = stg + 2 * 2048;
            if (lane == 0) bulk_wait_read<2>();
            __syncwarp();
            uint8_t* rowp = buf + lane * 64;
            const int sw = (lane >> 1) & 3;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 u;
              u.x = pack_bf16(f[8 * j], f[8 * j + 1]); u.y = pack_bf16(f[8 * j + 2], f[8 * j + 3]);
              u.z = pack_bf16(f[8 * j + 4], f[8 * j + 5]); u.w = pack_bf16(f[8 * j + 6], f[8 * j + 7]);
              *reinterpret_cast<uint4*>(rowp + ((j ^ sw) << 4)) = u;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmD, buf, col0, row0);
              bulk_commit();
            }
          }
        }
      }
      if (MODE == EPI_F32 && g.ln_stats_out && row_ok) {
        const int grp = n_blk * 2 + half;
        if (grp < g.ln_groups_out) {
          float a0, a1, q0, q1;
          unpack_f32x2(st_sum, a0, a1);
          unpack_f32x2(st_sq, q0, q1);
          g.ln_stats_out[(size_t)row * g.ln_groups_out + grp] = make_float2(a0 + a1, q0 + q1);
        }
      }
    }
    if (lane == 0) bulk_wait_all();
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();          // nobody exits (or frees TMEM) while the peer may still touch this CTA's smem / barriers
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}


static int make_f32_box_map(CUtensorMap* m, const float* base, int M, int N, int ld) {
  uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
  uint64_t strides[2] = {0, (uint64_t)ld * 4};
  uint32_t box[2] = {16, 32};
  return make_tmap(m, base, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
}

int launch_gemm_2cta(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const desta_gemm_epilogue* ep,
                     cudaStream_t stream) {
  using namespace g2;
  CUtensorMap tmA, tmB, tmC, tmR, tmD;
  if (make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK)) return DESTA_ERR_CUDA;
  if (make_tmap_2d_bf16(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, BNH, BK)) return DESTA_ERR_CUDA;
  const int mode = ep->out_f32 ? EPI_F32 : EPI_BF16;
  Gemm2Args g;
  g.M = M; g.N = N; g.K = K;
  g.out_f32 = ep->out_f32; g.ld_f32 = ep->ld_f32;
  g.out_bf16 = nullptr; g.ld_bf16 = 0; g.has_bf16 = 0;
  if (mode == EPI_BF16) {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    uint64_t strides[2] = {0, (uint64_t)ep->ld_bf16 * 2};
    uint32_t box[2] = {32, 32};
    if (make_tmap(&tmC, ep->out_bf16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B))
      return DESTA_ERR_CUDA;
    tmR = tmC;
    tmD = tmC;
    g.has_bf16 = 1;
  } else {
    if (make_f32_box_map(&tmC, ep->out_f32, M, N, ep->ld_f32)) return DESTA_ERR_CUDA;
    if (ep->resid) { if (make_f32_box_map(&tmR, ep->resid, M, N, ep->ld_resid)) return DESTA_ERR_CUDA; }
    else tmR = tmC;
    tmD = tmC;
    g.out_bf16 = (__nv_bfloat16*)ep->out_bf16; g.ld_bf16 = ep->ld_bf16;
    if (ep->out_bf16) {
      uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
      uint64_t strides[2] = {0, (uint64_t)ep->ld_bf16 * 2};
      uint32_t box[2] = {32, 32};
      if (make_tmap(&tmD, ep->out_bf16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B))
        return DESTA_ERR_CUDA;
    }
  }
  g.bias = ep->bias; g.resid = ep->resid; g.ld_resid = ep->ld_resid;
  g.pos = ep->pos; g.pos_rows = ep->pos_rows > 0 ? ep->pos_rows : 1;
  g.scale = ep->scale; g.scale_cols = ep->scale_cols; g.gelu = ep->gelu;
  g.ln_stats_out = nullptr; g.ln_groups_out = 0; g.ln_stats_in = nullptr; g.ln_groups_in = 0; g.ln_inv_cols = 0.f; g.ln_eps = 0.f;
  g.ln_colsum = nullptr;
  if (ep->ln_stats_out) {
    if (mode != EPI_F32) { set_error("desta_gemm_bf16: ln_stats_out needs the fp32 output"); return DESTA_ERR_ARG; }
    g.ln_stats_out = (float2*)ep->ln_stats_out; g.ln_groups_out = (N + 127) / 128;
  }
  if (ep->ln_stats_in) {
    if (mode != EPI_BF16 || !ep->ln_colsum || !ep->bias || ep->ln_groups <= 0 || ep->ln_cols <= 0) {
      set_error("desta_gemm_bf16: ln_stats_in needs the bf16 output, ln_colsum, bias, ln_groups and ln_cols");
      return DESTA_ERR_ARG;
    }
    g.ln_stats_in = (const float2*)ep->ln_stats_in; g.ln_groups_in = ep->ln_groups; g.ln_inv_cols = 1.0f / (float)ep->ln_cols;
    g.ln_eps = ep->ln_eps; g.ln_colsum = ep->ln_colsum;
  }
  bool& attr_set = *device_once_flag(1);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_2cta_kernel<EPI_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(EPI_BF16));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_bf16_tn_2cta_kernel<EPI_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(EPI_F32));
    if (e != cudaSuccess) { set_error("gemm 2cta smem attribute: %s", cudaGetErrorString(e)); return DESTA_ERR_CUDA; }
    attr_set = true;
  }
  const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
  int pairs = num_sms() / 2;
  if (tiles < pairs) pairs = tiles;
  if (mode == EPI_BF16) gemm_bf16_tn_2cta_kernel<EPI_BF16><<<2 * pairs, THREADS, smem_of(EPI_BF16), stream>>>(tmA, tmB, tmC, tmR, tmD, g);
  else gemm_bf16_tn_2cta_kernel<EPI_F32><<<2 * pairs, THREADS, smem_of(EPI_F32), stream>>>(tmA, tmB, tmC, tmR, tmD, g);
  return check_launch("gemm_bf16_tn_2cta_kernel");
}

}  // namespace desta
