// Attention backward for the Q-Former (<= 128 queries per sequence; 64 in DeSTA2.5) on tcgen05/TMEM.
// Autograd counterpart of the eager Bert self / cross attention ($TF/models/bert/modeling_bert.py:168-207, 234-284):
//   P = softmax(scale Q K^T),  dV = P^T dO,  dP = dO V^T,  dS = scale * P o (dP - rowsum(dO o O)),  dQ = dS K,  dK = dS^T Q.
// With train-mode dropout Z (keep / (1 - p), regenerated from the forward's Philox counter; modeling_bert.py:198-201):
//   O = (P o Z) V  ->  dV = (P o Z)^T dO,  dS = scale * P o (Z o dP - rowsum(dO o O))   (rowsum(dO o O) is unchanged).
//
// One CTA per (sequence, head); keys are streamed in blocks of 128 through a 2-stage TMA ring.  All five products run
// on the tensor cores with operands read as they lie in shared memory:
//   S  = Q K^T      A = Q  (K-major)            B = K  (K-major)        -> TMEM [  0,128)
//   dP = dO V^T     A = dO (K-major)            B = V  (K-major)        -> TMEM [128,256)
//   dV = P^T dO     A = P^T (MN-major view of the P tile)   B = dO (MN-major) -> TMEM [256,320)
//   dK = dS^T Q     A = dS^T (MN-major view of the dS tile) B = Q  (MN-major) -> TMEM [320,384)
//   dQ += dS K      A = dS (K-major)            B = K  (MN-major)       -> TMEM [384,448)
// warp 0 = TMA, warp 1 = TMEM alloc + MMA issue, warps 2-5 = one query row per thread (P / dS) and, for the dK / dV
// read-out, one key row per thread.
#include "common.cuh"
#include "desta_b200.h"

namespace desta {

constexpr int AB_T = 128, AB_HD = 64;
constexpr int AB_TILE = AB_T * AB_HD * 2;          // 16 KB: Q, dO, K-stage, V-stage
constexpr int AB_PS = AB_T * AB_T * 2;             // 32 KB: P, dS
constexpr int AB_SMEM = 2 * AB_TILE + 2 * 2 * AB_TILE + 2 * AB_PS + 128;   // 163,968 B
constexpr int AB_THREADS = 192;

struct AttnBwdArgs {
  const __nv_bfloat16* o; int ldo;
  const __nv_bfloat16* dout; int lddo;
  const float* lse;
  __nv_bfloat16* dq; int lddq;
  __nv_bfloat16* dk; int lddk;
  __nv_bfloat16* dv; int lddv;
  int Tq, Tk;
  float scale, scale_log2;
  int q_rows, kv_rows;
  DropoutCfg drop;
};

__device__ __forceinline__ void store_swizzled_row32(uint8_t* tile_row, int rx, int ch, const uint32_t (&pk)[16]) {
  // 32 columns (64 B) of one 128-column bf16 row: atom = ch / 2 (64 columns, 16 KB), chunks (ch & 1) * 4 .. + 3
  uint8_t* atom = tile_row + (ch >> 1) * (AB_T * 128);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int chunk = ((ch & 1) * 4 + q) ^ rx;
    *reinterpret_cast<uint4*>(atom + chunk * 16) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
  }
}

template <bool DROP>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const AttnBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023) __trap();
  uint8_t* sQ = smem;
  uint8_t* sDO = sQ + AB_TILE;
  uint8_t* sKV = sDO + AB_TILE;                 // stage s: K at s*32K, V at s*32K + 16K
  uint8_t* sP = sKV + 4 * AB_TILE;
  uint8_t* sDS = sP + AB_PS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + AB_PS);
  uint64_t* q_full = bars;          // Q + dO landed
  uint64_t* kv_full = bars + 1;     // [2]
  uint64_t* kv_empty = bars + 3;    // [2]
  uint64_t* sdp_full = bars + 5;    // S and dP in TMEM
  uint64_t* pds_ready = bars + 6;   // P and dS tiles in smem (128 arrivals)
  uint64_t* dkv_full = bars + 7;    // dV, dK (and the dQ update) complete
  uint64_t* dkv_empty = bars + 8;   // dV / dK read out of TMEM (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x, batch = blockIdx.y;
  const int n_kv = (a.Tk + AB_T - 1) / AB_T;

  // rows the TMA boxes never write must read as zeros: they are reduction rows of the MN-major products
  if (a.q_rows < AB_T || a.kv_rows < AB_T) {
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < 6 * AB_TILE / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(sdp_full, 1); mbar_init(pds_ready, 128); mbar_init(dkv_full, 1); mbar_init(dkv_empty, 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 320, tDQ = tmem_base + 384;

  // Warps 0 and 1 run converged; each TMA / tcgen05 group is handed to one lane with elect.sync (see attention.cu).
  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * a.q_rows * AB_HD * 2);
      tma_load_3d(sQ, &tmQ, q_full, head * AB_HD, 0, batch);
      tma_load_3d(sDO, &tmDO, q_full, head * AB_HD, 0, batch);
    }
    for (int j = 0; j < n_kv; ++j) {
      const int st = j & 1;
      mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&kv_full[st], 2 * a.kv_rows * AB_HD * 2);
        tma_load_3d(sKV + st * 2 * AB_TILE, &tmK, &kv_full[st], head * AB_HD, j * AB_T, batch);
        tma_load_3d(sKV + st * 2 * AB_TILE + AB_TILE, &tmV, &kv_full[st], head * AB_HD, j * AB_T, batch);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t id_kk_128 = umma_idesc_bf16(128, 128, 0, 0);    // S, dP
    constexpr uint32_t id_mm_64 = umma_idesc_bf16(128, 64, 1, 1);      // dV, dK: A and B MN-major
    constexpr uint32_t id_km_64 = umma_idesc_bf16(128, 64, 0, 1);      // dQ: A K-major, B MN-major
    const uint64_t q_k = umma_desc_sw128(smem_u32(sQ), 16, 1024);      // Q as K-major A
    const uint64_t do_k = umma_desc_sw128(smem_u32(sDO), 16, 1024);    // dO as K-major A
    const uint64_t q_mn = umma_desc_sw128(smem_u32(sQ), 1024, 1024);   // Q as MN-major B (reduction = query rows)
    const uint64_t do_mn = umma_desc_sw128(smem_u32(sDO), 1024, 1024);
    const uint64_t p_mn = umma_desc_sw128(smem_u32(sP), AB_T * 128, 1024);    // P^T: M = key columns (2 atoms of 16 KB)
    const uint64_t ds_mn = umma_desc_sw128(smem_u32(sDS), AB_T * 128, 1024);  // dS^T
    const uint64_t ds_k = umma_desc_sw128(smem_u32(sDS), 16, 1024);           // dS as K-major A
    mbar_wait(q_full, 0);
    for (int j = 0; j < n_kv; ++j) {
      const int st = j & 1;
      const uint32_t sK = smem_u32(sKV + st * 2 * AB_TILE), sV = sK + AB_TILE;
      mbar_wait(&kv_full[st], (j >> 1) & 1);
      tcgen05_fence_after();
      const uint64_t k_k = umma_desc_sw128(sK, 16, 1024), v_k = umma_desc_sw128(sV, 16, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < AB_HD / 16; ++k) umma_bf16_ss(tS, q_k + 2 * k, k_k + 2 * k, id_kk_128, k != 0);
#pragma unroll
        for (int k = 0; k < AB_HD / 16; ++k) umma_bf16_ss(tDP, do_k + 2 * k, v_k + 2 * k, id_kk_128, k != 0);
        umma_commit(sdp_full);
      }
      __syncwarp();
      mbar_wait(pds_ready, j & 1);                       // P, dS tiles written; S, dP consumed
      if (j > 0) mbar_wait(dkv_empty, (j - 1) & 1);      // previous dV / dK drained from TMEM
      tcgen05_fence_after();
      const uint64_t k_mn = umma_desc_sw128(sK, 1024, 1024);           // K as MN-major B (reduction = key rows)
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < AB_T / 16; ++k) {               // reduction over the 128 query rows, 16 per step = 2048 B
          umma_bf16_ss(tDV, p_mn + 128 * k, do_mn + 128 * k, id_mm_64, k != 0);
          umma_bf16_ss(tDK, ds_mn + 128 * k, q_mn + 128 * k, id_mm_64, k != 0);
        }
#pragma unroll
        for (int k = 0; k < AB_T / 16; ++k) {               // reduction over the 128 keys of this block
          const uint64_t dsd = ds_k + (uint64_t)((k >> 2) * (AB_T * 128 / 16) + (k & 3) * 2);
          umma_bf16_ss(tDQ, dsd, k_mn + 128 * k, id_km_64, (j | k) != 0);
        }
        umma_commit(&kv_empty[st]);
        umma_commit(dkv_full);
      }
      __syncwarp();
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                       // query row (P / dS phase) or key row (dK / dV phase)
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const bool q_ok = r < a.Tq;
    // D = rowsum(dO o O), LSE
    float Drow = 0.f, lse = 0.f;
    if (q_ok) {
      const size_t grow = (size_t)batch * a.Tq + r;
      const uint4* po = reinterpret_cast<const uint4*>(a.o + grow * a.ldo + head * AB_HD);
      const uint4* pd = reinterpret_cast<const uint4*>(a.dout + grow * a.lddo + head * AB_HD);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 x = __ldg(po + i), y = __ldg(pd + i);
        const __nv_bfloat162* xb = reinterpret_cast<const __nv_bfloat162*>(&x);
        const __nv_bfloat162* yb = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 xf = __bfloat1622float2(xb[t]), yf = __bfloat1622float2(yb[t]);
          Drow = fmaf(xf.x, yf.x, fmaf(xf.y, yf.y, Drow));
        }
      }
      lse = a.lse[((size_t)batch * gridDim.x + head) * a.Tq + r];
    }
    uint8_t* prow = sP + r * 128;
    uint8_t* dsrow = sDS + r * 128;
    const int rx = r & 7;
    const uint64_t grp_row = DROP ? (((uint64_t)batch * gridDim.x + head) * a.Tq + r) * (uint64_t)((a.Tk + 7) >> 3) : 0;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(sdp_full, j & 1);
      tcgen05_fence_after();
      const int kv_valid = a.Tk - j * AB_T;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t s[32], dp[32];
        tmem_ld_32x32(tS + lane_off + ch * 32, s);
        tmem_ld_32x32(tDP + lane_off + ch * 32, dp);
        tmem_wait_ld();
        uint32_t pk[16], dk[16];
        uint32_t keep = 0;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = 0.f, p1 = 0.f;
          if (q_ok && ch * 32 + i < kv_valid) p0 = fast_exp2(fmaf(__uint_as_float(s[i]), a.scale_log2, -lse));
          if (q_ok && ch * 32 + i + 1 < kv_valid) p1 = fast_exp2(fmaf(__uint_as_float(s[i + 1]), a.scale_log2, -lse));
          float z0 = 1.f, z1 = 1.f;
          if (DROP) {
            if ((i & 7) == 0) keep = dropout_keep8(a.drop, grp_row + (uint64_t)((j * AB_T + ch * 32 + i) >> 3));
            z0 = ((keep >> (i & 7)) & 1) ? a.drop.scale : 0.f;
            z1 = ((keep >> (i & 7)) & 2) ? a.drop.scale : 0.f;
          }
          const float d0 = p0 * (z0 * __uint_as_float(dp[i]) - Drow) * a.scale;
          const float d1 = p1 * (z1 * __uint_as_float(dp[i + 1]) - Drow) * a.scale;
          pk[i >> 1] = pack_bf16(p0 * z0, p1 * z1);
          dk[i >> 1] = pack_bf16(d0, d1);
        }
        store_swizzled_row32(prow, rx, ch, pk);
        store_swizzled_row32(dsrow, rx, ch, dk);
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      mbar_arrive(pds_ready);
      // dV / dK of this key block: thread = key row
      mbar_wait(dkv_full, j & 1);
      tcgen05_fence_after();
      const int krow = j * AB_T + r;
      const bool k_ok = krow < a.Tk;
      const size_t gk = (size_t)batch * a.Tk + krow;
#pragma unroll 1
      for (int which = 0; which < 2; ++which) {
        __nv_bfloat16* dst = which == 0 ? a.dv + gk * a.lddv + head * AB_HD : a.dk + gk * a.lddk + head * AB_HD;
        const uint32_t t = which == 0 ? tDV : tDK;
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32(t + lane_off + ch * 32, v);
          tmem_wait_ld();
          if (k_ok) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              uint4 u;
              u.x = pack_bf16(__uint_as_float(v[i]), __uint_as_float(v[i + 1]));
              u.y = pack_bf16(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
              u.z = pack_bf16(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
              u.w = pack_bf16(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
              *reinterpret_cast<uint4*>(dst + ch * 32 + i) = u;
            }
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(dkv_empty);
    }
    // dQ (accumulated over all key blocks; dkv_full of the last block also covers the last dQ MMAs)
    {
      __nv_bfloat16* dst = a.dq + ((size_t)batch * a.Tq + r) * a.lddq + head * AB_HD;
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32(tDQ + lane_off + ch * 32, v);
        tmem_wait_ld();
        if (q_ok) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 u;
            u.x = pack_bf16(__uint_as_float(v[i]), __uint_as_float(v[i + 1]));
            u.y = pack_bf16(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
            u.z = pack_bf16(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
            u.w = pack_bf16(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
            *reinterpret_cast<uint4*>(dst + ch * 32 + i) = u;
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Transposed variant for short query sequences (Tq <= 64: the Q-Former's 64 learned queries).  With the queries on the M
// axis half of every 128-row MMA tile is padding; here the KEYS are the M axis, so all 128 TMEM lanes do useful work:
//   S^T  = K Q^T      A = K block (K-major)          B = Q  (K-major, N = 64 queries)        -> TMEM ST[j & 1]  (64 cols)
//   dP^T = V dO^T     A = V block (K-major)          B = dO (K-major)                        -> TMEM DPT[j & 1] (64 cols)
//   dV   = P^T dO     A = P^T tile (K-major, K = queries)   B = dO (MN-major)                -> TMEM DV (64 cols)
//   dK   = dS^T Q     A = dS^T tile                         B = Q  (MN-major)                -> TMEM DK
//   dQ  += dS K       A = dS^T tile read MN-major (M = queries; the second 64-row M atom points at the P^T tile, its
//                     output rows 64..127 are never read)   B = K block (MN-major)           -> TMEM DQ
// Two threads per KEY row (32 queries each) compute p = exp2(c s - lse[q]) and ds (lse / D per query held in registers) and writes the bf16 P^T / dS^T tiles (128 keys x 128 B).  S^T / dP^T are double-buffered in TMEM
// and P^T / dS^T in shared memory: the MMA warp computes block j+2's scores while the threads work on block j+1 and the
// tensor core finishes dV / dK / dQ of block j, whose read-out is interleaved one block later.
#ifdef DESTA_ATTN_TRACE
__device__ long long g_bwd_trace[8];
#define BT_T0() const long long _t0 = clock64()
#define BT_ADD(slot) do { bt_acc[slot] += clock64() - _t0; } while (0)
#else
#define BT_T0()
#define BT_ADD(slot)
#endif

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

namespace t64 {
constexpr int BK = 128, HD = 64, BQ = 64, KSTAGES = 4, VSTAGES = 2;
constexpr int Q_BYTES = BQ * HD * 2;                 // 8 KB: Q, dO
constexpr int KV_BYTES = BK * HD * 2;                // 16 KB: K stage, V stage
constexpr int PT_BYTES = BK * BQ * 2;                // 16 KB: P^T, dS^T (x2 buffers)
constexpr int OUT_BYTES = BK * HD * 2;               // 16 KB: dV / dK staging tile for the TMA store
constexpr int SMEM = 2 * Q_BYTES + (KSTAGES + VSTAGES) * KV_BYTES + 4 * PT_BYTES + 2 * OUT_BYTES + 1024;   // 214,016 B
constexpr int T_ST = 0, T_DPT = 128, T_DV = 256, T_DK = 320, T_DQ = 384;
constexpr int THREADS = 320;                         // TMA warp, MMA warp, 8 compute warps
}  // namespace t64

template <bool DROP>
__global__ void __launch_bounds__(t64::THREADS, 1)
attn_bwd_t64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                    const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                    const __grid_constant__ CUtensorMap tmDK, const __grid_constant__ CUtensorMap tmDV, const AttnBwdArgs a) {
  using namespace t64;
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023) __trap();
  uint8_t* sQ = smem;
  uint8_t* sDO = sQ + Q_BYTES;
  uint8_t* sK = sDO + Q_BYTES;                         // K ring: 4 stages (a K block lives from S^T_j to dQ_j, two blocks later)
  uint8_t* sV = sK + KSTAGES * KV_BYTES;               // V ring: 2 stages (a V block is only read by dP^T_j)
  uint8_t* sPT = sV + VSTAGES * KV_BYTES;              // buffer b: dS^T at b*32K, P^T at b*32K + 16K
  uint8_t* sOut = sPT + 4 * PT_BYTES;                  // dV tile, dK tile (bf16, 128-byte rows, SWIZZLE_128B) for the TMA stores
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + 2 * OUT_BYTES);
  uint64_t* q_full = bars;            // Q + dO landed
  uint64_t* k_full = bars + 1;        // [4]
  uint64_t* k_empty = bars + 5;       // [4]  released after dQ_j
  uint64_t* v_full = bars + 9;        // [2]
  uint64_t* v_empty = bars + 11;      // [2]  released after dP^T_j
  uint64_t* sdp_full = bars + 13;     // [2]  S^T_j and dP^T_j in TMEM buffer j & 1
  uint64_t* pds_ready = bars + 15;    // [2]  P^T_j / dS^T_j tiles in smem buffer j & 1 (256 arrivals)
  uint64_t* dkv_full = bars + 17;     // dV_j, dK_j (and the dQ update) complete
  uint64_t* dkv_empty = bars + 18;    // dV / dK read out of TMEM (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);
  float* s_lse = reinterpret_cast<float*>(bars + 24);  // [64] log2-domain LSE per query (+inf for q >= Tq)
  float* s_D = s_lse + 64;                             // [64] rowsum(dO o O)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x, batch = blockIdx.y;
  const int n_kv = (a.Tk + BK - 1) / BK;
#ifdef DESTA_ATTN_TRACE
  long long bt_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long bt_start = clock64();
#endif

  // rows the TMA boxes never write must read as zeros (queries >= Tq; keys beyond a short Tk)
  if (a.q_rows < BQ || a.kv_rows < BK) {
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < (2 * Q_BYTES + (KSTAGES + VSTAGES) * KV_BYTES) / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < KSTAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < VSTAGES; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sdp_full[i], 1); mbar_init(&pds_ready[i], 256); }
    mbar_init(dkv_full, 1); mbar_init(dkv_empty, 256);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();                    // barriers + TMEM ready: the TMA producer starts loading right away ...
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // ... while threads 64..127 fetch the per-query constants LSE and D = rowsum(dO o O) (one query each); the compute
  // warps meet again on a named barrier before they read them
  if (threadIdx.x >= 64 && threadIdx.x < 128) {
    const int q = threadIdx.x - 64;
    float Drow = 0.f, lse = INFINITY;
    if (q < a.Tq) {
      const size_t grow = (size_t)batch * a.Tq + q;
      const uint4* po = reinterpret_cast<const uint4*>(a.o + grow * a.ldo + head * HD);
      const uint4* pd = reinterpret_cast<const uint4*>(a.dout + grow * a.lddo + head * HD);
      uint4 x[8], y[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { x[i] = __ldg(po + i); y[i] = __ldg(pd + i); }
      lse = a.lse[((size_t)batch * gridDim.x + head) * a.Tq + q];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const __nv_bfloat162* xb = reinterpret_cast<const __nv_bfloat162*>(&x[i]);
        const __nv_bfloat162* yb = reinterpret_cast<const __nv_bfloat162*>(&y[i]);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 xf = __bfloat1622float2(xb[t]), yf = __bfloat1622float2(yb[t]);
          Drow = fmaf(xf.x, yf.x, fmaf(xf.y, yf.y, Drow));
        }
      }
    }
    s_lse[q] = lse;
    s_D[q] = Drow;
  }

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * a.q_rows * HD * 2);
      tma_load_3d(sQ, &tmQ, q_full, head * HD, 0, batch);
      tma_load_3d(sDO, &tmDO, q_full, head * HD, 0, batch);
    }
    int ks = 0, vs = 0; uint32_t kph = 1, vph = 1;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&v_empty[vs], vph);
      if (elect_one()) {
        mbar_arrive_expect_tx(&v_full[vs], a.kv_rows * HD * 2);
        tma_load_3d(sV + vs * KV_BYTES, &tmV, &v_full[vs], head * HD, j * BK, batch);
      }
      __syncwarp();
      mbar_wait(&k_empty[ks], kph);
      if (elect_one()) {
        mbar_arrive_expect_tx(&k_full[ks], a.kv_rows * HD * 2);
        tma_load_3d(sK + ks * KV_BYTES, &tmK, &k_full[ks], head * HD, j * BK, batch);
      }
      __syncwarp();
      if (++ks == KSTAGES) { ks = 0; kph ^= 1; }
      if (++vs == VSTAGES) { vs = 0; vph ^= 1; }
    }
  } else if (warp == 1) {
    constexpr uint32_t id_sdp = umma_idesc_bf16(128, BQ, 0, 0);      // S^T, dP^T: A, B K-major, N = 64 queries
    constexpr uint32_t id_dkv = umma_idesc_bf16(128, HD, 0, 1);      // dV, dK: A K-major, B MN-major
    constexpr uint32_t id_dq = umma_idesc_bf16(128, HD, 1, 1);       // dQ: A, B MN-major
    const uint64_t q_k = umma_desc_sw128(smem_u32(sQ), 16, 1024);    // Q as K-major B (N = queries)
    const uint64_t do_k = umma_desc_sw128(smem_u32(sDO), 16, 1024);
    const uint64_t q_mn = umma_desc_sw128(smem_u32(sQ), 1024, 1024);   // Q as MN-major B (reduction = query rows)
    const uint64_t do_mn = umma_desc_sw128(smem_u32(sDO), 1024, 1024);
    int s_ks = 0, s_vs = 0; uint32_t s_kph = 0, s_vph = 0;
    auto issue_sdp = [&](int j) {
      { BT_T0(); mbar_wait(&k_full[s_ks], s_kph); mbar_wait(&v_full[s_vs], s_vph); BT_ADD(0); }
      tcgen05_fence_after();
      const uint64_t k_k = umma_desc_sw128(smem_u32(sK + s_ks * KV_BYTES), 16, 1024);
      const uint64_t v_k = umma_desc_sw128(smem_u32(sV + s_vs * KV_BYTES), 16, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16_ss(tmem_base + T_DPT + (j & 1) * 64, v_k + 2 * k, do_k + 2 * k, id_sdp, k != 0);
        umma_commit(&v_empty[s_vs]);                      // the V block is only read by dP^T
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16_ss(tmem_base + T_ST + (j & 1) * 64, k_k + 2 * k, q_k + 2 * k, id_sdp, k != 0);
        umma_commit(&sdp_full[j & 1]);
      }
      __syncwarp();
      if (++s_ks == KSTAGES) { s_ks = 0; s_kph ^= 1; }
      if (++s_vs == VSTAGES) { s_vs = 0; s_vph ^= 1; }
    };
    mbar_wait(q_full, 0);
    issue_sdp(0);
    if (n_kv > 1) issue_sdp(1);
    int st = 0;
    for (int j = 0; j < n_kv; ++j) {
      { BT_T0(); mbar_wait(&pds_ready[j & 1], (j >> 1) & 1); BT_ADD(1); }        // P^T_j, dS^T_j written; S^T_j, dP^T_j consumed
      if (j > 0) { BT_T0(); mbar_wait(dkv_empty, (j - 1) & 1); BT_ADD(2); }      // dV / dK of the previous block drained from TMEM
      tcgen05_fence_after();
      const uint32_t sKb = smem_u32(sK + st * KV_BYTES);
      const uint32_t sDS = smem_u32(sPT + (j & 1) * 2 * PT_BYTES), sP = sDS + PT_BYTES;
      const uint64_t p_k = umma_desc_sw128(sP, 16, 1024), ds_k = umma_desc_sw128(sDS, 16, 1024);
      // dS^T tile as MN-major A (M = queries, 64 per 128-byte row; K = key rows, 8-row groups 1024 B apart).  The
      // instruction needs M = 128: the second 64-row M atom (LBO = 16 KB) is the P^T tile right behind it - output rows
      // 64..127 of dQ are finite garbage by construction and never read.
      const uint64_t ds_mn = umma_desc_sw128(sDS, PT_BYTES, 1024);
      const uint64_t k_mn = umma_desc_sw128(sKb, 1024, 1024);          // K block as MN-major B (reduction = key rows)
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < BQ / 16; ++k) {                // reduction over the 64 queries, 16 per step
          umma_bf16_ss(tmem_base + T_DV, p_k + 2 * k, do_mn + 128 * k, id_dkv, k != 0);
          umma_bf16_ss(tmem_base + T_DK, ds_k + 2 * k, q_mn + 128 * k, id_dkv, k != 0);
        }
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)                  // reduction over the 128 keys of this block, 16 rows = 2048 B per step
          umma_bf16_ss(tmem_base + T_DQ, ds_mn + 128 * k, k_mn + 128 * k, id_dq, (j | k) != 0);
        umma_commit(&k_empty[st]);
        umma_commit(dkv_full);
      }
      __syncwarp();
      if (j + 2 < n_kv) issue_sdp(j + 2);
      if (++st == KSTAGES) st = 0;
    }
  } else {
    // 8 compute warps = two per SMSP: warp pair (w, w + 4) shares a TMEM lane quarter (128 key rows in all) and splits the
    // 64 query columns; there is no row-wise reduction in the backward, so the halves never talk to each other.
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;                       // key row inside the block
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int rx = r & 7;
    const uint64_t row_groups = (uint64_t)((a.Tk + 7) >> 3);
    const uint64_t grp_base = DROP ? ((uint64_t)batch * gridDim.x + head) * a.Tq * row_groups : 0;
    asm volatile("bar.sync 2, 256;" ::: "memory");           // s_lse / s_D written by the first two compute warps
    float lse_r[32], D_r[32];                                // per-query constants of this thread's 32 columns
#pragma unroll
    for (int i = 0; i < 32; ++i) { lse_r[i] = s_lse[half * 32 + i]; D_r[i] = s_D[half * 32 + i]; }
    auto readout = [&](int j) {       // key block j: half 0 drains dV, half 1 drains dK; one key row per thread
      { BT_T0(); mbar_wait(dkv_full, j & 1); if (warp == 2) BT_ADD(4); }
      tcgen05_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(tmem_base + (half == 0 ? T_DV : T_DK) + lane_off, v0);
      tmem_ld_32x32(tmem_base + (half == 0 ? T_DV : T_DK) + lane_off + 32, v1);
      tmem_wait_ld();
      tcgen05_fence_before();
      mbar_arrive(dkv_empty);          // accumulator is in registers: the next block's dV / dK MMAs may start
      // Row-per-thread global stores would touch 32 different 128-byte lines per instruction; instead the bf16 tile is
      // staged in shared memory (swizzled 128-byte rows) and written by ONE TMA store per tensor, which also clips the
      // ragged last key block at Tk.
      if (threadIdx.x == 64) bulk_wait_read0();                      // the previous block's stores have read the staging tiles
      asm volatile("bar.sync 1, 256;" ::: "memory");
      uint8_t* orow = sOut + half * OUT_BYTES + r * 128;
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8)
        *reinterpret_cast<uint4*>(orow + ((c8 ^ rx) << 4)) =
            make_uint4(pack_bf16(__uint_as_float(v0[8 * c8]), __uint_as_float(v0[8 * c8 + 1])), pack_bf16(__uint_as_float(v0[8 * c8 + 2]), __uint_as_float(v0[8 * c8 + 3])),
                       pack_bf16(__uint_as_float(v0[8 * c8 + 4]), __uint_as_float(v0[8 * c8 + 5])), pack_bf16(__uint_as_float(v0[8 * c8 + 6]), __uint_as_float(v0[8 * c8 + 7])));
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8)
        *reinterpret_cast<uint4*>(orow + (((4 + c8) ^ rx) << 4)) =
            make_uint4(pack_bf16(__uint_as_float(v1[8 * c8]), __uint_as_float(v1[8 * c8 + 1])), pack_bf16(__uint_as_float(v1[8 * c8 + 2]), __uint_as_float(v1[8 * c8 + 3])),
                       pack_bf16(__uint_as_float(v1[8 * c8 + 4]), __uint_as_float(v1[8 * c8 + 5])), pack_bf16(__uint_as_float(v1[8 * c8 + 6]), __uint_as_float(v1[8 * c8 + 7])));
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 64) {
        tma_store_3d(&tmDV, sOut, head * HD, j * BK, batch);
        tma_store_3d(&tmDK, sOut + OUT_BYTES, head * HD, j * BK, batch);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    };
    for (int j = 0; j < n_kv; ++j) {
      { BT_T0(); mbar_wait(&sdp_full[j & 1], (j >> 1) & 1); if (warp == 2) BT_ADD(3); }
      tcgen05_fence_after();
      const bool k_ok = j * BK + r < a.Tk;
      uint8_t* dsrow = sPT + (j & 1) * 2 * PT_BYTES + r * 128;
      uint8_t* prow = dsrow + PT_BYTES;
      // dropout masks of this thread's octet of keys: lane o of an octet evaluates the Philox group (query o + 8 i, this
      // octet's 8 keys) for the four i of this warp's column half; bit (lane % 8) of the mask held by lane (q % 8) is this
      // thread's keep bit for query q
      uint32_t km[4];
      if (DROP) {
        const uint64_t kgrp = (uint64_t)((j * BK + r) >> 3);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          km[i] = dropout_keep8(a.drop, grp_base + (uint64_t)((lane & 7) + 8 * (half * 4 + i)) * row_groups + kgrp);
      }
      uint32_t s[32], dp[32];
      tmem_ld_32x32(tmem_base + T_ST + (j & 1) * 64 + lane_off + half * 32, s);
      tmem_ld_32x32(tmem_base + T_DPT + (j & 1) * 64 + lane_off + half * 32, dp);
      tmem_wait_ld();
      uint32_t pk[16], dk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float p0 = k_ok ? fast_exp2(fmaf(__uint_as_float(s[i]), a.scale_log2, -lse_r[i])) : 0.f;
        float p1 = k_ok ? fast_exp2(fmaf(__uint_as_float(s[i + 1]), a.scale_log2, -lse_r[i + 1])) : 0.f;
        float z0 = 1.f, z1 = 1.f;
        if (DROP) {
          const uint32_t m0 = __shfl_sync(0xffffffffu, km[i >> 3], (lane & 24) | (i & 7));
          const uint32_t m1 = __shfl_sync(0xffffffffu, km[(i + 1) >> 3], (lane & 24) | ((i + 1) & 7));
          z0 = ((m0 >> (lane & 7)) & 1) ? a.drop.scale : 0.f;
          z1 = ((m1 >> (lane & 7)) & 1) ? a.drop.scale : 0.f;
        }
        const float ds0 = p0 * (z0 * __uint_as_float(dp[i]) - D_r[i]) * a.scale;
        const float ds1 = p1 * (z1 * __uint_as_float(dp[i + 1]) - D_r[i + 1]) * a.scale;
        pk[i >> 1] = pack_bf16(p0 * z0, p1 * z1);
        dk[i >> 1] = pack_bf16(ds0, ds1);
      }
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {              // this half's 64 bytes = 16-byte chunks half*4 .. half*4+3 of the 128-byte row
        const int chunk = (half * 4 + c4) ^ rx;
        *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk[4 * c4], pk[4 * c4 + 1], pk[4 * c4 + 2], pk[4 * c4 + 3]);
        *reinterpret_cast<uint4*>(dsrow + chunk * 16) = make_uint4(dk[4 * c4], dk[4 * c4 + 1], dk[4 * c4 + 2], dk[4 * c4 + 3]);
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      mbar_arrive(&pds_ready[j & 1]);
      if (j > 0) readout(j - 1);          // overlaps the tensor core working on block j
    }
    readout(n_kv - 1);
    if (threadIdx.x == 64) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // dV / dK stores complete before exit
    // dQ (accumulated over all key blocks; the last dkv_full also covers the last dQ MMAs): rows = queries = lanes 0..63
    if (quarter < 2) {
      __nv_bfloat16* dst = a.dq + ((size_t)batch * a.Tq + r) * a.lddq + head * HD + half * 32;
      uint32_t v0[32];
      tmem_ld_32x32(tmem_base + T_DQ + lane_off + half * 32, v0);
      tmem_wait_ld();
      if (r < a.Tq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8)
          *reinterpret_cast<uint4*>(dst + i) = make_uint4(pack_bf16(__uint_as_float(v0[i]), __uint_as_float(v0[i + 1])),
                                                          pack_bf16(__uint_as_float(v0[i + 2]), __uint_as_float(v0[i + 3])),
                                                          pack_bf16(__uint_as_float(v0[i + 4]), __uint_as_float(v0[i + 5])),
                                                          pack_bf16(__uint_as_float(v0[i + 6]), __uint_as_float(v0[i + 7])));
      }
    }
  }
#ifdef DESTA_ATTN_TRACE
  if (warp == 1 || warp == 2) bt_acc[warp == 1 ? 5 : 6] = clock64() - bt_start;
  if (lane == 0) for (int i = 0; i < 8; ++i) if (bt_acc[i]) atomicAdd((unsigned long long*)&g_bwd_trace[i], (unsigned long long)bt_acc[i]);
#endif
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

static int make_tmap3(CUtensorMap* m, const void* base, int heads, int T, int ld, int batch, int box_rows) {
  uint64_t dims[3] = {(uint64_t)heads * AB_HD, (uint64_t)T, (uint64_t)batch};
  uint64_t strides[3] = {0, (uint64_t)ld * 2, (uint64_t)T * ld * 2};
  uint32_t box[3] = {AB_HD, (uint32_t)box_rows, 1};
  return make_tmap(m, base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace desta

using namespace desta;

#ifdef DESTA_ATTN_TRACE
extern "C" int desta_attention_bwd_trace(long long* host8, int reset) {
  long long z[8] = {0};
  if (reset) return cudaMemcpyToSymbol(g_bwd_trace, z, sizeof(z)) == cudaSuccess ? 0 : -2;
  return cudaMemcpyFromSymbol(host8, g_bwd_trace, sizeof(z)) == cudaSuccess ? 0 : -2;
}
#endif

extern "C" int desta_attention_bwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const void* o,
                                   int ldo, const void* dout, int lddo, const float* lse, void* dq, int lddq, void* dk,
                                   int lddk, void* dv, int lddv, int batch, int heads, int Tq, int Tk, float scale,
                                   const desta_dropout* drop, void* stream) {
  if (!q || !k || !v || !o || !dout || !lse || !dq || !dk || !dv) { set_error("desta_attention_bwd: null pointer"); return DESTA_ERR_ARG; }
  if (batch <= 0 || heads <= 0 || Tq <= 0 || Tk <= 0) { set_error("desta_attention_bwd: bad shape"); return DESTA_ERR_ARG; }
  if (Tq > AB_T) { set_error("desta_attention_bwd: at most 128 queries per sequence (Q-Former prompts), got %d", Tq); return DESTA_ERR_SHAPE; }
  const bool t64k = Tq <= t64::BQ;      // short query sequences: transposed kernel (keys on the M axis)
  const int lds[8] = {ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv};
  for (int i = 0; i < 8; ++i)
    if ((lds[i] % 8) || lds[i] < heads * 64) { set_error("desta_attention_bwd: leading dims must be multiples of 8 and >= heads*64"); return DESTA_ERR_ARG; }
  if (heads > 65535 || batch > 65535) { set_error("desta_attention_bwd: heads/batch too large"); return DESTA_ERR_ARG; }
  AttnBwdArgs a;
  a.o = (const __nv_bfloat16*)o; a.ldo = ldo; a.dout = (const __nv_bfloat16*)dout; a.lddo = lddo; a.lse = lse;
  a.dq = (__nv_bfloat16*)dq; a.lddq = lddq; a.dk = (__nv_bfloat16*)dk; a.lddk = lddk; a.dv = (__nv_bfloat16*)dv; a.lddv = lddv;
  a.Tq = Tq; a.Tk = Tk; a.scale = scale; a.scale_log2 = scale * 1.4426950408889634f;
  a.q_rows = Tq < AB_T ? Tq : AB_T;
  a.kv_rows = Tk < AB_T ? Tk : AB_T;
  a.drop = make_dropout_cfg(drop);
  if (t64k) {
    CUtensorMap tmQ, tmDO, tmK, tmV;
    if (make_tmap3(&tmQ, q, heads, Tq, ldq, batch, a.q_rows)) return DESTA_ERR_CUDA;
    if (make_tmap3(&tmDO, dout, heads, Tq, lddo, batch, a.q_rows)) return DESTA_ERR_CUDA;
    if (make_tmap3(&tmK, k, heads, Tk, ldk, batch, a.kv_rows)) return DESTA_ERR_CUDA;
    if (make_tmap3(&tmV, v, heads, Tk, ldv, batch, a.kv_rows)) return DESTA_ERR_CUDA;
    CUtensorMap tmDK, tmDV;
    if (make_tmap3(&tmDK, dk, heads, Tk, lddk, batch, a.kv_rows)) return DESTA_ERR_CUDA;
    if (make_tmap3(&tmDV, dv, heads, Tk, lddv, batch, a.kv_rows)) return DESTA_ERR_CUDA;
    bool& attr64 = *device_once_flag(3);
    if (!attr64) {
      cudaError_t e = cudaFuncSetAttribute(attn_bwd_t64_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, t64::SMEM);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_t64_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, t64::SMEM);
      if (e != cudaSuccess) { set_error("attention bwd (t64) smem attribute: %s", cudaGetErrorString(e)); return DESTA_ERR_CUDA; }
      attr64 = true;
    }
    if (a.drop.thr16) attn_bwd_t64_kernel<true><<<dim3(heads, batch), t64::THREADS, t64::SMEM, (cudaStream_t)stream>>>(tmQ, tmDO, tmK, tmV, tmDK, tmDV, a);
    else attn_bwd_t64_kernel<false><<<dim3(heads, batch), t64::THREADS, t64::SMEM, (cudaStream_t)stream>>>(tmQ, tmDO, tmK, tmV, tmDK, tmDV, a);
    return check_launch("attn_bwd_t64_kernel");
  }
  CUtensorMap tmQ, tmDO, tmK, tmV;
  if (make_tmap3(&tmQ, q, heads, Tq, ldq, batch, a.q_rows)) return DESTA_ERR_CUDA;
  if (make_tmap3(&tmDO, dout, heads, Tq, lddo, batch, a.q_rows)) return DESTA_ERR_CUDA;
  if (make_tmap3(&tmK, k, heads, Tk, ldk, batch, a.kv_rows)) return DESTA_ERR_CUDA;
  if (make_tmap3(&tmV, v, heads, Tk, ldv, batch, a.kv_rows)) return DESTA_ERR_CUDA;
  bool& attr_set = *device_once_flag(4);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
    if (e != cudaSuccess) { set_error("attention bwd smem attribute: %s", cudaGetErrorString(e)); return DESTA_ERR_CUDA; }
    attr_set = true;
  }
  if (a.drop.thr16) attn_bwd_kernel<true><<<dim3(heads, batch), AB_THREADS, AB_SMEM, (cudaStream_t)stream>>>(tmQ, tmDO, tmK, tmV, a);
  else attn_bwd_kernel<false><<<dim3(heads, batch), AB_THREADS, AB_SMEM, (cudaStream_t)stream>>>(tmQ, tmDO, tmK, tmV, a);
  return check_launch("attn_bwd_kernel");
}
