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

static int make_tmap3(CUtensorMap* m, const void* base, int heads, int T, int ld, int batch, int box_rows) {
  uint64_t dims[3] = {(uint64_t)heads * AB_HD, (uint64_t)T, (uint64_t)batch};
  uint64_t strides[3] = {0, (uint64_t)ld * 2, (uint64_t)T * ld * 2};
  uint32_t box[3] = {AB_HD, (uint32_t)box_rows, 1};
  return make_tmap(m, base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace desta

using namespace desta;

extern "C" int desta_attention_bwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const void* o,
                                   int ldo, const void* dout, int lddo, const float* lse, void* dq, int lddq, void* dk,
                                   int lddk, void* dv, int lddv, int batch, int heads, int Tq, int Tk, float scale,
                                   const desta_dropout* drop, void* stream) {
  if (!q || !k || !v || !o || !dout || !lse || !dq || !dk || !dv) { set_error("desta_attention_bwd: null pointer"); return DESTA_ERR_ARG; }
  if (batch <= 0 || heads <= 0 || Tq <= 0 || Tk <= 0) { set_error("desta_attention_bwd: bad shape"); return DESTA_ERR_ARG; }
  if (Tq > AB_T) { set_error("desta_attention_bwd: at most 128 queries per sequence (Q-Former prompts), got %d", Tq); return DESTA_ERR_SHAPE; }
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
  CUtensorMap tmQ, tmDO, tmK, tmV;
  if (make_tmap3(&tmQ, q, heads, Tq, ldq, batch, a.q_rows)) return DESTA_ERR_CUDA;
  if (make_tmap3(&tmDO, dout, heads, Tq, lddo, batch, a.q_rows)) return DESTA_ERR_CUDA;
  if (make_tmap3(&tmK, k, heads, Tk, ldk, batch, a.kv_rows)) return DESTA_ERR_CUDA;
  if (make_tmap3(&tmV, v, heads, Tk, ldv, batch, a.kv_rows)) return DESTA_ERR_CUDA;
  static bool attr_set = false;
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
