// <|AUDIO|> placeholder splice as ONE gather/scatter kernel driven by an integer row map.
// Replaces the gathers + N-iteration Python scatter loop of DeSTA25AudioModel._prepare_inputs_for_llm
// (ref: desta/models/modeling_desta25.py:940-1050; embed_tokens gathers :975-982,1009; scatter loop :1014-1045).
//
// HBM-bound: one warp per output row; 16-byte vector loads from the embedding table / audio features,
//   fp32 -> bf16 round-to-nearest-even for audio rows (the implicit cast of the slice assignment at :1040-1041),
//   transcript / text rows copied bit-for-bit.
#include "common.cuh"
#include "desta_b200.h"

namespace desta {

// Each warp owns output rows; the 32 lanes scan the clip descriptors (staged in shared memory) to find the LAST clip
// whose span [start, start + Q + T_i) covers the row - "later clips win", exactly the reference's loop order - so no
// row-map pass and no scratch buffer are needed: ONE launch.
// A token id outside [0, vocab) makes the reference's embedding lookup raise (IndexError on the CPU, a device-side assert on
// CUDA; modeling_desta25.py:981,1009).  The kernel stays memory-safe (the row is read from a clamped id) and REPORTS it:
// err_word[0] counts offending rows, err_word[1] keeps one offending id; the host raises from that word (ops.splice).
__device__ __forceinline__ int64_t checked_id(int64_t id, int64_t vocab, int64_t* err_word, int lane) {
  if (id >= 0 && id < vocab) return id;
  if (err_word && lane == 0) {
    atomicAdd(reinterpret_cast<unsigned long long*>(err_word), 1ull);
    err_word[1] = id;
  }
  return id < 0 ? 0 : vocab - 1;
}

template <bool FEAT_F32>
__global__ void __launch_bounds__(256)
splice_rows_kernel(const int64_t* __restrict__ input_ids, const uint4* __restrict__ table, int64_t vocab,
                   const desta_splice_desc* __restrict__ descs, int n_desc, const int64_t* __restrict__ trans_ids,
                   const void* __restrict__ feats, int Q, int S, int d8, int rows, uint4* __restrict__ out,
                   int64_t* __restrict__ err_word) {
  extern __shared__ desta_splice_desc s_desc[];
  for (int i = threadIdx.x; i < n_desc; i += blockDim.x) s_desc[i] = descs[i];
  __syncthreads();
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows; row += gridDim.x * warps_per_block) {
    const int b = row / S, pos = row - b * S;
    int hit = -1;
    for (int c = lane; c < n_desc; c += 32) {
      const desta_splice_desc dd = s_desc[c];
      if (dd.prompt == b && pos >= dd.start && pos < dd.start + Q + dd.n_trans) hit = c;   // ascending c: last one wins
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) hit = max(hit, __shfl_xor_sync(0xffffffffu, hit, o));
    uint4* o = out + (size_t)row * d8;
    const uint4* src = nullptr;
    if (hit < 0) {
      src = table + (size_t)checked_id(input_ids[row], vocab, err_word, lane) * d8;
    } else {
      const int j = pos - s_desc[hit].start;
      if (j >= Q) {
        src = table + (size_t)checked_id(trans_ids[s_desc[hit].trans_off + (j - Q)], vocab, err_word, lane) * d8;
      } else if (!FEAT_F32) {
        src = reinterpret_cast<const uint4*>(feats) + ((size_t)hit * Q + j) * d8;
      } else {
        const float4* f = reinterpret_cast<const float4*>(feats) + ((size_t)hit * Q + j) * d8 * 2;
        for (int i = lane; i < d8; i += 32) {
          const float4 a = f[2 * i], bb = f[2 * i + 1];
          uint4 u;
          u.x = pack_bf16(a.x, a.y); u.y = pack_bf16(a.z, a.w);
          u.z = pack_bf16(bb.x, bb.y); u.w = pack_bf16(bb.z, bb.w);
          o[i] = u;
        }
        continue;
      }
    }
    for (int i = lane; i < d8; i += 32) o[i] = __ldg(src + i);
  }
}

__global__ void __launch_bounds__(256)
splice_bwd_kernel(const uint4* __restrict__ grad_out, const desta_splice_desc* __restrict__ descs, int Q, int S, int d8,
                  float4* __restrict__ dfeats) {
  const int c = blockIdx.y;
  const desta_splice_desc dd = descs[c];
  for (int j = blockIdx.x; j < Q; j += gridDim.x) {
    const uint4* src = grad_out + ((size_t)dd.prompt * S + dd.start + j) * d8;
    float4* dst = dfeats + ((size_t)c * Q + j) * d8 * 2;
    for (int i = threadIdx.x; i < d8; i += blockDim.x) {
      const uint4 u = __ldg(src + i);
      const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&u);
      const float2 f0 = __bfloat1622float2(b[0]), f1 = __bfloat1622float2(b[1]), f2 = __bfloat1622float2(b[2]), f3 = __bfloat1622float2(b[3]);
      dst[2 * i] = make_float4(f0.x, f0.y, f1.x, f1.y);
      dst[2 * i + 1] = make_float4(f2.x, f2.y, f3.x, f3.y);
    }
  }
}


}  // namespace desta

using namespace desta;

extern "C" int desta_splice(const int64_t* input_ids, const void* table, int64_t vocab, const desta_splice_desc* descs_dev,
                            const desta_splice_desc* descs_host, int n_desc, const int64_t* trans_ids,
                            const void* audio_feats, int feat_dtype, int Q, void* out, int B, int S, int d,
                            int64_t* err_word, void* stream) {
  if (!input_ids || !table || !out) { set_error("desta_splice: null pointer"); return DESTA_ERR_ARG; }
  if (B <= 0 || S <= 0) return DESTA_OK;
  if (d % 8) { set_error("desta_splice: hidden size must be a multiple of 8 (got %d)", d); return DESTA_ERR_ARG; }
  if (n_desc > 0 && (!descs_dev || !descs_host || !audio_feats)) { set_error("desta_splice: null descs/feats"); return DESTA_ERR_ARG; }
  if (n_desc >= 32768) { set_error("desta_splice: too many clips (%d)", n_desc); return DESTA_ERR_ARG; }
  for (int c = 0; c < n_desc; ++c) {
    const desta_splice_desc& dd = descs_host[c];
    const long len = (long)Q + dd.n_trans;
    if (dd.prompt < 0 || dd.prompt >= B || dd.start < 0 || dd.n_trans < 0 || dd.start + len > S || len >= 65536) {
      set_error("desta_splice: clip %d span [%d, %ld) of prompt %d leaves the sequence (B=%d, S=%d)", c, dd.start,
                dd.start + len, dd.prompt, B, S);
      return DESTA_ERR_SHAPE;
    }
    if (dd.n_trans > 0 && !trans_ids) { set_error("desta_splice: null trans_ids"); return DESTA_ERR_ARG; }
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int rows = B * S;
  int grid = (rows + 7) / 8;
  const int cap = num_sms() * 8;
  if (grid > cap) grid = cap;
  const size_t sm = (size_t)(n_desc > 0 ? n_desc : 1) * sizeof(desta_splice_desc);
  if (sm > 200 * 1024) { set_error("desta_splice: too many clips for one launch (%d)", n_desc); return DESTA_ERR_ARG; }
  if (sm > 48 * 1024) {
    cudaFuncSetAttribute(splice_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(splice_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  }
  if (feat_dtype == 0)
    splice_rows_kernel<true><<<grid, 256, sm, s>>>(input_ids, (const uint4*)table, vocab, descs_dev, n_desc, trans_ids, audio_feats, Q,
                                                   S, d / 8, rows, (uint4*)out, err_word);
  else
    splice_rows_kernel<false><<<grid, 256, sm, s>>>(input_ids, (const uint4*)table, vocab, descs_dev, n_desc, trans_ids, audio_feats, Q,
                                                    S, d / 8, rows, (uint4*)out, err_word);
  return check_launch("splice kernels");
}

extern "C" int desta_splice_bwd(const void* grad_out, const desta_splice_desc* descs_dev, int n_desc, int Q, int S, int d,
                                float* dfeats, void* stream) {
  if (!grad_out || !descs_dev || !dfeats) { set_error("desta_splice_bwd: null pointer"); return DESTA_ERR_ARG; }
  if (n_desc <= 0) return DESTA_OK;
  if (d % 8) { set_error("desta_splice_bwd: hidden size must be a multiple of 8"); return DESTA_ERR_ARG; }
  splice_bwd_kernel<<<dim3(Q < 64 ? Q : 64, n_desc), 256, 0, (cudaStream_t)stream>>>((const uint4*)grad_out, descs_dev, Q, S, d / 8,
                                                                                   (float4*)dfeats);
  return check_launch("splice_bwd_kernel");
}
