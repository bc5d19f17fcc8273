// Pair epilogues shared by the weight-streaming GEMV (skinny_linear.cuh) and the tensor-core GEMMs.
// Every linear on the hot path hands the epilogue two adjacent fp32 accumulators (columns n, n+1 of
// token t).  That is the natural unit: RoPE rotates the interleaved pair (x[2i], x[2i+1]) (rope.py:18-23)
// and the packed gate/up weight puts w1[i], w3[i] on rows 2i, 2i+1.
#pragma once
#include "common.cuh"

namespace mb200 {

enum EpiMode : int {
  EPI_STORE = 0,     // out[t, n] = bf16(acc)
  EPI_RESIDUAL = 1,  // out[t, n] = bf16( bf16(acc) + residual[t, n] )           (transformer_layers.py:166,168)
  EPI_F32 = 2,       // logits[t, n] = float(bf16(acc))                           (transformer.py:235,240)
  EPI_SWIGLU = 3,    // g[t, n/2] = bf16( bf16(silu(bf16(acc0))) * bf16(acc1) )   (transformer_layers.py:106)
  EPI_QKV_ROPE = 4,  // split into q/k/v, rotate q,k pairs, optional ring scatter  (transformer_layers.py:66-70, cache.py:91-92)
  EPI_MOE_SCALE = 5, // yw[row, n] = bf16( w[row] * bf16(acc) ), stored on every rank of an expert-parallel group  (moe.py:31)
};

constexpr int kMaxPeers = 8;

struct EpiParams {
  void* out = nullptr;             // bf16 [T, ld_out]
  const void* residual = nullptr;  // bf16 [T, ld_out]
  float* out_f32 = nullptr;        // fp32 [T, ld_out]
  int64_t ld_out = 0;
  // EPI_QKV_ROPE
  void* q_out = nullptr;  // [T, q_dim]
  void* k_out = nullptr;  // [T, kv_dim]
  void* v_out = nullptr;  // [T, kv_dim]
  void* cache_k = nullptr;  // [rows, kv_dim]
  void* cache_v = nullptr;
  const int32_t* positions = nullptr;   // [T]
  const int32_t* cache_rows = nullptr;  // [T] or null
  const float* rope = nullptr;          // [n_pos, 64, 2]
  int q_dim = 0, kv_dim = 0;
  // EPI_MOE_SCALE: routing weight of each (token, expert) row; the weighted expert output goes to `out` and to the same offset
  // of the mapped buffers of the other ranks (NVLink peer stores; n_peers == 0 when unsharded)
  const void* row_w = nullptr;  // bf16 [rows]
  void* peer_out[kMaxPeers] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int n_peers = 0;
};

template <int MODE>
__device__ __forceinline__ void epi_pair(const EpiParams& p, int t, int n, float acc0, float acc1) {
  // the Linear's own output rounding (bf16 result of nn.Linear)
  const float y0 = round_bf16(acc0), y1 = round_bf16(acc1);
  if constexpr (MODE == EPI_STORE) {
    *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.out) + (int64_t)t * p.ld_out + n) = pack_bf16x2(y0, y1);
  } else if constexpr (MODE == EPI_RESIDUAL) {
    const uint32_t r = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint16_t*>(p.residual) + (int64_t)t * p.ld_out + n);
    *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.out) + (int64_t)t * p.ld_out + n) =
        pack_bf16x2(y0 + bf16lo(r), y1 + bf16hi(r));
  } else if constexpr (MODE == EPI_F32) {
    *reinterpret_cast<float2*>(p.out_f32 + (int64_t)t * p.ld_out + n) = make_float2(y0, y1);
  } else if constexpr (MODE == EPI_SWIGLU) {
    const float s = round_bf16(ref_silu(y0));
    reinterpret_cast<bf16*>(p.out)[(int64_t)t * p.ld_out + (n >> 1)] = __float2bfloat16_rn(s * y1);
  } else if constexpr (MODE == EPI_MOE_SCALE) {
    const float w = bf16_to_float(reinterpret_cast<const uint16_t*>(p.row_w)[t]);
    const uint32_t packed = pack_bf16x2(w * y0, w * y1);
    const int64_t off = (int64_t)t * p.ld_out + n;
    *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.out) + off) = packed;
    for (int r = 0; r < p.n_peers; ++r) *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.peer_out[r]) + off) = packed;
  } else if constexpr (MODE == EPI_QKV_ROPE) {
    if (n < p.q_dim + p.kv_dim) {  // q or k: rotate
      const int pos = p.positions[t];
      const int i = (n & (kHeadDim - 1)) >> 1;
      const float2 cs = *reinterpret_cast<const float2*>(p.rope + ((int64_t)pos * (kHeadDim / 2) + i) * 2);
      float re, im;
      ref_cmul(y0, y1, cs.x, cs.y, re, im);
      const uint32_t packed = pack_bf16x2(re, im);
      if (n < p.q_dim) {
        *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.q_out) + (int64_t)t * p.q_dim + n) = packed;
      } else {
        const int c = n - p.q_dim;
        *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.k_out) + (int64_t)t * p.kv_dim + c) = packed;
        if (p.cache_rows != nullptr) {
          const int row = p.cache_rows[t];
          if (row >= 0) *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.cache_k) + (int64_t)row * p.kv_dim + c) = packed;
        }
      }
    } else {  // v: stored as projected
      const int c = n - p.q_dim - p.kv_dim;
      const uint32_t packed = pack_bf16x2(y0, y1);
      *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.v_out) + (int64_t)t * p.kv_dim + c) = packed;
      if (p.cache_rows != nullptr) {
        const int row = p.cache_rows[t];
        if (row >= 0) *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.cache_v) + (int64_t)row * p.kv_dim + c) = packed;
      }
    }
  }
}

// ---- row-chunk epilogues for the tcgen05 GEMM -------------------------------------------------------------------------------
// There one thread owns one output row and reads its accumulator out of TMEM 32 columns at a time, so the natural unit is
// (row t, columns n .. n+31) with n a multiple of 32: the chunk never straddles a head or the q|k|v boundaries, every store is
// a 16-byte vector, and the per-row metadata (position, ring row) is read once per chunk instead of once per pair.  The
// arithmetic is the same as epi_pair's, except that SiLU uses the hardware exp2/reciprocal (relative error ~1e-6, i.e. a
// one-ulp bf16 flip in ~0.03 % of the gate values -- far below what the accumulation order already does).
__device__ __forceinline__ uint32_t pack2_rn(float lo, float hi) {
  const __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);  // one F2FP: .x (low half) = lo
  return *reinterpret_cast<const uint32_t*>(&b);
}
__device__ __forceinline__ float fast_silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// What epi_chunk32 will READ for row t, columns n0 .. n0 + 127, requested ahead of time (the thread is about to wait for its
// accumulator): the residual row segment, or the row's position and its RoPE table lines.  Prefetches only -- no registers held.
__device__ __forceinline__ void prefetch_l1(const void* ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }
template <int MODE>
__device__ __forceinline__ void epi_prefetch128(const EpiParams& p, int t, int n0) {
  if constexpr (MODE == EPI_RESIDUAL) {
    const uint8_t* src = reinterpret_cast<const uint8_t*>(reinterpret_cast<const uint16_t*>(p.residual) + (int64_t)t * p.ld_out + n0);
    prefetch_l1(src);
    prefetch_l1(src + 128);
  } else if constexpr (MODE == EPI_QKV_ROPE) {
    if (n0 < p.q_dim + p.kv_dim) {
      const int pos = p.positions[t];
      const uint8_t* cs = reinterpret_cast<const uint8_t*>(p.rope + (int64_t)pos * kHeadDim);  // [n_pos, 64, 2] floats: 512 B per position
#pragma unroll
      for (int i = 0; i < 4; ++i) prefetch_l1(cs + 128 * i);
    }
    if (n0 >= p.q_dim && p.cache_rows != nullptr) prefetch_l1(p.cache_rows + t);
  }
}

template <int MODE>
__device__ __forceinline__ void epi_chunk32(const EpiParams& p, int t, int n, const uint32_t (&v)[32]) {
  if constexpr (MODE == EPI_STORE || MODE == EPI_RESIDUAL) {
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + (int64_t)t * p.ld_out + n);
    uint32_t o[16];
    if constexpr (MODE == EPI_RESIDUAL) {
      const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.residual) + (int64_t)t * p.ld_out + n);
      uint4 r4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) r4[q] = src[q];
      const uint32_t* r = reinterpret_cast<const uint32_t*>(r4);
#pragma unroll
      for (int j = 0; j < 16; ++j)
        o[j] = pack2_rn(round_bf16(__uint_as_float(v[2 * j])) + bf16lo(r[j]), round_bf16(__uint_as_float(v[2 * j + 1])) + bf16hi(r[j]));
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = pack2_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
  } else if constexpr (MODE == EPI_MOE_SCALE) {
    const float w = bf16_to_float(reinterpret_cast<const uint16_t*>(p.row_w)[t]);
    uint32_t o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = pack2_rn(w * round_bf16(__uint_as_float(v[2 * j])), w * round_bf16(__uint_as_float(v[2 * j + 1])));
    const int64_t off = (int64_t)t * p.ld_out + n;
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + off);
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
    for (int r = 0; r < p.n_peers; ++r) {  // the same row on the other ranks (NVLink peer stores, 64 B per thread and chunk)
      uint4* pd = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.peer_out[r]) + off);
#pragma unroll
      for (int q = 0; q < 4; ++q) pd[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
    }
  } else if constexpr (MODE == EPI_F32) {
    float4* dst = reinterpret_cast<float4*>(p.out_f32 + (int64_t)t * p.ld_out + n);
#pragma unroll
    for (int q = 0; q < 8; ++q)
      dst[q] = make_float4(round_bf16(__uint_as_float(v[4 * q])), round_bf16(__uint_as_float(v[4 * q + 1])),
                           round_bf16(__uint_as_float(v[4 * q + 2])), round_bf16(__uint_as_float(v[4 * q + 3])));
  } else if constexpr (MODE == EPI_SWIGLU) {
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + (int64_t)t * p.ld_out + (n >> 1));
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a0 = round_bf16(__uint_as_float(v[4 * j])), b0 = round_bf16(__uint_as_float(v[4 * j + 1]));
      const float a1 = round_bf16(__uint_as_float(v[4 * j + 2])), b1 = round_bf16(__uint_as_float(v[4 * j + 3]));
      o[j] = pack2_rn(round_bf16(fast_silu(a0)) * b0, round_bf16(fast_silu(a1)) * b1);
    }
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
  } else if constexpr (MODE == EPI_QKV_ROPE) {
    uint32_t o[16];
    uint16_t* dst;
    uint16_t* ring = nullptr;
    if (n < p.q_dim + p.kv_dim) {  // q or k: rotate the 16 interleaved pairs of this chunk
      const int pos = p.positions[t];
      const float4* cs4 = reinterpret_cast<const float4*>(p.rope + ((int64_t)pos * (kHeadDim / 2) + ((n & (kHeadDim - 1)) >> 1)) * 2);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 cs = __ldg(cs4 + j);  // (cos, sin) of pairs 2j, 2j+1
        float re, im;
        ref_cmul(round_bf16(__uint_as_float(v[4 * j])), round_bf16(__uint_as_float(v[4 * j + 1])), cs.x, cs.y, re, im);
        o[2 * j] = pack2_rn(re, im);
        ref_cmul(round_bf16(__uint_as_float(v[4 * j + 2])), round_bf16(__uint_as_float(v[4 * j + 3])), cs.z, cs.w, re, im);
        o[2 * j + 1] = pack2_rn(re, im);
      }
      if (n < p.q_dim) {
        dst = reinterpret_cast<uint16_t*>(p.q_out) + (int64_t)t * p.q_dim + n;
      } else {
        dst = reinterpret_cast<uint16_t*>(p.k_out) + (int64_t)t * p.kv_dim + (n - p.q_dim);
        if (p.cache_rows != nullptr) {
          const int row = p.cache_rows[t];
          if (row >= 0) ring = reinterpret_cast<uint16_t*>(p.cache_k) + (int64_t)row * p.kv_dim + (n - p.q_dim);
        }
      }
    } else {  // v: stored as projected
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = pack2_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
      const int c = n - p.q_dim - p.kv_dim;
      dst = reinterpret_cast<uint16_t*>(p.v_out) + (int64_t)t * p.kv_dim + c;
      if (p.cache_rows != nullptr) {
        const int row = p.cache_rows[t];
        if (row >= 0) ring = reinterpret_cast<uint16_t*>(p.cache_v) + (int64_t)row * p.kv_dim + c;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) reinterpret_cast<uint4*>(dst)[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
    if (ring != nullptr) {
#pragma unroll
      for (int q = 0; q < 4; ++q) reinterpret_cast<uint4*>(ring)[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
    }
  }
}

}  // namespace mb200
