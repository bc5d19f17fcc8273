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
};

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

}  // namespace mb200
