// GQA decode attention over the rotating KV cache: one query token per sequence.
//
// Roofline: HBM.  Algorithmic bytes per (sequence, layer) = 2 (K,V) * kv_len * KV * hd * 2 B; q/out are noise.
// One CTA = (split s, kv head g, sequence b): it owns ring slots [s*C, (s+1)*C) of that head, C = ceil(kv_len/S),
// and serves all H/KV query heads of the group from the same K/V bytes, so nothing is repeated or
// materialised (the reference's repeat_kv, transformer_layers.py:84, multiplies the traffic by H/KV).
// Softmax over the ring is order-free (RoPE was applied with absolute positions before caching), so slots are
// consumed in slot order exactly like the reference's padded-keys mask (cache.py:250-254); slots >= kv_len are
// uninitialised memory (cache.py:166) and are never read.
// Inside the CTA: half a warp per key (16 lanes x 16 B = one 256-B row), online softmax per half-warp in fp32,
// merged across the CTA through shared memory; splits are merged by the last CTA to arrive per (b, g).
#pragma once
#include "common.cuh"

namespace mb200 {

constexpr int AD_THREADS = 128;
constexpr int AD_WARPS = AD_THREADS / 32;
constexpr int AD_MAX_REP = 8;  // query heads per kv head (4 for 7B/Nemo/8x7B, 6 for 8x22B)
constexpr float kLog2e = 1.4426950408889634f;

struct AttnDecodeParams {
  const bf16* q;        // [B, H*hd]
  const bf16* cache_k;  // [max_batch, W, KV, hd]
  const bf16* cache_v;
  const int32_t* kv_len;  // [B]
  bf16* out;              // [B, H*hd]
  float* partial;         // [B, KV, S, REP, hd + 2] (m, l, acc) when S > 1
  int* counters;          // [B, KV] zero-initialised once; self-resetting
  int B, W, H, KV, S;
  float scale;
};

// One online-softmax state per (half-warp, query head): running max m, denominator l, 8 output dims per lane.
template <int REP>
__global__ void __launch_bounds__(AD_THREADS) attn_decode_kernel(const AttnDecodeParams p) {
  const int s = blockIdx.x, g = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int half = lane >> 4, hl = lane & 15;  // which key of the pair, 16-byte column inside the row
  const int len = p.kv_len[b];
  const int C = (len + p.S - 1) / p.S;
  const int k_begin = min(s * C, len), k_end = min(k_begin + C, len);

  // q for the REP heads of this group, this lane's 8 dims, pre-scaled into the exp2 domain is NOT done: the
  // reference scales the fp32 scores (q k^T * hd^-0.5), so do exactly that.
  float qf[REP][8];
#pragma unroll
  for (int r = 0; r < REP; ++r) {
    const uint4 v = *reinterpret_cast<const uint4*>(p.q + ((int64_t)b * p.H + g * REP + r) * kHeadDim + hl * 8);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      qf[r][2 * j] = bf16lo(u[j]);
      qf[r][2 * j + 1] = bf16hi(u[j]);
    }
  }
  constexpr float kMasked = -1.0e30f;  // finite "minus infinity" keeps the update branch-free (see the loop below)
  float m[REP], l[REP], acc[REP][8];
#pragma unroll
  for (int r = 0; r < REP; ++r) {
    m[r] = kMasked;
    l[r] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[r][j] = 0.f;
  }

  const int64_t row_stride = (int64_t)p.KV * kHeadDim;  // elements between consecutive slots
  const bf16* kbase = p.cache_k + ((int64_t)b * p.W) * row_stride + (int64_t)g * kHeadDim + hl * 8;
  const bf16* vbase = p.cache_v + ((int64_t)b * p.W) * row_stride + (int64_t)g * kHeadDim + hl * 8;

  // keys of this CTA are dealt to half-warps round-robin: slot = k_begin + it * 8 + warp * 2 + half
  constexpr int STRIDE = AD_WARPS * 2;
  int slot = k_begin + warp * 2 + half;
  uint4 kq = make_uint4(0, 0, 0, 0), vq = kq;
  if (slot < k_end) {
    kq = ldg_stream16(kbase + slot * row_stride);
    vq = ldg_stream16(vbase + slot * row_stride);
  }
  // NOTE: all 32 lanes stay in the loop while either half has work (shuffles are warp-wide).
  const int iters = (k_end - k_begin + STRIDE - 1) / STRIDE;
  for (int it = 0; it < iters; ++it) {
    const bool valid = slot < k_end;
    const uint4 kc = kq, vc = vq;
    const int nslot = slot + STRIDE;
    if (nslot < k_end) {  // prefetch the next pair while this one is reduced
      kq = ldg_stream16(kbase + nslot * row_stride);
      vq = ldg_stream16(vbase + nslot * row_stride);
    }
    const uint32_t ku[4] = {kc.x, kc.y, kc.z, kc.w};
    const uint32_t vu[4] = {vc.x, vc.y, vc.z, vc.w};
    float kf[8], vf[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      kf[2 * j] = bf16lo(ku[j]);
      kf[2 * j + 1] = bf16hi(ku[j]);
      vf[2 * j] = bf16lo(vu[j]);
      vf[2 * j + 1] = bf16hi(vu[j]);
    }
#pragma unroll
    for (int r = 0; r < REP; ++r) {
      float d = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) d = fmaf(qf[r][j], kf[j], d);
      // reduce over the 16 lanes of the half-warp
      d += __shfl_xor_sync(0xffffffffu, d, 8);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      // branch-free: a branch here fences each head's FFMA -> shuffle -> exp chain into its own reconvergence region
      // and serialises the heads (measured 1500 cycles per key pair); as selects the chains interleave
      const float sc = valid ? d * p.scale : kMasked;
      const float mn = fmaxf(m[r], sc);
      const float corr = exp2f((m[r] - mn) * kLog2e);
      const float pe = exp2f((sc - mn) * kLog2e);
      const float pr = valid ? pe : 0.f;
      l[r] = l[r] * corr + pr;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[r][j] = fmaf(pr, vf[j], acc[r][j] * corr);
      m[r] = mn;
    }
    slot = nslot;
  }

  // ---- merge the two half-warps, then the warps of the CTA ----
  __shared__ float sm_m[AD_WARPS][REP], sm_l[AD_WARPS][REP];
  __shared__ float sm_acc[AD_WARPS][REP][kHeadDim];
#pragma unroll
  for (int r = 0; r < REP; ++r) {
    const float mo = __shfl_xor_sync(0xffffffffu, m[r], 16);
    const float lo = __shfl_xor_sync(0xffffffffu, l[r], 16);
    const float mn = fmaxf(m[r], mo);
    const float cs = exp2f((m[r] - mn) * kLog2e);  // both empty: 1 * (l = 0)
    const float co = exp2f((mo - mn) * kLog2e);
    l[r] = l[r] * cs + lo * co;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float ao = __shfl_xor_sync(0xffffffffu, acc[r][j], 16);
      acc[r][j] = acc[r][j] * cs + ao * co;
    }
    m[r] = mn;
    if (half == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sm_acc[warp][r][hl * 8 + j] = acc[r][j];
      if (hl == 0) {
        sm_m[warp][r] = m[r];
        sm_l[warp][r] = l[r];
      }
    }
  }
  __syncthreads();

  // thread d (0..127) finishes dim d of every head of the group
  const int d = tid;
  float fm[REP], fl[REP], fa[REP];
#pragma unroll
  for (int r = 0; r < REP; ++r) {
    float mn = -INFINITY;
#pragma unroll
    for (int w = 0; w < AD_WARPS; ++w) mn = fmaxf(mn, sm_m[w][r]);
    float lt = 0.f, at = 0.f;
#pragma unroll
    for (int w = 0; w < AD_WARPS; ++w) {
      const float c = (sm_m[w][r] == -INFINITY) ? 0.f : exp2f((sm_m[w][r] - mn) * kLog2e);
      lt += sm_l[w][r] * c;
      at += sm_acc[w][r][d] * c;
    }
    fm[r] = mn;
    fl[r] = lt;
    fa[r] = at;
  }

  if (p.S == 1) {
#pragma unroll
    for (int r = 0; r < REP; ++r) p.out[((int64_t)b * p.H + g * REP + r) * kHeadDim + d] = __float2bfloat16_rn(fa[r] / fl[r]);
    return;
  }

  // ---- publish the partial; the last split of this (b, g) to arrive combines all of them ----
  const int PSTRIDE = kHeadDim + 2;
  float* mine = p.partial + ((((int64_t)b * p.KV + g) * p.S + s) * REP) * PSTRIDE;
#pragma unroll
  for (int r = 0; r < REP; ++r) {
    mine[r * PSTRIDE + 2 + d] = fa[r];
    if (d == 0) {
      mine[r * PSTRIDE + 0] = fm[r];
      mine[r * PSTRIDE + 1] = fl[r];
    }
  }
  __threadfence();
  __syncthreads();
  __shared__ int is_last;
  if (tid == 0) {
    const int prev = atomicAdd(&p.counters[b * p.KV + g], 1);
    is_last = (prev == p.S - 1);
    if (is_last) p.counters[b * p.KV + g] = 0;  // self-reset for the next launch (stream-ordered)
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // combine: every load independent (see decode_megakernel.cuh): (m, l) of all splits -> shared memory, then one
  // (head, dim) output per thread with the split loop unrolled
  const float* all = p.partial + (((int64_t)b * p.KV + g) * p.S) * REP * PSTRIDE;
  __shared__ float cm[64 * REP], cl[64 * REP];
  for (int i = tid; i < p.S * REP; i += AD_THREADS) {
    cm[i] = __ldcg(all + (int64_t)i * PSTRIDE);
    cl[i] = __ldcg(all + (int64_t)i * PSTRIDE + 1);
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < REP; ++r) {
    float mn = -INFINITY;
    for (int t = 0; t < p.S; ++t) mn = fmaxf(mn, cm[t * REP + r]);
    float lt = 0.f, at = 0.f;
#pragma unroll 8
    for (int t = 0; t < p.S; ++t) {
      const float mt = cm[t * REP + r];
      const float c = (mt == -INFINITY) ? 0.f : exp2f((mt - mn) * kLog2e);
      lt += cl[t * REP + r] * c;
      at += __ldcg(all + ((int64_t)t * REP + r) * PSTRIDE + 2 + d) * c;
    }
    p.out[((int64_t)b * p.H + g * REP + r) * kHeadDim + d] = __float2bfloat16_rn(at / lt);
  }
}

inline size_t attn_decode_workspace(int64_t B, int64_t KV, int64_t S, int64_t rep) {
  // partials + counters (counters live at the end, 256-B aligned)
  size_t part = (size_t)B * KV * S * rep * (kHeadDim + 2) * sizeof(float);
  part = (part + 255) & ~(size_t)255;
  return part + (((size_t)B * KV * sizeof(int)) + 255 & ~(size_t)255);
}

}  // namespace mb200
