// Row-wise RMSNorm (prefill path; the decode path fuses it into the consuming GEMV) and the KV ring scatter.
#pragma once
#include "common.cuh"

namespace mb200 {

// out = bf16( bf16( x * rsqrt(mean(x^2) + eps) ) * w ), one CTA per token (transformer_layers.py:115-120)
__global__ void __launch_bounds__(256) rmsnorm_kernel(const uint4* __restrict__ x, const uint4* __restrict__ w, uint4* __restrict__ out,
                                                      int dim, float eps) {
  pdl_trigger();
  pdl_wait();  // x is the previous kernel's output
  const int kc = dim >> 3;
  const uint4* xr = x + (int64_t)blockIdx.x * kc;
  uint4* orow = out + (int64_t)blockIdx.x * kc;
  __shared__ float red[8];
  float ss = 0.f;
  for (int c = threadIdx.x; c < kc; c += 256) {
    const uint4 v = xr[c];
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = bf16lo(u[j]), b = bf16hi(u[j]);
      ss = fmaf(a, a, ss);
      ss = fmaf(b, b, ss);
    }
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float r = ref_rsqrt(tot / (float)dim + eps);
  for (int c = threadIdx.x; c < kc; c += 256) {
    const uint4 v = xr[c], g = w[c];
    const uint32_t u[4] = {v.x, v.y, v.z, v.w}, gw[4] = {g.x, g.y, g.z, g.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[j] = pack_bf16x2(round_bf16(bf16lo(u[j]) * r) * bf16lo(gw[j]), round_bf16(bf16hi(u[j]) * r) * bf16hi(gw[j]));
    orow[c] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// cache[rows[t]] = src[t] for rows[t] >= 0 (cache.py:91-92); one warp per (token, K|V) row of KV*hd bf16
__global__ void kv_ring_write_kernel(const uint4* __restrict__ k_new, const uint4* __restrict__ v_new, uint4* __restrict__ cache_k,
                                     uint4* __restrict__ cache_v, const int32_t* __restrict__ rows, int T, int row_chunks) {
  const int t = blockIdx.x;
  const int row = rows[t];
  if (row < 0) return;
  for (int c = threadIdx.x; c < row_chunks; c += blockDim.x) {
    cache_k[(int64_t)row * row_chunks + c] = k_new[(int64_t)t * row_chunks + c];
    cache_v[(int64_t)row * row_chunks + c] = v_new[(int64_t)t * row_chunks + c];
  }
}

// Device-side step state of the batched decode loop (replaces cache.py:197-263 for one-token steps).  Reads the
// per-sequence positions, writes the metadata block every layer's kernels read -- same layout as the host-built block of
// mistral_inference_b200/cache.py::build_metadata_host for seqlens = [1] * B:
//   positions[B] | q_start[B + 1] | seqpos[B] | per distinct window W: cache_rows[B], kv_len[B]
// -- and advances the positions, so a captured CUDA graph replays correct metadata without any host write.
constexpr int kMaxWindows = 8;
struct DecodeMetaParams {
  int32_t* seqpos;  // [B] tokens cached so far (in/out: incremented)
  int32_t* meta;    // [3B + 1 + n_w * 2B]
  int B, n_w;
  int windows[kMaxWindows];
};
__global__ void decode_meta_kernel(const DecodeMetaParams p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > p.B) return;
  int32_t* q_start = p.meta + p.B;
  if (b == p.B) {
    q_start[b] = p.B;
    return;
  }
  const int pos = p.seqpos[b];
  p.meta[b] = pos;
  q_start[b] = b;
  p.meta[2 * p.B + 1 + b] = pos;
  int32_t* o = p.meta + 3 * p.B + 1;
  for (int j = 0; j < p.n_w; ++j) {
    const int W = p.windows[j];
    o[b] = pos % W + b * W;             // cache.py:235
    o[p.B + b] = min(pos + 1, W);       // cache.py:250-254: kv_seqlen = (seqpos + 1).clamp(max = W)
    o += 2 * p.B;
  }
  p.seqpos[b] = pos + 1;
}

}  // namespace mb200
