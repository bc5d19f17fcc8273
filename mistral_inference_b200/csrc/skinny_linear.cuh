// Weight-streaming linear for T <= 4 tokens (decode):  y[t, n] = sum_k x[t, k] * W[n, k]
//
// Roofline: HBM.  Algorithmic bytes = N*K*2 (every weight read exactly once); the activations
// (T*K*2 bytes) are re-read by every CTA from L2.  Each warp owns 2 adjacent weight rows (one RoPE pair /
// one gate-up pair) and streams them with 16-byte no-allocate loads, 4 in flight per row per lane; x lives in
// shared memory (optionally RMS-normalised in place by every CTA -- 8..28 KB, cheaper than a kernel boundary).
#pragma once
#include "epilogue.cuh"

namespace mb200 {

constexpr int kSkinnyThreads = 256;
constexpr int kSkinnyWarps = kSkinnyThreads / 32;
constexpr int kSkinnyRowsPerWarp = 2;
constexpr int kSkinnyRowsPerCta = kSkinnyWarps * kSkinnyRowsPerWarp;

struct SkinnyParams {
  const void* x;       // [T, K] bf16
  const void* norm_w;  // [K] bf16 (NORM only)
  const void* w;       // [N, K] bf16
  int N, K;
  float eps;
  EpiParams epi;
};

template <int T, int MODE, bool NORM>
__global__ void __launch_bounds__(kSkinnyThreads) skinny_linear_kernel(const SkinnyParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint4* xs = reinterpret_cast<uint4*>(smem_raw);  // [T][K/8] 16-byte chunks
  __shared__ float red[T][kSkinnyWarps];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kc = p.K >> 3;  // chunks per row

  // ---- stage x (and normalise) ----
  const uint4* xg = reinterpret_cast<const uint4*>(p.x);
  for (int i = tid; i < T * kc; i += kSkinnyThreads) xs[i] = xg[i];
  if constexpr (NORM) {
    __syncthreads();
    float ss[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      ss[t] = 0.f;
      for (int c = tid; c < kc; c += kSkinnyThreads) {
        const uint4 v = xs[t * kc + c];
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = bf16lo(u[j]), b = bf16hi(u[j]);
          ss[t] = fmaf(a, a, ss[t]);
          ss[t] = fmaf(b, b, ss[t]);
        }
      }
      ss[t] = warp_sum(ss[t]);
      if (lane == 0) red[t][warp] = ss[t];
    }
    __syncthreads();
    const uint4* wn = reinterpret_cast<const uint4*>(p.norm_w);
#pragma unroll
    for (int t = 0; t < T; ++t) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < kSkinnyWarps; ++w) tot += red[t][w];
      const float r = ref_rsqrt(tot / (float)p.K + p.eps);
      for (int c = tid; c < kc; c += kSkinnyThreads) {
        const uint4 v = xs[t * kc + c];
        const uint4 g = wn[c];
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
        const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // bf16( bf16(x * r) * w )
          const float a = round_bf16(bf16lo(u[j]) * r) * bf16lo(gw[j]);
          const float b = round_bf16(bf16hi(u[j]) * r) * bf16hi(gw[j]);
          o[j] = pack_bf16x2(a, b);
        }
        xs[t * kc + c] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  __syncthreads();

  // ---- stream the two rows of this warp ----
  const int n0 = (blockIdx.x * kSkinnyWarps + warp) * kSkinnyRowsPerWarp;
  if (n0 >= p.N) return;
  const uint4* w0 = reinterpret_cast<const uint4*>(p.w) + (int64_t)n0 * kc;
  const uint4* w1 = w0 + kc;

  float acc[2][T];
#pragma unroll
  for (int t = 0; t < T; ++t) acc[0][t] = acc[1][t] = 0.f;

  constexpr int U = 4;  // 16-byte loads in flight per row per lane
  int c = lane;
  for (; c + (U - 1) * 32 < kc; c += U * 32) {
    uint4 a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      a[u] = ldg_stream16(w0 + c + u * 32);
      b[u] = ldg_stream16(w1 + c + u * 32);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t aw[4] = {a[u].x, a[u].y, a[u].z, a[u].w};
      const uint32_t bw[4] = {b[u].x, b[u].y, b[u].z, b[u].w};
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const uint4 xv = xs[t * kc + c + u * 32];
        const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float xl = bf16lo(xw[j]), xh = bf16hi(xw[j]);
          acc[0][t] = fmaf(bf16lo(aw[j]), xl, acc[0][t]);
          acc[0][t] = fmaf(bf16hi(aw[j]), xh, acc[0][t]);
          acc[1][t] = fmaf(bf16lo(bw[j]), xl, acc[1][t]);
          acc[1][t] = fmaf(bf16hi(bw[j]), xh, acc[1][t]);
        }
      }
    }
  }
  for (; c < kc; c += 32) {  // tail (K/8 not a multiple of 128)
    const uint4 a = ldg_stream16(w0 + c), b = ldg_stream16(w1 + c);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
    const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const uint4 xv = xs[t * kc + c];
      const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xl = bf16lo(xw[j]), xh = bf16hi(xw[j]);
        acc[0][t] = fmaf(bf16lo(aw[j]), xl, acc[0][t]);
        acc[0][t] = fmaf(bf16hi(aw[j]), xh, acc[0][t]);
        acc[1][t] = fmaf(bf16lo(bw[j]), xl, acc[1][t]);
        acc[1][t] = fmaf(bf16hi(bw[j]), xh, acc[1][t]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < T; ++t) {
    acc[0][t] = warp_sum(acc[0][t]);
    acc[1][t] = warp_sum(acc[1][t]);
  }
#pragma unroll
  for (int t = 0; t < T; ++t)
    if (lane == t) epi_pair<MODE>(p.epi, t, n0, acc[0][t], acc[1][t]);
}

template <int MODE, bool NORM>
int launch_skinny(const SkinnyParams& p, int T, cudaStream_t stream) {
  MB_CHECK_ARG(T >= 1 && T <= MB200_SKINNY_MAX_T, "skinny linear: T=%d out of range", T);
  MB_CHECK_ARG(p.K % 8 == 0 && p.N % kSkinnyRowsPerWarp == 0, "skinny linear: K=%d must be a multiple of 8, N=%d even", p.K, p.N);
  const size_t smem = (size_t)T * p.K * 2;
  MB_CHECK_ARG(smem <= 200 * 1024, "skinny linear: T*K too large for shared memory (%zu B)", smem);
  const dim3 grid(ceil_div(p.N, kSkinnyRowsPerCta));
  auto go = [&](auto kernel) -> int {
    if (smem > 48 * 1024) MB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, kSkinnyThreads, smem, stream>>>(p);
    MB_CHECK_LAUNCH("skinny_linear_kernel");
    return MB200_OK;
  };
  switch (T) {
    case 1: return go(skinny_linear_kernel<1, MODE, NORM>);
    case 2: return go(skinny_linear_kernel<2, MODE, NORM>);
    case 3: return go(skinny_linear_kernel<3, MODE, NORM>);
    default: return go(skinny_linear_kernel<4, MODE, NORM>);
  }
}

}  // namespace mb200
