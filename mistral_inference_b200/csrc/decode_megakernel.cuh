// Persistent decode step: ONE cooperative kernel per generated token (batch 1).
//
// Why: a batch-1 decode step streams ~14.2 GB of weights through ~160 dependent matrix-vector products of
// 5-36 us each.  As separate kernels every boundary drains the memory pipe (launch + ramp-up + tail is of the
// same order as the kernels themselves), which is what caps the per-op path near 50 % of the HBM roofline.
// Here one CTA per SM lives for the whole token:
//   * a PRODUCER warp walks the CTA's static weight schedule for the whole model (every layer's QKV / wo /
//     gate-up / down slice and the lm-head slice are contiguous row ranges known up front) and streams it with
//     cp.async.bulk (TMA bulk copies, completion on mbarriers) through a ~176 KB shared-memory ring.  Weights do
//     not depend on activations, so the producer never waits for a phase boundary: while the consumers sit in a
//     grid barrier the ring keeps filling, and HBM stays busy across all ~160 dependencies.
//   * 8 CONSUMER warps do the math out of shared memory (fp32 FMA on bf16 pairs; batch 1 is ~0.1 flop/byte, far
//     below the CUDA-core roof, tensor cores would not help), with the same fused prologues/epilogues as the
//     per-op kernels (RMSNorm, RoPE + ring scatter, SiLU*mul, residual adds) and the same rounding points as the
//     reference.
//   * phases are separated by a self-resetting sense-reversing grid barrier (5 per layer).
// Row pairs (2 rows = one RoPE pair / one gate-up pair) are dealt to CTAs as contiguous ranges:
// CTA c owns pairs [c*P/G, (c+1)*P/G) of each matrix, so its slice of every weight matrix is one contiguous byte
// range and load imbalance is at most one pair.
// Attention (phase 2) is flash-decoding over the ring with (kv head, split) work items, read directly from
// global memory by the consumers while the producer is already prefetching the wo slice.
#pragma once
#include "attn_decode.cuh"
#include "common.cuh"

namespace mb200 {

constexpr int MK_CONSUMER_WARPS = 8;
constexpr int MK_CONSUMERS = MK_CONSUMER_WARPS * 32;
constexpr int MK_THREADS = MK_CONSUMERS + 32;  // + producer warp
constexpr int MK_STAGE_BYTES = 16 * 1024;
constexpr int MK_MAX_KC = MK_STAGE_BYTES / 4;  // elements per row chunk (2 rows x KC x 2 B per stage)
constexpr int MK_MAX_STAGES = 12;
constexpr int MK_MAX_SPLITS = 32;

struct MkLayer {  // 64 bytes, device array prepared by the caller (include/mistral_b200.h: mb200_layer_desc)
  const bf16* wqkv;
  const bf16* wo;
  const bf16* w13;
  const bf16* w2;
  const bf16* attn_norm;
  const bf16* ffn_norm;
  bf16* cache_k;  // [max_batch, W, KV, hd]
  bf16* cache_v;
};

struct MkParams {
  const MkLayer* layers;
  const int32_t* windows;  // [n_layers] ring size per layer
  int n_layers;
  const bf16* emb;         // [V, dim]
  const bf16* final_norm;  // [dim]
  const bf16* w_out;       // [V, dim]
  const float* rope;       // [n_pos, 64, 2]
  const int64_t* token;    // device scalar: the token to embed
  int pos;                 // absolute position of that token
  int batch_row;           // which row of the cache this sequence uses
  float* logits;           // [V] fp32
  int dim, hidden, H, KV, vocab;
  float eps;
  int n_stages, xs_bytes;
  // scratch (global)
  unsigned* bar_count;  // grid barrier arrival counter (self-resetting)
  unsigned* bar_gen;    // grid barrier generation
  int* attn_counters;   // [KV]
  bf16* xbuf;           // [2][dim] residual stream ping-pong
  bf16* hbuf;           // [dim]
  bf16* qbuf;           // [H*hd]
  bf16* abuf;           // [H*hd] attention output
  bf16* gbuf;           // [hidden]
  float* partial;       // [KV][splits][REP][hd+2]
  unsigned long long* prof;  // optional [n_layers][12] globaltimer stamps written by CTA 0 (debug timeline), or null
};

__device__ __forceinline__ void mk_stamp(const MkParams& p, int tid, int layer, int idx) {
  if (p.prof != nullptr && blockIdx.x == 0 && tid == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    p.prof[layer * 12 + idx] = t;
  }
}

// ---- PTX: mbarrier + bulk copy ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(MK_CONSUMERS) : "memory"); }

__device__ __forceinline__ uint4 ldcg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Sense-reversing grid barrier among the consumer threads of all CTAs (self-resetting: usable across launches).
__device__ __forceinline__ void grid_barrier(const MkParams& p, int tid) {
  consumer_sync();
  if (tid == 0) {
    const unsigned gen = ld_acquire_u32(p.bar_gen);
    __threadfence();
    const unsigned arrived = atomicAdd(p.bar_count, 1u);
    if (arrived == gridDim.x - 1) {
      atomicExch(p.bar_count, 0u);
      __threadfence();
      atomicAdd(p.bar_gen, 1u);
    } else {
      while (ld_acquire_u32(p.bar_gen) == gen) {
      }
    }
    __threadfence();
  }
  consumer_sync();
}

// How a [N, K] matrix is cut for the ring: pairs of rows, K in `nch` chunks of `kc` elements.
struct MatCut {
  int pairs, nch, kc, p0, p1;
};
__device__ __forceinline__ MatCut cut_matrix(int N, int K) {
  MatCut c;
  c.pairs = N >> 1;
  c.nch = (K + MK_MAX_KC - 1) / MK_MAX_KC;
  c.kc = K / c.nch;
  c.p0 = (int)(((long long)blockIdx.x * c.pairs) / gridDim.x);
  c.p1 = (int)(((long long)(blockIdx.x + 1) * c.pairs) / gridDim.x);
  return c;
}

struct RingState {
  uint32_t it;  // running stage counter (same sequence in producer and consumers)
};

// Stage order inside a matrix (same in producer and consumers): the CTA's pairs are taken in GROUPS of up to 8 (one pair
// per consumer warp); inside a group the stages are chunk-major: (pair g0+0, ch 0) ... (pair g0+g-1, ch 0), (pair g0+0, ch 1) ...
// So warp w touches stages base + ch*g + w: never more than 8 apart, i.e. always within one lap of the (>= 9-stage) ring --
// which is what makes parity-based mbarrier waits safe (a waiter two laps ahead would alias and pass early).

// ---- producer: stream this CTA's slice of one matrix --------------------------------------------
__device__ __forceinline__ void produce_matrix(const bf16* W, int N, int K, uint8_t* ring, uint64_t* full, uint64_t* empty, int n_stages,
                                               RingState& rs) {
  const MatCut c = cut_matrix(N, K);
  const uint32_t row_bytes = (uint32_t)c.kc * 2;
  for (int g0 = c.p0; g0 < c.p1; g0 += MK_CONSUMER_WARPS) {
    const int g = min(MK_CONSUMER_WARPS, c.p1 - g0);
    for (int ch = 0; ch < c.nch; ++ch) {
      for (int w = 0; w < g; ++w) {
        const bf16* r0 = W + (int64_t)(2 * (g0 + w)) * K;
        const uint32_t slot = rs.it % n_stages, par = (rs.it / n_stages) & 1;
        mbar_wait(&empty[slot], par ^ 1);
        uint8_t* dst = ring + (size_t)slot * MK_STAGE_BYTES;
        mbar_arrive_expect_tx(&full[slot], 2 * row_bytes);
        if (c.nch == 1) {
          bulk_g2s(dst, r0, 2 * row_bytes, &full[slot]);  // the two rows are contiguous
        } else {
          bulk_g2s(dst, r0 + ch * c.kc, row_bytes, &full[slot]);
          bulk_g2s(dst + row_bytes, r0 + K + ch * c.kc, row_bytes, &full[slot]);
        }
        ++rs.it;
      }
    }
  }
}

// ---- consumers: y[pair] = W[pair rows] . xs, epilogue(pair, acc0, acc1) on one lane ----------------
// Warp-per-pair inside a group: warp w owns pair g0+w and consumes its `nch` stages by itself (32 lanes x 16 B per step,
// unrolled -> plenty of ILP); only the owning warp releases a slot (empty barriers have arrival count 1).  One block
// barrier per GROUP keeps all warps within a group of each other (see the stage-order note above).
template <class Epi>
__device__ __forceinline__ void consume_matrix(int N, int K, const uint8_t* ring, uint64_t* full, uint64_t* empty, int n_stages, RingState& rs,
                                               const uint4* xs, int tid, Epi epi) {
  const MatCut c = cut_matrix(N, K);
  const int lane = tid & 31, warp = tid >> 5;
  const int kc8 = c.kc >> 3;  // 16-byte chunks per row chunk
  for (int g0 = c.p0; g0 < c.p1; g0 += MK_CONSUMER_WARPS) {
    const int g = min(MK_CONSUMER_WARPS, c.p1 - g0);
    if (warp < g) {
      float a0 = 0.f, a1 = 0.f;
      for (int ch = 0; ch < c.nch; ++ch) {
        const uint32_t it = rs.it + (uint32_t)(ch * g + warp);
        const uint32_t slot = it % n_stages, par = (it / n_stages) & 1;
        // Guard (tests/test_megakernel_protocol.py): bulk copies land out of order, so this warp may get here before the
        // slot's PREVIOUS fill (owned by another warp) has landed; `full` would then still be one phase behind and a
        // parity wait would alias and pass early.  Waiting first until that previous fill has been CONSUMED (same
        // condition the producer waits for before refilling) pins `full` to phase {r, r+1} when it is tested.
        mbar_wait(&empty[slot], par ^ 1);
        mbar_wait(&full[slot], par);
        const uint4* w0 = reinterpret_cast<const uint4*>(ring + (size_t)slot * MK_STAGE_BYTES);
        const uint4* w1 = w0 + kc8;
        const uint4* xc = xs + ch * kc8;
#pragma unroll 4
        for (int i = lane; i < kc8; i += 32) {
          const uint4 a = w0[i], b = w1[i], x = xc[i];
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w}, xw[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float xl = bf16lo(xw[j]), xh = bf16hi(xw[j]);
            a0 = fmaf(bf16lo(aw[j]), xl, a0);
            a0 = fmaf(bf16hi(aw[j]), xh, a0);
            a1 = fmaf(bf16lo(bw[j]), xl, a1);
            a1 = fmaf(bf16hi(bw[j]), xh, a1);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);  // this warp is the only reader of the slot
      }
      a0 = warp_sum(a0);
      a1 = warp_sum(a1);
      if (lane == 0) epi(2 * (g0 + warp), a0, a1);
    }
    rs.it += (uint32_t)(g * c.nch);
    consumer_sync();
  }
}

// ---- consumers: stage an activation vector (written by other CTAs: L2 loads) and optionally RMS-normalise it ----
__device__ __forceinline__ void stage_x(uint4* xs, const bf16* src, const bf16* norm_w, int K, float eps, float* red, int tid) {
  const int kc = K >> 3;
#pragma unroll 4
  for (int i = tid; i < kc; i += MK_CONSUMERS) xs[i] = ldcg16(reinterpret_cast<const uint4*>(src) + i);
  consumer_sync();
  if (norm_w == nullptr) return;
  const int lane = tid & 31, warp = tid >> 5;
  float ss = 0.f;
  for (int i = tid; i < kc; i += MK_CONSUMERS) {
    const uint4 v = xs[i];
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = bf16lo(u[j]), b = bf16hi(u[j]);
      ss = fmaf(a, a, ss);
      ss = fmaf(b, b, ss);
    }
  }
  ss = warp_sum(ss);
  if (lane == 0) red[warp] = ss;
  consumer_sync();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < MK_CONSUMER_WARPS; ++w) tot += red[w];
  const float r = ref_rsqrt(tot / (float)K + eps);
  const uint4* wn = reinterpret_cast<const uint4*>(norm_w);
  for (int i = tid; i < kc; i += MK_CONSUMERS) {
    const uint4 v = xs[i], g = wn[i];
    const uint32_t u[4] = {v.x, v.y, v.z, v.w}, gw[4] = {g.x, g.y, g.z, g.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[j] = pack_bf16x2(round_bf16(bf16lo(u[j]) * r) * bf16lo(gw[j]), round_bf16(bf16hi(u[j]) * r) * bf16hi(gw[j]));
    xs[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
  consumer_sync();
}

__device__ __forceinline__ uint32_t ldcg_u32(const void* p) { return __ldcg(reinterpret_cast<const unsigned int*>(p)); }

// ---- phase 2: flash-decoding over the ring for one (kv head, split) item, 8 warps ---------------------
template <int REP>
__device__ __forceinline__ void mk_attention(const MkParams& p, const MkLayer& L, int W, int tid, float* sm_m, float* sm_l, float* sm_acc,
                                             int* sm_flag) {
  const int S = min(max((int)gridDim.x / p.KV, 1), MK_MAX_SPLITS);
  const int item = blockIdx.x;
  if (item >= p.KV * S) return;
  const int g = item / S, s = item % S;
  const int lane = tid & 31, warp = tid >> 5, half = lane >> 4, hl = lane & 15;
  const int len = min(p.pos + 1, W);
  const int C = (len + S - 1) / S;
  const int k_begin = min(s * C, len), k_end = min(k_begin + C, len);
  const float scale = 0.08838834764831845f;

  float qf[REP][8];
#pragma unroll
  for (int r = 0; r < REP; ++r) {
    const uint4 v = ldcg16(p.qbuf + (g * REP + r) * kHeadDim + hl * 8);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      qf[r][2 * j] = bf16lo(u[j]);
      qf[r][2 * j + 1] = bf16hi(u[j]);
    }
  }
  float m[REP], l[REP], acc[REP][8];
#pragma unroll
  for (int r = 0; r < REP; ++r) {
    m[r] = -INFINITY;
    l[r] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[r][j] = 0.f;
  }
  const int64_t row_stride = (int64_t)p.KV * kHeadDim;
  const bf16* kbase = L.cache_k + ((int64_t)p.batch_row * W) * row_stride + (int64_t)g * kHeadDim + hl * 8;
  const bf16* vbase = L.cache_v + ((int64_t)p.batch_row * W) * row_stride + (int64_t)g * kHeadDim + hl * 8;
  constexpr int STRIDE = MK_CONSUMER_WARPS * 2;
  constexpr int PF = 5;  // keys in flight per half-warp beyond the current one (16 half-warps x 5 x 512 B = 40 KB per SM)
  int slot = k_begin + warp * 2 + half;
  uint4 kq[PF], vq[PF];
#pragma unroll
  for (int u = 0; u < PF; ++u) {
    kq[u] = vq[u] = make_uint4(0, 0, 0, 0);
    const int sl = slot + u * STRIDE;
    if (sl < k_end) {  // the slot of the current token was written by another CTA in phase 1: bypass L1
      kq[u] = ldcg16(kbase + sl * row_stride);
      vq[u] = ldcg16(vbase + sl * row_stride);
    }
  }
  const int iters = (k_end - k_begin + STRIDE - 1) / STRIDE;
  for (int it = 0; it < iters; ++it) {
    const bool valid = slot < k_end;
    const uint4 kc = kq[0], vc = vq[0];
#pragma unroll
    for (int u = 0; u + 1 < PF; ++u) {
      kq[u] = kq[u + 1];
      vq[u] = vq[u + 1];
    }
    const int nslot = slot + PF * STRIDE;
    if (nslot < k_end) {
      kq[PF - 1] = ldcg16(kbase + nslot * row_stride);
      vq[PF - 1] = ldcg16(vbase + nslot * row_stride);
    }
    const uint32_t ku[4] = {kc.x, kc.y, kc.z, kc.w}, vu[4] = {vc.x, vc.y, vc.z, vc.w};
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
      d += __shfl_xor_sync(0xffffffffu, d, 8);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      if (valid) {
        const float sc = d * scale;
        const float mn = fmaxf(m[r], sc);
        const float corr = exp2f((m[r] - mn) * kLog2e);
        const float pr = exp2f((sc - mn) * kLog2e);
        l[r] = l[r] * corr + pr;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[r][j] = fmaf(pr, vf[j], acc[r][j] * corr);
        m[r] = mn;
      }
    }
    slot += STRIDE;
  }
  // merge half-warps, then the 8 warps through shared memory
#pragma unroll
  for (int r = 0; r < REP; ++r) {
    const float mo = __shfl_xor_sync(0xffffffffu, m[r], 16);
    const float lo = __shfl_xor_sync(0xffffffffu, l[r], 16);
    const float mn = fmaxf(m[r], mo);
    const float cs = (m[r] == -INFINITY) ? 0.f : exp2f((m[r] - mn) * kLog2e);
    const float co = (mo == -INFINITY) ? 0.f : exp2f((mo - mn) * kLog2e);
    l[r] = l[r] * cs + lo * co;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float ao = __shfl_xor_sync(0xffffffffu, acc[r][j], 16);
      acc[r][j] = acc[r][j] * cs + ao * co;
    }
    m[r] = mn;
    if (half == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sm_acc[(warp * REP + r) * kHeadDim + hl * 8 + j] = acc[r][j];
      if (hl == 0) {
        sm_m[warp * REP + r] = m[r];
        sm_l[warp * REP + r] = l[r];
      }
    }
  }
  consumer_sync();
  const int PSTRIDE = kHeadDim + 2;
  float* mine = p.partial + (((int64_t)g * S + s) * REP) * PSTRIDE;
  if (tid < kHeadDim) {
    const int d = tid;
#pragma unroll
    for (int r = 0; r < REP; ++r) {
      float mn = -INFINITY;
#pragma unroll
      for (int w = 0; w < MK_CONSUMER_WARPS; ++w) mn = fmaxf(mn, sm_m[w * REP + r]);
      float lt = 0.f, at = 0.f;
#pragma unroll
      for (int w = 0; w < MK_CONSUMER_WARPS; ++w) {
        const float mw = sm_m[w * REP + r];
        const float c = (mw == -INFINITY) ? 0.f : exp2f((mw - mn) * kLog2e);
        lt += sm_l[w * REP + r] * c;
        at += sm_acc[(w * REP + r) * kHeadDim + d] * c;
      }
      mine[r * PSTRIDE + 2 + d] = at;
      if (d == 0) {
        mine[r * PSTRIDE + 0] = mn;
        mine[r * PSTRIDE + 1] = lt;
      }
    }
    __threadfence();
  }
  consumer_sync();
  if (tid == 0) {
    const int prev = atomicAdd(&p.attn_counters[g], 1);
    const int last = (prev == S - 1);
    if (last) p.attn_counters[g] = 0;
    *sm_flag = last;
  }
  consumer_sync();
  if (!*sm_flag) return;
  __threadfence();
  if (tid < kHeadDim) {
    const int d = tid;
    const float* all = p.partial + ((int64_t)g * S) * REP * PSTRIDE;
#pragma unroll
    for (int r = 0; r < REP; ++r) {
      float mn = -INFINITY;
      for (int t = 0; t < S; ++t) mn = fmaxf(mn, __ldcg(all + ((int64_t)t * REP + r) * PSTRIDE));
      float lt = 0.f, at = 0.f;
      for (int t = 0; t < S; ++t) {
        const float* pp = all + ((int64_t)t * REP + r) * PSTRIDE;
        const float mt = __ldcg(pp);
        const float c = (mt == -INFINITY) ? 0.f : exp2f((mt - mn) * kLog2e);
        lt += __ldcg(pp + 1) * c;
        at += __ldcg(pp + 2 + d) * c;
      }
      p.abuf[(g * REP + r) * kHeadDim + d] = __float2bfloat16_rn(at / lt);
    }
  }
}

template <int REP>
__global__ void __launch_bounds__(MK_THREADS, 1) decode_megakernel(const MkParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  // layout: [ring: n_stages x 16 KB][xs: xs_bytes][barriers][reduction scratch]
  uint8_t* ring = smem;
  uint4* xs = reinterpret_cast<uint4*>(smem + (size_t)p.n_stages * MK_STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xs) + p.xs_bytes);
  uint64_t* empty = full + MK_MAX_STAGES;
  float* red = reinterpret_cast<float*>(empty + MK_MAX_STAGES);                       // [8]
  int* sm_flag = reinterpret_cast<int*>(red + 8 + 32);
  // attention merge scratch aliases the xs buffer (xs is dead during phase 2): m, l [8*REP], acc [8*REP*128]
  float* sm_m = reinterpret_cast<float*>(xs);
  float* sm_l = sm_m + MK_CONSUMER_WARPS * AD_MAX_REP;
  float* sm_acc = sm_l + MK_CONSUMER_WARPS * AD_MAX_REP;

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int q_dim = p.H * kHeadDim, kv_dim = p.KV * kHeadDim;
  RingState rs;
  rs.it = 0;

  if (tid >= MK_CONSUMERS) {
    // ================= producer warp (one lane issues; weights never wait for activations) =================
    if (tid == MK_CONSUMERS) {
      for (int l = 0; l < p.n_layers; ++l) {
        const MkLayer L = p.layers[l];
        produce_matrix(L.wqkv, q_dim + 2 * kv_dim, p.dim, ring, full, empty, p.n_stages, rs);
        produce_matrix(L.wo, p.dim, q_dim, ring, full, empty, p.n_stages, rs);
        produce_matrix(L.w13, 2 * p.hidden, p.dim, ring, full, empty, p.n_stages, rs);
        produce_matrix(L.w2, p.dim, p.hidden, ring, full, empty, p.n_stages, rs);
      }
      produce_matrix(p.w_out, p.vocab, p.dim, ring, full, empty, p.n_stages, rs);
    }
    return;
  }

  // ================= consumer warps =================
  const int64_t token = *p.token;
  for (int l = 0; l < p.n_layers; ++l) {
    const MkLayer L = p.layers[l];
    const int W = p.windows[l];
    const bf16* x_in = (l == 0) ? p.emb + token * p.dim : p.xbuf + (size_t)(l & 1) * p.dim;
    bf16* x_out = p.xbuf + (size_t)((l + 1) & 1) * p.dim;

    // ---- phase 1: RMSNorm + QKV + RoPE + ring scatter ----
    mk_stamp(p, tid, l, 0);
    stage_x(xs, x_in, L.attn_norm, p.dim, p.eps, red, tid);
    mk_stamp(p, tid, l, 1);
    {
      const int slot_row = p.batch_row * W + p.pos % W;
      const float* rope_row = p.rope + (int64_t)p.pos * (kHeadDim / 2) * 2;
      bf16* ck = L.cache_k + (int64_t)slot_row * kv_dim;
      bf16* cv = L.cache_v + (int64_t)slot_row * kv_dim;
      consume_matrix(q_dim + 2 * kv_dim, p.dim, ring, full, empty, p.n_stages, rs, xs, tid, [&](int n, float a0, float a1) {
        const float y0 = round_bf16(a0), y1 = round_bf16(a1);
        if (n < q_dim + kv_dim) {
          const float2 cs = *reinterpret_cast<const float2*>(rope_row + ((n & (kHeadDim - 1)) >> 1) * 2);
          float re, im;
          ref_cmul(y0, y1, cs.x, cs.y, re, im);
          const uint32_t packed = pack_bf16x2(re, im);
          if (n < q_dim)
            *reinterpret_cast<uint32_t*>(p.qbuf + n) = packed;
          else
            *reinterpret_cast<uint32_t*>(ck + (n - q_dim)) = packed;
        } else {
          *reinterpret_cast<uint32_t*>(cv + (n - q_dim - kv_dim)) = pack_bf16x2(y0, y1);
        }
      });
    }
    mk_stamp(p, tid, l, 2);
    grid_barrier(p, tid);
    mk_stamp(p, tid, l, 3);

    // ---- phase 2: attention over the ring ----
    mk_attention<REP>(p, L, W, tid, sm_m, sm_l, sm_acc, sm_flag);
    mk_stamp(p, tid, l, 4);
    grid_barrier(p, tid);
    mk_stamp(p, tid, l, 5);

    // ---- phase 3: wo + residual ----
    stage_x(xs, p.abuf, nullptr, q_dim, 0.f, red, tid);
    consume_matrix(p.dim, q_dim, ring, full, empty, p.n_stages, rs, xs, tid, [&](int n, float a0, float a1) {
      const uint32_t r = ldcg_u32(x_in + n);
      *reinterpret_cast<uint32_t*>(p.hbuf + n) = pack_bf16x2(round_bf16(a0) + bf16lo(r), round_bf16(a1) + bf16hi(r));
    });
    mk_stamp(p, tid, l, 6);
    grid_barrier(p, tid);
    mk_stamp(p, tid, l, 7);

    // ---- phase 4: RMSNorm + gate/up + SiLU*mul ----
    stage_x(xs, p.hbuf, L.ffn_norm, p.dim, p.eps, red, tid);
    consume_matrix(2 * p.hidden, p.dim, ring, full, empty, p.n_stages, rs, xs, tid, [&](int n, float a0, float a1) {
      const float s = round_bf16(ref_silu(round_bf16(a0)));
      p.gbuf[n >> 1] = __float2bfloat16_rn(s * round_bf16(a1));
    });
    mk_stamp(p, tid, l, 8);
    grid_barrier(p, tid);
    mk_stamp(p, tid, l, 9);

    // ---- phase 5: down + residual ----
    stage_x(xs, p.gbuf, nullptr, p.hidden, 0.f, red, tid);
    consume_matrix(p.dim, p.hidden, ring, full, empty, p.n_stages, rs, xs, tid, [&](int n, float a0, float a1) {
      const uint32_t r = ldcg_u32(p.hbuf + n);
      *reinterpret_cast<uint32_t*>(x_out + n) = pack_bf16x2(round_bf16(a0) + bf16lo(r), round_bf16(a1) + bf16hi(r));
    });
    mk_stamp(p, tid, l, 10);
    grid_barrier(p, tid);
    mk_stamp(p, tid, l, 11);
  }

  // ---- final RMSNorm + lm head (fp32 logits, each a bf16-rounded value) ----
  stage_x(xs, p.xbuf + (size_t)(p.n_layers & 1) * p.dim, p.final_norm, p.dim, p.eps, red, tid);
  consume_matrix(p.vocab, p.dim, ring, full, empty, p.n_stages, rs, xs, tid, [&](int n, float a0, float a1) {
    *reinterpret_cast<float2*>(p.logits + n) = make_float2(round_bf16(a0), round_bf16(a1));
  });
}

}  // namespace mb200
