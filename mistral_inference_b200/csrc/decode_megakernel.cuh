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
// Attention (phases 2a/2b) is flash-decoding split by POSITION: CTA c owns ring slots [c*C, (c+1)*C) of the sequence for ALL
// kv heads, which is one contiguous byte range of the [W, KV, hd] ring for K and one for V -- so it streams through the same
// shared-memory ring as the weights (large bulk copies, issued by the producer long before phase 1 ends; DRAM-friendly, unlike
// per-head 256-byte rows at a 2 KB stride).  Warp w serves kv head w (its H/KV query heads) over the slice, so no cross-warp merge
// is needed; every slice publishes (m, l, acc) per head and phase 2b merges the slices with all loads in flight at once.
// Only the row of the token being decoded (written in phase 1 by other CTAs) is read from global memory after the barrier.
#pragma once
#include "attn_decode.cuh"
#include "common.cuh"
#include "gemm_mma.cuh"

namespace mb200 {

constexpr int MK_CONSUMER_WARPS = 8;
constexpr int MK_CONSUMERS = MK_CONSUMER_WARPS * 32;
#ifndef MB200_MK_PRODUCERS
#define MB200_MK_PRODUCERS 2
#endif
constexpr int MK_PRODUCER_WARPS = MB200_MK_PRODUCERS;  // one issuing thread each, stages dealt round-robin (a single thread is ~700 cycles per stage: the stage period)
#ifndef MB200_MK_WG
#define MB200_MK_WG 0
#endif
#if MB200_MK_WG
// Experiment (round 2): producers in their own warpgroup so that setmaxnreg can move registers to the consumers (384 threads launch
// with 168 registers; the pool a CTA can re-acquire is only what it released: (168 - 120) x 128 = (192 - 168) x 256).
constexpr int MK_THREADS = MK_CONSUMERS + 128;
#else
constexpr int MK_THREADS = MK_CONSUMERS + 32 * MK_PRODUCER_WARPS;
#endif
constexpr int MK_WEIGHT_STAGE_BYTES = 16 * 1024;   // a weight stage: 2 rows x KC elements x 2 B
constexpr int MK_MAX_KC = MK_WEIGHT_STAGE_BYTES / 4;  // elements per row chunk
constexpr int MK_KV_PAD = 16;                      // K/V position rows are laid out with a 16-byte pad (ldmatrix bank spread)
constexpr int MK_STAGE_BYTES = 16 * 1024 + 16 * MK_KV_PAD;  // ring slot stride: also fits 8 padded 2 KB K/V position rows
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
  // Mixture of experts (moe.py:16-32): n_experts == 0 -> dense FeedForward (layers[l].w13 / w2)
  int n_experts, top_k;
  const bf16* const* moe_gate;  // [n_layers]              router weight [E, dim]
  const bf16* const* moe_w13;   // [n_layers * n_experts]  expert gate/up, packed like w13
  const bf16* const* moe_w2;    // [n_layers * n_experts]  expert down
  int n_layers;
  const bf16* emb;         // [V, dim]
  const bf16* final_norm;  // [dim]
  const bf16* w_out;       // [V, dim]
  const float* rope;       // [n_pos, 64, 2]
  const int64_t* token;    // device scalar: the token to embed
  int pos;                 // absolute position of that token
  int batch_row;           // which row of the cache this sequence uses
  float* logits;           // [V] fp32
  long long* next_token;   // optional: greedy argmax of the logits (first index on ties, like torch.argmax), or null
  unsigned long long* argmax_slots;  // [gridDim] per-CTA (value, index) keys
  int* argmax_counter;     // self-resetting
  int dim, hidden, H, KV, vocab;
  float eps;
  int n_stages, xs_bytes;
  int inflight_cap;  // max ring stages with outstanding bulk copies (< n_stages)
  int kv_uncapped;
  // scratch (global)
  unsigned* bar_flags;  // grid barrier counter (monotonic, never reset)
  unsigned* bar_epoch;  // device word: number of barriers completed by previous launches (published by the last CTA to finish)
  int* done_counter;    // self-resetting: CTAs that have finished this launch
  int* attn_counters;   // [KV]
  bf16* xbuf;           // [2][dim] residual stream ping-pong
  bf16* hbuf;           // [dim]
  bf16* qbuf;           // [H*hd]
  bf16* abuf;           // [H*hd] attention output
  bf16* gbuf;           // [hidden]
  float* partial;       // [KV][splits][REP][hd+2]
  unsigned long long* prof_bar;  // optional [gridDim][n_layers][6][2] arrive/leave %globaltimer of every CTA at every barrier
  unsigned long long* prof;  // optional [n_layers][12] globaltimer stamps written by CTA 0 (debug timeline), or null
};

// debug timeline: 8 sampled CTAs (every 21st) record %globaltimer at each phase boundary: prof[sample][layer][16] (12..15: inside attention)
__device__ __forceinline__ void mk_stamp(const MkParams& p, int tid, int layer, int idx) {
  if (p.prof != nullptr && tid == 0 && blockIdx.x % 21 == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    p.prof[((blockIdx.x / 21) * p.n_layers + layer) * 16 + idx] = t;
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
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Watchdog: a protocol bug in a persistent cooperative kernel is a GPU hang; every spin loop therefore gives up after
// ~seconds, prints what it was waiting for and traps, which turns the hang into a reportable launch failure.
#ifndef MB200_WATCHDOG_SPINS
#define MB200_WATCHDOG_SPINS (1u << 22)
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0, uint32_t it = 0) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins == MB200_WATCHDOG_SPINS) {
      printf("[mb200 watchdog] block %d thread %d stuck in mbarrier wait tag=%d it=%u parity=%u\n", (int)blockIdx.x, (int)threadIdx.x, tag, it,
             parity);
      __trap();
    }
  }
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

// Grid barrier among the consumer threads of all CTAs: ONE monotonically increasing counter, never reset.  Thread 0 of each
// CTA arrives with a single release atomic; it alone polls the word with acquire loads.  The k-th barrier of a launch is
// complete when the counter reaches base + k * gridDim, where `base` comes from the host (launch count x barriers per launch
// x gridDim), so nothing has to be read or reset on the device.
// (Measured alternatives on 148 CTAs: counter + generation word with fences 5 us; one flag per CTA polled by 148 threads of
// every CTA 1.5 us when arrivals are spread out but 4-5 us when all CTAs arrive together -- 22 K simultaneous polls.)
// Arrivals are spread over MK_BAR_WORDS counters on different 128-byte lines (CTA c -> word c % 8): when all CTAs arrive
// within ~0.5 us (after the short wo / down phases) 148 same-address atomics serialise at one L2 slice (~3 us measured).
constexpr int MK_BAR_WORDS = 8;
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_release_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void bar_stamp(const MkParams& p, int tid, int layer, int which, int leave) {
  if (p.prof_bar != nullptr && tid == 0 && layer >= 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    p.prof_bar[(((int64_t)blockIdx.x * p.n_layers + layer) * 6 + which) * 2 + leave] = t;
  }
}
// Word j collects the arrivals of a CONTIGUOUS range of CTAs (c -> c * 8 / gridDim), which lets a phase whose input chunk was
// produced by a known CTA range wait for just that range (see the down projection).
__device__ __forceinline__ int bar_word_of(int cta) { return (cta * MK_BAR_WORDS) / (int)gridDim.x; }
__device__ __forceinline__ unsigned bar_word_count(int j) {  // number of CTAs mapping to word j
  const int G = (int)gridDim.x;
  // first cta with c*8/G >= j  is ceil(j*G/8)
  const int lo = (j * G + MK_BAR_WORDS - 1) / MK_BAR_WORDS, hi = ((j + 1) * G + MK_BAR_WORDS - 1) / MK_BAR_WORDS;
  return (unsigned)(hi - lo);
}
__device__ __forceinline__ void grid_arrive(const MkParams& p, int tid, unsigned& epoch, int layer = -1, int which = 0) {
  ++epoch;  // number of barriers completed once this one is
  bar_stamp(p, tid, layer, which, 0);
  consumer_sync();  // every consumer thread's global writes of this phase happen-before thread 0's release below
  if (tid == 0) red_add_release_u32(p.bar_flags + bar_word_of(blockIdx.x) * 32, 1u);
}
// wait until every CTA in words [w_lo, w_hi] has arrived at barrier number `epoch`
__device__ __forceinline__ void grid_wait(const MkParams& p, int tid, unsigned epoch, int w_lo, int w_hi) {
  if (tid >= w_lo && tid <= w_hi) {
    const unsigned target = epoch * bar_word_count(tid);
    unsigned spins = 0;
    while ((int)(ld_acquire_u32(p.bar_flags + tid * 32) - target) < 0) {
      if (++spins == MB200_WATCHDOG_SPINS) {
        printf("[mb200 watchdog] block %d stuck in grid barrier word %d target=%u counter=%u\n", (int)blockIdx.x, tid, target,
               ld_acquire_u32(p.bar_flags + tid * 32));
        __trap();
      }
    }
  }
  consumer_sync();
}
__device__ __forceinline__ void grid_barrier(const MkParams& p, int tid, unsigned& epoch, int layer = -1, int which = 0) {
  grid_arrive(p, tid, epoch, layer, which);
  grid_wait(p, tid, epoch, 0, MK_BAR_WORDS - 1);
  bar_stamp(p, tid, layer, which, 1);
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

// routing decision of one MoE layer (moe.py:24-32), produced on every CTA by moe_route
constexpr int MK_MAX_TOPK = 4;
struct MoeRoute {
  int e[MK_MAX_TOPK];    // selected experts in ASCENDING expert index (the order `results +=` runs in, moe.py:29-31)
  float w[MK_MAX_TOPK];  // their routing weights (bf16 values)
};

struct RingState {
  uint32_t it;  // running stage counter (same sequence in producer and consumers)
};

// Stage order inside a matrix (same in producer and consumers): the CTA's pairs are taken in GROUPS of up to 8 (one pair
// per consumer warp); inside a group the stages are chunk-major: (pair g0+0, ch 0) ... (pair g0+g-1, ch 0), (pair g0+0, ch 1) ...
// So warp w touches stages base + ch*g + w: never more than 8 apart, i.e. always within one lap of the (>= 9-stage) ring --
// which is what makes parity-based mbarrier waits safe (a waiter two laps ahead would alias and pass early).

constexpr float kMaskedScore = -1.0e30f;

// ---- attention slice of this CTA (same arithmetic in producer and consumers) ---------------------------------
struct AttnSlice {
  int C, k_begin, k_end, n_kvst, pps, n_slices;  // positions per CTA, my range, my K (= V) stage count, positions per stage
};
__device__ __forceinline__ AttnSlice attn_slice(const MkParams& p, int W) {
  AttnSlice a;
  const int len = min(p.pos + 1, W);
  a.C = (len + (int)gridDim.x - 1) / (int)gridDim.x;
  a.n_slices = (len + a.C - 1) / a.C;
  a.k_begin = min((int)blockIdx.x * a.C, len);
  a.k_end = min(a.k_begin + a.C, len);
  a.pps = min(16, (MK_STAGE_BYTES / (p.KV * kHeadDim * 2 + MK_KV_PAD)) & ~7);  // 8 or 16 positions per stage (one MMA key block)
  a.n_kvst = (a.k_end - a.k_begin + a.pps - 1) / a.pps;
  return a;
}

// ---- producer (one thread) ---------------------------------------------------------------------------------------------
// The CTA's whole schedule (every layer: QKV slice, K/V slice, wo, gate/up, down slices; then the lm-head slice) is a pure
// function of (blockIdx, shapes, pos), so the producer simply walks it, blocking only on free ring slots.  Its per-stage
// budget is ~350 cycles (5.5 stages/us per SM at the HBM share of one SM): keep these loops lean -- a generic "schedule
// iterator" version (needed for a second, L2-prefetch cursor) cost 5 % end to end on its own.
// Experiment on record: running cp.async.bulk.prefetch.L2 D stages ahead to use the L2 as a second-level ring
// made things WORSE on B200 (3.4 -> 4.3 ms/token at D = 16, 5.0 at D = 64): the prefetched lines do not survive until the copy.
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}

struct Producer {
  uint8_t* ring;
  uint64_t* full;
  uint64_t* empty;
  int n_stages;
  uint32_t it;
  uint64_t policy;  // L2 evict-first: weights and old K/V rows are read exactly once per token

  // In-flight cap: stage `it` is only issued once stage it - cap has LANDED.  All n_stages slots still buffer data through the
  // phase boundaries, but the SM never has more than `cap` stages of read requests queued: right after a short phase the
  // consumers have drained the whole ring, and an uncapped producer then fires 12 x 16 KB at once -- the grid barrier's own
  // atomic / polls queue behind that burst in the SM's memory request path (measured: barrier latency 4 us after the wo
  // phase vs 1.5 us in steady state).  The bandwidth-delay product of one SM's HBM share is only ~3 stages.
  int cap;
  int me, n_prod;          // this producer issues the stages with it % n_prod == me
  bool kv_uncapped;        // K/V slice stages ignore the in-flight cap (they are needed at once in phase 2a)
  uint32_t slot, par;      // ring slot / parity of stage `it`, kept incrementally (no division in the issue loop)
  uint32_t cslot, cpar;    // same for stage it - cap

  __device__ __forceinline__ void advance() {
    ++it;
    if (++slot == (uint32_t)n_stages) {
      slot = 0;
      par ^= 1;
    }
    if (it > (uint32_t)cap && ++cslot == (uint32_t)n_stages) {
      cslot = 0;
      cpar ^= 1;
    }
  }
  // returns the slot's buffer (and its full barrier, armed for `bytes`) or nullptr when the stage belongs to another producer
  __device__ __forceinline__ uint8_t* acquire(uint32_t bytes, uint64_t*& bar, bool capped = true) {
    if ((int)(it % (uint32_t)n_prod) != me) {
      advance();
      return nullptr;
    }
    if (capped && it >= (uint32_t)cap) mbar_wait(&full[cslot], cpar, 7, it);  // stage it - cap has landed (its slot cannot have been refilled yet)
    mbar_wait(&empty[slot], par ^ 1, 1, it);
    bar = &full[slot];
    mbar_arrive_expect_tx(bar, bytes);
    uint8_t* dst = ring + (size_t)slot * MK_STAGE_BYTES;
    advance();
    return dst;
  }

  // this CTA's slice of one [N, K] weight matrix, in the stage order consume_matrix expects
  __device__ __forceinline__ void matrix(const bf16* W, int N, int K) {
    const MatCut c = cut_matrix(N, K);
    const uint32_t row_bytes = (uint32_t)c.kc * 2;
    for (int g0 = c.p0; g0 < c.p1; g0 += MK_CONSUMER_WARPS) {
      const int g = min(MK_CONSUMER_WARPS, c.p1 - g0);
      for (int ch = 0; ch < c.nch; ++ch) {
        for (int w = 0; w < g; ++w) {
          const bf16* r0 = W + (int64_t)(2 * (g0 + w)) * K;
          uint64_t* bar;
          uint8_t* dst = acquire(2 * row_bytes, bar);
          if (dst == nullptr) continue;
          if (c.nch == 1) {
            bulk_g2s_hint(dst, r0, 2 * row_bytes, bar, policy);  // the two rows are contiguous
          } else {
            bulk_g2s_hint(dst, r0 + ch * c.kc, row_bytes, bar, policy);
            bulk_g2s_hint(dst + row_bytes, r0 + K + ch * c.kc, row_bytes, bar, policy);
          }
        }
      }
    }
  }

  // expert down projections of one MoE layer: per group of 8 pairs expert-major, chunk-major (see consume_moe_down)
  __device__ __forceinline__ void moe_down(const bf16* const* w2, const int* sel, int top_k, int N, int K) {
    const MatCut c = cut_matrix(N, K);
    const uint32_t row_bytes = (uint32_t)c.kc * 2;
    for (int g0 = c.p0; g0 < c.p1; g0 += MK_CONSUMER_WARPS) {
      const int g = min(MK_CONSUMER_WARPS, c.p1 - g0);
      for (int j = 0; j < top_k; ++j) {
        const bf16* W = w2[sel[j]];
        for (int ch = 0; ch < c.nch; ++ch) {
          for (int w = 0; w < g; ++w) {
            const bf16* r0 = W + (int64_t)(2 * (g0 + w)) * K;
            uint64_t* bar;
            uint8_t* dst = acquire(2 * row_bytes, bar);
            if (dst == nullptr) continue;
            if (c.nch == 1) {
              bulk_g2s_hint(dst, r0, 2 * row_bytes, bar, policy);
            } else {
              bulk_g2s_hint(dst, r0 + ch * c.kc, row_bytes, bar, policy);
              bulk_g2s_hint(dst + row_bytes, r0 + K + ch * c.kc, row_bytes, bar, policy);
            }
          }
        }
      }
    }
  }

  // this CTA's K and V slice: alternating K / V stages of `pps` positions; one copy per position row so that rows sit
  // (row_bytes + 16) apart in shared memory (8 consecutive rows then cover all 32 banks for ldmatrix)
  __device__ __forceinline__ void kv_slice(const MkParams& p, const MkLayer& L, int W) {
    const AttnSlice a = attn_slice(p, W);
    const int64_t row_elems = (int64_t)p.KV * kHeadDim;
    const uint32_t row_bytes = (uint32_t)row_elems * 2;
    const bf16* kbase = L.cache_k + ((int64_t)p.batch_row * W) * row_elems;
    const bf16* vbase = L.cache_v + ((int64_t)p.batch_row * W) * row_elems;
    for (int j = 0; j < a.n_kvst; ++j) {
      const int k0 = a.k_begin + j * a.pps;
      const int rows = min(a.pps, a.k_end - k0);
#pragma unroll
      for (int kv = 0; kv < 2; ++kv) {
        uint64_t* bar;
        uint8_t* dst = acquire((uint32_t)rows * row_bytes, bar, !kv_uncapped);
        if (dst == nullptr) continue;
        const bf16* src = (kv ? vbase : kbase) + (int64_t)k0 * row_elems;
        for (int r = 0; r < rows; ++r) bulk_g2s_hint(dst + r * (row_bytes + MK_KV_PAD), src + (int64_t)r * row_elems, row_bytes, bar, policy);
      }
    }
  }
};

__device__ __forceinline__ void producer_main(const MkParams& p, uint8_t* ring, uint64_t* full, uint64_t* empty, int me, const MoeRoute* route,
                                              uint64_t* route_bar) {
  Producer pr;
  pr.ring = ring;
  pr.full = full;
  pr.empty = empty;
  pr.n_stages = p.n_stages;
  pr.it = 0;
  pr.cap = p.inflight_cap;
  pr.me = me;
  pr.n_prod = MK_PRODUCER_WARPS;
  pr.kv_uncapped = p.kv_uncapped != 0;
  pr.slot = pr.par = pr.cslot = pr.cpar = 0;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pr.policy));
  const int q_dim = p.H * kHeadDim, kv_dim = p.KV * kHeadDim;
  for (int l = 0; l < p.n_layers; ++l) {
    const MkLayer L = p.layers[l];
    pr.matrix(L.wqkv, q_dim + 2 * kv_dim, p.dim);
    pr.kv_slice(p, L, p.windows[l]);
    pr.matrix(L.wo, p.dim, q_dim);
    if (p.n_experts == 0) {
      pr.matrix(L.w13, 2 * p.hidden, p.dim);
      pr.matrix(L.w2, p.dim, p.hidden);
    } else {
      // expert weights are data dependent: wait for this layer's routing decision (the only point where the weight stream
      // cannot run ahead of the activations)
      mbar_wait(route_bar, (uint32_t)(l & 1), 8, (uint32_t)l);
      int sel[MK_MAX_TOPK];
      for (int j = 0; j < p.top_k; ++j) sel[j] = route->e[j];
      for (int j = 0; j < p.top_k; ++j) pr.matrix(p.moe_w13[l * p.n_experts + sel[j]], 2 * p.hidden, p.dim);
      pr.moe_down(p.moe_w2 + l * p.n_experts, sel, p.top_k, p.dim, p.hidden);
    }
  }
  pr.matrix(p.w_out, p.vocab, p.dim);
}

// ---- consumers: y[pair] = W[pair rows] . xs, epilogue(pair, acc0, acc1) on one lane ----------------
// Warp-per-pair inside a group: warp w owns pair g0+w and consumes its `nch` stages by itself (32 lanes x 16 B per step,
// unrolled -> plenty of ILP); only the owning warp releases a slot (empty barriers have arrival count 1).  One block
// barrier per GROUP keeps all warps within a group of each other (see the stage-order note above).
// `pre(n)` runs on the finishing lane BEFORE the pair's stages are consumed and its result is handed to `epi`: loads the
// epilogue needs (the residual) are then off the critical path of the phase's last pair (an L2 round trip right before the
// barrier's release store: measured 3.2-4.2 us barrier latency after wo / down vs 1.75 us after gate/up, which loads nothing).
// `ready(ch)` is called by ALL consumer threads before K-chunk `ch` of the FIRST group is touched: a phase can then stage its
// input chunk by chunk as the CTAs that produce it finish (pass a no-op when the input was staged up front).
template <class Ready, class Pre, class Epi>
__device__ __forceinline__ void consume_matrix(int N, int K, const uint8_t* ring, uint64_t* full, uint64_t* empty, int n_stages, RingState& rs,
                                               const uint4* xs, int tid, Ready ready, Pre pre, Epi epi) {
  const MatCut c = cut_matrix(N, K);
  const int lane = tid & 31, warp = tid >> 5;
  const int kc8 = c.kc >> 3;  // 16-byte chunks per row chunk
  for (int g0 = c.p0; g0 < c.p1; g0 += MK_CONSUMER_WARPS) {
    const int g = min(MK_CONSUMER_WARPS, c.p1 - g0);
    float a0 = 0.f, a1 = 0.f;
    uint2 prefetched = make_uint2(0u, 0u);
    if (warp < g && lane == 0) prefetched = pre(2 * (g0 + warp));
    for (int ch = 0; ch < c.nch; ++ch) {
      if (g0 == c.p0) ready(ch);
      if (warp < g) {
        const uint32_t it = rs.it + (uint32_t)(ch * g + warp);
        const uint32_t slot = it % n_stages, par = (it / n_stages) & 1;
        // Guard (tests/test_megakernel_protocol.py): bulk copies land out of order, so this warp may get here before the
        // slot's PREVIOUS fill (owned by another warp) has landed; `full` would then still be one phase behind and a
        // parity wait would alias and pass early.  Waiting first until that previous fill has been CONSUMED (same
        // condition the producer waits for before refilling) pins `full` to phase {r, r+1} when it is tested.
        mbar_wait(&empty[slot], par ^ 1, 2, it);
        mbar_wait(&full[slot], par, 3, it);
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
        if (lane == 0) mbar_arrive_n(&empty[slot], MK_CONSUMER_WARPS);  // this warp is the only reader of the slot
      }
    }
    if (warp < g) {
      a0 = warp_sum(a0);
      a1 = warp_sum(a1);
      if (lane == 0) epi(2 * (g0 + warp), a0, a1, prefetched);
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

// ---- phase 2a: partial attention of this CTA's position slice, all heads, out of the ring, on tensor cores --------------
// Warp w serves kv head w.  Per K/V stage pair (PPS = 8 or 16 positions = one key block) it computes
//   S[16 x PPS] = Q[16 x 128] K^T   with the REP query heads of the group as MMA rows 0..REP-1 (the other rows are zero),
//   online softmax on the accumulator fragments (fp32), P rounded to bf16 as the A operand, O[16 x 128] += P V.
// A CUDA-core version of the same loop (half a warp per key, shuffle-reduced dot products) is ISSUE bound: ~48 instructions
// per (key, head), 5.6 us per layer at kv_len 4096; this form is ~25x fewer instructions.  mma.sync (not tcgen05): the tiles
// are tiny and the softmax lives in registers; the tensor pipe is idle otherwise.
// All 8 warps wait for (and release) every K/V stage; a warp touches only its own head's 256-byte segment of each row, so the
// fresh row of the token being decoded and the zero-fill of rows past the slice are patched per warp, without block syncs.
template <int REP, int PPS>
__device__ __forceinline__ void mk_attention_slice(const MkParams& p, const MkLayer& L, int W, uint8_t* ring, uint64_t* full, uint64_t* empty,
                                                   int n_stages, RingState& rs, int tid, int layer) {
  static_assert(REP <= 8, "query heads of a group are MMA rows 0..7");
  const AttnSlice a = attn_slice(p, W);
  if (a.n_kvst == 0) return;
  const int lane = tid & 31, warp = tid >> 5;
  const bool has_head = warp < p.KV;
  const int g = has_head ? warp : 0;
  const int row = lane >> 2, cq = lane & 3;  // accumulator fragment: row = lane/4, column pair = lane%4
  const float sl2 = 0.08838834764831845f * kLog2e;  // scores are scaled by hd^-0.5; softmax in the exp2 domain
  const int64_t row_elems = (int64_t)p.KV * kHeadDim;
  const uint32_t row_stride = (uint32_t)row_elems * 2 + MK_KV_PAD;  // bytes between position rows in a stage

  // Q as A fragments: a0 = (row, k..k+1), a2 = (row, k+8..k+9); rows >= REP and the row+8 halves are zero
  uint32_t qa[8][4];
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    qa[ks][0] = qa[ks][1] = qa[ks][2] = qa[ks][3] = 0u;
    if (row < REP) {
      const bf16* qp = p.qbuf + (g * REP + row) * kHeadDim + ks * 16 + cq * 2;
      qa[ks][0] = ldcg_u32(qp);
      qa[ks][2] = ldcg_u32(qp + 8);
    }
  }
  // the row of the token being decoded was written in phase 1: lanes 0-15 hold its K segment, lanes 16-31 its V segment
  const int cur = p.pos % W;
  uint4 cur_kv = make_uint4(0, 0, 0, 0);
  if (cur >= a.k_begin && cur < a.k_end) {
    const bf16* base = (lane < 16 ? L.cache_k : L.cache_v);
    cur_kv = ldcg16(base + ((int64_t)p.batch_row * W + cur) * row_elems + (int64_t)g * kHeadDim + (lane & 15) * 8);
  }
  float o[16][4];
#pragma unroll
  for (int n = 0; n < 16; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  float m_run = kMaskedScore, l_run = 0.f;  // state of row `row` (this lane's share of l; reduced over the 4 lanes at the end)

  for (int j = 0; j < a.n_kvst; ++j) {
    const uint32_t itk = rs.it + 2u * j, itv = itk + 1;
    const uint32_t sk = itk % n_stages, pk = (itk / n_stages) & 1, sv = itv % n_stages, pv = (itv / n_stages) & 1;
    mbar_wait(&full[sk], pk, 5, itk);  // every warp visits every K/V stage in order: never more than one lap from the barrier
    mbar_wait(&full[sv], pv, 6, itv);
    if (j == 0) mk_stamp(p, tid, layer, 14);
    if (j == a.n_kvst - 1) mk_stamp(p, tid, layer, 15);
    if (has_head) {
      const int k0 = a.k_begin + j * PPS;
      const int nk = min(PPS, a.k_end - k0);
      uint8_t* kst = ring + (size_t)sk * MK_STAGE_BYTES + g * (kHeadDim * 2);
      uint8_t* vst = ring + (size_t)sv * MK_STAGE_BYTES + g * (kHeadDim * 2);
      // patches (own 256-byte segments only): fresh current-token row; zero the V rows past the slice (P = 0 there, but 0 * NaN)
      if (cur >= k0 && cur < k0 + nk)
        *reinterpret_cast<uint4*>((lane < 16 ? kst : vst) + (size_t)(cur - k0) * row_stride + (lane & 15) * 16) = cur_kv;
      if (nk < PPS)
        for (int r = nk + (lane >> 4); r < PPS; r += 2) *reinterpret_cast<uint4*>(vst + (size_t)r * row_stride + (lane & 15) * 16) = make_uint4(0, 0, 0, 0);
      // generic-proxy writes to a slot that the producer will refill through the async proxy (TMA): order them
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();

      // ---- S = Q K^T: PPS/8 key tiles x 8 k-steps; one ldmatrix.x4 = the B fragments of 2 k-steps for one key tile
      float sc[PPS / 8][4];
#pragma unroll
      for (int t = 0; t < PPS / 8; ++t) {
        sc[t][0] = sc[t][1] = sc[t][2] = sc[t][3] = 0.f;
        const uint32_t kaddr = smem_u32(kst) + (uint32_t)(t * 8 + (lane & 7)) * row_stride + (uint32_t)(lane >> 3) * 16;
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4(kaddr + k2 * 64, b0, b1, b2, b3);
          mma_bf16_16816(sc[t], qa[2 * k2], b0, b1);
          mma_bf16_16816(sc[t], qa[2 * k2 + 1], b2, b3);
        }
      }
      // ---- online softmax for row `row` (values c0, c1 of each key tile); keys >= nk are masked
      float mx = m_run;
#pragma unroll
      for (int t = 0; t < PPS / 8; ++t) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int key = t * 8 + cq * 2 + c;
          sc[t][c] = key < nk ? sc[t][c] * sl2 : kMaskedScore;
          mx = fmaxf(mx, sc[t][c]);
        }
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float corr = exp2f(m_run - mx);
      m_run = mx;
      l_run *= corr;
      uint32_t pa[4] = {0u, 0u, 0u, 0u};  // P as the A fragment of one k16 step: a0 = keys 0-7, a2 = keys 8-15; rows + 8 stay zero
#pragma unroll
      for (int t = 0; t < PPS / 8; ++t) {
        const float e0 = exp2f(sc[t][0] - mx), e1 = exp2f(sc[t][1] - mx);  // masked: exp2(-1e30) = 0
        l_run += e0 + e1;
        pa[2 * t] = pack_bf16x2(e0, e1);
      }
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        o[n][0] *= corr;
        o[n][1] *= corr;
      }
      // ---- O += P V: B = V^T fragments via ldmatrix.trans (rows = keys)
      if constexpr (PPS == 16) {
        const uint32_t vaddr = smem_u32(vst) + (uint32_t)((lane & 7) + ((lane >> 3) & 1) * 8) * row_stride + (uint32_t)(lane >> 4) * 16;
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) {  // dim tiles 2*n2, 2*n2+1
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4_trans(vaddr + n2 * 32, b0, b1, b2, b3);
          mma_bf16_16816(o[2 * n2], pa, b0, b1);
          mma_bf16_16816(o[2 * n2 + 1], pa, b2, b3);
        }
      } else {  // 8 keys: the k16 step's upper half (keys 8-15) is zero on both operands
        const uint32_t vaddr = smem_u32(vst) + (uint32_t)(lane & 7) * row_stride + (uint32_t)(lane >> 3) * 16;
#pragma unroll
        for (int n4 = 0; n4 < 4; ++n4) {  // dim tiles 4*n4 .. 4*n4+3
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4_trans(vaddr + n4 * 64, b0, b1, b2, b3);
          mma_bf16_16816(o[4 * n4], pa, b0, 0u);
          mma_bf16_16816(o[4 * n4 + 1], pa, b1, 0u);
          mma_bf16_16816(o[4 * n4 + 2], pa, b2, 0u);
          mma_bf16_16816(o[4 * n4 + 3], pa, b3, 0u);
        }
      }
    }
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&empty[sk]);
      mbar_arrive(&empty[sv]);
    }
  }
  rs.it += 2u * a.n_kvst;
  if (!has_head) return;
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);  // the 4 lanes of a row each summed their own columns
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
  if (row >= REP) return;
  // publish this slice's partial for query head g*REP + row; the merge works in the natural-exp domain: m / log2(e)
  m_run *= 0.6931471805599453f;
  const int PSTRIDE = kHeadDim + 2;
  float* mine = p.partial + ((int64_t)blockIdx.x * p.H + g * REP + row) * PSTRIDE;
#pragma unroll
  for (int n = 0; n < 16; ++n) *reinterpret_cast<float2*>(mine + 2 + n * 8 + cq * 2) = make_float2(o[n][0], o[n][1]);
  if (cq == 0) *reinterpret_cast<float2*>(mine) = make_float2(m_run, l_run);
}

// ---- phase 2b: merge the slices.  Unit u = (query head, 32-dim quarter); warp w folds slices w, w+8, ... (all loads issued
// before the first use), the 8 warps are folded through shared memory, lane = dim.
constexpr int MK_CMB_MAX = 20;  // slices per warp held in registers at once
__device__ __forceinline__ void mk_attention_combine(const MkParams& p, int W, int tid, float* scratch) {
  const AttnSlice a = attn_slice(p, W);
  const int lane = tid & 31, warp = tid >> 5;
  const int PSTRIDE = kHeadDim + 2;
  float* sm = scratch;  // [8 warps][34]: m, l, acc[32]
  for (int u = blockIdx.x; u < p.H * 4; u += gridDim.x) {
    const int h = u >> 2, q4 = u & 3;
    float m = kMaskedScore, l = 0.f, acc = 0.f;
    for (int a0 = warp; a0 < a.n_slices; a0 += MK_CONSUMER_WARPS * MK_CMB_MAX) {
      float mt[MK_CMB_MAX], lt[MK_CMB_MAX], at[MK_CMB_MAX];
#pragma unroll
      for (int i = 0; i < MK_CMB_MAX; ++i) {
        const int sl = a0 + i * MK_CONSUMER_WARPS;
        mt[i] = kMaskedScore;
        lt[i] = at[i] = 0.f;
        if (sl < a.n_slices) {
          const float* pp = p.partial + ((int64_t)sl * p.H + h) * PSTRIDE;
          const float2 ml = __ldcg(reinterpret_cast<const float2*>(pp));
          mt[i] = ml.x;
          lt[i] = ml.y;
          at[i] = __ldcg(pp + 2 + q4 * 32 + lane);
        }
      }
      // two passes: the max first, then every slice's weight is independent (no sequential rescale chain)
      float mn = m;
#pragma unroll
      for (int i = 0; i < MK_CMB_MAX; ++i) mn = fmaxf(mn, mt[i]);
      const float c0 = exp2f((m - mn) * kLog2e);
      l *= c0;
      acc *= c0;
#pragma unroll
      for (int i = 0; i < MK_CMB_MAX; ++i) {
        const float c1 = exp2f((mt[i] - mn) * kLog2e);
        l = fmaf(lt[i], c1, l);
        acc = fmaf(at[i], c1, acc);
      }
      m = mn;
    }
    sm[warp * 34 + 2 + lane] = acc;
    if (lane == 0) {
      sm[warp * 34] = m;
      sm[warp * 34 + 1] = l;
    }
    consumer_sync();
    if (warp == 0) {
      float mn = kMaskedScore;
#pragma unroll
      for (int w = 0; w < MK_CONSUMER_WARPS; ++w) mn = fmaxf(mn, sm[w * 34]);
      float lt = 0.f, at = 0.f;
#pragma unroll
      for (int w = 0; w < MK_CONSUMER_WARPS; ++w) {
        const float mw = sm[w * 34];
        const float c = exp2f((mw - mn) * kLog2e);
        lt += sm[w * 34 + 1] * c;
        at += sm[w * 34 + 2 + lane] * c;
      }
      p.abuf[h * kHeadDim + q4 * 32 + lane] = __float2bfloat16_rn(at / lt);
    }
    consumer_sync();
  }
}

// ---- mixture of experts (moe.py:24-32), batch 1 ------------------------------------------------------------------------------

// Router on every CTA (identical, deterministic): logits = bf16(hn . gate^T) (one warp per expert), top-k on the bf16 logits,
// softmax over the k selected in fp32, rounded to bf16 (moe.py:25-27).  Ties: lower expert index first.
__device__ __forceinline__ void moe_route(const MkParams& p, int layer, const uint4* xs, float* red, MoeRoute* route, uint64_t* route_bar, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  const bf16* gate = p.moe_gate[layer];
  const int kc = p.dim >> 3;
  for (int e = warp; e < p.n_experts; e += MK_CONSUMER_WARPS) {
    const uint4* wrow = reinterpret_cast<const uint4*>(gate + (int64_t)e * p.dim);
    float acc = 0.f;
    for (int i = lane; i < kc; i += 32) {
      const uint4 a = __ldg(wrow + i), x = xs[i];
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, xw[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc = fmaf(bf16lo(aw[j]), bf16lo(xw[j]), acc);
        acc = fmaf(bf16hi(aw[j]), bf16hi(xw[j]), acc);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) red[8 + e] = round_bf16(acc);  // red[8..8+E): router logits (E <= 32)
  }
  consumer_sync();
  if (tid == 0) {
    int sel[MK_MAX_TOPK];
    float val[MK_MAX_TOPK];
    unsigned taken = 0;
    for (int j = 0; j < p.top_k; ++j) {
      int best = -1;
      for (int e = 0; e < p.n_experts; ++e)
        if (!((taken >> e) & 1u) && (best < 0 || red[8 + e] > red[8 + best])) best = e;
      taken |= 1u << best;
      sel[j] = best;
      val[j] = red[8 + best];
    }
    float den = 0.f, ex[MK_MAX_TOPK];
    for (int j = 0; j < p.top_k; ++j) {
      ex[j] = expf(val[j] - val[0]);  // val[0] is the maximum
      den += ex[j];
    }
    for (int j = 0; j < p.top_k; ++j) val[j] = round_bf16(ex[j] / den);
    // ascending expert order, weights travelling with their experts
    for (int a = 1; a < p.top_k; ++a)
      for (int b = a; b > 0 && sel[b] < sel[b - 1]; --b) {
        const int ts = sel[b];
        sel[b] = sel[b - 1];
        sel[b - 1] = ts;
        const float tv = val[b];
        val[b] = val[b - 1];
        val[b - 1] = tv;
      }
    for (int j = 0; j < p.top_k; ++j) {
      route->e[j] = sel[j];
      route->w[j] = val[j];
    }
    mbar_arrive(route_bar);  // release: the producers may read `route` and start streaming the selected experts
  }
  consumer_sync();
}

// Expert down projections with the reference's accumulation: for the selected experts in ascending index,
//   y_e = bf16(W2_e g_e);  t_e = bf16(w_e * y_e);  res = (first ? t_e : bf16(res + t_e));   out = bf16(h + res)
// Stage order per group of 8 pairs: expert-major, then chunk-major (mirrored by Producer::moe_down).
template <class Pre, class Epi>
__device__ __forceinline__ void consume_moe_down(const MkParams& p, const MoeRoute& rt, const uint8_t* ring, uint64_t* full, uint64_t* empty,
                                                 int n_stages, RingState& rs, const uint4* xs, int tid, Pre pre, Epi epi) {
  const MatCut c = cut_matrix(p.dim, p.hidden);
  const int lane = tid & 31, warp = tid >> 5;
  const int kc8 = c.kc >> 3;
  const int hid8 = p.hidden >> 3;
  for (int g0 = c.p0; g0 < c.p1; g0 += MK_CONSUMER_WARPS) {
    const int g = min(MK_CONSUMER_WARPS, c.p1 - g0);
    if (warp < g) {
      uint2 prefetched = make_uint2(0u, 0u);
      if (lane == 0) prefetched = pre(2 * (g0 + warp));
      float r0 = 0.f, r1 = 0.f;
      for (int j = 0; j < p.top_k; ++j) {
        float a0 = 0.f, a1 = 0.f;
        for (int ch = 0; ch < c.nch; ++ch) {
          const uint32_t it = rs.it + (uint32_t)((j * c.nch + ch) * g + warp);
          const uint32_t slot = it % n_stages, par = (it / n_stages) & 1;
          mbar_wait(&empty[slot], par ^ 1, 2, it);
          mbar_wait(&full[slot], par, 3, it);
          const uint4* w0 = reinterpret_cast<const uint4*>(ring + (size_t)slot * MK_STAGE_BYTES);
          const uint4* w1 = w0 + kc8;
          const uint4* xc = xs + j * hid8 + ch * kc8;
#pragma unroll 4
          for (int i = lane; i < kc8; i += 32) {
            const uint4 a = w0[i], b = w1[i], x = xc[i];
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w}, xw[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float xl = bf16lo(xw[q]), xh = bf16hi(xw[q]);
              a0 = fmaf(bf16lo(aw[q]), xl, a0);
              a0 = fmaf(bf16hi(aw[q]), xh, a0);
              a1 = fmaf(bf16lo(bw[q]), xl, a1);
              a1 = fmaf(bf16hi(bw[q]), xh, a1);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive_n(&empty[slot], MK_CONSUMER_WARPS);
        }
        a0 = warp_sum(a0);
        a1 = warp_sum(a1);
        const float t0 = round_bf16(rt.w[j] * round_bf16(a0)), t1 = round_bf16(rt.w[j] * round_bf16(a1));
        r0 = (j == 0) ? t0 : round_bf16(r0 + t0);  // results starts at zero: bf16(0 + t) == t
        r1 = (j == 0) ? t1 : round_bf16(r1 + t1);
      }
      if (lane == 0) epi(2 * (g0 + warp), r0, r1, prefetched);
    }
    rs.it += (uint32_t)(g * c.nch * p.top_k);
    consumer_sync();
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
  MoeRoute* route = reinterpret_cast<MoeRoute*>(red + 48);                            // routing decision of the current MoE layer
  uint64_t* route_bar = reinterpret_cast<uint64_t*>(route + 1);                       // consumers -> producers: "route is valid"

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], MK_CONSUMER_WARPS);  // weight stages: the owning warp arrives x8; K/V stages: every warp x1
    }
    mbar_init(route_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int q_dim = p.H * kHeadDim, kv_dim = p.KV * kHeadDim;
  RingState rs;
  rs.it = 0;

  if (tid >= MK_CONSUMERS) {
    // ================= producers (one thread per producer warp; weights and old K/V rows never wait for activations) =================
#if MB200_MK_WG
    asm volatile("setmaxnreg.dec.sync.aligned.u32 120;");
    if ((tid & 31) == 0 && ((tid - MK_CONSUMERS) >> 5) < MK_PRODUCER_WARPS)
#else
    if ((tid & 31) == 0)
#endif
      producer_main(p, ring, full, empty, (tid - MK_CONSUMERS) >> 5, route, route_bar);
    return;
  }
#if MB200_MK_WG
  asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
#endif

  // ================= consumer warps =================
  // barriers completed by previous launches on this workspace.  Nobody writes the word until every CTA of this launch has
  // finished (see the end of the kernel), and launches are stream ordered, so this read cannot race.
  unsigned epoch = ld_acquire_u32(p.bar_epoch);
  const unsigned epoch0 = epoch;
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
      consume_matrix(q_dim + 2 * kv_dim, p.dim, ring, full, empty, p.n_stages, rs, xs, tid, [&](int) {},
                     [&](int n) { return *reinterpret_cast<const uint2*>(rope_row + ((n & (kHeadDim - 1)) >> 1) * 2); },
                     [&](int n, float a0, float a1, uint2 pf) {
        const float y0 = round_bf16(a0), y1 = round_bf16(a1);
        if (n < q_dim + kv_dim) {
          const float2 cs = make_float2(__uint_as_float(pf.x), __uint_as_float(pf.y));
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
    grid_barrier(p, tid, epoch, l, 0);
    mk_stamp(p, tid, l, 3);

    // ---- phase 2a: partial attention of my position slice;  2b: merge the slices ----
    if (attn_slice(p, W).pps == 16)
      mk_attention_slice<REP, 16>(p, L, W, ring, full, empty, p.n_stages, rs, tid, l);
    else
      mk_attention_slice<REP, 8>(p, L, W, ring, full, empty, p.n_stages, rs, tid, l);
    mk_stamp(p, tid, l, 12);
    grid_barrier(p, tid, epoch, l, 1);
    mk_stamp(p, tid, l, 13);
    mk_attention_combine(p, W, tid, reinterpret_cast<float*>(xs));
    mk_stamp(p, tid, l, 4);
    grid_barrier(p, tid, epoch, l, 2);
    mk_stamp(p, tid, l, 5);

    // ---- phase 3: wo + residual ----
    stage_x(xs, p.abuf, nullptr, q_dim, 0.f, red, tid);
    consume_matrix(p.dim, q_dim, ring, full, empty, p.n_stages, rs, xs, tid, [&](int) {}, [&](int n) { return make_uint2(ldcg_u32(x_in + n), 0u); }, [&](int n, float a0, float a1, uint2 pf) {
      const uint32_t r = pf.x;
      *reinterpret_cast<uint32_t*>(p.hbuf + n) = pack_bf16x2(round_bf16(a0) + bf16lo(r), round_bf16(a1) + bf16hi(r));
    });
    mk_stamp(p, tid, l, 6);
    grid_barrier(p, tid, epoch, l, 3);
    mk_stamp(p, tid, l, 7);

    if (p.n_experts == 0) {
    // ---- phase 4: RMSNorm + gate/up + SiLU*mul ----
      stage_x(xs, p.hbuf, L.ffn_norm, p.dim, p.eps, red, tid);
      consume_matrix(2 * p.hidden, p.dim, ring, full, empty, p.n_stages, rs, xs, tid, [&](int) {}, [&](int) { return make_uint2(0u, 0u); }, [&](int n, float a0, float a1, uint2) {
        const float s = round_bf16(ref_silu(round_bf16(a0)));
        p.gbuf[n >> 1] = __float2bfloat16_rn(s * round_bf16(a1));
      });
      mk_stamp(p, tid, l, 8);
      grid_barrier(p, tid, epoch, l, 4);
      mk_stamp(p, tid, l, 9);

      // ---- phase 5: down + residual ----
      // (Tried: no full barrier here -- stage g chunk by chunk as the barrier words of the CTA range that produced each K-chunk
      //  complete, via grid_arrive / grid_wait + the `ready` hook.  Correct, but 4 polling rounds + 4 block syncs cost more than
      //  the ~5 us gate/up arrival skew they hide: 345 vs 351 tok/s.)
      stage_x(xs, p.gbuf, nullptr, p.hidden, 0.f, red, tid);
      consume_matrix(p.dim, p.hidden, ring, full, empty, p.n_stages, rs, xs, tid, [&](int) {}, [&](int n) { return make_uint2(ldcg_u32(p.hbuf + n), 0u); },
                     [&](int n, float a0, float a1, uint2 pf) {
                       const uint32_t r = pf.x;
                       *reinterpret_cast<uint32_t*>(x_out + n) = pack_bf16x2(round_bf16(a0) + bf16lo(r), round_bf16(a1) + bf16hi(r));
                     });
    } else {
      // ---- phase 4 (MoE): RMSNorm + router; gate/up + SiLU*mul of the selected experts (ascending expert index) ----
      stage_x(xs, p.hbuf, L.ffn_norm, p.dim, p.eps, red, tid);
      moe_route(p, l, xs, red, route, route_bar, tid);
      const MoeRoute rt = *route;
      for (int j = 0; j < p.top_k; ++j) {
        bf16* gj = p.gbuf + (size_t)j * p.hidden;
        consume_matrix(2 * p.hidden, p.dim, ring, full, empty, p.n_stages, rs, xs, tid, [&](int) {}, [&](int) { return make_uint2(0u, 0u); },
                       [&](int n, float a0, float a1, uint2) {
                         const float sv = round_bf16(ref_silu(round_bf16(a0)));
                         gj[n >> 1] = __float2bfloat16_rn(sv * round_bf16(a1));
                       });
      }
      mk_stamp(p, tid, l, 8);
      grid_barrier(p, tid, epoch, l, 4);
      mk_stamp(p, tid, l, 9);
      // ---- phase 5 (MoE): expert down projections, weighted bf16 accumulation in expert order, + residual ----
      stage_x(xs, p.gbuf, nullptr, p.top_k * p.hidden, 0.f, red, tid);
      consume_moe_down(p, rt, ring, full, empty, p.n_stages, rs, xs, tid, [&](int n) { return make_uint2(ldcg_u32(p.hbuf + n), 0u); },
                       [&](int n, float r0, float r1, uint2 pf) {
                         *reinterpret_cast<uint32_t*>(x_out + n) = pack_bf16x2(r0 + bf16lo(pf.x), r1 + bf16hi(pf.x));
                       });
    }
    mk_stamp(p, tid, l, 10);
    grid_barrier(p, tid, epoch, l, 5);
    mk_stamp(p, tid, l, 11);
  }

  // ---- final RMSNorm + lm head (fp32 logits, each a bf16-rounded value) + greedy argmax ----
  stage_x(xs, p.xbuf + (size_t)(p.n_layers & 1) * p.dim, p.final_norm, p.dim, p.eps, red, tid);
  // argmax key: order-preserving map of the fp32 logit in the high word, ~index in the low word, so that the maximum key is
  // the largest logit and, among equal logits, the SMALLEST index (what torch.argmax returns; generate.py:156)
  unsigned long long best = 0ull;
  auto key_of = [](float v, int idx) {
    unsigned u = __float_as_uint(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned)(0x7fffffff - idx);
  };
  consume_matrix(p.vocab, p.dim, ring, full, empty, p.n_stages, rs, xs, tid, [&](int) {}, [&](int) { return make_uint2(0u, 0u); },
                 [&](int n, float a0, float a1, uint2) {
                   const float y0 = round_bf16(a0), y1 = round_bf16(a1);
                   *reinterpret_cast<float2*>(p.logits + n) = make_float2(y0, y1);
                   const unsigned long long k0 = key_of(y0, n), k1 = key_of(y1, n + 1);
                   best = max(best, max(k0, k1));
                 });
  if (p.next_token != nullptr) {
    // lane 0 of every warp holds its pairs' best; CTA reduce through shared memory, then the last CTA to arrive reduces all
    unsigned long long* sm_best = reinterpret_cast<unsigned long long*>(xs);
    consumer_sync();
    if ((tid & 31) == 0) sm_best[tid >> 5] = best;
    consumer_sync();
    if (tid == 0) {
      unsigned long long b = 0ull;
#pragma unroll
      for (int w = 0; w < MK_CONSUMER_WARPS; ++w) b = max(b, sm_best[w]);
      p.argmax_slots[blockIdx.x] = b;
      __threadfence();
      const int prev = atomicAdd(p.argmax_counter, 1);
      if (prev == (int)gridDim.x - 1) {
        *p.argmax_counter = 0;
        __threadfence();
        unsigned long long g = 0ull;
        for (int c = 0; c < (int)gridDim.x; ++c) g = max(g, __ldcg(p.argmax_slots + c));
        *p.next_token = (long long)(0x7fffffff - (int)(g & 0xffffffffull));
      }
    }
  }
  // publish the barrier epoch for the next launch: the last CTA to get here (all CTAs are past every barrier by then)
  consumer_sync();
  if (tid == 0) {
    __threadfence();
    const int prev = atomicAdd(p.done_counter, 1);
    if (prev == (int)gridDim.x - 1) {
      *p.done_counter = 0;
      st_release_u32(p.bar_epoch, epoch);
    }
  }
  (void)epoch0;
}

}  // namespace mb200
