// One persistent kernel for the second half of a dense transformer block in a batched decode step (5 <= T <= 64 tokens):
//
//     h   = x + attn_out . wo^T                     (transformer_layers.py:93,166)
//     hn  = RMSNorm(h, ffn_norm)                    (:167, :115-120)
//     g   = silu(hn . w1^T) * (hn . w3^T)           (:106)
//     out = h + g . w2^T                            (:106,168)
//
// Why: as four launches (stream-K GEMM, rmsnorm, stream-K GEMM, stream-K GEMM) each one pays launch latency, a pipeline fill and
// a drain -- measured 11-15 us of fixed cost per weight-streaming GEMM at Nemo-12B shapes against 6-45 us of streaming
// (scripts/bench_linear.py), which programmatic dependent launch only partly hides.  Here 148 CTAs stay resident through the
// four phases: the phases are separated by a grid barrier (a release/acquire counter), and the TMA producer of every CTA runs AHEAD
// across it -- weights depend on nothing, so while the epilogue warps reduce split tiles, wait in the barrier or normalise rows,
// the producer is already filling the ring with the next matrix's W tiles; only the A tiles (the previous phase's output) wait
// for the barrier.  Each GEMM phase is the stream-K scheme of gemm_streamk.cuh: unit = (128-wide tile, 64-deep k block), CTA c
// takes the c-th contiguous 1/G of the units, partial tiles are reduced deterministically through per-CTA workspace slots.
// Same roles as there: warp 0 TMA producer, warp 1 tcgen05.mma issuer (TMEM accumulators, double-buffered), warps 2-5 epilogue.
// Results are bit-identical to the four separate launches (same unit partition, same accumulation order).
#pragma once
#include "gemm_streamk.cuh"

namespace mb200 {

constexpr size_t FB_BAR_OFFSET = 26624;    // workspace header: barrier counter (monotonic), +128: epoch, +256: done counter
constexpr int FB_PHASES = 3;               // GEMM phases (wo, gate/up, down); the RMSNorm rows ride between the first two

struct FbParams {
  int T, dim, q_dim, hidden;
  float eps;
  const bf16* x;         // [T, dim] residual stream entering the block
  const bf16* norm_w;    // [dim] ffn_norm
  bf16* h;               // [T, dim]
  bf16* hn;              // [T, dim] workspace
  bf16* g;               // [T, hidden] workspace
  bf16* out;             // [T, dim]
  float* partials;       // [gridDim][TA][128] fp32
  unsigned* flags;       // [gridDim]
  unsigned* bar;         // barrier counter; bar + 32: epoch; bar + 64: done counter
};

struct FbCtx {
  uint8_t* smem;
  uint64_t *full, *empty, *tmem_full, *tmem_empty;
  uint32_t tmem_base;
  int G, cta;
  unsigned bar_base;  // value of the barrier counter when this launch started
};

__device__ __forceinline__ long long fb_first(long long total, int G, int c) {
  const long long per = total / G, rem = total % G;
  return (long long)c * per + (c < rem ? c : rem);
}

// ---- producer: W tiles of the first ring before `dep()` (the grid barrier that makes this phase's A valid), then the rest ----
template <int TA, int STAGES, class Dep>
__device__ __forceinline__ void fb_produce(const FbCtx& cx, const CUtensorMap* map_a, const CUtensorMap* map_w, int N, int K, uint32_t& it, Dep dep) {
  using Cfg = TgCfg<SK_BN, TA>;
  constexpr int STAGE_BYTES = Cfg::kStageBytes, A_BYTES = Cfg::kABytes;
  const int num_k = K / TG_BK;
  const long long total = (long long)(N / SK_BN) * num_k;
  const uint32_t u0 = (uint32_t)fb_first(total, cx.G, cx.cta), n_it = (uint32_t)(fb_first(total, cx.G, cx.cta + 1) - u0);
  auto issue = [&](uint32_t j, bool do_a, bool do_w) {
    const uint32_t u = u0 + j, s = (it + j) % STAGES;
    const int tile = (int)(u / (uint32_t)num_k), kb = (int)(u % (uint32_t)num_k);
    uint8_t* sa = cx.smem + s * STAGE_BYTES;
    if (do_w) tma_load_2d(sa + A_BYTES, map_w, &cx.full[s], kb * TG_BK, tile * SK_BN);
    if (do_a) tma_load_2d(sa, map_a, &cx.full[s], kb * TG_BK, 0);
  };
  const uint32_t head = n_it < (uint32_t)STAGES ? n_it : (uint32_t)STAGES;
  for (uint32_t j = 0; j < head; ++j) {
    const uint32_t g = it + j, s = g % STAGES, par = (g / STAGES) & 1;
    mbar_wait(&cx.empty[s], par ^ 1, 41, g);  // the slot may still hold a stage of the previous phase
    mbar_arrive_expect_tx(&cx.full[s], STAGE_BYTES);
    issue(j, false, true);
  }
  dep();
  for (uint32_t j = 0; j < head; ++j) issue(j, true, false);
  for (uint32_t j = head; j < n_it; ++j) {
    const uint32_t g = it + j, s = g % STAGES, par = (g / STAGES) & 1;
    mbar_wait(&cx.empty[s], par ^ 1, 42, g);
    mbar_arrive_expect_tx(&cx.full[s], STAGE_BYTES);
    issue(j, true, true);
  }
  it += n_it;
}

template <int TA, int STAGES>
__device__ __forceinline__ void fb_mma(const FbCtx& cx, int N, int K, uint32_t& it, uint32_t& seg) {
  using Cfg = TgCfg<SK_BN, TA>;
  constexpr int STAGE_BYTES = Cfg::kStageBytes, A_BYTES = Cfg::kABytes;
  const int num_k = K / TG_BK;
  const long long total = (long long)(N / SK_BN) * num_k;
  const long long u_begin = fb_first(total, cx.G, cx.cta), u_end = fb_first(total, cx.G, cx.cta + 1);
  for (long long u = u_begin; u < u_end; ++seg) {
    const int kb0 = (int)(u % num_k);
    const int kb1 = (int)min((long long)num_k, kb0 + (u_end - u));
    const uint32_t acc = seg & 1, acc_par = (seg >> 1) & 1;
    mbar_wait(&cx.tmem_empty[acc], acc_par ^ 1, 43, seg);
    tc_fence_after();
    const uint32_t d_tmem = cx.tmem_base + acc * SK_BN;
    for (int kb = kb0; kb < kb1; ++kb, ++it) {
      const uint32_t s = it % STAGES, par = (it / STAGES) & 1;
      mbar_wait(&cx.full[s], par, 44, it);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(cx.smem + s * STAGE_BYTES);
      const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + A_BYTES);
#pragma unroll
      for (int k = 0; k < TG_BK / 16; ++k) umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), Cfg::kIdesc, (kb > kb0 || k) ? 1u : 0u);
      umma_commit(&cx.empty[s]);
    }
    umma_commit(&cx.tmem_full[acc]);
    u += kb1 - kb0;
  }
}

template <int MODE, int TA>
__device__ __forceinline__ void fb_epilogue(const FbCtx& cx, const FbParams& p, const EpiParams& epi, int N, int K, uint32_t& seg) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lane_base = (warp & 3) * 32, etid = (int)threadIdx.x - 64;
  const bool row_ok = TA == 128 || lane_base + lane < TA;
  const int num_k = K / TG_BK;
  const long long total = (long long)(N / SK_BN) * num_k;
  const long long u_begin = fb_first(total, cx.G, cx.cta), u_end = fb_first(total, cx.G, cx.cta + 1);
  for (long long u = u_begin; u < u_end; ++seg) {
    const int tile = (int)(u / num_k), kb0 = (int)(u % num_k);
    const int kb1 = (int)min((long long)num_k, kb0 + (u_end - u));
    const int n0 = tile * SK_BN;
    const uint32_t acc = seg & 1, acc_par = (seg >> 1) & 1;
    mbar_wait(&cx.tmem_full[acc], acc_par, 45, seg);
    tc_fence_after();
    const int t = row_ok ? lane_base + lane : 0x7fffffff;
    const uint32_t trow = cx.tmem_base + ((uint32_t)lane_base << 16) + acc * SK_BN;
    uint32_t v[32];
    if (kb0 != 0) {  // contributor
      if (lane_base < TA) {
        float* mine = p.partials + ((size_t)cx.cta * TA + lane_base + lane) * SK_BN;
#pragma unroll 1
        for (int c = 0; c < SK_BN / 32; ++c) {
          tmem_ld_32x32b_x32(trow + c * 32, v);
#pragma unroll
          for (int q = 0; q < 8; ++q) reinterpret_cast<uint4*>(mine + c * 32)[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      }
      __threadfence();
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (etid == 0) st_release_u32(p.flags + cx.cta, 1u);
    } else if (kb1 != num_k) {  // owner of a tile that others finish
      const long long tile_end = (long long)(tile + 1) * num_k;
      int last = cx.cta;
      while (last + 1 < cx.G && fb_first(total, cx.G, last + 1) < tile_end) ++last;
      if (etid == 0) {
        for (int c = cx.cta + 1; c <= last; ++c) {
          unsigned spins = 0;
          while (ld_acquire_u32(p.flags + c) == 0u) {
            if (++spins == MB200_WATCHDOG_SPINS) {
              printf("[mb200 watchdog] ffn block %d waits for the partial of block %d (tile %d)\n", cx.cta, c, tile);
              __trap();
            }
          }
        }
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");
#pragma unroll 1
      for (int c = 0; c < SK_BN / 32; ++c) {
        tmem_ld_32x32b_x32(trow + c * 32, v);
        if (lane_base < TA) {
          for (int o = cx.cta + 1; o <= last; ++o) {
            const float* theirs = p.partials + ((size_t)o * TA + lane_base + lane) * SK_BN + c * 32;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint4 w = __ldcg(reinterpret_cast<const uint4*>(theirs) + q);
              v[4 * q] = __float_as_uint(__uint_as_float(v[4 * q]) + __uint_as_float(w.x));
              v[4 * q + 1] = __float_as_uint(__uint_as_float(v[4 * q + 1]) + __uint_as_float(w.y));
              v[4 * q + 2] = __float_as_uint(__uint_as_float(v[4 * q + 2]) + __uint_as_float(w.z));
              v[4 * q + 3] = __float_as_uint(__uint_as_float(v[4 * q + 3]) + __uint_as_float(w.w));
            }
          }
        }
        if (t < p.T) epi_chunk32<MODE>(epi, t, n0 + c * 32, v);
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (etid == 0)
        for (int c = cx.cta + 1; c <= last; ++c) p.flags[c] = 0u;
    } else {
#pragma unroll 1
      for (int c = 0; c < SK_BN / 32; ++c) {
        tmem_ld_32x32b_x32(trow + c * 32, v);
        if (t < p.T) epi_chunk32<MODE>(epi, t, n0 + c * 32, v);
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&cx.tmem_empty[acc]);
    u += kb1 - kb0;
  }
}

// grid barrier among the epilogue warps of all CTAs (they own every global write); `k` = 1-based index of the barrier in this launch
__device__ __forceinline__ void fb_barrier_arrive_wait(const FbCtx& cx, const FbParams& p, int k) {
  const int etid = (int)threadIdx.x - 64;
  asm volatile("fence.proxy.async;" ::: "memory");  // this phase's stores are read by other CTAs' TMA loads (async proxy)
  __threadfence();
  asm volatile("bar.sync 2, 128;" ::: "memory");
  if (etid == 0) {
    red_add_release_u32(p.bar, 1u);
    const unsigned target = cx.bar_base + (unsigned)k * (unsigned)cx.G;
    unsigned spins = 0;
    while ((int)(ld_acquire_u32(p.bar) - target) < 0) {
      if (++spins == MB200_WATCHDOG_SPINS * 4u) {
        printf("[mb200 watchdog] ffn block %d stuck in grid barrier %d (counter %u, target %u)\n", cx.cta, k, ld_acquire_u32(p.bar), target);
        __trap();
      }
    }
  }
  asm volatile("bar.sync 2, 128;" ::: "memory");
}
// the producer thread only WAITS for a barrier (it owns no global writes): then the A tiles of the next phase may be fetched
__device__ __forceinline__ void fb_barrier_wait_only(const FbCtx& cx, const FbParams& p, int k) {
  const unsigned target = cx.bar_base + (unsigned)k * (unsigned)cx.G;
  unsigned spins = 0;
  while ((int)(ld_acquire_u32(p.bar) - target) < 0) {
    if (++spins == MB200_WATCHDOG_SPINS * 4u) {
      printf("[mb200 watchdog] ffn block %d producer stuck before barrier %d\n", cx.cta, k);
      __trap();
    }
  }
  asm volatile("fence.proxy.async;" ::: "memory");  // other CTAs' generic-proxy stores -> this thread's TMA (async proxy) reads
}

template <int TA, int STAGES>
__global__ void __launch_bounds__(TG_THREADS, 1)
    ffn_block_kernel(const __grid_constant__ CUtensorMap map_attn, const __grid_constant__ CUtensorMap map_wo, const __grid_constant__ CUtensorMap map_hn,
                     const __grid_constant__ CUtensorMap map_w13, const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_w2,
                     const FbParams p) {
  using Cfg = TgCfg<SK_BN, TA>;
  constexpr int STAGE_BYTES = Cfg::kStageBytes, TMEM_COLS = Cfg::kTmemCols;
  extern __shared__ uint8_t smem_raw[];
  FbCtx cx;
  cx.smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  cx.full = reinterpret_cast<uint64_t*>(cx.smem + STAGES * STAGE_BYTES + Cfg::kSlack);
  cx.empty = cx.full + STAGES;
  cx.tmem_full = cx.empty + STAGES;
  cx.tmem_empty = cx.tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(cx.tmem_empty + 2);
  __shared__ float red[8];
  cx.G = (int)gridDim.x;
  cx.cta = (int)blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&cx.full[i], 1);
      mbar_init(&cx.empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&cx.tmem_full[i], 1);
      mbar_init(&cx.tmem_empty[i], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_base_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cx.tmem_base = *tmem_base_slot;
  if (threadIdx.x == 0) pdl_trigger();

  if (warp == 0) {
    // ================= TMA producer: runs ahead of the barriers with weight tiles =================
    if (lane == 0) {
      uint32_t it = 0;
      fb_produce<TA, STAGES>(cx, &map_attn, &map_wo, p.dim, p.q_dim, it, [&]() {
        pdl_wait();  // attn_out is the preceding kernel's output
        cx.bar_base = ld_acquire_u32(p.bar + 32);
      });
      fb_produce<TA, STAGES>(cx, &map_hn, &map_w13, 2 * p.hidden, p.dim, it, [&]() { fb_barrier_wait_only(cx, p, 2); });
      fb_produce<TA, STAGES>(cx, &map_g, &map_w2, p.dim, p.hidden, it, [&]() { fb_barrier_wait_only(cx, p, 3); });
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      uint32_t it = 0, seg = 0;
      fb_mma<TA, STAGES>(cx, p.dim, p.q_dim, it, seg);
      fb_mma<TA, STAGES>(cx, 2 * p.hidden, p.dim, it, seg);
      fb_mma<TA, STAGES>(cx, p.dim, p.hidden, it, seg);
    }
  } else {
    // ================= epilogue warps: GEMM epilogues, split-tile reduction, RMSNorm rows, grid barriers =================
    pdl_wait();
    cx.bar_base = ld_acquire_u32(p.bar + 32);  // completed barriers of earlier launches (published by their last CTA)
    const int etid = (int)threadIdx.x - 64;
    uint32_t seg = 0;
    EpiParams e;
    // ---- phase 0: h = x + attn_out . wo^T ----
    e.out = p.h;
    e.residual = p.x;
    e.ld_out = p.dim;
    fb_epilogue<EPI_RESIDUAL, TA>(cx, p, e, p.dim, p.q_dim, seg);
    fb_barrier_arrive_wait(cx, p, 1);
    // ---- RMSNorm: CTA r normalises row r (transformer_layers.py:115-120) ----
    if (cx.cta < p.T) {
      const int kc = p.dim >> 3;
      const uint4* xr = reinterpret_cast<const uint4*>(p.h + (size_t)cx.cta * p.dim);
      uint4* orow = reinterpret_cast<uint4*>(p.hn + (size_t)cx.cta * p.dim);
      const uint4* wn = reinterpret_cast<const uint4*>(p.norm_w);
      // same partition and summation order as rmsnorm_kernel (256 threads): thread e plays its threads e and e + 128
      float ssa = 0.f, ssb = 0.f;
      for (int c = etid; c < kc; c += 256) {
        const uint4 v4 = __ldcg(xr + c);
        const uint32_t u[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = bf16lo(u[j]), b = bf16hi(u[j]);
          ssa = fmaf(a, a, ssa);
          ssa = fmaf(b, b, ssa);
        }
      }
      for (int c = etid + 128; c < kc; c += 256) {
        const uint4 v4 = __ldcg(xr + c);
        const uint32_t u[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = bf16lo(u[j]), b = bf16hi(u[j]);
          ssb = fmaf(a, a, ssb);
          ssb = fmaf(b, b, ssb);
        }
      }
      ssa = warp_sum(ssa);
      ssb = warp_sum(ssb);
      if (lane == 0) {
        red[warp - 2] = ssa;
        red[warp + 2] = ssb;
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");
      float tot = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) tot += red[i];
      const float r = ref_rsqrt(tot / (float)p.dim + p.eps);
      for (int c = etid; c < kc; c += 128) {
        const uint4 v4 = __ldcg(xr + c), g4 = wn[c];
        const uint32_t u[4] = {v4.x, v4.y, v4.z, v4.w}, gw[4] = {g4.x, g4.y, g4.z, g4.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = pack_bf16x2(round_bf16(bf16lo(u[j]) * r) * bf16lo(gw[j]), round_bf16(bf16hi(u[j]) * r) * bf16hi(gw[j]));
        orow[c] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    fb_barrier_arrive_wait(cx, p, 2);
    // ---- phase 1: g = silu(hn . w1^T) * (hn . w3^T) ----
    e = EpiParams();
    e.out = p.g;
    e.ld_out = p.hidden;
    fb_epilogue<EPI_SWIGLU, TA>(cx, p, e, 2 * p.hidden, p.dim, seg);
    fb_barrier_arrive_wait(cx, p, 3);
    // ---- phase 2: out = h + g . w2^T ----
    e = EpiParams();
    e.out = p.out;
    e.residual = p.h;
    e.ld_out = p.dim;
    fb_epilogue<EPI_RESIDUAL, TA>(cx, p, e, p.dim, p.hidden, seg);
    // publish the barrier count for the next launch: the last CTA to get here (every CTA is past the third barrier by then)
    asm volatile("bar.sync 2, 128;" ::: "memory");
    if (etid == 0) {
      __threadfence();
      const int prev = atomicAdd(reinterpret_cast<int*>(p.bar + 64), 1);
      if (prev == cx.G - 1) {
        *reinterpret_cast<int*>(p.bar + 64) = 0;
        st_release_u32(p.bar + 32, cx.bar_base + 3u * (unsigned)cx.G);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(cx.tmem_base, TMEM_COLS);
}

inline int fb_sm_count() {
  static int sms = [] {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return n > SK_MAX_CTAS ? SK_MAX_CTAS : n;
  }();
  return sms;
}

inline bool ffn_block_eligible(int64_t T, int64_t dim, int64_t q_dim, int64_t hidden) {
  const char* e = getenv("MB200_FFN_BLOCK");  // opt-in: measured 5-7 % SLOWER than the separate launches
  if (e == nullptr || e[0] != '1') return false;
  const int sms = fb_sm_count();
  if (!(T >= 5 && T <= 64 && T <= sms && dim % SK_BN == 0 && (2 * hidden) % SK_BN == 0 && dim % TG_BK == 0 && q_dim % TG_BK == 0 && hidden % TG_BK == 0))
    return false;
  // every phase must give each CTA at least one (tile, k block) unit: then the unit partition equals the separate launches'
  const int64_t u0 = (dim / SK_BN) * (q_dim / TG_BK), u1 = (2 * hidden / SK_BN) * (dim / TG_BK), u2 = (dim / SK_BN) * (hidden / TG_BK);
  return u0 >= sms && u1 >= sms && u2 >= sms;
}

template <int TA>
int launch_ffn_block_ta(const void* attn_out, const void* wo, const void* x, const void* norm_w, const void* w13, const void* w2, void* h, void* g, void* out,
                        int64_t T, int64_t dim, int64_t q_dim, int64_t hidden, float eps, void* workspace, size_t workspace_bytes, size_t header,
                        cudaStream_t stream) {
  using Cfg = TgCfg<SK_BN, TA>;
  constexpr int STAGES = Cfg::kStages;
  constexpr int SMEM = STAGES * Cfg::kStageBytes + Cfg::kSlack + 1024 + 512;
  const int sms = fb_sm_count();
  const size_t hn_off = header + SK_PARTIAL_BYTES;
  const size_t need = hn_off + (size_t)T * dim * 2;
  if (workspace == nullptr || workspace_bytes < need) return fail(MB200_E_WORKSPACE, "ffn_block: workspace %zu < %zu", workspace_bytes, need);
  uint8_t* ws = (uint8_t*)workspace;
  FbParams p;
  p.T = (int)T;
  p.dim = (int)dim;
  p.q_dim = (int)q_dim;
  p.hidden = (int)hidden;
  p.eps = eps;
  p.x = (const bf16*)x;
  p.norm_w = (const bf16*)norm_w;
  p.h = (bf16*)h;
  p.hn = (bf16*)(ws + hn_off);
  p.g = (bf16*)g;
  p.out = (bf16*)out;
  p.partials = reinterpret_cast<float*>(ws + header);
  p.flags = reinterpret_cast<unsigned*>(ws + SK_FLAGS_OFFSET);
  p.bar = reinterpret_cast<unsigned*>(ws + FB_BAR_OFFSET);
  CUtensorMap m_attn, m_wo, m_hn, m_w13, m_g, m_w2;
  int rc;
  if ((rc = make_tensor_map_2d(&m_attn, attn_out, T, q_dim, TA))) return rc;
  if ((rc = make_tensor_map_2d(&m_wo, wo, dim, q_dim, SK_BN))) return rc;
  if ((rc = make_tensor_map_2d(&m_hn, p.hn, T, dim, TA))) return rc;
  if ((rc = make_tensor_map_2d(&m_w13, w13, 2 * hidden, dim, SK_BN))) return rc;
  if ((rc = make_tensor_map_2d(&m_g, p.g, T, hidden, TA))) return rc;
  if ((rc = make_tensor_map_2d(&m_w2, w2, dim, hidden, SK_BN))) return rc;
  auto kern = ffn_block_kernel<TA, STAGES>;
  MB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  MB_CHECK_CUDA(launch_pdl(kern, dim3((unsigned)sms), dim3(TG_THREADS), (size_t)SMEM, stream, m_attn, m_wo, m_hn, m_w13, m_g, m_w2, p));
  return MB200_OK;
}

inline int launch_ffn_block(const void* attn_out, const void* wo, const void* x, const void* norm_w, const void* w13, const void* w2, void* h, void* g, void* out,
                            int64_t T, int64_t dim, int64_t q_dim, int64_t hidden, float eps, void* workspace, size_t workspace_bytes, size_t header,
                            cudaStream_t stream) {
  if (T <= 32) return launch_ffn_block_ta<32>(attn_out, wo, x, norm_w, w13, w2, h, g, out, T, dim, q_dim, hidden, eps, workspace, workspace_bytes, header, stream);
  return launch_ffn_block_ta<64>(attn_out, wo, x, norm_w, w13, w2, h, g, out, T, dim, q_dim, hidden, eps, workspace, workspace_bytes, header, stream);
}

}  // namespace mb200
