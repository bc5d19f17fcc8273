// Mixture of experts for T > 1 tokens (prefill and batched decode): device-side router + grouped tensor-core expert GEMMs.
// Replaces MoeLayer.forward (moe.py:24-32) -- gate Linear, torch.topk, softmax, and the per-expert `torch.where` / index /
// FeedForward / weighted `+=` loop with its host synchronisations -- by five launches with NO host round trip:
//
//   moe_route_kernel    logits = bf16(hn . gate^T); top-k on the bf16 logits (ties: lower expert index first); fp32 softmax over
//                       the k selected -> bf16 weights (moe.py:25-27); the k (expert, weight) pairs of a token are stored in
//                       ASCENDING expert index: the order the reference's `results[idx] += w * expert(x)` loop visits them.
//   moe_plan_kernel     one CTA: per-expert row counts, segment starts (each expert's segment padded to a multiple of the GEMM's
//                       m-tile), a DETERMINISTIC slot for every (token, expert) pair (token order inside a segment: identical on
//                       every rank of an expert-parallel group, which is what lets ranks write each other's rows), and the list
//                       of m tiles of the experts this rank owns.
//   moe_gather_kernel   xs[slot] = hn[token], row_w[slot] = routing weight  (rows of local experts only)
//   gemm_tcgen05_grouped_kernel x2 (gemm_tcgen05.cuh): g = silu(xs W1_e^T) * (xs W3_e^T);  yw = bf16(w * bf16(g W2_e^T)), the
//                       second one storing every row on ALL ranks of the group (peer stores over NVLink) -- the exchange step of
//                       expert parallelism is the epilogue of the down projection, an all-gather of rows with no reduction.
//   moe_combine_kernel  (expert parallel: signal the peers / wait for theirs) out[t] = bf16(h[t] + sum_j yw[slot(t, j)]) with the
//                       sum taken in ascending expert index, each step rounded to bf16 like `results +=` -- bit-identical to the
//                       reference for any k and any number of ranks.
// Roofline: tensor pipe for prefill (2 * rows * 3 * dim * hidden flop), HBM for decode (every touched expert's weights once).
#pragma once
#include "gemm_streamk.cuh"
#include "gemm_tcgen05.cuh"

namespace mb200 {

constexpr int MOE_MAX_TOPK = 8;

// ---- router: one warp per token (WIDE = false: prefill) or one CTA per token with the 8 warps splitting the row (WIDE = true:
// decode-sized batches, where a single warp walking 4096 dims x 8 experts is pure load latency: 36 us measured for 8 tokens) ------
template <int E, bool WIDE>
__global__ void __launch_bounds__(256) moe_route_kernel(const bf16* __restrict__ hn, const bf16* __restrict__ gate_w, int T, int dim, int k,
                                                        int32_t* __restrict__ sel, bf16* __restrict__ wts) {
  pdl_trigger();
  pdl_wait();  // hn is the preceding RMSNorm's output
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = WIDE ? (int)blockIdx.x : (int)blockIdx.x * 8 + warp;
  if (t >= T) return;
  const int kc = dim >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(hn + (int64_t)t * dim);
  float acc[E];
#pragma unroll
  for (int e = 0; e < E; ++e) acc[e] = 0.f;
#pragma unroll 2
  for (int c = WIDE ? (int)threadIdx.x : lane; c < kc; c += WIDE ? 256 : 32) {
    const uint4 xv = xr[c];
    const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const uint4 gv = __ldg(reinterpret_cast<const uint4*>(gate_w + (int64_t)e * dim) + c);
      const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[e] = fmaf(bf16lo(gw[j]), bf16lo(xw[j]), acc[e]);
        acc[e] = fmaf(bf16hi(gw[j]), bf16hi(xw[j]), acc[e]);
      }
    }
  }
  if constexpr (WIDE) {  // fold the 8 warps' partial sums in a fixed order
    __shared__ float part[8][E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const float v = warp_sum(acc[e]);
      if (lane == 0) part[warp][e] = v;
    }
    __syncthreads();
    if (warp != 0) return;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += part[w][e];
      acc[e] = round_bf16(v);  // the router Linear's bf16 output (moe.py:25)
    }
  } else {
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = round_bf16(warp_sum(acc[e]));  // the router Linear's bf16 output (moe.py:25)
  }
  if (lane != 0) return;
  int se[MOE_MAX_TOPK];
  float sv[MOE_MAX_TOPK];
  unsigned taken = 0u;
  for (int j = 0; j < k; ++j) {  // top-k on the bf16 logits, ties -> lower index
    int best = -1;
    float bv = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (!((taken >> e) & 1u) && (best < 0 || acc[e] > bv)) {
        best = e;
        bv = acc[e];
      }
    taken |= 1u << best;
    se[j] = best;
    sv[j] = bv;
  }
  float den = 0.f, ex[MOE_MAX_TOPK];
  for (int j = 0; j < k; ++j) {
    ex[j] = expf(sv[j] - sv[0]);  // sv[0] is the maximum
    den += ex[j];
  }
  for (int j = 0; j < k; ++j) sv[j] = round_bf16(ex[j] / den);  // softmax in fp32, then .to(bf16) (moe.py:27)
  for (int a = 1; a < k; ++a)  // ascending expert index, weights travelling with their experts
    for (int b = a; b > 0 && se[b] < se[b - 1]; --b) {
      const int ts = se[b];
      se[b] = se[b - 1];
      se[b - 1] = ts;
      const float tv = sv[b];
      sv[b] = sv[b - 1];
      sv[b - 1] = tv;
    }
  for (int j = 0; j < k; ++j) {
    sel[(int64_t)t * k + j] = se[j];
    wts[(int64_t)t * k + j] = __float2bfloat16_rn(sv[j]);
  }
}

// ---- plan: one CTA ---------------------------------------------------------------------------------------------------------
// plan layout (int32): [0] m tiles owned by this rank, [1] padded rows in total, [2] tile capacity, [3] pairs, [4]/[5] statistics, [8 + e] start of
// expert e's segment (e = 0..E), then tile_expert[capacity], tile_row0[capacity] from word MOE_PLAN_HEADER.
constexpr int MP_THREADS = 1024;
__global__ void __launch_bounds__(MP_THREADS) moe_plan_kernel(const int32_t* __restrict__ sel, int pairs, int E, int tile_rows, int shard_rank,
                                                              int shard_world, int tile_cap, int32_t* __restrict__ slot, int32_t* __restrict__ plan) {
  extern __shared__ int32_t sm[];  // cnt[E][MP_THREADS], then seg[E + 1], total[E]
  int32_t* cnt = sm;
  int32_t* seg = sm + E * MP_THREADS;
  int32_t* total = seg + E + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (pairs <= 512) {
    // decode-sized batch: thread e walks the whole (short) pair list for expert e -- same deterministic plan, a few hundred cycles
    if (tid < E) {
      int n = 0;
      for (int i = 0; i < pairs; ++i)
        if (sel[i] == tid) slot[i] = n++;  // position inside the expert's segment; the segment start is added below
      total[tid] = n;
    }
    __syncthreads();
    if (tid == 0) {
      int rows = 0, n = 0, touched = 0;
      for (int e = 0; e < E; ++e) {
        seg[e] = rows;
        plan[8 + e] = rows;
        const int m_tiles = (total[e] + tile_rows - 1) / tile_rows;
        touched += total[e] > 0;
        if (e % shard_world == shard_rank)
          for (int m = 0; m < m_tiles && n < tile_cap; ++m, ++n) {
            plan[MOE_PLAN_HEADER + n] = e;
            plan[MOE_PLAN_HEADER + tile_cap + n] = rows + m * tile_rows;
          }
        rows += m_tiles * tile_rows;
      }
      plan[8 + E] = rows;
      plan[0] = n;
      plan[1] = rows;
      plan[2] = tile_cap;
      plan[3] = pairs;
      plan[4] += touched;
      plan[5] += 1;
      plan[6] = 0;  // decode-sized calls never take the cluster variant
    }
    __syncthreads();
    for (int i = tid; i < pairs; i += MP_THREADS) slot[i] += seg[sel[i]];
    return;
  }
  const int per = (pairs + MP_THREADS - 1) / MP_THREADS;
  const int p0 = min(tid * per, pairs), p1 = min(p0 + per, pairs);
  for (int e = 0; e < E; ++e) cnt[e * MP_THREADS + tid] = 0;
  for (int i = p0; i < p1; ++i) cnt[sel[i] * MP_THREADS + tid] += 1;
  __syncthreads();
  // exclusive scan of every expert's 1024 counts: warp w takes experts w, w + 32, ...; lane l scans entries [32 l, 32 l + 32)
  for (int e = warp; e < E; e += MP_THREADS / 32) {
    int32_t* c = cnt + e * MP_THREADS + lane * 32;
    int run = 0;
    for (int i = 0; i < 32; ++i) {
      const int v = c[i];
      c[i] = run;
      run += v;
    }
    int incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    const int base = incl - run;
    for (int i = 0; i < 32; ++i) c[i] += base;
    if (lane == 31) total[e] = incl;
  }
  __syncthreads();
  if (tid == 0) {
    int rows = 0, n = 0, np = 0;
    for (int e = 0; e < E; ++e) {
      seg[e] = rows;
      plan[8 + e] = rows;
      const int m_tiles = (total[e] + tile_rows - 1) / tile_rows;
      if (e % shard_world == shard_rank) {
        for (int m = 0; m < m_tiles && n < tile_cap; ++m, ++n) {
          plan[MOE_PLAN_HEADER + n] = e;
          plan[MOE_PLAN_HEADER + tile_cap + n] = rows + m * tile_rows;
        }
        for (int m = 0; m < m_tiles && np < tile_cap; m += 2, ++np) {  // pairs of vertically adjacent tiles for the 2-CTA cluster GEMM
          plan[MOE_PLAN_HEADER + 2 * tile_cap + np] = e;
          plan[MOE_PLAN_HEADER + 3 * tile_cap + np] = (rows + m * tile_rows) | (m + 1 < m_tiles ? MOE_PAIR_SECOND : 0);
        }
      }
      rows += m_tiles * tile_rows;
    }
    plan[6] = np;
    seg[E] = rows;
    plan[8 + E] = rows;
    plan[0] = n;
    plan[1] = rows;
    plan[2] = tile_cap;
    plan[3] = pairs;
    int touched = 0;
    for (int e = 0; e < E; ++e) touched += total[e] > 0;
    plan[4] += touched;  // statistics (never reset by the kernel): experts with at least one row, summed over calls ...
    plan[5] += 1;        // ... and the number of calls: the measured "distinct experts per layer" of bench.py
  }
  __syncthreads();
  int run[MOE_MAX_EXPERTS];
#pragma unroll
  for (int e = 0; e < MOE_MAX_EXPERTS; ++e) run[e] = 0;
  for (int i = p0; i < p1; ++i) {
    const int e = sel[i];
    int r = 0;
#pragma unroll
    for (int q = 0; q < MOE_MAX_EXPERTS; ++q)  // static indexing keeps `run` in registers
      if (q == e) r = run[q]++;
    slot[i] = seg[e] + cnt[e * MP_THREADS + tid] + r;
  }
}

// ---- gather: one warp per (token, expert) pair ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) moe_gather_kernel(const uint4* __restrict__ hn, const int32_t* __restrict__ sel, const bf16* __restrict__ wts,
                                                         const int32_t* __restrict__ slot, int pairs, int k, int row_chunks, int shard_rank,
                                                         int shard_world, uint4* __restrict__ xs, bf16* __restrict__ row_w) {
  const int pair = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (pair >= pairs) return;
  if (sel[pair] % shard_world != shard_rank) return;  // another rank's expert
  const int t = pair / k, s = slot[pair];
  if (lane == 0) row_w[s] = wts[pair];
  const uint4* src = hn + (int64_t)t * row_chunks;
  uint4* dst = xs + (int64_t)s * row_chunks;
  for (int c = lane; c < row_chunks; c += 32) dst[c] = src[c];
}

// ---- combine (+ the expert-parallel handshake) --------------------------------------------------------------------------------
struct MoeCombineParams {
  const uint4* yw;        // [rows, dim] weighted expert outputs (all ranks' rows once the handshake is through)
  const int32_t* slot;    // [T, k]
  const uint4* residual;  // h [T, dim] or null
  uint4* out;             // [T, dim]
  int T, k, row_chunks;
  // expert parallel (n_ranks > 1): flags[r] of THIS rank is written by rank r when all its rows of this call have landed here
  int n_ranks, my_rank;
  unsigned* my_flags;               // [n_ranks] in this rank's comm buffer
  unsigned* peer_flags[kMaxPeers];  // the same array on the other ranks (mapped)
  unsigned* epoch;                  // local device word: completed calls on this buffer
  int* done_counter;                // local, self-resetting
};
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(128) moe_combine_kernel(const MoeCombineParams p) {
  const int t = blockIdx.x;
  if (p.n_ranks > 1) {
    // This kernel starts after this rank's down projection has finished (stream order), i.e. after its peer stores were issued
    // and completed.  Block 0 tells every peer; every block then waits until every peer has told this rank.
    __shared__ unsigned target;
    if (threadIdx.x == 0) {
      const unsigned e = *reinterpret_cast<volatile unsigned*>(p.epoch) + 1u;
      target = e;
      if (blockIdx.x == 0) {
        __threadfence_system();
        for (int r = 0; r < p.n_ranks - 1; ++r) st_release_sys_u32(p.peer_flags[r] + p.my_rank, e);
        st_release_sys_u32(p.my_flags + p.my_rank, e);
      }
      for (int r = 0; r < p.n_ranks; ++r) {
        unsigned long long spins = 0;
        while ((int)(ld_acquire_sys_u32(p.my_flags + r) - e) < 0) {
          if (++spins == (1ull << 26)) {
            printf("[mb200 watchdog] rank %d block %d: no signal from rank %d for MoE exchange %u (have %u)\n", p.my_rank, (int)blockIdx.x, r, e,
                   ld_acquire_sys_u32(p.my_flags + r));
            __trap();
          }
        }
      }
    }
    __syncthreads();
    (void)target;
  }
  const int32_t* sl = p.slot + (int64_t)t * p.k;
  for (int c = threadIdx.x; c < p.row_chunks; c += 128) {
    float r[8];
    {
      const uint4 v = p.yw[(int64_t)sl[0] * p.row_chunks + c];
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        r[2 * j] = bf16lo(u[j]);  // results starts at zero: bf16(0 + t) == t
        r[2 * j + 1] = bf16hi(u[j]);
      }
    }
    for (int q = 1; q < p.k; ++q) {
      const uint4 v = p.yw[(int64_t)sl[q] * p.row_chunks + c];
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        r[2 * j] = round_bf16(r[2 * j] + bf16lo(u[j]));  // results[idx] += ... in bf16 (moe.py:31)
        r[2 * j + 1] = round_bf16(r[2 * j + 1] + bf16hi(u[j]));
      }
    }
    uint32_t o[4];
    if (p.residual != nullptr) {
      const uint4 h = p.residual[(int64_t)t * p.row_chunks + c];
      const uint32_t hu[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = pack_bf16x2(bf16lo(hu[j]) + r[2 * j], bf16hi(hu[j]) + r[2 * j + 1]);  // h + r (transformer_layers.py:168)
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = pack_bf16x2(r[2 * j], r[2 * j + 1]);
    }
    p.out[(int64_t)t * p.row_chunks + c] = make_uint4(o[0], o[1], o[2], o[3]);
  }
  if (p.n_ranks > 1) {  // the last block to finish publishes the epoch for the next call on this buffer
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const int prev = atomicAdd(p.done_counter, 1);
      if (prev == (int)gridDim.x - 1) {
        *p.done_counter = 0;
        __threadfence();
        *reinterpret_cast<volatile unsigned*>(p.epoch) = *reinterpret_cast<volatile unsigned*>(p.epoch) + 1u;
      }
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------------------------------
inline int moe_tile_rows(int64_t T) { return T <= 32 ? 32 : (T <= 64 ? 64 : 128); }  // an expert gets at most one row per token: decode batches fit ONE short m tile
inline int64_t moe_tile_cap(int64_t pairs, int64_t E, int tile_rows) { return (pairs + tile_rows - 1) / tile_rows + E; }
inline int64_t moe_plan_words(int64_t pairs, int64_t E, int tile_rows) { return MOE_PLAN_HEADER + 4 * moe_tile_cap(pairs, E, tile_rows); }
inline int64_t moe_row_cap(int64_t pairs, int64_t E, int tile_rows) { return moe_tile_cap(pairs, E, tile_rows) * tile_rows; }

template <int MODE, int BN, int TA, int CL = 1>
int launch_grouped_bn(const void* a, int64_t rows_cap, int64_t K, int64_t N, const void* const* w_host, int E, const int32_t* plan, const EpiParams& epi,
                      int sms, cudaStream_t stream) {
  using Cfg = TgCfg<BN, TA>;
  CUtensorMap map_a;
  MoeWeightMaps maps;
  int rc = make_tensor_map_2d(&map_a, a, rows_cap, K, TA);
  if (rc) return rc;
  for (int e = 0; e < MOE_MAX_EXPERTS; ++e) {
    const void* w = e < E && w_host[e] != nullptr ? w_host[e] : nullptr;
    if (w == nullptr) {  // an expert of another rank: never referenced by this rank's tiles
      maps.m[e] = map_a;
      continue;
    }
    rc = make_tensor_map_2d(&maps.m[e], w, N, K, BN / CL);  // cluster pairs: each CTA fetches half of the W tile and multicasts it
    if (rc) return rc;
  }
  TcGemmParams p;
  p.T = (int)rows_cap;
  p.N = (int)N;
  p.K = (int)K;
  p.epi = epi;
  MB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_grouped_kernel<MODE, CL, BN, TA>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
  if (CL == 1) {
    gemm_tcgen05_grouped_kernel<MODE, CL, BN, TA><<<sms, TG_THREADS, Cfg::kSmem, stream>>>(map_a, maps, p, plan);
    MB_CHECK_LAUNCH("gemm_tcgen05_grouped_kernel");
    return MB200_OK;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2u * (unsigned)(sms / 2));
  cfg.blockDim = dim3(TG_THREADS);
  cfg.dynamicSmemBytes = Cfg::kSmem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  MB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_grouped_kernel<MODE, CL, BN, TA>, map_a, maps, p, plan));
  MB_CHECK_LAUNCH("gemm_tcgen05_grouped_kernel<cluster 2>");
  return MB200_OK;
}

// tile width: prefill (128-row tiles) takes the widest tile that divides N; decode-sized calls (32-row tiles, HBM-bound) pick the
// width that fills the rounds of the persistent schedule best for the EXPECTED number of touched experts
// decode-sized calls: the stream-K weight-streaming kernel over (expert segment, n tile, k block) units
template <int MODE, int TA>
int launch_grouped_streamk(const void* a, int64_t rows_cap, int64_t K, int64_t N, const void* const* w_host, int E, const int32_t* plan,
                           const EpiParams& epi, void* workspace, size_t workspace_bytes, size_t header, int sms, cudaStream_t stream) {
  using Cfg = TgCfg<SK_BN, TA>;
  if (sms > SK_MAX_CTAS) sms = SK_MAX_CTAS;
  if (workspace == nullptr || workspace_bytes < header + SK_PARTIAL_BYTES) return fail(MB200_E_WORKSPACE, "grouped stream-K gemm: workspace %zu < %zu", workspace_bytes, header + SK_PARTIAL_BYTES);
  CUtensorMap map_a;
  MoeWeightMaps maps;
  int rc = make_tensor_map_2d(&map_a, a, rows_cap, K, TA);
  if (rc) return rc;
  for (int e = 0; e < MOE_MAX_EXPERTS; ++e) {
    const void* w = e < E && w_host[e] != nullptr ? w_host[e] : nullptr;
    if (w == nullptr) {
      maps.m[e] = map_a;
      continue;
    }
    rc = make_tensor_map_2d(&maps.m[e], w, N, K, SK_BN);
    if (rc) return rc;
  }
  SkParams p;
  p.T = (int)rows_cap;
  p.N = (int)N;
  p.K = (int)K;
  p.epi = epi;
  p.partials = reinterpret_cast<float*>((uint8_t*)workspace + header);
  p.flags = reinterpret_cast<unsigned*>((uint8_t*)workspace + SK_FLAGS_OFFSET);
  MB_CHECK_CUDA(cudaFuncSetAttribute(gemm_streamk_grouped_kernel<MODE, TA>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
  MB_CHECK_CUDA(launch_pdl(gemm_streamk_grouped_kernel<MODE, TA>, dim3((unsigned)sms), dim3(TG_THREADS), (size_t)Cfg::kSmem, stream, map_a, maps, p, plan));
  return MB200_OK;
}

template <int MODE>
int launch_grouped(const void* a, int64_t rows_cap, int64_t K, int64_t N, const void* const* w_host, int E, int est_mtiles, int tile_rows,
                   const int32_t* plan, const EpiParams& epi, void* workspace, size_t workspace_bytes, size_t header, cudaStream_t stream) {
  int dev = 0, sms = 0;
  MB_CHECK_CUDA(cudaGetDevice(&dev));
  MB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  MB_CHECK_ARG(K % TG_BK == 0 && N % 32 == 0, "grouped gemm: K=%lld must be a multiple of 64, N=%lld of 32", (long long)K, (long long)N);
  if (tile_rows < 128 && streamk_eligible(tile_rows, N, K)) {
    if (tile_rows == 32) return launch_grouped_streamk<MODE, 32>(a, rows_cap, K, N, w_host, E, plan, epi, workspace, workspace_bytes, header, sms, stream);
    return launch_grouped_streamk<MODE, 64>(a, rows_cap, K, N, w_host, E, plan, epi, workspace, workspace_bytes, header, sms, stream);
  }
  if (tile_rows == 128) {
    // enough rows per expert for vertically adjacent tile pairs: the 2-CTA cluster kernel (W tile multicast, 2/3 of the L2 -> SM traffic)
    if (N % 256 == 0 && tcgen05_cluster_enabled() && rows_cap >= (int64_t)E * 512)
      return launch_grouped_bn<MODE, 256, 128, 2>(a, rows_cap, K, N, w_host, E, plan, epi, sms, stream);
    if (N % 256 == 0) return launch_grouped_bn<MODE, 256, 128>(a, rows_cap, K, N, w_host, E, plan, epi, sms, stream);
    if (N % 128 == 0) return launch_grouped_bn<MODE, 128, 128>(a, rows_cap, K, N, w_host, E, plan, epi, sms, stream);
    if (N % 64 == 0) return launch_grouped_bn<MODE, 64, 128>(a, rows_cap, K, N, w_host, E, plan, epi, sms, stream);
    return launch_grouped_bn<MODE, 32, 128>(a, rows_cap, K, N, w_host, E, plan, epi, sms, stream);
  }
  int best = 32;
  double best_score = -1.0;
  const int forced = tcgen05_forced_bn();
  const int cand[4] = {256, 128, 64, 32};
  for (int i = 0; i < 4; ++i) {
    const int bn = cand[i];
    if (N % bn != 0) continue;
    if (forced == bn) {
      best = bn;
      break;
    }
    const int64_t tiles = (int64_t)est_mtiles * (N / bn), rounds = (tiles + sms - 1) / sms;
    const double score = (double)tiles / (double)(rounds * sms);
    if (score > best_score + 0.02) {  // wider tiles (bigger TMA boxes) unless a narrower one fills the rounds clearly better
      best_score = score;
      best = bn;
    }
  }
  if (tile_rows == 64) {
    switch (best) {
      case 256: return launch_grouped_bn<MODE, 256, 64>(a, rows_cap, K, N, w_host, E, plan, epi, sms, stream);
      case 128: return launch_grouped_bn<MODE, 128, 64>(a, rows_cap, K, N, w_host, E, plan, epi, sms, stream);
      case 64: return launch_grouped_bn<MODE, 64, 64>(a, rows_cap, K, N, w_host, E, plan, epi, sms, stream);
      default: return launch_grouped_bn<MODE, 32, 64>(a, rows_cap, K, N, w_host, E, plan, epi, sms, stream);
    }
  }
  switch (best) {
    case 256: return launch_grouped_bn<MODE, 256, 32>(a, rows_cap, K, N, w_host, E, plan, epi, sms, stream);
    case 128: return launch_grouped_bn<MODE, 128, 32>(a, rows_cap, K, N, w_host, E, plan, epi, sms, stream);
    case 64: return launch_grouped_bn<MODE, 64, 32>(a, rows_cap, K, N, w_host, E, plan, epi, sms, stream);
    default: return launch_grouped_bn<MODE, 32, 32>(a, rows_cap, K, N, w_host, E, plan, epi, sms, stream);
  }
}

}  // namespace mb200
