// Weight-streaming GEMM for decode-sized batches (5 <= T <= 128 tokens):  C[T, N] = A[T, K] * W[N, K]^T, bound by HBM.
//
// At these sizes the step reads every weight byte once and does ~2T flop per byte: the kernel's only job is to keep all 148 SMs
// pulling their share of W at the HBM rate.  Two measured facts shape it (B200, Nemo-12B shapes, batch 32):
//   * tile width: a W stage of [128 rows x 64] = 16 KB streams at the full HBM rate (lm head: 1.34 GB in 205 us), [64 x 64] gets
//     ~40 % and [32 x 64] ~26 %: per stage the one-thread TMA producer and the one-thread MMA issuer each spend a few hundred
//     cycles, so a stage must carry >= 16 KB.  Tiles are therefore always 128 wide;
//   * N / 128 tiles do not divide over 148 SMs (wo of Nemo: 40 tiles; gate/up: 224 = 1.5 rounds), so the work is cut stream-K
//     style instead: the (tile, k-block) units of the whole problem are one sequence, and CTA c takes the c-th contiguous 1/G of
//     it -- every SM streams the same number of bytes (+- one 16 KB stage) in one pass, whatever N and K are.
// A CTA's range covers the tail of one tile, some whole tiles and the head of another (or, when tiles are longer than ranges,
// a piece from the middle of one).  Partial tiles are reduced deterministically: a contributor that does not hold the tile's
// first k-block writes its fp32 accumulator to its own slot of the workspace and raises its flag -- always as the FIRST segment
// of its range; the owner of the first k-block -- always as the LAST segment of its range -- waits for the flags, sums the slots
// in ascending k order (all 128 epilogue threads, up to four contributors per L2 round trip), adds its own accumulator and runs
// the epilogue.  No atomics on data, no host-visible state: flags are consumed (reset) by their single reader.  The partition
// and the flag protocol are model-checked on the CPU in tests/test_streamk_protocol.py; scripts/trace_streamk.py (tracing build)
// prints the time line of a chain of launches.
//
// Structure per CTA is that of gemm_tcgen05.cuh: warp 0 TMA producer ([TA x 64] A box + [128 x 64] W box per stage, deep ring),
// warp 1 tcgen05.mma issuer (M = 128 -- rows >= TA read past the short A box and are never stored --, N = 128, fp32 accumulators
// double-buffered in TMEM), warps 2-5 epilogue (epilogue.cuh row-chunk epilogues).  GROUPED: the mixture-of-experts variant -- m
// tiles (expert segments of <= TA rows) come from the device-side plan of csrc/moe.cuh, each with its own weight tensor map.
#pragma once
#include "gemm_tcgen05.cuh"

namespace mb200 {

constexpr int SK_BN = 128;
constexpr int SK_MAX_CTAS = 160;
constexpr size_t SK_PARTIAL_BYTES = (size_t)SK_MAX_CTAS * 128 * SK_BN * sizeof(float);  // one [128 x 128] fp32 slot per CTA
constexpr size_t SK_FLAGS_OFFSET = 24576;  // inside the zero-initialised workspace header: uint32 flags[SK_MAX_CTAS]

// -DMB200_SK_TRACE: every CTA of the dense kernel stamps %globaltimer / %clock64 at eight points of its life into a device ring
// (scripts/trace_streamk.py reads it through mb200_debug_sk_trace) -- how the microseconds between dependent launches are spent.
#ifdef MB200_SK_TRACE
constexpr int SK_TRACE_LAUNCHES = 64, SK_TRACE_POINTS = 8;
__device__ unsigned long long sk_trace_buf[SK_TRACE_LAUNCHES][SK_MAX_CTAS][SK_TRACE_POINTS][2];
__device__ unsigned sk_trace_count;
__device__ __forceinline__ void sk_stamp(unsigned launch, int point) {
  unsigned long long g, c;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(c));
  sk_trace_buf[launch % SK_TRACE_LAUNCHES][blockIdx.x][point][0] = g;
  sk_trace_buf[launch % SK_TRACE_LAUNCHES][blockIdx.x][point][1] = c;
}
#define SK_STAMP(point) sk_stamp(*trace_launch, point)
#else
#define SK_STAMP(point) ((void)0)
#endif

struct SkParams {
  int T, N, K;
  EpiParams epi;
  float* partials;  // [gridDim][TA][128] fp32
  unsigned* flags;  // [gridDim], zero between launches
};

template <int MODE, int TA, bool GROUPED>
__device__ __forceinline__ void sk_gemm_body(const CUtensorMap& map_a, const CUtensorMap* map_w_base, const SkParams& p, const int32_t* plan) {
  using Cfg = TgCfg<SK_BN, TA>;
  constexpr int STAGES = Cfg::kStages, STAGE_BYTES = Cfg::kStageBytes, A_BYTES = Cfg::kABytes, TMEM_COLS = Cfg::kTmemCols;
  constexpr uint32_t kIdesc = Cfg::kIdesc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + Cfg::kSlack);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;  // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (int)gridDim.x, cta = (int)blockIdx.x;
  if (GROUPED) pdl_wait();  // the routing plan (m tiles) is the preceding kernels' output
  const int num_m = GROUPED ? plan[0] : 1, num_n = p.N / SK_BN, num_k = p.K / TG_BK;
  const int32_t* tile_expert = GROUPED ? plan + MOE_PLAN_HEADER : nullptr;
  const int32_t* tile_row0 = GROUPED ? plan + MOE_PLAN_HEADER + plan[2] : nullptr;
  // unit u = (tile, k-block) = (u / num_k, u % num_k); CTA c owns units [first(c), first(c + 1))
  const long long total = (long long)num_m * num_n * num_k;
  const long long per = total / G, rem = total % G;
  auto first = [&](int c) { return (long long)c * per + (c < rem ? c : rem); };
  const long long u_begin = first(cta), u_end = first(cta + 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_base_slot, TMEM_COLS);
#ifdef MB200_SK_TRACE
  __shared__ unsigned trace_launch_slot;
  if (threadIdx.x == 96) trace_launch_slot = atomicAdd(&sk_trace_count, 1u) / gridDim.x;  // all CTAs of a launch start before any of the next
  volatile unsigned* trace_launch = &trace_launch_slot;
#endif
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  if (threadIdx.x == 0) {
    SK_STAMP(0);  // CTA set up
    pdl_trigger();
  }

  if (warp == 0) {
    // ================= TMA producer =================
    // Iteration `it` handles unit u_begin + it.  Weights are never written by any kernel: the W tiles of the first ring are
    // requested BEFORE the programmatic-dependent-launch wait (they stream in while the previous kernel drains); the A tiles of
    // those stages, which the previous kernel produced, follow after it.
    if (lane == 0) {
      const uint32_t n_it = (uint32_t)(u_end - u_begin);
      auto issue = [&](uint32_t it, bool do_a, bool do_w) {
        const uint32_t u = (uint32_t)u_begin + it;
        const int tile = (int)(u / (uint32_t)num_k), kb = (int)(u % (uint32_t)num_k);
        const uint32_t s = it % STAGES;
        uint8_t* sa = smem + s * STAGE_BYTES;
        if (do_w) {
          const int n0 = (tile / num_m) * SK_BN;
          const CUtensorMap* wmap = GROUPED ? map_w_base + tile_expert[tile % num_m] : map_w_base;
          tma_load_2d(sa + A_BYTES, wmap, &full[s], kb * TG_BK, n0);
        }
        if (do_a) tma_load_2d(sa, &map_a, &full[s], kb * TG_BK, GROUPED ? tile_row0[tile % num_m] : 0);
      };
      const uint32_t head = GROUPED ? 0u : (n_it < (uint32_t)STAGES ? n_it : (uint32_t)STAGES);  // (grouped: the tile list itself is the previous kernel's output)
      for (uint32_t it = 0; it < head; ++it) {
        mbar_arrive_expect_tx(&full[it % STAGES], STAGE_BYTES);  // first lap: every slot is free
        issue(it, false, true);
      }
      pdl_wait();
      SK_STAMP(1);  // predecessor complete: A tiles may be requested
      for (uint32_t it = 0; it < head; ++it) issue(it, true, false);
      for (uint32_t it = head; it < n_it; ++it) {
        const uint32_t s = it % STAGES, par = (it / STAGES) & 1;
        mbar_wait(&empty[s], par ^ 1, 21, it);
        mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
        issue(it, true, true);
      }
      SK_STAMP(3);  // last tile requested
    }
  } else if (warp == 1) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      uint32_t it = 0, seg = 0;
      for (long long u = u_begin; u < u_end; ++seg) {
        const int kb0 = (int)(u % num_k);
        const int kb1 = (int)min((long long)num_k, kb0 + (u_end - u));
        const uint32_t acc = seg & 1, acc_par = (seg >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_par ^ 1, 22, seg);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * SK_BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t s = it % STAGES, par = (it / STAGES) & 1;
          mbar_wait(&full[s], par, 23, it);
#ifdef MB200_SK_TRACE
          if (it == 0) SK_STAMP(2);  // first stage landed: first MMA
#endif
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + A_BYTES);
#pragma unroll
          for (int k = 0; k < TG_BK / 16; ++k) umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc, (kb > kb0 || k) ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        umma_commit(&tmem_full[acc]);
        u += kb1 - kb0;
      }
      SK_STAMP(4);  // last MMA issued
    }
  } else {
    // ================= epilogue warps 2..5: TMEM lanes 32 * (warp % 4) .. + 31 =================
    const int lane_base = (warp & 3) * 32;
    if (!GROUPED) pdl_wait();  // stores below must not pass the predecessor's reads of the same buffers (returns at once when it is done)
    const bool row_ok = TA == 128 || lane_base + lane < TA;  // accumulator rows >= TA come from beyond the short A box
    const int etid = (int)threadIdx.x - 64;                  // 0..127 among the epilogue threads
    uint32_t seg = 0;
    for (long long u = u_begin; u < u_end; ++seg) {
      const int tile = (int)(u / num_k), kb0 = (int)(u % num_k);
      const int kb1 = (int)min((long long)num_k, kb0 + (u_end - u));
      const int m0 = GROUPED ? tile_row0[tile % num_m] : 0, n0 = (tile / num_m) * SK_BN;
      const uint32_t acc = seg & 1, acc_par = (seg >> 1) & 1;
      // Split tiles.  A contributor that does not own the tile's first k-block parks its fp32 accumulator in its workspace slot;
      // the owner adds the slots in ascending k order.  The slots are summed by ALL 128 epilogue threads, 32 rows per pass: thread
      // (r, cg) takes row r, columns 32 cg .. + 31 of every contributor, two contributors per round trip to L2.  (One thread per
      // row walking 4 chunks x n contributors was a chain of ~12 dependent L2 round trips: 6-8 us at the end of every launch.)
      const bool owner = kb0 == 0 && kb1 != num_k;
      int last = cta;
      float sum[32];
      auto sum_pass = [&](int ps) {
        const int r = etid & 31, cg = etid >> 5, n_o = last - cta;
        const size_t slot = (size_t)TA * SK_BN;
        const uint4* src = reinterpret_cast<const uint4*>(p.partials + ((size_t)(cta + 1) * TA + ps * 32 + r) * SK_BN + cg * 32);
        // four contributors (usually all of them) per round trip: the 8 column groups are independent, so their loads overlap;
        // the sum runs left to right over the contributors, i.e. in ascending k order
#pragma unroll 1
        for (int ob = 0; ob < n_o; ob += 4) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            uint4 w[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
              if (ob + jj < n_o) w[jj] = __ldcg(src + (size_t)(ob + jj) * (slot / 4) + q);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              if (ob + jj < n_o) {
                if (ob + jj == 0) {
                  sum[4 * q] = __uint_as_float(w[jj].x), sum[4 * q + 1] = __uint_as_float(w[jj].y), sum[4 * q + 2] = __uint_as_float(w[jj].z), sum[4 * q + 3] = __uint_as_float(w[jj].w);
                } else {
                  sum[4 * q] += __uint_as_float(w[jj].x), sum[4 * q + 1] += __uint_as_float(w[jj].y), sum[4 * q + 2] += __uint_as_float(w[jj].z), sum[4 * q + 3] += __uint_as_float(w[jj].w);
                }
              }
            }
          }
        }
      };
      if (owner) {
        // Contributors holding the tile's tail did it FIRST in their ranges; those whose whole range lies inside this tile finish
        // together with this CTA.  The flags are polled in parallel (one thread per contributor) and the first 32 rows are summed
        // before this CTA's own accumulator is waited for.
        const long long tile_end = (long long)(tile + 1) * num_k;
        while (last + 1 < G && first(last + 1) < tile_end) ++last;
        for (int c = cta + 1 + etid; c <= last; c += 128) {
          unsigned spins = 0;
          while (ld_acquire_u32(p.flags + c) == 0u) {
            if (++spins == MB200_WATCHDOG_SPINS) {
              printf("[mb200 watchdog] stream-K block %d waits for the partial of block %d (tile %d)\n", cta, c, tile);
              __trap();
            }
          }
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        sum_pass(0);
      }
      const int t = row_ok ? m0 + lane_base + lane : 0x7fffffff;
      if (kb0 == 0 && t < p.T) epi_prefetch128<MODE>(p.epi, t, n0);  // this segment ends in an epilogue: its operands, while the MMAs run
      mbar_wait(&tmem_full[acc], acc_par, 24, seg);
#ifdef MB200_SK_TRACE
      if (threadIdx.x == 64 && u + (kb1 - kb0) >= u_end) SK_STAMP(5);  // last accumulator complete
#endif
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)lane_base << 16) + acc * SK_BN;
      uint32_t v[32];
      if (kb0 != 0) {
        // ---- contributor: park the fp32 accumulator in this CTA's slot, then raise the flag ----
        if (lane_base < TA) {
          float* mine = p.partials + ((size_t)cta * TA + lane_base + lane) * SK_BN;
#pragma unroll 1
          for (int c = 0; c < SK_BN / 32; ++c) {
            tmem_ld_32x32b_x32(trow + c * 32, v);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              reinterpret_cast<uint4*>(mine + c * 32)[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        }
        __threadfence();
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (etid == 0) st_release_u32(p.flags + cta, 1u);
      } else if (owner) {
        // ---- owner: park the sums in shared memory (this is the CTA's last segment: every MMA has completed, the ring is
        // free); the warp that owns the pass's TMEM lanes adds its accumulator and runs the epilogue ----
        constexpr int SROW = SK_BN + 4;  // floats: a 528-byte row stride keeps the 16-byte accesses of both sides conflict-free
        float* stage = reinterpret_cast<float*>(smem);
#pragma unroll 1
        for (int ps = 0; ps < TA / 32; ++ps) {
          if (ps > 0) sum_pass(ps);
          float4* dst = reinterpret_cast<float4*>(stage + (etid & 31) * SROW + (etid >> 5) * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) dst[q] = make_float4(sum[4 * q], sum[4 * q + 1], sum[4 * q + 2], sum[4 * q + 3]);
          asm volatile("bar.sync 2, 128;" ::: "memory");
          if (lane_base == ps * 32) {  // warp-uniform: this warp's TMEM lanes are the rows of this pass
#pragma unroll 1
            for (int c = 0; c < SK_BN / 32; ++c) {
              tmem_ld_32x32b_x32(trow + c * 32, v);
              const float4* rest = reinterpret_cast<const float4*>(stage + lane * SROW + c * 32);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 w = rest[q];
                v[4 * q] = __float_as_uint(__uint_as_float(v[4 * q]) + w.x);
                v[4 * q + 1] = __float_as_uint(__uint_as_float(v[4 * q + 1]) + w.y);
                v[4 * q + 2] = __float_as_uint(__uint_as_float(v[4 * q + 2]) + w.z);
                v[4 * q + 3] = __float_as_uint(__uint_as_float(v[4 * q + 3]) + w.w);
              }
              if (t < p.T) epi_chunk32<MODE>(p.epi, t, n0 + c * 32, v);
            }
          }
          if (ps + 1 < TA / 32) asm volatile("bar.sync 2, 128;" ::: "memory");  // the next pass reuses the staging rows
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (etid == 0)
          for (int c = cta + 1; c <= last; ++c) p.flags[c] = 0u;  // consumed: ready for the next launch
      } else {
        // ---- whole tile in this CTA ----
#pragma unroll 1
        for (int c = 0; c < SK_BN / 32; ++c) {
          tmem_ld_32x32b_x32(trow + c * 32, v);
          if (t < p.T) epi_chunk32<MODE>(p.epi, t, n0 + c * 32, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      u += kb1 - kb0;
    }
#ifdef MB200_SK_TRACE
    if (threadIdx.x == 64) SK_STAMP(6);  // epilogues done
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
#ifdef MB200_SK_TRACE
  if (threadIdx.x == 0) SK_STAMP(7);  // exit
#endif
}

template <int MODE, int TA>
__global__ void __launch_bounds__(TG_THREADS, 2)  // <= 168 registers: the reduction's loads must not grow the footprint other decode kernels' CTAs share the SM with
    gemm_streamk_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const SkParams p) {
  sk_gemm_body<MODE, TA, false>(map_a, &map_w, p, nullptr);
}

template <int MODE, int TA>
__global__ void __launch_bounds__(TG_THREADS, 2)
    gemm_streamk_grouped_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ MoeWeightMaps maps_w, const SkParams p,
                                const int32_t* __restrict__ plan) {
  sk_gemm_body<MODE, TA, true>(map_a, maps_w.m, p, plan);
}

inline bool streamk_eligible(int64_t T, int64_t N, int64_t K) {
  const char* e = getenv("MB200_STREAMK");
  if (e != nullptr && e[0] == '0') return false;
  return T >= 1 && T <= 128 && N % SK_BN == 0 && K % TG_BK == 0;
}

// workspace: partial slots at `ws + header`, flags in the header (both overlap scratch of other, stream-ordered entry points)
template <int MODE, int TA>
int launch_streamk_ta(const GemmParams& g, void* workspace, size_t workspace_bytes, size_t header, cudaStream_t stream) {
  using Cfg = TgCfg<SK_BN, TA>;
  int dev = 0, sms = 0;
  MB_CHECK_CUDA(cudaGetDevice(&dev));
  MB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (sms > SK_MAX_CTAS) sms = SK_MAX_CTAS;
  if (workspace == nullptr || workspace_bytes < header + SK_PARTIAL_BYTES) return fail(MB200_E_WORKSPACE, "stream-K gemm: workspace %zu < %zu", workspace_bytes, header + SK_PARTIAL_BYTES);
  CUtensorMap map_a, map_w;
  int rc = make_tensor_map_2d(&map_a, g.a, g.T, g.K, TA);
  if (rc) return rc;
  rc = make_tensor_map_2d(&map_w, g.w, g.N, g.K, SK_BN);
  if (rc) return rc;
  SkParams p;
  p.T = g.T;
  p.N = g.N;
  p.K = g.K;
  p.epi = g.epi;
  p.partials = reinterpret_cast<float*>((uint8_t*)workspace + header);
  p.flags = reinterpret_cast<unsigned*>((uint8_t*)workspace + SK_FLAGS_OFFSET);
  const long long units = (long long)(g.N / SK_BN) * (g.K / TG_BK);
  const int grid = (int)(units < sms ? units : sms);
  MB_CHECK_CUDA(cudaFuncSetAttribute(gemm_streamk_kernel<MODE, TA>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
  MB_CHECK_CUDA(launch_pdl(gemm_streamk_kernel<MODE, TA>, dim3((unsigned)grid), dim3(TG_THREADS), (size_t)Cfg::kSmem, stream, map_a, map_w, p));
  return MB200_OK;
}

template <int MODE>
int launch_streamk(const GemmParams& g, void* workspace, size_t workspace_bytes, size_t header, cudaStream_t stream) {
  if (g.T <= 32) return launch_streamk_ta<MODE, 32>(g, workspace, workspace_bytes, header, stream);
  if (g.T <= 64) return launch_streamk_ta<MODE, 64>(g, workspace, workspace_bytes, header, stream);
  return launch_streamk_ta<MODE, 128>(g, workspace, workspace_bytes, header, stream);
}

}  // namespace mb200
