// C[T, N] = A[T, K] * W[N, K]^T on the 5th-generation tensor cores: tcgen05.mma with TMEM accumulators, TMA-fed.
//
// Roofline: tensor pipe (2*T*N*K flops).  Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0   TMA producer: cp.async.bulk.tensor (128B-swizzled [128 x 64] A tile + [BN x 64] W tile per stage) -> 4- or 6-stage
//            ring.  For T >= 512 the CTAs run as clusters of two on vertically adjacent tiles and each fetches half of the
//            shared W tile, multicast into both shared memories (2/3 of the L2 -> SM traffic of the single-CTA kernel)
//   warp 1   MMA issuer: one elected thread issues tcgen05.mma (M = 128, N = BN, K = 16 x 4 per stage); tcgen05.commit
//            releases the shared-memory stage (in both CTAs of a pair) and, after the last k-block, publishes the accumulator
//   warps 2-5 epilogue: one thread per output row reads the fp32 accumulator out of TMEM 32 columns at a time (two TMEM
//            buffers, so the MMAs of the next tile overlap this one's epilogue; two register buffers, so the next tcgen05.ld
//            overlaps this chunk's arithmetic) and runs the row-chunk epilogues of epilogue.cuh (bf16 rounding, then RoPE +
//            ring scatter / SiLU*mul / residual add / fp32 logits) with 16-byte stores
// Both operands are K-major ([rows, K] row-major): the canonical TN GEMM, no transposes anywhere.
// Tiles are walked m-fastest so that the CTAs that share a W tile run together and hit it in L2.
#pragma once
#include <cuda.h>

#include <cstdlib>

#include <mutex>
#include <unordered_map>

#include "decode_megakernel.cuh"  // mbarrier helpers with the watchdog
#include "epilogue.cuh"
#include "gemm_mma.cuh"

namespace mb200 {

constexpr int TG_BM = 128, TG_BN = 256, TG_BK = 64;
constexpr int TG_THREADS = 192;  // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2..5 epilogue
constexpr int TG_A_BYTES = TG_BM * TG_BK * 2;
// The tile width BN is 256 (4 stages of 48 KB), 192 (4 stages of 40 KB) or 128 (6 stages of 32 KB).  256 has the least W traffic
// per flop; 128 is chosen by the launcher for problems too small to give every SM a 256-wide tile (and for N that is a multiple
// of 128 only); 192 only for N that is a multiple of 192 but not of 128.  Narrower tiles to even out the last round of the
// persistent schedule were measured and lose: L2 -> SM bandwidth per flop, not tile quantisation, is what binds at T = 4096.
// TA = rows of the A (token) box.  128 for prefill.  For small batches (decode with 5..64 tokens) the box is 32 or 64 rows:
// the MMA still has M = 128 and reads 16 KB from the A tile's base, i.e. it runs into whatever follows the tile in shared
// memory -- accumulator row i depends on A row i only, and rows >= T are never stored.  The stage shrinks to (TA + BN) * 128
// bytes, so the ring gets deep enough (16-20 stages) to keep a whole SM's share of HBM bandwidth in flight: those launches are
// weight-streaming GEMMs, bound by HBM, with BN chosen by the launcher to give every SM at most one (equal) tile per round.
template <int BN, int TA = 128>
struct TgCfg {
  static constexpr int kABytes = TA * TG_BK * 2;
  static constexpr int kBBytes = BN * TG_BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // The MMA always reads M = 128 rows (16 KB) from a stage's A base; with a short A box that runs on into the stage's W tile
  // (harmless: accumulator row i depends on A row i only) and, in the LAST stage, past it when the stage is smaller than 16 KB.
  // A full 16 KB of slack is more than the 20-24 KB stages of the stream-K kernel need, and it is what keeps its CTAs at 117.5 KB:
  // TWO of them do not fit an SM.  That is deliberate: with the exact slack (-DMB200_TIGHT_SLACK: 101.5 KB, the successor launch's
  // CTA sits beside the running one from its start) the Nemo-12B batch-32 step measured 7.48 ms against 6.85 ms on the same box
  // (call 22) -- the early CTA's ring fill and polling take bandwidth and issue slots from the launch that is on the critical path.
  // As it is, a successor CTA starts when the predecessor's CTA on that SM exits, i.e. during the predecessor's reduction tail.
#ifdef MB200_TIGHT_SLACK
  static constexpr int kSlack = (TA < 128 && kStageBytes < TG_A_BYTES) ? TG_A_BYTES - kStageBytes : 0;
#else
  static constexpr int kSlack = TA < 128 ? TG_A_BYTES : 0;
#endif
  static constexpr int kMaxStages = (227 * 1024 - 1024 - 512 - kSlack) / kStageBytes;
  // decode-sized variants: a ring of ~100 KB (80 KB of weights in flight per CTA, above the ~45-65 KB bandwidth-delay product of
  // one SM's HBM share) instead of the whole 227 KB: measured faster in a chain of dependent launches (scripts/bench_linear.py:
  // qkv 24.0 -> 21.3 us, down 39.5 -> 36.8 us; only the 1.3 GB lm head loses 5 %) -- a successor's CTA, which starts when the
  // predecessor's CTA on its SM exits, has its first ring filled sooner, and other decode kernels' CTAs fit beside it.
  static constexpr int kShortRingStages = (100 * 1024) / kStageBytes;
  static constexpr int kStages = (TA < 128 || BN < 128) ? (kShortRingStages < 3 ? 3 : (kShortRingStages > kMaxStages ? kMaxStages : kShortRingStages))
                                                        : (BN == 128 ? 6 : 4);
  static constexpr int kSmem = kStages * kStageBytes + kSlack + 1024 /*align*/ + 512 /*barriers*/;
  static constexpr int kTmemCols = 2 * BN <= 32 ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));  // 2 accumulators
  // Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
  static constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TG_BM >> 4) << 24);
};

struct TcGemmParams {
  int T, N, K;
  EpiParams epi;
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
// Same, delivered to the same shared-memory offset (and signalled on the same barrier offset) of every CTA in cta_mask.
__device__ __forceinline__ void tma_load_2d_multicast(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the same arrival on the barrier at this offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
        "=r"(v[31])
      : "r"(taddr));
}

__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  tmem_ld_32x32b_x32_nowait(taddr, v);
  tmem_wait_ld();
}

// Shared-memory matrix descriptor of a K-major, 128B-swizzled tile whose rows are 64 bf16 = 128 bytes:
// 8-row swizzle atoms of 1024 B (SBO), version 1 (sm_100), layout type 2 (SWIZZLE_128B).  (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3ffff) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// CL = 2: thread-block clusters of two CTAs that work on vertically adjacent tiles (same W columns, consecutive 128-row blocks).
// Each CTA fetches HALF of the shared [256 x 64] W tile and TMA-multicasts it into both shared memories, so the L2 -> SM traffic
// per CTA and k-block drops from 48 KB to 32 KB (at T = 4096 the single-CTA kernel pulls ~20 TB/s out of L2).  A stage may be
// refilled only when BOTH CTAs' MMAs have read it: the empty barriers count two commits, each multicast to the pair.
//
// GROUPED = true (mixture of experts, csrc/moe.cuh): the A rows are the (token, expert) pairs sorted by expert, every expert's
// segment padded to a multiple of TA rows; m tile i covers rows plan.tile_row0[i] .. + TA of expert plan.tile_expert[i], whose
// weight matrix has its own tensor map (map_w[expert]).  The number of m tiles is DEVICE data (no host sync after routing).
constexpr int MOE_PLAN_HEADER = 64;  // int32 words: [0] m tiles of this rank, [1] padded rows in total, [2] tile capacity, [6] tile pairs, [8..] segment starts
// after the header: tile_expert[cap], tile_row0[cap], pair_expert[cap], pair_info[cap] (cap = plan[2]).  A PAIR is two vertically
// adjacent m tiles of ONE expert (or a single last tile: bit 30 of pair_info clear), the unit of the 2-CTA cluster variant.
constexpr int MOE_PAIR_SECOND = 1 << 30;
constexpr int MOE_MAX_EXPERTS = 16;  // tensor maps travel as kernel parameters (128 B each)
struct MoeWeightMaps {
  CUtensorMap m[MOE_MAX_EXPERTS];
};

template <int MODE, int CL, int BN, int TA, bool GROUPED>
__device__ __forceinline__ void tc_gemm_body(const CUtensorMap& map_a, const CUtensorMap* map_w_base, const TcGemmParams& p, const int32_t* plan) {
  static_assert(TA == 128 || CL == 1, "small-batch variant is single-CTA");
  using Cfg = TgCfg<BN, TA>;
  constexpr int TG_STAGES = Cfg::kStages, TG_B_BYTES = Cfg::kBBytes, TG_STAGE_BYTES = Cfg::kStageBytes, TG_TMEM_COLS = Cfg::kTmemCols;
  constexpr int TG_A_BYTES = Cfg::kABytes;  // shadows the 128-row constant
  constexpr int TG_BN = BN;
  constexpr uint32_t kUmmaIdesc = Cfg::kIdesc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);  // SW128 wants 1024-B tiles
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + TG_STAGES * TG_STAGE_BYTES + Cfg::kSlack);
  uint64_t* empty = full + TG_STAGES;
  uint64_t* tmem_full = empty + TG_STAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;     // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int cta = (int)blockIdx.x / CL, n_cta = (int)gridDim.x / CL;  // cluster index / clusters in the grid
  // a cluster walks "super tiles" of CL vertically adjacent tiles; rows past T read as zeros (TMA) and are never stored
  // grouped: m units are the plan's tiles (single CTA) or tile pairs (cluster of two)
  const int num_m = GROUPED ? (CL > 1 ? plan[6] : plan[0]) : ((p.T + TG_BM - 1) / TG_BM + CL - 1) / CL;
  const int num_n = p.N / TG_BN, num_tiles = num_m * num_n, num_k = p.K / TG_BK;
  const int32_t* tile_expert = GROUPED ? plan + MOE_PLAN_HEADER + (CL > 1 ? 2 * plan[2] : 0) : nullptr;
  const int32_t* tile_row0 = GROUPED ? tile_expert + plan[2] : nullptr;
  // Tile order.  The persistent CTAs take tiles t = cta, cta + n_cta, ...: at any moment ~n_cta consecutive tile ids are in flight,
  // in lock step through K.  With few m units (T <= 4096: the whole A matrix fits the 126 MB L2) the walk is m-fastest: the CTAs
  // that share a W tile run together.  With many (a 32 x 1024-token prefill: A = 335 MB; a Mixtral prefill: 277 MB of gathered rows)
  // m-fastest makes every CTA stream its own A tile from DRAM once per n tile -- measured with ncu: 32.7 GB of DRAM reads for a
  // grouped gate/up GEMM whose operands total 2.2 GB, 4.9 TB/s, DRAM-bound at 1.15 PFLOP/s.  There the in-flight set is shaped as a
  // GM x GN block instead (GM m units share each W tile, GN n tiles share each A tile): DRAM traffic per tile drops ~4x.
  constexpr int kGM = CL > 1 ? 8 : 12, kGN = CL > 1 ? 9 : 12;
  auto tile_mn = [&](int tile, int& mu, int& nt) {
    if (num_m <= 16) {
      mu = tile % num_m;
      nt = tile / num_m;
      return;
    }
    const int per_row = kGN * num_m, rows_full = num_n / kGN;  // an "n-block row": all m units x kGN n tiles
    int nb, r, gn;
    if (tile < rows_full * per_row) {
      nb = tile / per_row;
      r = tile % per_row;
      gn = kGN;
    } else {
      nb = rows_full;
      r = tile - rows_full * per_row;
      gn = num_n - rows_full * kGN;
    }
    const int blk = kGM * gn, mb_full = num_m / kGM;
    int mblk, q, gm;
    if (r < mb_full * blk) {
      mblk = r / blk;
      q = r % blk;
      gm = kGM;
    } else {
      mblk = mb_full;
      q = r - mb_full * blk;
      gm = num_m - mb_full * kGM;
    }
    mu = mblk * kGM + q % gm;
    nt = nb * kGN + q / gm;
  };
  // first row of this CTA's m tile of unit `u` (cluster rank 1 takes the pair's second tile; a missing second tile is recomputed
  // from the first one's rows and not stored)
  auto grouped_m0 = [&](int u, bool& store) {
    const int info = tile_row0[u];
    store = true;
    if (CL == 1) return info;
    const int row0 = info & (MOE_PAIR_SECOND - 1);
    if (rank == 0) return row0;
    store = (info & MOE_PAIR_SECOND) != 0;
    return store ? row0 + TG_BM : row0;
  };
  const CUtensorMap& map_w = *map_w_base;

  if (threadIdx.x == 0) {
    for (int i = 0; i < TG_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], CL);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    if (!GROUPED) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_base_slot, TG_TMEM_COLS);
  tc_fence_before();
  if (CL > 1)
    cluster_sync_all();  // the peer's barriers must be initialised before any multicast copy or commit reaches them
  else
    __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = cta; tile < num_tiles; tile += n_cta) {
        bool store_unused;
        int mu, nt;
        tile_mn(tile, mu, nt);
        const int m0 = GROUPED ? grouped_m0(mu, store_unused) : (mu * CL + rank) * TG_BM, n0 = nt * TG_BN;
        const CUtensorMap* wmap = GROUPED ? map_w_base + tile_expert[mu] : map_w_base;
        for (int kb = 0; kb < num_k; ++kb, ++it) {
          const uint32_t s = it % TG_STAGES, par = (it / TG_STAGES) & 1;
          mbar_wait(&empty[s], par ^ 1, 11, it);
          mbar_arrive_expect_tx(&full[s], TG_STAGE_BYTES);  // A + both halves of W (the peer's half may land first: tx-count goes negative)
          uint8_t* sa = smem + s * TG_STAGE_BYTES;
          tma_load_2d(sa, &map_a, &full[s], kb * TG_BK, m0);
          if (CL > 1)
            tma_load_2d_multicast(sa + TG_A_BYTES + rank * (TG_B_BYTES / CL), wmap, &full[s], kb * TG_BK, n0 + rank * (TG_BN / CL),
                                  (uint16_t)((1u << CL) - 1));
          else
            tma_load_2d(sa + TG_A_BYTES, wmap, &full[s], kb * TG_BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      uint32_t it = 0, acc_it = 0;
      for (int tile = cta; tile < num_tiles; tile += n_cta, ++acc_it) {
        const uint32_t acc = acc_it & 1, acc_par = (acc_it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_par ^ 1, 12, acc_it);  // the epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TG_BN;
        for (int kb = 0; kb < num_k; ++kb, ++it) {
          const uint32_t s = it % TG_STAGES, par = (it / TG_STAGES) & 1;
          mbar_wait(&full[s], par, 13, it);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * TG_STAGE_BYTES);
          const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + TG_A_BYTES);
#pragma unroll
          for (int k = 0; k < TG_BK / 16; ++k)  // +32 bytes (= 2 in 16-byte units) per K = 16 step inside the 128-byte swizzled row
            umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kUmmaIdesc, (kb | k) ? 1u : 0u);
          if (CL > 1)  // frees the stage in both CTAs (each producer writes into both) when these MMAs have read it
            umma_commit_multicast(&empty[s], (uint16_t)((1u << CL) - 1));
          else
            umma_commit(&empty[s]);
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete
      }
    }
  } else {
    // ================= epilogue warps 2..5: TMEM lanes 32*(warp%4) .. +31 =================
    const int lane_base = (warp & 3) * 32;
    uint32_t acc_it = 0;
    for (int tile = cta; tile < num_tiles; tile += n_cta, ++acc_it) {
      bool store = true;
      int mu, nt;
      tile_mn(tile, mu, nt);
      const int m0 = GROUPED ? grouped_m0(mu, store) : (mu * CL + rank) * TG_BM, n0 = nt * TG_BN;
      const uint32_t acc = acc_it & 1, acc_par = (acc_it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_par, 14, acc_it);
      tc_fence_after();
      // accumulator rows >= TA were computed from whatever follows the short A box in shared memory: never stored
      const int t = (store && (TA == 128 || lane_base + lane < TA)) ? m0 + lane_base + lane : 0x7fffffff;
      // two register buffers: the TMEM load of chunk c+1 is in flight while chunk c goes through the epilogue
      const uint32_t trow = tmem_base + ((uint32_t)lane_base << 16) + acc * TG_BN;
      uint32_t va[32], vb[32];
      tmem_ld_32x32b_x32_nowait(trow, va);
      if constexpr (TG_BN == 32) {
        tmem_wait_ld();
        if (t < p.T) epi_chunk32<MODE>(p.epi, t, n0, va);
      } else {
#pragma unroll 1
        for (int c = 0; c < TG_BN / 32; c += 2) {
          tmem_wait_ld();
          tmem_ld_32x32b_x32_nowait(trow + (c + 1) * 32, vb);
          if (t < p.T) epi_chunk32<MODE>(p.epi, t, n0 + c * 32, va);
          tmem_wait_ld();
          if (c + 2 < TG_BN / 32) tmem_ld_32x32b_x32_nowait(trow + (c + 2) * 32, va);
          if (t < p.T) epi_chunk32<MODE>(p.epi, t, n0 + (c + 1) * 32, vb);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }
  tc_fence_before();
  if (CL > 1)
    cluster_sync_all();  // the peer's last commits still arrive on this CTA's barriers
  else
    __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TG_TMEM_COLS);
}

template <int MODE, int CL, int BN, int TA = 128>
__global__ void __launch_bounds__(TG_THREADS, 1)
    gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const TcGemmParams p) {
  tc_gemm_body<MODE, CL, BN, TA, false>(map_a, &map_w, p, nullptr);
}

template <int MODE, int CL, int BN, int TA>
__global__ void __launch_bounds__(TG_THREADS, 1)
    gemm_tcgen05_grouped_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ MoeWeightMaps maps_w, const TcGemmParams p,
                                const int32_t* __restrict__ plan) {
  tc_gemm_body<MODE, CL, BN, TA, true>(map_a, maps_w.m, p, plan);
}

// ---- host: tensor maps (driver API through the runtime's entry-point lookup, no libcuda link dependency) ----
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// [rows, K] bf16 row-major, box = [box_rows x 64] elements, 128-byte swizzle, out-of-range rows read as zero
inline int make_tensor_map_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t K, int box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return fail(MB200_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {(cuuint32_t)TG_BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MB200_E_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld K=%lld", (int)r, (long long)rows, (long long)K);
  return MB200_OK;
}

// MB200_GEMM=mma forces the mma.sync kernel (A/B comparisons, debugging)
inline bool tcgen05_gemm_eligible(int64_t T, int64_t N, int64_t K) {
  static int forced_mma = -1;
  if (forced_mma < 0) {
    const char* e = getenv("MB200_GEMM");
    forced_mma = (e != nullptr && e[0] == 'm') ? 1 : 0;
  }
  if (forced_mma || K % TG_BK != 0) return false;
  if (T >= TG_BM) return N % 128 == 0 || N % 192 == 0;
  return N % 32 == 0;  // small-batch weight-streaming variant (T <= 64 uses short A boxes; 65..127 the 128-row box)
}

// MB200_GEMM_CLUSTER=0 forces the single-CTA kernel; MB200_GEMM_BN=128|192|256 forces the tile width.  Read at every launch
// (a getenv is ~100 ns) so the tests can switch variants inside one process.
inline bool tcgen05_cluster_enabled() {
  const char* e = getenv("MB200_GEMM_CLUSTER");
  return !(e != nullptr && e[0] == '0');
}
inline int tcgen05_forced_bn() {
  const char* e = getenv("MB200_GEMM_BN");
  return e != nullptr ? atoi(e) : 0;
}

template <int MODE, int BN>
int launch_gemm_tcgen05_bn(const GemmParams& g, bool pair, int sms, cudaStream_t stream) {
  using Cfg = TgCfg<BN>;
  CUtensorMap map_a, map_w;
  int rc = make_tensor_map_2d(&map_a, g.a, g.T, g.K, TG_BM);
  if (rc) return rc;
  rc = make_tensor_map_2d(&map_w, g.w, g.N, g.K, pair ? BN / 2 : BN);
  if (rc) return rc;
  TcGemmParams p;
  p.T = g.T;
  p.N = g.N;
  p.K = g.K;
  p.epi = g.epi;
  if (pair) {
    const int supers = ceil_div(ceil_div(g.T, TG_BM), 2) * (g.N / BN), pairs = sms / 2;
    MB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<MODE, 2, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2u * (unsigned)(supers < pairs ? supers : pairs));
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
    MB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<MODE, 2, BN>, map_a, map_w, p));
    MB_CHECK_LAUNCH("gemm_tcgen05_kernel<cluster 2>");
    return MB200_OK;
  }
  const int tiles = ceil_div(g.T, TG_BM) * (g.N / BN);
  MB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<MODE, 1, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
  gemm_tcgen05_kernel<MODE, 1, BN><<<tiles < sms ? tiles : sms, TG_THREADS, Cfg::kSmem, stream>>>(map_a, map_w, p);
  MB_CHECK_LAUNCH("gemm_tcgen05_kernel");
  return MB200_OK;
}

// ---- small-batch (decode, T < 128) launcher: weight-streaming, HBM-bound ------------------------------------------------------
// One M tile; the N tiles are dealt to persistent CTAs.  BN is chosen so that the rounds of the persistent schedule are as full
// as possible (every CTA streams the same number of weight rows), preferring wider tiles (bigger TMA boxes) on ties.
inline int tcgen05_small_bn(int64_t N, int sms) {
  const int forced = tcgen05_forced_bn();
  if ((forced == 32 || forced == 64 || forced == 128 || forced == 256) && N % forced == 0) return forced;
  int best = 0;
  double best_score = -1.0;
  const int cand[4] = {256, 128, 64, 32};
  for (int i = 0; i < 4; ++i) {
    const int bn = cand[i];
    if (N % bn != 0) continue;
    const int64_t tiles = N / bn, rounds = (tiles + sms - 1) / sms;
    double score = (double)tiles / (double)(rounds * sms);
    if (tiles < sms / 2) score *= 0.5;  // too few SMs pulling: per-SM ingest, not HBM, would bind
    if (score > best_score + 1e-9) {
      best_score = score;
      best = bn;
    }
  }
  return best;
}

template <int MODE, int BN, int TA>
int launch_gemm_tcgen05_small_bn(const GemmParams& g, int sms, cudaStream_t stream) {
  using Cfg = TgCfg<BN, TA>;
  CUtensorMap map_a, map_w;
  int rc = make_tensor_map_2d(&map_a, g.a, g.T, g.K, TA);
  if (rc) return rc;
  rc = make_tensor_map_2d(&map_w, g.w, g.N, g.K, BN);
  if (rc) return rc;
  TcGemmParams p;
  p.T = g.T;
  p.N = g.N;
  p.K = g.K;
  p.epi = g.epi;
  const int tiles = g.N / BN;
  MB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<MODE, 1, BN, TA>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
  gemm_tcgen05_kernel<MODE, 1, BN, TA><<<tiles < sms ? tiles : sms, TG_THREADS, Cfg::kSmem, stream>>>(map_a, map_w, p);
  MB_CHECK_LAUNCH("gemm_tcgen05_kernel<small batch>");
  return MB200_OK;
}

template <int MODE, int TA>
int launch_gemm_tcgen05_small_ta(const GemmParams& g, int sms, cudaStream_t stream) {
  switch (tcgen05_small_bn(g.N, sms)) {
    case 256: return launch_gemm_tcgen05_small_bn<MODE, 256, TA>(g, sms, stream);
    case 128: return launch_gemm_tcgen05_small_bn<MODE, 128, TA>(g, sms, stream);
    case 64: return launch_gemm_tcgen05_small_bn<MODE, 64, TA>(g, sms, stream);
    case 32: return launch_gemm_tcgen05_small_bn<MODE, 32, TA>(g, sms, stream);
    default: return fail(MB200_E_INVALID, "small-batch GEMM: N=%d is not a multiple of 32", g.N);
  }
}

template <int MODE>
int launch_gemm_tcgen05(const GemmParams& g, cudaStream_t stream) {
  const bool pair = tcgen05_cluster_enabled() && g.T >= 4 * TG_BM;
  int dev = 0, sms = 0;
  MB_CHECK_CUDA(cudaGetDevice(&dev));
  MB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (g.T <= 32) return launch_gemm_tcgen05_small_ta<MODE, 32>(g, sms, stream);
  if (g.T <= 64) return launch_gemm_tcgen05_small_ta<MODE, 64>(g, sms, stream);
  if (g.T < TG_BM) return launch_gemm_tcgen05_small_ta<MODE, 128>(g, sms, stream);
  // 128-wide tiles only when 256-wide ones cannot fill the machine once (small T or N).  Measured at T = 4096: narrowing the
  // N = 4096 GEMMs (512 tiles = 3.46 rounds -> 1024 tiles = 6.9 rounds) makes them SLOWER (wo+w2 237 -> 343 us on average):
  // per flop the narrow tile pulls 1.5x the bytes out of L2, and that, not tile quantisation, is the binding limit there.
  const int units = pair ? sms / 2 : sms, m_units = pair ? ceil_div(ceil_div(g.T, TG_BM), 2) : ceil_div(g.T, TG_BM);
  int bn = 256;
  if (g.N % 256 != 0 || (int64_t)m_units * (g.N / 256) < units) {
    bn = g.N % 128 == 0 ? 128 : 192;
  }
  // 192-wide tiles (N = 6144: 5.19 rounds of 256 -> 6.92 rounds at 3/4 of the cost) were measured too: qkv+RoPE at T = 4096
  // 156 -> 169 us, slower for the same reason.  They stay available for N that is a multiple of 192 only and for the tests.
  const int forced = tcgen05_forced_bn();
  if ((forced == 128 || forced == 192 || forced == 256) && g.N % forced == 0) bn = forced;
  if (bn == 192) return launch_gemm_tcgen05_bn<MODE, 192>(g, pair, sms, stream);
  const bool narrow = bn == 128;
  return narrow ? launch_gemm_tcgen05_bn<MODE, 128>(g, pair, sms, stream) : launch_gemm_tcgen05_bn<MODE, 256>(g, pair, sms, stream);
}

}  // namespace mb200
