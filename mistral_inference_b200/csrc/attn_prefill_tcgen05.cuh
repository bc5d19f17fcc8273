// First-prefill attention (varlen, causal, sliding window; every key comes from the new chunk) on tcgen05 / TMEM / TMA.
//
// Roofline: tensor pipe / MUFU (one exp2 per score).  CTA = (query head, 128-query tile, sequence); 10 warps:
//   warp 0    TMA producer: Q tile once, then K and V tiles of 128 keys into a 3-stage ring (128B-swizzled [128 x 64] boxes)
//   warp 1    MMA issuer: S[128 x 128] = Q K^T (both operands K-major from shared memory), and after the softmax
//             O[128 x 128] += P V with P read from TENSOR MEMORY (A operand) and V as an MN-major shared-memory operand
//   warps 2-9 softmax: two threads per query row (TMEM lane), 64 score columns each: tcgen05.ld the half row, mask (causal +
//             window + sequence end; edge tiles only), online softmax in fp32 (row maximum exchanged through shared memory),
//             P -> bf16 -> tcgen05.st back into TMEM, rescale O in TMEM when a maximum grew, final O / l -> bf16 -> global
// TMEM columns: S buffer 0 0..127 | S buffer 1 128..255 | O 256..383 | P 384..447 (bf16 pairs).  S is double buffered so the
// QK^T MMAs of tile t+1 run while the softmax warps work on tile t; each score is read from TMEM once.
// The mma.sync kernel (attn_prefill.cuh) remains for chunks that also read the ring (seqpos > 0) and for the cache-less mode.
#pragma once
#include "gemm_tcgen05.cuh"

namespace mb200 {

constexpr int FA_BM = 128, FA_BN = 128, FA_THREADS = 320, FA_STAGES = 3;  // warp 0 TMA, warp 1 MMA, warps 2..9 softmax
constexpr int FA_TILE_BYTES = 128 * kHeadDim * 2;  // one [128 x 128] bf16 tile = two swizzled [128 x 64] boxes = 32 KB
constexpr int FA_XCH_BYTES = 2 * 2 * FA_BM * 4;  // row max / row sum exchange between the two softmax threads of a row, double buffered
constexpr int FA_SMEM = FA_TILE_BYTES * (1 + 2 * FA_STAGES) + 128 /*barriers*/ + FA_XCH_BYTES;  // 231,552 of 232,448 bytes
constexpr int FA_COL_S = 0 /* two buffers: 0 and 128 */, FA_COL_O = 256, FA_COL_P = 384, FA_TMEM_COLS = 512;

struct FaParams {
  const int32_t* q_start;  // [B+1]
  bf16* out;               // [T, H*hd]
  int T, B, W, H, KV;
  float scale_log2;
};

// P (TMEM, A operand) x V (smem, MN-major B operand) and Q x K^T instruction descriptors: M = 128, N = 128, bf16 -> fp32
constexpr uint32_t kIdescQK = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FA_BN >> 3) << 17) | ((uint32_t)(FA_BM >> 4) << 24);
constexpr uint32_t kIdescPV = kIdescQK | (1u << 16);  // b_major = MN: V tile rows are keys (K), contiguous along head_dim (N)

// MN-major, 128B-swizzled operand made of [keys x 64 dims] boxes: 8-key groups 1024 B apart (SBO), 64-dim blocks one box apart (LBO)
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t box_bytes) {
  return (uint64_t)((smem_addr & 0x3ffff) >> 4) | ((uint64_t)(box_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
      "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(FA_THREADS, 1)
    attn_prefill_tcgen05_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                                const __grid_constant__ CUtensorMap map_v, const FaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];  // SW128 tiles want 1024-byte alignment; there is no room left to pad by hand
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + FA_TILE_BYTES;  // stage s: K at sKV + s*2*TILE, V right after
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FA_TILE_BYTES * (1 + 2 * FA_STAGES));
  uint64_t* q_full = bars;               // TMA -> MMA
  uint64_t* kv_full = bars + 1;          // [3] TMA -> MMA
  uint64_t* kv_empty = bars + 4;         // [3] MMA (commit) -> TMA
  uint64_t* s_full = bars + 7;           // [2] MMA (commit) -> softmax, per S buffer
  uint64_t* p_full = bars + 9;           // softmax (4 warps) -> MMA
  uint64_t* pv_done = bars + 10;         // MMA (commit) -> softmax: the PV product of a tile has finished reading P / writing O
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  float* xch = reinterpret_cast<float*>(smem + FA_TILE_BYTES * (1 + 2 * FA_STAGES) + 128);  // [2 buffers][2 halves][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.z, g = h / (p.H / p.KV);
  const int tok0 = p.q_start[b], s_len = p.q_start[b + 1] - tok0;
  // Heads vary fastest and the heavier (later) query tiles are dispatched first: longest-processing-time order for the causal
  // triangle, and the query heads of one KV group read the same K/V tiles at the same time (L2 locality).
  const int qt = (int)gridDim.y - 1 - (int)blockIdx.y;
  const int i0 = qt * FA_BM;
  if (i0 >= s_len) return;
  const int i_end = min(i0 + FA_BM, s_len);
  const int key_lo = max(0, i0 - p.W + 1), key_hi = i_end - 1;  // visible keys of this tile: [key_lo, key_hi]
  const int n_tiles = (key_hi - key_lo + FA_BN) / FA_BN;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < FA_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(p_full, 8);  // one arrive per softmax warp
    mbar_init(pv_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, FA_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, FA_TILE_BYTES);
      tma_load_2d(sQ, &map_q, q_full, h * kHeadDim, tok0 + i0);
      tma_load_2d(sQ + FA_TILE_BYTES / 2, &map_q, q_full, h * kHeadDim + 64, tok0 + i0);
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % FA_STAGES, par = (t / FA_STAGES) & 1;
        mbar_wait(&kv_empty[s], par ^ 1, 21, t);
        mbar_arrive_expect_tx(&kv_full[s], 2 * FA_TILE_BYTES);
        uint8_t* sk = sKV + s * 2 * FA_TILE_BYTES;
        const int row = tok0 + key_lo + t * FA_BN;
        tma_load_2d(sk, &map_k, &kv_full[s], g * kHeadDim, row);
        tma_load_2d(sk + FA_TILE_BYTES / 2, &map_k, &kv_full[s], g * kHeadDim + 64, row);
        tma_load_2d(sk + FA_TILE_BYTES, &map_v, &kv_full[s], g * kHeadDim, row);
        tma_load_2d(sk + FA_TILE_BYTES + FA_TILE_BYTES / 2, &map_v, &kv_full[s], g * kHeadDim + 64, row);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      mbar_wait(q_full, 0, 22, 0);
      const uint32_t q_addr = smem_u32(sQ);
      auto issue_s = [&](int t) {  // S buffer t&1 = Q K_t^T: 8 k-steps of 16 dims; dims 0..63 / 64..127 in two boxes 16 KB apart
        const int s = t % FA_STAGES, par = (t / FA_STAGES) & 1;
        mbar_wait(&kv_full[s], par, 23, t);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sKV + s * 2 * FA_TILE_BYTES);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (ks >> 2) * (FA_TILE_BYTES / 2) + (ks & 3) * 32;
          umma_bf16(tmem + FA_COL_S + (t & 1) * 128, umma_desc_sw128(q_addr + off), umma_desc_sw128(k_addr + off), kIdescQK, ks ? 1u : 0u);
        }
        umma_commit(&s_full[t & 1]);
      };
      issue_s(0);
      for (int t = 0; t < n_tiles; ++t) {
        if (t + 1 < n_tiles) issue_s(t + 1);  // runs on the tensor pipe while the softmax warps work on tile t
        // O += P V once the softmax warps have written P (and rescaled O)
        mbar_wait(p_full, t & 1, 24, t);
        tc_fence_after();
        const int s = t % FA_STAGES;
        const uint32_t v_addr = smem_u32(sKV + s * 2 * FA_TILE_BYTES) + FA_TILE_BYTES;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)  // 16 keys per step: P columns 8*ks (bf16 pairs), V rows 16*ks (128 B each)
          umma_bf16_ts(tmem + FA_COL_O, tmem + FA_COL_P + ks * 8, umma_desc_sw128_mn(v_addr + ks * 2048, FA_TILE_BYTES / 2), kIdescPV,
                       (t | ks) ? 1u : 0u);
        umma_commit(&kv_empty[s]);  // K/V stage free once S and PV of this tile have read it
        umma_commit(pv_done);       // P may be overwritten / O rescaled / (last tile) O read
      }
    }
  } else {
    // ================= softmax warps: two threads per query row =================
    // Warps 2..5 and 6..9 each cover the 128 TMEM lanes (a warp may touch lanes 32*(warp%4)..+31); the first group takes score
    // columns 0..63 of every tile, the second 64..127, and the same halves of O.  The row maximum is exchanged through shared
    // memory (one named barrier per tile); each thread keeps the partial row sum of its own columns.
    const int lane_base = (warp & 3) * 32;
    const int half = (warp - 2) >> 2;     // which 64 columns
    const int r = lane_base + lane;       // row in the tile = TMEM lane
    const int i = i0 + r;                 // local query index; position == i (first prefill: seqpos = 0)
    const bool row_valid = i < s_len;
    const uint32_t trow = tmem + ((uint32_t)lane_base << 16);
    float m_run = -1.0e30f, l_run = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      mbar_wait(&s_full[t & 1], (t >> 1) & 1, 25, t);
      tc_fence_after();
      const int j0 = key_lo + t * FA_BN + half * 64;
      const uint32_t scol = trow + FA_COL_S + (t & 1) * 128 + half * 64;
      // Only edge tiles need per-element masks (tile-uniform test): the causal diagonal, the window's lower edge, the ragged
      // end of the sequence.  Interior tiles take the straight-line path.
      const bool edge = (key_lo + t * FA_BN + FA_BN - 1 > i0) || (key_lo + t * FA_BN <= i0 + FA_BM - 1 - p.W) || (i0 + FA_BM > s_len);
      // the score row is read from TMEM ONCE and kept in registers for both passes
      uint32_t sv[2][32];
      float mraw = -3.0e38f;
      tmem_ld_32x32b_x32_nowait(scol, sv[0]);
      tmem_ld_32x32b_x32_nowait(scol + 32, sv[1]);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (edge) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int j = j0 + c * 32 + e;
            const bool ok = row_valid && j <= i && j > i - p.W;
            sv[c][e] = ok ? sv[c][e] : 0xff800000u;  // -inf: exp2 gives exactly 0 in pass 2
            mraw = fmaxf(mraw, ok ? __uint_as_float(sv[c][e]) : -3.0e38f);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) mraw = fmaxf(mraw, __uint_as_float(sv[c][e]));
        }
      }
      // exchange with the thread that holds the other 64 columns of this row (buffer t&1: its previous use was two tiles ago)
      float* xb = xch + (t & 1) * 2 * FA_BM;
      xb[half * FA_BM + r] = mraw;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mraw = fmaxf(mraw, xb[(half ^ 1) * FA_BM + r]);
      const float mx = fmaxf(m_run, mraw > -1.0e38f ? mraw * p.scale_log2 : -1.0e30f);
      const bool grew = mx > m_run;
      // P and O belong to the PV product of the previous tile until it has completed
      if (t > 0) {
        mbar_wait(pv_done, (t - 1) & 1, 27, t);
        tc_fence_after();
      }
      // rescale O (TMEM) only when some row of this warp raised its maximum (after the first tiles that is rare); both threads of
      // a row see the same maximum, so the two warps that share these rows take the same branch
      if (t > 0 && __any_sync(0xffffffffu, grew)) {
        const float corr = ex2_approx(m_run - mx);
        l_run *= corr;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(trow + FA_COL_O + half * 64 + c * 32, v);
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * corr);
          tmem_st_32x32b_x32(trow + FA_COL_O + half * 64 + c * 32, v);
        }
      }
      m_run = mx;
      // P = exp2(s * scale - m) as bf16 pairs -> TMEM (A operand of the PV product)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float e0 = ex2_approx(fmaf(__uint_as_float(sv[c][e]), p.scale_log2, -mx));
          const float e1 = ex2_approx(fmaf(__uint_as_float(sv[c][e + 1]), p.scale_log2, -mx));
          l_run += e0 + e1;
          pk[e >> 1] = pack2_rn(e0, e1);
        }
        // 16 packed columns per chunk of 32 keys
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
                         trow + FA_COL_P + half * 32 + c * 16),
                     "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]), "r"(pk[8]), "r"(pk[9]),
                     "r"(pk[10]), "r"(pk[11]), "r"(pk[12]), "r"(pk[13]), "r"(pk[14]), "r"(pk[15])
                     : "memory");
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // final: O / l -> bf16 -> global (this thread's half row: 128 contiguous bytes).  The row sum is the sum of the two partial
    // sums; the exchange uses the buffer the last tile did NOT use (its last readers passed the last tile's barrier).
    float* xb = xch + (n_tiles & 1) * 2 * FA_BM;
    xb[half * FA_BM + r] = l_run;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l_run += xb[(half ^ 1) * FA_BM + r];
    mbar_wait(pv_done, (n_tiles - 1) & 1, 26, 0);
    tc_fence_after();
    const float inv = row_valid ? 1.f / l_run : 0.f;
    bf16* dst = p.out + (int64_t)(tok0 + i) * p.H * kHeadDim + (int64_t)h * kHeadDim + half * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(trow + FA_COL_O + half * 64 + c * 32, v);
      if (row_valid) {
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          uint4 o;
          o.x = pack2_rn(__uint_as_float(v[e]) * inv, __uint_as_float(v[e + 1]) * inv);
          o.y = pack2_rn(__uint_as_float(v[e + 2]) * inv, __uint_as_float(v[e + 3]) * inv);
          o.z = pack2_rn(__uint_as_float(v[e + 4]) * inv, __uint_as_float(v[e + 5]) * inv);
          o.w = pack2_rn(__uint_as_float(v[e + 6]) * inv, __uint_as_float(v[e + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + c * 32 + e) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, FA_TMEM_COLS);
}

// [rows, cols] bf16 row-major, box = [128 rows x 64 cols], 128-byte swizzle
inline int make_tensor_map_rows(CUtensorMap* map, const void* base, int64_t rows, int64_t cols) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return fail(MB200_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  const cuuint32_t box[2] = {64, 128};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MB200_E_CUDA, "cuTensorMapEncodeTiled (attention) failed (%d)", (int)r);
  return MB200_OK;
}

inline bool tcgen05_attn_eligible(int64_t T, int64_t max_seqlen) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("MB200_ATTN");
    forced = (e != nullptr && e[0] == 'm') ? 1 : 0;  // MB200_ATTN=mma forces the mma.sync kernel
  }
  return !forced && T >= 128 && max_seqlen >= 128;
}

inline int launch_attn_prefill_tcgen05(const void* q, const void* k_new, const void* v_new, const int32_t* q_start, void* out, int64_t T, int64_t B,
                                       int64_t max_seqlen, int64_t W, int64_t H, int64_t KV, cudaStream_t stream) {
  CUtensorMap mq, mk, mv;
  int rc = make_tensor_map_rows(&mq, q, T, H * kHeadDim);
  if (rc) return rc;
  rc = make_tensor_map_rows(&mk, k_new, T, KV * kHeadDim);
  if (rc) return rc;
  rc = make_tensor_map_rows(&mv, v_new, T, KV * kHeadDim);
  if (rc) return rc;
  FaParams p;
  p.q_start = q_start;
  p.out = (bf16*)out;
  p.T = (int)T;
  p.B = (int)B;
  p.W = (int)W;
  p.H = (int)H;
  p.KV = (int)KV;
  p.scale_log2 = 0.08838834764831845f * 1.4426950408889634f;
  MB_CHECK_CUDA(cudaFuncSetAttribute(attn_prefill_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
  const dim3 grid((unsigned)H, (unsigned)ceil_div(max_seqlen, FA_BM), (unsigned)B);
  attn_prefill_tcgen05_kernel<<<grid, FA_THREADS, FA_SMEM, stream>>>(mq, mk, mv, p);
  MB_CHECK_LAUNCH("attn_prefill_tcgen05_kernel");
  return MB200_OK;
}

}  // namespace mb200
