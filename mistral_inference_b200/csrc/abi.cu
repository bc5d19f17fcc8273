// libmb200.so -- the C ABI declared in include/mistral_b200.h.  Argument checking + kernel dispatch only.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "attn_decode.cuh"
#include "attn_decode_tma.cuh"
#include "attn_prefill.cuh"
#include "attn_prefill_tcgen05.cuh"
#include "decode_megakernel.cuh"
#include "elementwise.cuh"
#include "gemm_mma.cuh"
#include "gemm_streamk.cuh"
#include "gemm_tcgen05.cuh"
#include "moe.cuh"
#include "sampling.cuh"
#include "skinny_linear.cuh"

namespace mb200 {
thread_local char g_err[512] = "";
static unsigned long long* g_mk_prof = nullptr;
static unsigned long long* g_mk_prof_bar = nullptr;  // debug: decode megakernel phase timeline buffer (device)

constexpr size_t kWsHeader = 64 * 1024;  // persistent, zero-initialised by the caller once: self-resetting counters
inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static int run_rmsnorm(const void* x, const void* w, void* out, int64_t T, int64_t dim, float eps, cudaStream_t st) {
  MB_CHECK_ARG(dim % 8 == 0 && T >= 0, "rmsnorm: dim=%lld must be a multiple of 8", (long long)dim);
  if (T == 0) return MB200_OK;
  MB_CHECK_CUDA(launch_pdl(rmsnorm_kernel, dim3((unsigned)T), dim3(256), 0, st, (const uint4*)x, (const uint4*)w, (uint4*)out, (int)dim, eps));
  return MB200_OK;
}

template <int MODE>
static int run_linear(const void* x, const void* norm_w, const void* w, const EpiParams& epi, int64_t T, int64_t N, int64_t K, float eps,
                      void* workspace, size_t workspace_bytes, cudaStream_t st) {
  MB_CHECK_ARG(T >= 1, "linear: T=%lld", (long long)T);
  if (T <= MB200_SKINNY_MAX_T) {
    SkinnyParams p;
    p.x = x;
    p.norm_w = norm_w;
    p.w = w;
    p.N = (int)N;
    p.K = (int)K;
    p.eps = eps;
    p.epi = epi;
    return norm_w ? launch_skinny<MODE, true>(p, (int)T, st) : launch_skinny<MODE, false>(p, (int)T, st);
  }
  const void* a = x;
  if (norm_w) {
    const size_t need = kWsHeader + SK_PARTIAL_BYTES + align256((size_t)T * K * 2);
    if (workspace == nullptr || workspace_bytes < need) return fail(MB200_E_WORKSPACE, "linear: workspace %zu < %zu", workspace_bytes, need);
    void* normed = (uint8_t*)workspace + kWsHeader + SK_PARTIAL_BYTES;  // after the stream-K partial slots
    int rc = run_rmsnorm(x, norm_w, normed, T, K, eps, st);
    if (rc) return rc;
    a = normed;
  }
  GemmParams g;
  g.a = a;
  g.w = w;
  g.T = (int)T;
  g.N = (int)N;
  g.K = (int)K;
  g.epi = epi;
  if (streamk_eligible(T, N, K)) return launch_streamk<MODE>(g, workspace, workspace_bytes, kWsHeader, st);  // decode-sized batches: HBM-bound
  if (tcgen05_gemm_eligible(T, N, K)) return launch_gemm_tcgen05<MODE>(g, st);
  return launch_gemm_mma<MODE>(g, st);
}
}  // namespace mb200

using namespace mb200;

extern "C" {

int mb200_abi_version(void) { return MB200_ABI_VERSION; }
const char* mb200_last_error(void) { return g_err; }

int mb200_device_info(int* sm_count, int* max_smem_optin) {
  int dev = 0;
  MB_CHECK_CUDA(cudaGetDevice(&dev));
  if (sm_count) MB_CHECK_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  if (max_smem_optin) MB_CHECK_CUDA(cudaDeviceGetAttribute(max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  return MB200_OK;
}

size_t mb200_workspace_bytes(int64_t T, int64_t dim, int64_t n_heads, int64_t n_kv_heads, int64_t head_dim, int64_t hidden, int64_t vocab,
                             int64_t max_batch) {
  (void)vocab;
  (void)head_dim;
  const int64_t widest = dim > hidden ? dim : hidden;
  const int64_t rep = n_kv_heads > 0 ? n_heads / n_kv_heads : 1;
  size_t s = kWsHeader;
  s += SK_PARTIAL_BYTES;                                                  // stream-K partial accumulators (decode-sized GEMMs)
  s += align256((size_t)T * widest * 2);                                  // normed activations
  s += attn_decode_workspace(max_batch, n_kv_heads, 64, rep);             // split-KV partials (n_splits <= 64)
  // decode_step scratch (residual ping-pong, h, q, attn, g) lives in the same region as the normed activations
  const size_t mk = 6 * 256 + (size_t)(3 * dim + 2 * n_heads * head_dim + MK_MAX_TOPK * hidden) * 2 +
                    (size_t)256 * n_heads * (kHeadDim + 2) * sizeof(float) + 256;  // up to 256 SMs worth of attention slices
  if (s < kWsHeader + mk) s = kWsHeader + mk;
  return s;
}

int mb200_rmsnorm(const void* x, const void* w, void* out, int64_t T, int64_t dim, float eps, void* stream) {
  MB_CHECK_ARG(x && w && out, "rmsnorm: null pointer");
  return run_rmsnorm(x, w, out, T, dim, eps, (cudaStream_t)stream);
}

int mb200_attn_qkv(const void* x, const void* norm_w, const void* wqkv, const float* rope, const int32_t* positions, void* q_out, void* k_out,
                   void* v_out, void* cache_k, void* cache_v, const int32_t* cache_rows, int64_t T, int64_t dim, int64_t n_heads,
                   int64_t n_kv_heads, int64_t head_dim, float eps, void* workspace, size_t workspace_bytes, void* stream) {
  MB_CHECK_ARG(x && norm_w && wqkv && rope && positions && q_out && k_out && v_out, "attn_qkv: null pointer");
  MB_CHECK_ARG(head_dim == kHeadDim, "attn_qkv: head_dim=%lld unsupported (128 only)", (long long)head_dim);
  MB_CHECK_ARG(cache_rows == nullptr || (cache_k && cache_v), "attn_qkv: cache_rows without cache pointers");
  EpiParams e;
  e.q_out = q_out;
  e.k_out = k_out;
  e.v_out = v_out;
  e.cache_k = cache_k;
  e.cache_v = cache_v;
  e.positions = positions;
  e.cache_rows = cache_rows;
  e.rope = rope;
  e.q_dim = (int)(n_heads * head_dim);
  e.kv_dim = (int)(n_kv_heads * head_dim);
  const int64_t N = (n_heads + 2 * n_kv_heads) * head_dim;
  return run_linear<EPI_QKV_ROPE>(x, norm_w, wqkv, e, T, N, dim, eps, workspace, workspace_bytes, (cudaStream_t)stream);
}

int mb200_decode_meta(int32_t* seqpos_dev, int32_t* meta_dev, int64_t B, const int32_t* windows_host, int64_t n_windows, void* stream) {
  MB_CHECK_ARG(seqpos_dev && meta_dev && windows_host, "decode_meta: null pointer");
  MB_CHECK_ARG(B >= 1 && n_windows >= 1 && n_windows <= kMaxWindows, "decode_meta: B=%lld, n_windows=%lld (max %d)", (long long)B,
               (long long)n_windows, kMaxWindows);
  DecodeMetaParams p;
  p.seqpos = seqpos_dev;
  p.meta = meta_dev;
  p.B = (int)B;
  p.n_w = (int)n_windows;
  for (int j = 0; j < (int)n_windows; ++j) {
    MB_CHECK_ARG(windows_host[j] >= 1, "decode_meta: window %d", (int)windows_host[j]);
    p.windows[j] = windows_host[j];
  }
  decode_meta_kernel<<<(unsigned)ceil_div(B + 1, 128), 128, 0, (cudaStream_t)stream>>>(p);
  MB_CHECK_LAUNCH("decode_meta_kernel");
  return MB200_OK;
}

int mb200_kv_ring_write(const void* k_new, const void* v_new, void* cache_k, void* cache_v, const int32_t* cache_rows, int64_t T,
                        int64_t n_kv_heads, int64_t head_dim, void* stream) {
  MB_CHECK_ARG(k_new && v_new && cache_k && cache_v && cache_rows, "kv_ring_write: null pointer");
  MB_CHECK_ARG((n_kv_heads * head_dim) % 8 == 0, "kv_ring_write: row not 16-byte aligned");
  if (T == 0) return MB200_OK;
  kv_ring_write_kernel<<<(unsigned)T, 128, 0, (cudaStream_t)stream>>>((const uint4*)k_new, (const uint4*)v_new, (uint4*)cache_k, (uint4*)cache_v,
                                                                      cache_rows, (int)T, (int)(n_kv_heads * head_dim / 8));
  MB_CHECK_LAUNCH("kv_ring_write_kernel");
  return MB200_OK;
}

int mb200_attn_decode(const void* q, const void* cache_k, const void* cache_v, const int32_t* kv_len, void* out, int64_t B, int64_t W,
                      int64_t n_heads, int64_t n_kv_heads, int64_t head_dim, int64_t n_splits, void* workspace, size_t workspace_bytes,
                      void* stream) {
  MB_CHECK_ARG(q && cache_k && cache_v && kv_len && out, "attn_decode: null pointer");
  MB_CHECK_ARG(head_dim == kHeadDim, "attn_decode: head_dim=%lld unsupported (128 only)", (long long)head_dim);
  MB_CHECK_ARG(n_heads % n_kv_heads == 0, "attn_decode: H %% KV != 0");
  const int rep = (int)(n_heads / n_kv_heads);
  MB_CHECK_ARG(n_splits >= 1 && n_splits <= 64, "attn_decode: n_splits=%lld out of [1, 64]", (long long)n_splits);
  MB_CHECK_ARG((size_t)B * n_kv_heads * sizeof(int) <= 8192, "attn_decode: B*KV too large for the counter block");
  AttnDecodeParams p;
  p.q = (const bf16*)q;
  p.cache_k = (const bf16*)cache_k;
  p.cache_v = (const bf16*)cache_v;
  p.kv_len = kv_len;
  p.out = (bf16*)out;
  p.B = (int)B;
  p.W = (int)W;
  p.H = (int)n_heads;
  p.KV = (int)n_kv_heads;
  p.S = (int)n_splits;
  p.scale = 0.08838834764831845f;  // 128^-0.5 (xformers default scale; Attention.scale is unused, SURVEY E-3)
  p.partial = nullptr;
  p.counters = nullptr;
  if (n_splits > 1) {
    const size_t need = kWsHeader + (size_t)B * n_kv_heads * n_splits * rep * (kHeadDim + 2) * sizeof(float);
    if (workspace == nullptr || workspace_bytes < need) return fail(MB200_E_WORKSPACE, "attn_decode: workspace %zu < %zu", workspace_bytes, need);
    p.counters = (int*)workspace;
    p.partial = (float*)((uint8_t*)workspace + kWsHeader);
  }
  const dim3 grid((unsigned)n_splits, (unsigned)n_kv_heads, (unsigned)B);
  cudaStream_t st = (cudaStream_t)stream;
  const char* which = getenv("MB200_ATTN_DECODE");  // "plain": the register-staged kernel of attn_decode.cuh (A/B comparisons)
  if (!(which != nullptr && which[0] == 'p')) {
    if (n_splits > 1) p.counters = (int*)workspace;
    switch (rep) {
      case 1: return launch_attn_decode_tma<1>(p, B * W, st);
      case 2: return launch_attn_decode_tma<2>(p, B * W, st);
      case 4: return launch_attn_decode_tma<4>(p, B * W, st);
      case 6: return launch_attn_decode_tma<6>(p, B * W, st);
      case 8: return launch_attn_decode_tma<8>(p, B * W, st);
      default: return fail(MB200_E_INVALID, "attn_decode: H/KV=%d unsupported (1,2,4,6,8)", rep);
    }
  }
  switch (rep) {
    case 1: attn_decode_kernel<1><<<grid, AD_THREADS, 0, st>>>(p); break;
    case 2: attn_decode_kernel<2><<<grid, AD_THREADS, 0, st>>>(p); break;
    case 4: attn_decode_kernel<4><<<grid, AD_THREADS, 0, st>>>(p); break;
    case 6: attn_decode_kernel<6><<<grid, AD_THREADS, 0, st>>>(p); break;
    case 8: attn_decode_kernel<8><<<grid, AD_THREADS, 0, st>>>(p); break;
    default: return fail(MB200_E_INVALID, "attn_decode: H/KV=%d unsupported (1,2,4,6,8)", rep);
  }
  MB_CHECK_LAUNCH("attn_decode_kernel");
  return MB200_OK;
}

int mb200_attn_prefill(const void* q, const void* k_new, const void* v_new, const void* cache_k, const void* cache_v, const int32_t* q_start,
                       const int32_t* seqpos, void* out, int64_t T, int64_t B, int64_t max_seqlen, int64_t W, int64_t n_heads,
                       int64_t n_kv_heads, int64_t head_dim, int causal, void* stream) {
  MB_CHECK_ARG(q && k_new && v_new && out, "attn_prefill: null pointer");
  MB_CHECK_ARG(!causal || (cache_k && cache_v && q_start && seqpos), "attn_prefill: causal mode needs ring + metadata");
  MB_CHECK_ARG(head_dim == kHeadDim, "attn_prefill: head_dim=%lld unsupported (128 only)", (long long)head_dim);
  MB_CHECK_ARG(n_heads % n_kv_heads == 0, "attn_prefill: H %% KV != 0");
  if (T == 0) return MB200_OK;
  if (causal == 2 && tcgen05_attn_eligible(T, max_seqlen))  // first prefill: every key comes from the chunk -> tcgen05 / TMEM / TMA kernel
    return launch_attn_prefill_tcgen05(q, k_new, v_new, q_start, out, T, B, max_seqlen, W, n_heads, n_kv_heads, (cudaStream_t)stream);
  AttnPrefillParams p;
  p.q = (const bf16*)q;
  p.k_new = (const bf16*)k_new;
  p.v_new = (const bf16*)v_new;
  p.cache_k = (const bf16*)cache_k;
  p.cache_v = (const bf16*)cache_v;
  p.q_start = q_start;
  p.seqpos = seqpos;
  p.out = (bf16*)out;
  p.T = (int)T;
  p.B = (int)B;
  p.W = (int)W;
  p.H = (int)n_heads;
  p.KV = (int)n_kv_heads;
  p.causal = causal;
  p.scale_log2 = 0.08838834764831845f * 1.4426950408889634f;
  const int64_t span = causal ? max_seqlen : T;
  const dim3 grid((unsigned)ceil_div(span, AP_BQ), (unsigned)n_heads, (unsigned)(causal ? B : 1));
  MB_CHECK_CUDA(cudaFuncSetAttribute(attn_prefill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AP_SMEM));
  attn_prefill_kernel<<<grid, AP_THREADS, AP_SMEM, (cudaStream_t)stream>>>(p);
  MB_CHECK_LAUNCH("attn_prefill_kernel");
  return MB200_OK;
}

int mb200_linear_residual(const void* x, const void* w, const void* residual, void* out, int64_t T, int64_t N, int64_t K, void* workspace,
                          size_t workspace_bytes, void* stream) {
  MB_CHECK_ARG(x && w && out, "linear_residual: null pointer");
  EpiParams e;
  e.out = out;
  e.residual = residual;
  e.ld_out = N;
  if (residual) return run_linear<EPI_RESIDUAL>(x, nullptr, w, e, T, N, K, 0.f, workspace, workspace_bytes, (cudaStream_t)stream);
  return run_linear<EPI_STORE>(x, nullptr, w, e, T, N, K, 0.f, workspace, workspace_bytes, (cudaStream_t)stream);
}

int mb200_ffn_gateup(const void* x, const void* norm_w, const void* w13, void* g_out, int64_t T, int64_t dim, int64_t hidden, float eps,
                     void* workspace, size_t workspace_bytes, void* stream) {
  MB_CHECK_ARG(x && w13 && g_out, "ffn_gateup: null pointer");
  EpiParams e;
  e.out = g_out;
  e.ld_out = hidden;
  return run_linear<EPI_SWIGLU>(x, norm_w, w13, e, T, 2 * hidden, dim, eps, workspace, workspace_bytes, (cudaStream_t)stream);
}

int mb200_lm_head(const void* x, const void* norm_w, const void* w_out, float* logits, int64_t T, int64_t dim, int64_t vocab, float eps,
                  void* workspace, size_t workspace_bytes, void* stream) {
  MB_CHECK_ARG(x && norm_w && w_out && logits, "lm_head: null pointer");
  EpiParams e;
  e.out_f32 = logits;
  e.ld_out = vocab;
  return run_linear<EPI_F32>(x, norm_w, w_out, e, T, vocab, dim, eps, workspace, workspace_bytes, (cudaStream_t)stream);
}

int mb200_argmax_rows(const float* logits, int64_t* out_dev, int64_t T, int64_t vocab, void* stream) {
  MB_CHECK_ARG(logits && out_dev && T >= 0 && vocab >= 1, "argmax_rows: bad arguments");
  if (T == 0) return MB200_OK;
  argmax_rows_kernel<<<(unsigned)T, SP_THREADS, 0, (cudaStream_t)stream>>>(logits, (long long*)out_dev, (int)vocab);
  MB_CHECK_LAUNCH("argmax_rows_kernel");
  return MB200_OK;
}

int mb200_logprob_gather(const float* logits, const int64_t* target_dev, float* out_dev, int64_t T, int64_t vocab, void* stream) {
  MB_CHECK_ARG(logits && target_dev && out_dev && T >= 0 && vocab >= 1, "logprob_gather: bad arguments");
  if (T == 0) return MB200_OK;
  logprob_gather_kernel<<<(unsigned)T, SP_THREADS, 0, (cudaStream_t)stream>>>(logits, (const long long*)target_dev, out_dev, (int)vocab);
  MB_CHECK_LAUNCH("logprob_gather_kernel");
  return MB200_OK;
}

int mb200_sample_top_p(const float* logits, const float* uniform_dev, int64_t* out_dev, int64_t T, int64_t vocab, float temperature,
                       float top_p, void* stream) {
  MB_CHECK_ARG(logits && uniform_dev && out_dev && T >= 0 && vocab >= 1, "sample_top_p: bad arguments");
  MB_CHECK_ARG(temperature > 0.f && top_p >= 0.f && top_p <= 1.f, "sample_top_p: temperature=%g must be > 0 and top_p=%g in [0, 1]",
               (double)temperature, (double)top_p);
  if (T == 0) return MB200_OK;
  sample_top_p_kernel<<<(unsigned)T, SP_THREADS, 0, (cudaStream_t)stream>>>(logits, uniform_dev, (long long*)out_dev, (int)vocab,
                                                                            1.0f / temperature, top_p);
  MB_CHECK_LAUNCH("sample_top_p_kernel");
  return MB200_OK;
}

// ---- mixture of experts (csrc/moe.cuh) ---------------------------------------------------------------------------------------
int mb200_moe_sizes(int64_t T, int64_t n_experts, int64_t top_k, int64_t* tile_rows, int64_t* rows_cap, int64_t* plan_words) {
  MB_CHECK_ARG(T >= 1 && n_experts >= 1 && top_k >= 1, "moe_sizes: bad arguments");
  const int tr = moe_tile_rows(T);
  const int64_t cap = moe_tile_cap(T * top_k, n_experts, tr);
  if (tile_rows) *tile_rows = tr;
  if (rows_cap) *rows_cap = cap * tr;
  if (plan_words) *plan_words = moe_plan_words(T * top_k, n_experts, tr);
  return MB200_OK;
}

int mb200_moe_route(const void* hn, const void* gate_w, int64_t T, int64_t dim, int64_t n_experts, int64_t top_k, int64_t shard_rank,
                    int64_t shard_world, int32_t* sel, void* wts, int32_t* slot, int32_t* plan, void* xs, void* row_w, void* stream) {
  MB_CHECK_ARG(hn && gate_w && sel && wts && slot && plan && xs && row_w, "moe_route: null pointer");
  MB_CHECK_ARG(T >= 1 && dim % 8 == 0, "moe_route: T=%lld dim=%lld", (long long)T, (long long)dim);
  MB_CHECK_ARG(top_k >= 1 && top_k <= MOE_MAX_TOPK && top_k <= n_experts && n_experts <= MOE_MAX_EXPERTS,
               "moe_route: E=%lld (max %d), k=%lld (max %d)", (long long)n_experts, MOE_MAX_EXPERTS, (long long)top_k, MOE_MAX_TOPK);
  MB_CHECK_ARG(shard_world >= 1 && shard_rank >= 0 && shard_rank < shard_world, "moe_route: shard %lld of %lld", (long long)shard_rank,
               (long long)shard_world);
  cudaStream_t st = (cudaStream_t)stream;
  const bool wide = T <= 256;  // decode-sized batches: one CTA per token
  const dim3 blocks(wide ? (unsigned)T : (unsigned)ceil_div(T, 8));
#define MB_ROUTE(EE)                                                                                                                              \
  if (wide)                                                                                                                                       \
    MB_CHECK_CUDA(launch_pdl(moe_route_kernel<EE, true>, blocks, dim3(256), 0, st, (const bf16*)hn, (const bf16*)gate_w, (int)T, (int)dim,      \
                             (int)top_k, sel, (bf16*)wts));                                                                                       \
  else                                                                                                                                            \
    MB_CHECK_CUDA(launch_pdl(moe_route_kernel<EE, false>, blocks, dim3(256), 0, st, (const bf16*)hn, (const bf16*)gate_w, (int)T, (int)dim,     \
                             (int)top_k, sel, (bf16*)wts));
  switch (n_experts) {
    case 2: MB_ROUTE(2) break;
    case 4: MB_ROUTE(4) break;
    case 8: MB_ROUTE(8) break;
    case 16: MB_ROUTE(16) break;
    default: return fail(MB200_E_INVALID, "moe_route: n_experts=%lld unsupported (2, 4, 8, 16)", (long long)n_experts);
  }
#undef MB_ROUTE
  MB_CHECK_LAUNCH("moe_route_kernel");
  const int tile_rows = moe_tile_rows(T);
  const int64_t pairs = T * top_k, cap = moe_tile_cap(pairs, n_experts, tile_rows);
  const size_t plan_smem = ((size_t)n_experts * MP_THREADS + 2 * n_experts + 1) * sizeof(int32_t);
  MB_CHECK_CUDA(cudaFuncSetAttribute(moe_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan_smem));
  moe_plan_kernel<<<1, MP_THREADS, plan_smem, st>>>(sel, (int)pairs, (int)n_experts, tile_rows, (int)shard_rank, (int)shard_world, (int)cap, slot, plan);
  MB_CHECK_LAUNCH("moe_plan_kernel");
  moe_gather_kernel<<<(unsigned)ceil_div(pairs, 8), 256, 0, st>>>((const uint4*)hn, sel, (const bf16*)wts, slot, (int)pairs, (int)top_k, (int)(dim / 8),
                                                                  (int)shard_rank, (int)shard_world, (uint4*)xs, (bf16*)row_w);
  MB_CHECK_LAUNCH("moe_gather_kernel");
  return MB200_OK;
}

int mb200_moe_grouped_ffn(const void* xs, const void* const* w13_host, const void* const* w2_host, const int32_t* plan, const void* row_w,
                          const int32_t* slot, const void* residual, void* g, void* yw, void* out, int64_t T, int64_t dim, int64_t hidden,
                          int64_t n_experts, int64_t top_k, const mb200_moe_comm* comm, void* workspace, size_t workspace_bytes, void* stream) {
  MB_CHECK_ARG(xs && w13_host && w2_host && plan && row_w && slot && g && yw && out, "moe_grouped_ffn: null pointer");
  MB_CHECK_ARG(T >= 1 && top_k >= 1 && top_k <= MOE_MAX_TOPK && n_experts <= MOE_MAX_EXPERTS && dim % 64 == 0 && hidden % 64 == 0,
               "moe_grouped_ffn: T=%lld k=%lld E=%lld dim=%lld hidden=%lld", (long long)T, (long long)top_k, (long long)n_experts, (long long)dim,
               (long long)hidden);
  const int n_ranks = comm ? comm->n_ranks : 1, my_rank = comm ? comm->my_rank : 0;
  MB_CHECK_ARG(n_ranks >= 1 && n_ranks <= kMaxPeers && my_rank >= 0 && my_rank < n_ranks, "moe_grouped_ffn: rank %d of %d", my_rank, n_ranks);
  cudaStream_t st = (cudaStream_t)stream;
  const int tile_rows = moe_tile_rows(T);
  const int64_t pairs = T * top_k, rows_cap = moe_row_cap(pairs, n_experts, tile_rows);
  int local = 0;
  for (int e = 0; e < (int)n_experts; ++e) local += (w13_host[e] != nullptr);
  MB_CHECK_ARG(local >= 1, "moe_grouped_ffn: this rank owns no expert");
  // expected number of this rank's experts that get at least one row (uniform routing): sizes the decode tile width
  double touched = (double)n_experts * (1.0 - pow(1.0 - (double)top_k / (double)n_experts, (double)T)) * ((double)local / (double)n_experts);
  int est = (int)(touched + 0.5);
  if (est < 1) est = 1;
  if (tile_rows == 128) est = (int)((pairs / n_ranks + tile_rows - 1) / tile_rows) + local;
  EpiParams e1;
  e1.out = g;
  e1.ld_out = hidden;
  int rc = launch_grouped<EPI_SWIGLU>(xs, rows_cap, dim, 2 * hidden, w13_host, (int)n_experts, est, tile_rows, plan, e1, workspace, workspace_bytes, kWsHeader, st);
  if (rc) return rc;
  EpiParams e2;
  e2.out = yw;
  e2.ld_out = dim;
  e2.row_w = row_w;
  e2.n_peers = n_ranks - 1;
  for (int r = 0; r < n_ranks - 1; ++r) {
    MB_CHECK_ARG(comm->peer_yw[r] != nullptr, "moe_grouped_ffn: peer buffer %d missing", r);
    e2.peer_out[r] = comm->peer_yw[r];
  }
  rc = launch_grouped<EPI_MOE_SCALE>(g, rows_cap, hidden, dim, w2_host, (int)n_experts, est, tile_rows, plan, e2, workspace, workspace_bytes, kWsHeader, st);
  if (rc) return rc;
  MoeCombineParams c;
  c.yw = (const uint4*)yw;
  c.slot = slot;
  c.residual = (const uint4*)residual;
  c.out = (uint4*)out;
  c.T = (int)T;
  c.k = (int)top_k;
  c.row_chunks = (int)(dim / 8);
  c.n_ranks = n_ranks;
  c.my_rank = my_rank;
  c.my_flags = nullptr;
  c.epoch = nullptr;
  c.done_counter = nullptr;
  for (int r = 0; r < kMaxPeers; ++r) c.peer_flags[r] = nullptr;
  if (n_ranks > 1) {
    MB_CHECK_ARG(comm->my_flags && comm->epoch && comm->done_counter, "moe_grouped_ffn: comm state missing");
    c.my_flags = (unsigned*)comm->my_flags;
    c.epoch = (unsigned*)comm->epoch;
    c.done_counter = (int*)comm->done_counter;
    for (int r = 0; r < n_ranks - 1; ++r) {
      MB_CHECK_ARG(comm->peer_flags[r] != nullptr, "moe_grouped_ffn: peer flags %d missing", r);
      c.peer_flags[r] = (unsigned*)comm->peer_flags[r];
    }
  }
  moe_combine_kernel<<<(unsigned)T, 128, 0, st>>>(c);
  MB_CHECK_LAUNCH("moe_combine_kernel");
  return MB200_OK;
}

// ---- NVLink peer buffers for the expert-parallel exchange: plain CUDA IPC on cudaMalloc'ed memory ------------------------------
int mb200_comm_alloc(size_t bytes, void** ptr_out) {
  MB_CHECK_ARG(ptr_out && bytes > 0, "comm_alloc: bad arguments");
  MB_CHECK_CUDA(cudaMalloc(ptr_out, bytes));
  MB_CHECK_CUDA(cudaMemset(*ptr_out, 0, bytes));
  MB_CHECK_CUDA(cudaDeviceSynchronize());
  return MB200_OK;
}
int mb200_comm_free(void* ptr) {
  if (ptr) MB_CHECK_CUDA(cudaFree(ptr));
  return MB200_OK;
}
int mb200_comm_export(void* ptr, void* handle_out64) {
  MB_CHECK_ARG(ptr && handle_out64, "comm_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  MB_CHECK_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle_out64, ptr));
  return MB200_OK;
}
int mb200_comm_open(const void* handle64, void** ptr_out) {
  MB_CHECK_ARG(handle64 && ptr_out, "comm_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  MB_CHECK_CUDA(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return MB200_OK;
}
int mb200_comm_close(void* ptr) {
  if (ptr) MB_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return MB200_OK;
}

int mb200_decode_step(const mb200_layer_desc* layers_dev, const int32_t* windows_dev, int64_t n_layers, const void* emb, const void* final_norm,
                      const void* w_out, const float* rope, const int64_t* token_dev, int64_t pos, int64_t batch_row, float* logits,
                      int64_t* next_token_dev, int64_t dim,
                      int64_t hidden, int64_t n_heads, int64_t n_kv_heads, int64_t head_dim, int64_t vocab, float eps, int64_t n_experts,
                      int64_t top_k, const void* const* moe_gate_dev, const void* const* moe_w13_dev, const void* const* moe_w2_dev,
                      void* workspace, size_t workspace_bytes, void* stream) {
  static_assert(sizeof(mb200_layer_desc) == sizeof(MkLayer), "layer descriptor layout");
  MB_CHECK_ARG(layers_dev && windows_dev && emb && final_norm && w_out && rope && token_dev && logits && workspace, "decode_step: null pointer");
  MB_CHECK_ARG(head_dim == kHeadDim, "decode_step: head_dim=%lld unsupported (128 only)", (long long)head_dim);
  MB_CHECK_ARG(n_heads % n_kv_heads == 0, "decode_step: H %% KV != 0");
  const int rep = (int)(n_heads / n_kv_heads);
  const int64_t q_dim = n_heads * head_dim;
  auto cut_ok = [](int64_t K) { const int64_t nch = (K + MK_MAX_KC - 1) / MK_MAX_KC; return K % (nch * 8) == 0; };
  MB_CHECK_ARG(n_kv_heads * kHeadDim * 2 + MK_KV_PAD <= MK_STAGE_BYTES / 8, "decode_step: a ring stage must hold 8 padded K/V position rows");
  MB_CHECK_ARG(cut_ok(dim) && cut_ok(hidden) && cut_ok(q_dim), "decode_step: dim/hidden/q_dim must split into 16-byte-aligned row chunks");
  MB_CHECK_ARG(vocab % 2 == 0 && hidden % 1 == 0, "decode_step: vocab must be even");
  int dev = 0, sms = 0, smem_max = 0, coop = 0;
  MB_CHECK_CUDA(cudaGetDevice(&dev));
  MB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  MB_CHECK_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  MB_CHECK_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  MB_CHECK_ARG(coop, "decode_step: device does not support cooperative launch");

  MkParams p;
  p.layers = reinterpret_cast<const MkLayer*>(layers_dev);
  p.windows = windows_dev;
  p.n_layers = (int)n_layers;
  p.emb = (const bf16*)emb;
  p.final_norm = (const bf16*)final_norm;
  p.w_out = (const bf16*)w_out;
  p.rope = rope;
  p.token = token_dev;
  p.pos = (int)pos;
  p.batch_row = (int)batch_row;
  p.logits = logits;
  p.next_token = (long long*)next_token_dev;
  MB_CHECK_ARG(n_experts == 0 || (moe_gate_dev && moe_w13_dev && moe_w2_dev && top_k >= 1 && top_k <= MK_MAX_TOPK && top_k <= n_experts && n_experts <= 32),
               "decode_step: bad MoE arguments (E=%lld, k=%lld)", (long long)n_experts, (long long)top_k);
  p.n_experts = (int)n_experts;
  p.top_k = (int)(n_experts ? top_k : 0);
  p.moe_gate = (const bf16* const*)moe_gate_dev;
  p.moe_w13 = (const bf16* const*)moe_w13_dev;
  p.moe_w2 = (const bf16* const*)moe_w2_dev;
  p.dim = (int)dim;
  p.hidden = (int)hidden;
  p.H = (int)n_heads;
  p.KV = (int)n_kv_heads;
  p.vocab = (int)vocab;
  p.eps = eps;
  // shared memory plan: x buffer (also the attention merge scratch), barriers + reduction scratch, the rest is the ring
  int64_t widest = dim > hidden ? dim : hidden;
  if (q_dim > widest) widest = q_dim;
  size_t xs_bytes = (size_t)widest * 2;
  if (n_experts && xs_bytes < (size_t)top_k * hidden * 2) xs_bytes = (size_t)top_k * hidden * 2;  // g of every selected expert
  if (xs_bytes < 2048) xs_bytes = 2048;  // also the slice-merge scratch of phase 2b
  xs_bytes = (xs_bytes + 127) & ~(size_t)127;
  const size_t tail = 2 * MK_MAX_STAGES * sizeof(uint64_t) + 48 * sizeof(float) + sizeof(MoeRoute) + 8 + 64;
  int n_stages = (int)(((size_t)smem_max - xs_bytes - tail) / MK_STAGE_BYTES);
  if (n_stages > MK_MAX_STAGES) n_stages = MK_MAX_STAGES;
  MB_CHECK_ARG(n_stages > MK_CONSUMER_WARPS, "decode_step: not enough shared memory for the weight ring (%d stages)", n_stages);
  p.n_stages = n_stages;
  p.xs_bytes = (int)xs_bytes;
  {
    static int cap = -1;
    if (cap < 0) {
      const char* e = getenv("MB200_MK_INFLIGHT");
      cap = e ? atoi(e) : 5;  // B200 sweep (7B): 3 -> 3.02 ms/token, 4 -> 2.87, 6 -> 2.87, 8 -> 2.89, uncapped -> 2.92
    }
    p.inflight_cap = cap < 2 ? 2 : (cap > 8 ? 8 : cap);
    static int kvu = -1;
    if (kvu < 0) {
      const char* e2 = getenv("MB200_MK_KV_UNCAPPED");
      kvu = e2 ? atoi(e2) : 0;
    }
    p.kv_uncapped = kvu;
  }
  const size_t smem = (size_t)n_stages * MK_STAGE_BYTES + xs_bytes + tail;

  MB_CHECK_ARG(n_kv_heads <= MK_CONSUMER_WARPS, "decode_step: n_kv_heads=%lld > %d (one consumer warp per kv head)", (long long)n_kv_heads,
               MK_CONSUMER_WARPS);
  // global scratch: header words + activations + per-slice attention partials
  uint8_t* ws = (uint8_t*)workspace;
  p.attn_counters = (int*)(ws + 8192);
  p.argmax_counter = (int*)(ws + 12288);
  p.argmax_slots = (unsigned long long*)(ws + 32768);  // sms x 8 B
  p.bar_flags = (unsigned*)(ws + 16384);
  p.bar_epoch = (unsigned*)(ws + 20480);
  p.done_counter = (int*)(ws + 20480 + 128);
  size_t off = kWsHeader;
  auto take = [&](size_t bytes) { uint8_t* r = ws + off; off += align256(bytes); return r; };
  p.xbuf = (bf16*)take((size_t)2 * dim * 2);
  p.hbuf = (bf16*)take((size_t)dim * 2);
  p.qbuf = (bf16*)take((size_t)q_dim * 2);
  p.abuf = (bf16*)take((size_t)q_dim * 2);
  p.gbuf = (bf16*)take((size_t)(n_experts ? top_k : 1) * hidden * 2);
  p.partial = (float*)take((size_t)sms * n_heads * (kHeadDim + 2) * sizeof(float));  // [slice = CTA][H][m, l, acc[128]]
  p.prof = g_mk_prof;
  p.prof_bar = g_mk_prof_bar;
  if (workspace_bytes < off) return fail(MB200_E_WORKSPACE, "decode_step: workspace %zu < %zu", workspace_bytes, off);

  void* args[] = {(void*)&p};
  const void* fn = nullptr;
  switch (rep) {
    case 1: fn = (const void*)decode_megakernel<1>; break;
    case 2: fn = (const void*)decode_megakernel<2>; break;
    case 4: fn = (const void*)decode_megakernel<4>; break;
    case 6: fn = (const void*)decode_megakernel<6>; break;
    case 8: fn = (const void*)decode_megakernel<8>; break;
    default: return fail(MB200_E_INVALID, "decode_step: H/KV=%d unsupported (1,2,4,6,8)", rep);
  }
  MB_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MB_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)sms), dim3(MK_THREADS), args, smem, (cudaStream_t)stream));
  return MB200_OK;
}

// Debug: device buffer of n_layers*12 uint64 that CTA 0 of the decode megakernel fills with %globaltimer stamps (NULL = off).
int mb200_debug_set_decode_timeline(void* device_buffer) {
  g_mk_prof = (unsigned long long*)device_buffer;
  return MB200_OK;
}
int mb200_debug_set_barrier_timeline(void* device_buffer) {
  g_mk_prof_bar = (unsigned long long*)device_buffer;
  return MB200_OK;
}

// Test-only: CUDA-core fp32-accumulate GEMM (c fp32 [T, N]) used to cross-check the tensor-core kernels on the GPU.
int mb200_test_gemm_naive(const void* a, const void* w, float* c, int64_t T, int64_t N, int64_t K, void* stream) {
  MB_CHECK_ARG(a && w && c, "test_gemm_naive: null pointer");
  const dim3 grid((unsigned)ceil_div(N, 128), (unsigned)T);
  gemm_naive_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const bf16*)a, (const bf16*)w, c, (int)T, (int)N, (int)K);
  MB_CHECK_LAUNCH("gemm_naive_kernel");
  return MB200_OK;
}

}  // extern "C"

#ifdef MB200_SK_TRACE
// tracing build only (scripts/trace_streamk.py): copies the stream-K stamp ring to the host and returns the stamp count
extern "C" int mb200_debug_sk_trace(void* host_out, size_t bytes, unsigned* count) {
  using namespace mb200;
  if (bytes < sizeof(sk_trace_buf)) return fail(MB200_E_INVALID, "sk trace: %zu < %zu", bytes, sizeof(sk_trace_buf));
  MB_CHECK_CUDA(cudaDeviceSynchronize());
  MB_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, sk_trace_buf, sizeof(sk_trace_buf)));
  MB_CHECK_CUDA(cudaMemcpyFromSymbol(count, sk_trace_count, sizeof(unsigned)));
  return MB200_OK;
}
#endif
