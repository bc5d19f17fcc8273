// libmb200.so -- the C ABI declared in include/mistral_b200.h.  Argument checking + kernel dispatch only.
#include "attn_decode.cuh"
#include "attn_prefill.cuh"
#include "elementwise.cuh"
#include "gemm_mma.cuh"
#include "skinny_linear.cuh"

namespace mb200 {
thread_local char g_err[512] = "";

constexpr size_t kWsHeader = 64 * 1024;  // persistent, zero-initialised by the caller once: self-resetting counters
inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static int run_rmsnorm(const void* x, const void* w, void* out, int64_t T, int64_t dim, float eps, cudaStream_t st) {
  MB_CHECK_ARG(dim % 8 == 0 && T >= 0, "rmsnorm: dim=%lld must be a multiple of 8", (long long)dim);
  if (T == 0) return MB200_OK;
  rmsnorm_kernel<<<(unsigned)T, 256, 0, st>>>((const uint4*)x, (const uint4*)w, (uint4*)out, (int)dim, eps);
  MB_CHECK_LAUNCH("rmsnorm_kernel");
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
    const size_t need = kWsHeader + align256((size_t)T * K * 2);
    if (workspace == nullptr || workspace_bytes < need) return fail(MB200_E_WORKSPACE, "linear: workspace %zu < %zu", workspace_bytes, need);
    void* normed = (uint8_t*)workspace + kWsHeader;
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
  s += align256((size_t)T * widest * 2);                                  // normed activations
  s += attn_decode_workspace(max_batch, n_kv_heads, 64, rep);             // split-KV partials (n_splits <= 64)
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
  MB_CHECK_ARG((size_t)B * n_kv_heads * sizeof(int) <= kWsHeader, "attn_decode: B*KV too large for the counter block");
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

// Test-only: CUDA-core fp32-accumulate GEMM (c fp32 [T, N]) used to cross-check the tensor-core kernels on the GPU.
int mb200_test_gemm_naive(const void* a, const void* w, float* c, int64_t T, int64_t N, int64_t K, void* stream) {
  MB_CHECK_ARG(a && w && c, "test_gemm_naive: null pointer");
  const dim3 grid((unsigned)ceil_div(N, 128), (unsigned)T);
  gemm_naive_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const bf16*)a, (const bf16*)w, c, (int)T, (int)N, (int)K);
  MB_CHECK_LAUNCH("gemm_naive_kernel");
  return MB200_OK;
}

}  // extern "C"
