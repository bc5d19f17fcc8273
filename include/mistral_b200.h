/*
 * mistral_b200.h -- C ABI of libmb200.so: the sm_100a implementation of the mistral-inference
 * transformer hot path (Attention block + FeedForward/MoE block + the norm / lm-head either side).
 *
 * The reference (mistralai/mistral-inference @ 2557e12) is pure Python and has NO FFI / plugin
 * boundary; its hot path is a sequence of torch / xformers library calls.
 * Each entry point below replaces one group of those call sites and cites them.  The reference-side
 * binding is a ctypes stub; in this repo the caller is
 * mistral_inference_b200/_abi.py, which mirrors the reference's Python API on top.
 *
 * Conventions (all entry points):
 *   - plain C types only.  Every `*_d` / `const void*` tensor argument is a DEVICE pointer
 *     (torch: tensor.data_ptr()); bf16 tensors are raw 16-bit words, row-major, innermost contiguous.
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream); all work is
 *     enqueued on it; no entry point synchronises, allocates or frees device memory, or touches host
 *     copies of the data.  Scratch comes from the caller-provided `workspace` (device, 256-B aligned).
 *   - returns 0 on success, a negative MB200_E_* code otherwise; mb200_last_error() gives the
 *     thread-local message.  Nothing throws across the boundary.
 *   - integer metadata (positions, rows, lengths) are int32 device arrays.
 *   - head_dim must be 128 (every config in BASELINE.json); dims must be multiples of 8 (16-byte rows).
 */
#ifndef MISTRAL_B200_H_
#define MISTRAL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB200_ABI_VERSION 1

#define MB200_OK 0
#define MB200_E_INVALID (-1)   /* bad argument / unsupported shape */
#define MB200_E_WORKSPACE (-2) /* workspace too small */
#define MB200_E_CUDA (-3)      /* CUDA runtime / driver error at launch */

int mb200_abi_version(void);
const char* mb200_last_error(void);
/* Number of SMs / max opt-in shared memory of the current device (for the host-side planners). */
int mb200_device_info(int* sm_count, int* max_smem_optin);

/* ---------------------------------------------------------------------------------------------
 * RMSNorm.  out = bf16( bf16( x_f32 * rsqrt(mean(x_f32^2) + eps) ) * w )
 * Replaces RMSNorm.forward (transformer_layers.py:115-120; call sites :165,:167, transformer.py:219).
 * x, out: [T, dim] bf16; w: [dim] bf16.
 */
int mb200_rmsnorm(const void* x, const void* w, void* out, int64_t T, int64_t dim, float eps, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused attention input: RMSNorm -> packed QKV projection -> interleaved-pair RoPE (q,k) -> optional
 * scatter of k,v into the rotating KV cache.
 * Replaces attention_norm + wq/wk/wv + apply_rotary_emb + CacheView.update
 * (transformer_layers.py:165,66-70; rope.py:13-23; cache.py:83-92).
 *   x          [T, dim] bf16 (block input, un-normed)
 *   norm_w     [dim] bf16
 *   wqkv       [(H + 2*KV) * hd, dim] bf16: rows = wq ++ wk ++ wv ([out, in] like nn.Linear)
 *   rope       [n_pos, hd/2, 2] fp32 = view_as_real(precompute_freqs_cis(...)) (rope.py:6-10)
 *   positions  [T] int32 absolute positions (cache.py:228-230)
 *   q_out      [T, H*hd] bf16; k_out, v_out [T, KV*hd] bf16 (rotated k, raw v)
 *   cache_k/v  [n_rows, KV, hd] bf16 flat ring (cache.py:88-89) and cache_rows [T] int32 = slot + b*W
 *              (cache.py:235) or -1 for tokens that are not cached (to_cache_mask false, cache.py:226).
 *              Pass cache_rows = NULL to skip the scatter (prefill reads the old ring first:
 *              transformer_layers.py:75-76; use mb200_kv_ring_write afterwards).
 * T <= MB200_SKINNY_MAX_T uses the weight-streaming GEMV path (HBM-bound), larger T the tensor-core path.
 * workspace: >= mb200_workspace_bytes(...) for this T.
 */
int mb200_attn_qkv(const void* x, const void* norm_w, const void* wqkv, const float* rope, const int32_t* positions,
                   void* q_out, void* k_out, void* v_out, void* cache_k, void* cache_v, const int32_t* cache_rows,
                   int64_t T, int64_t dim, int64_t n_heads, int64_t n_kv_heads, int64_t head_dim, float eps,
                   void* workspace, size_t workspace_bytes, void* stream);

/* CacheView.update on its own (cache.py:83-92): rows[t] >= 0 ? cache[rows[t]] = src[t]. */
int mb200_kv_ring_write(const void* k_new, const void* v_new, void* cache_k, void* cache_v, const int32_t* cache_rows,
                        int64_t T, int64_t n_kv_heads, int64_t head_dim, void* stream);

/* ---------------------------------------------------------------------------------------------
 * GQA decode attention over the rotating cache (one query token per sequence).
 * Replaces cache.key/value + repeat_kv + memory_efficient_attention with
 * BlockDiagonalCausalWithOffsetPaddedKeysMask (transformer_layers.py:78-88, cache.py:250-254):
 * sequence b attends to ring slots [0, kv_len[b]) of its ring, kv_len = min(pos+1, W); order-free softmax.
 *   q [B, H*hd] bf16; cache_k/v [max_batch, W, KV, hd] bf16; kv_len [B] int32 (device); out [B, H*hd] bf16
 *   n_splits: KV range is cut into this many CTAs per (b, kv head) (flash-decoding), combined in-kernel.
 */
int mb200_attn_decode(const void* q, const void* cache_k, const void* cache_v, const int32_t* kv_len, void* out,
                      int64_t B, int64_t W, int64_t n_heads, int64_t n_kv_heads, int64_t head_dim, int64_t n_splits,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Prefill attention (varlen, causal, sliding window) reading old keys from the ring and new keys from
 * k_new/v_new.  Replaces interleave_kv + repeat_kv + memory_efficient_attention with
 * BlockDiagonalCausalMask.make_local_attention / BlockDiagonalMask.make_local_attention_from_bottomright
 * (transformer_layers.py:75-76,84-88; cache.py:94-117,240,243-248).
 * Query i of sequence b sits at absolute position p = seqpos[b] + i and attends to absolute positions
 * (p - W, p]; positions < seqpos[b] come from ring slot (pos % W), the rest from the new chunk.
 *   q [T, H*hd]; k_new, v_new [T, KV*hd]; cache_k/v [max_batch, W, KV, hd]; out [T, H*hd] (all bf16)
 *   q_start [B+1] int32 (prefix sums of seqlens), seqpos [B] int32 (tokens already cached) -- device arrays
 *   max_seqlen: max over b of seqlens[b] (grid sizing);  window = W (cache size of this layer)
 *   causal = 2: like 1, and the caller guarantees seqpos[b] == 0 for every sequence (first prefill, cache.py:236-240): no key
 *               comes from the ring, which lets large chunks run on the tcgen05 / TMEM / TMA kernel.
 *   causal = 0: the cache-less forward (transformer_layers.py:72-73,88 with mask=None): every query attends
 *               to every new key of the whole flattened batch; ring, q_start, seqpos are ignored.
 */
int mb200_attn_prefill(const void* q, const void* k_new, const void* v_new, const void* cache_k, const void* cache_v,
                       const int32_t* q_start, const int32_t* seqpos, void* out, int64_t T, int64_t B, int64_t max_seqlen,
                       int64_t W, int64_t n_heads, int64_t n_kv_heads, int64_t head_dim, int causal, void* stream);

/* ---------------------------------------------------------------------------------------------
 * out = residual + bf16( x @ W^T )   (bf16 add, one more rounding).
 * Replaces wo + residual (transformer_layers.py:93,166) and w2 + residual (:106,:168).
 *   x [T, K]; w [N, K]; residual, out [T, N] (out may alias residual).  residual = NULL: plain linear.
 */
int mb200_linear_residual(const void* x, const void* w, const void* residual, void* out, int64_t T, int64_t N, int64_t K,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused FFN input: RMSNorm -> packed gate/up projection -> bf16(silu(a)) * b.
 * Replaces ffn_norm + w1, w3, silu, mul (transformer_layers.py:167,106).
 *   x [T, dim]; norm_w [dim] (NULL = x is already normed, used by MoE experts);
 *   w13 [2*hidden, dim]: row 2i = w1[i] (gate), row 2i+1 = w3[i] (up);  g_out [T, hidden]
 */
int mb200_ffn_gateup(const void* x, const void* norm_w, const void* w13, void* g_out, int64_t T, int64_t dim,
                     int64_t hidden, float eps, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Final RMSNorm + lm head, fp32 logits.  Replaces norm + output + .float() (transformer.py:219,235,240).
 *   x [T, dim]; norm_w [dim]; w_out [V, dim]; logits [T, V] fp32 (each value is a bf16-rounded number).
 */
int mb200_lm_head(const void* x, const void* norm_w, const void* w_out, float* logits, int64_t T, int64_t dim,
                  int64_t vocab, float eps, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Device-side step state of the batched decode loop.  Replaces BufferCache.get_input_metadata for one-token steps
 * (cache.py:197-263: positions, cache_positions, the padded-keys mask's kv_seqlen) and update_seqlens
 * (cache.py:191-192) with one tiny kernel, so a CUDA graph of the decode step replays without host writes.
 *   seqpos_dev [B] int32 (DEVICE, in/out): tokens cached so far per sequence; incremented by one.
 *   meta_dev   [3B + 1 + n_windows * 2B] int32 (DEVICE, out):
 *              positions[B] | q_start[B+1] | seqpos[B] | per distinct window W: cache_rows[B] (= pos %% W + b*W), kv_len[B]
 *   windows_host [n_windows] int32 (HOST array, n_windows <= 8): the distinct cache sizes (cache.py:13-24), ascending.
 */
int mb200_decode_meta(int32_t* seqpos_dev, int32_t* meta_dev, int64_t B, const int32_t* windows_host, int64_t n_windows,
                      void* stream);

/* ---------------------------------------------------------------------------------------------
 * Token selection and log-probabilities on the device (the per-token tail of generate()).
 *   logits [T, vocab] fp32 (the lm head's output), one CTA per row, nothing of size [T, vocab] is written.
 * mb200_argmax_rows:    out[t] = argmax(logits[t]) with the first index on ties -- `sample` with temperature == 0
 *                       (generate.py:154-158).
 * mb200_logprob_gather: out[t] = log_softmax(logits[t])[target[t]] in fp32 (generate.py:101-117,134-135); rows with
 *                       target[t] < 0 are skipped.
 * mb200_sample_top_p:   one draw per row from softmax(logits / temperature) restricted to the nucleus: a token is kept iff
 *                       the probability mass of the tokens ranked before it is <= top_p (generate.py:151-170; the reference
 *                       hard-codes top_p = 0.8, :126).  uniform [T] fp32 in [0, 1) supplies the randomness (torch.rand on the
 *                       device keeps torch.manual_seed semantics); the draw is the inverse CDF over the kept tokens in index
 *                       order -- same distribution as torch.multinomial on the sorted vector, not the same stream.
 *   out / target: int64 device arrays.
 */
int mb200_argmax_rows(const float* logits, int64_t* out_dev, int64_t T, int64_t vocab, void* stream);
int mb200_logprob_gather(const float* logits, const int64_t* target_dev, float* out_dev, int64_t T, int64_t vocab, void* stream);
int mb200_sample_top_p(const float* logits, const float* uniform_dev, int64_t* out_dev, int64_t T, int64_t vocab,
                       float temperature, float top_p, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Mixture of experts for T > 1 tokens (prefill, batched decode).  Replaces MoeLayer.forward (moe.py:24-32): the gate Linear,
 * torch.topk on the bf16 router logits, the fp32 softmax over the k selected, and the per-expert torch.where / gather /
 * FeedForward / weighted `results[idx] +=` loop (one host sync per expert) -- with no host round trip at all.
 *
 * mb200_moe_sizes:   buffer sizes for T tokens: tile_rows (m-tile height of the grouped GEMMs: 32 / 64 for decode-sized batches,
 *                    128 otherwise), rows_cap (rows of xs / g / yw / row_w: every expert's segment is padded to a multiple of
 *                    tile_rows), plan_words (int32 words of `plan`).
 * mb200_moe_route:   router + routing plan + gather.
 *     hn [T, dim] bf16 = ffn_norm(h); gate_w [E, dim] bf16 (moe.py:20)
 *     sel [T, k] int32, wts [T, k] bf16: the selected experts of each token in ASCENDING expert index with their routing weights
 *     slot [T, k] int32: row of each (token, expert) pair in the expert-sorted buffers (deterministic: token order per expert)
 *     plan [plan_words] int32: device-side description of the grouped GEMMs' m tiles (count, expert and first row of each)
 *     xs [rows_cap, dim] bf16: hn rows gathered by slot; row_w [rows_cap] bf16: routing weight of each row
 *     shard_rank / shard_world: expert parallelism, this rank owns the experts e % shard_world == shard_rank (1 rank: 0 / 1);
 *     slots are numbered over ALL experts on every rank, tiles and gathered rows cover the local experts only.
 * mb200_moe_grouped_ffn: grouped gate/up GEMM (+ SiLU*mul) -> grouped down GEMM whose epilogue rounds the expert output to bf16,
 *     scales by the routing weight, rounds again (moe.py:31) and stores the row locally AND on every peer (comm->peer_yw: NVLink
 *     stores, the expert-parallel exchange is this epilogue) -> combine: out[t] = residual[t] + sum over the token's k rows in
 *     ascending expert index, every step rounded to bf16 like the reference's `+=`.
 *     w13_host / w2_host: HOST arrays of E device pointers (packed gate/up [2*hidden, dim] and down [dim, hidden] of each expert;
 *     NULL for experts of other ranks).  g [rows_cap, hidden], yw [rows_cap, dim] bf16 scratch; out [T, dim]; residual may be NULL.
 *     comm: NULL when unsharded; otherwise the mapped peer buffers and the handshake words (see mb200_comm_* below).  The combine
 *     kernel signals every peer that this rank's rows are written and waits for every peer's signal; `epoch` counts the calls.
 */
typedef struct mb200_moe_comm {
  int32_t n_ranks, my_rank;
  void* peer_yw[8];     /* yw buffer of the other ranks (mapped), n_ranks - 1 entries */
  void* my_flags;       /* uint32 [n_ranks] in this rank's comm buffer: flags[r] written by rank r */
  void* peer_flags[8];  /* the same array on the other ranks (mapped), n_ranks - 1 entries */
  void* epoch;          /* uint32 device word, local: number of completed calls on this buffer */
  void* done_counter;   /* int32 device word, local, zero */
} mb200_moe_comm;
int mb200_moe_sizes(int64_t T, int64_t n_experts, int64_t top_k, int64_t* tile_rows, int64_t* rows_cap, int64_t* plan_words);
int mb200_moe_route(const void* hn, const void* gate_w, int64_t T, int64_t dim, int64_t n_experts, int64_t top_k, int64_t shard_rank,
                    int64_t shard_world, int32_t* sel, void* wts, int32_t* slot, int32_t* plan, void* xs, void* row_w, void* stream);
int mb200_moe_grouped_ffn(const void* xs, const void* const* w13_host, const void* const* w2_host, const int32_t* plan, const void* row_w,
                          const int32_t* slot, const void* residual, void* g, void* yw, void* out, int64_t T, int64_t dim,
                          int64_t hidden, int64_t n_experts, int64_t top_k, const mb200_moe_comm* comm, void* workspace,
                          size_t workspace_bytes, void* stream);

/* Buffers that other ranks (one process per GPU) can write: plain cudaMalloc + CUDA IPC.  alloc zero-fills and synchronises;
 * export writes the 64-byte IPC handle to pass to the other processes (e.g. torch.distributed.all_gather_object); open maps a
 * peer's buffer into this process (peer access over NVLink is enabled lazily).  These are the only entry points that allocate. */
int mb200_comm_alloc(size_t bytes, void** ptr_out);
int mb200_comm_free(void* ptr);
int mb200_comm_export(void* ptr, void* handle_out64);
int mb200_comm_open(const void* handle64, void** ptr_out);
int mb200_comm_close(void* ptr);

/* Upper bound of the scratch any entry point above needs for up to T tokens of this geometry. */
size_t mb200_workspace_bytes(int64_t T, int64_t dim, int64_t n_heads, int64_t n_kv_heads, int64_t head_dim,
                             int64_t hidden, int64_t vocab, int64_t max_batch);

/* ---------------------------------------------------------------------------------------------
 * One whole decode step (batch 1) as ONE persistent cooperative kernel: embedding row -> n_layers x
 * [RMSNorm+QKV+RoPE+ring write | GQA attention over the ring | wo+residual | RMSNorm+gate/up+SiLU*mul |
 * down+residual] -> final RMSNorm + lm head.  Replaces one Transformer.forward(next_token, seqlens=[1], cache)
 * call of the decode loop (generate.py:139 -> transformer.py:163-242) including every per-layer library call.
 * Weights stream through a shared-memory ring fed by TMA bulk copies that keep prefetching across the phase
 * (grid) barriers, which is what lets a batch-1 step approach the HBM roofline.
 *   layers_dev   [n_layers] mb200_layer_desc in DEVICE memory (pointers to this layer's packed weights + ring)
 *   windows_dev  [n_layers] int32 ring size W of each layer
 *   token_dev    device int64 scalar (e.g. the previous step's argmax); pos = its absolute position;
 *                batch_row = which row of the [max_batch, W, KV, hd] cache this sequence occupies
 *   logits       [vocab] fp32
 *   next_token_dev  optional device int64: greedy argmax of the logits, first index on ties (torch.argmax, generate.py:156);
 *                NULL to skip.  Feeding it back as token_dev makes the greedy loop one launch per token, nothing on the host.
 *   n_experts/top_k  0/0 for dense FeedForward layers.  Mixture of experts (moe.py:16-32): the layer descriptors' w13/w2 are
 *                ignored; moe_gate_dev [n_layers] (router weight [E, dim]), moe_w13_dev / moe_w2_dev [n_layers * E] are DEVICE arrays
 *                of device pointers.  Router, top-k, softmax over the k and the ascending-expert bf16 accumulation run in-kernel.
 * Requires a device that can co-schedule one CTA per SM (cooperative launch).
 */
typedef struct mb200_layer_desc {
  const void* wqkv;      /* [(H+2KV)*hd, dim] */
  const void* wo;        /* [dim, H*hd] */
  const void* w13;       /* [2*hidden, dim], rows interleaved w1/w3 */
  const void* w2;        /* [dim, hidden] */
  const void* attn_norm; /* [dim] */
  const void* ffn_norm;  /* [dim] */
  void* cache_k;         /* [max_batch, W, KV, hd] */
  void* cache_v;
} mb200_layer_desc;

int mb200_decode_step(const mb200_layer_desc* layers_dev, const int32_t* windows_dev, int64_t n_layers, const void* emb,
                      const void* final_norm, const void* w_out, const float* rope, const int64_t* token_dev, int64_t pos,
                      int64_t batch_row, float* logits, int64_t* next_token_dev, int64_t dim, int64_t hidden, int64_t n_heads, int64_t n_kv_heads,
                      int64_t head_dim, int64_t vocab, float eps, int64_t n_experts, int64_t top_k, const void* const* moe_gate_dev,
                      const void* const* moe_w13_dev, const void* const* moe_w2_dev, void* workspace, size_t workspace_bytes, void* stream);

#define MB200_SKINNY_MAX_T 4

/* Workspace contract: the first 64 KiB of `workspace` hold self-resetting counters (split-KV arrival counts, the decode
 * kernel's grid-barrier words and epoch, stream-K flags); the caller zero-fills the workspace ONCE when allocating it
 * (torch.zeros) and never writes to it afterwards.
 * Concurrency: a workspace carries state BETWEEN and DURING launches, so all calls that share one workspace must be ordered on
 * one stream (or by events); concurrent streams need one workspace each.  The library keeps no other mutable state that affects
 * results: process-wide state is limited to the debug hooks below, environment switches read once, and the driver entry point
 * for tensor-map encoding.  mb200_last_error() is thread-local. */
#define MB200_WORKSPACE_HEADER_BYTES (64 * 1024)

/* Debug: a device buffer of n_layers*12 uint64 that mb200_decode_step fills with %globaltimer stamps at every phase
 * boundary of CTA 0 (NULL switches it off).  Used by scripts/mk_timeline.py to see where a decode step spends time. */
int mb200_debug_set_decode_timeline(void* device_buffer);
/* Debug: [n_sm][n_layers][6][2] uint64 arrive/leave stamps of every CTA at every grid barrier (NULL = off). */
int mb200_debug_set_barrier_timeline(void* device_buffer);

/* Test-only: CUDA-core fp32 GEMM c[T, N] = a[T, K] w[N, K]^T used to cross-check the tensor-core kernels. */
int mb200_test_gemm_naive(const void* a, const void* w, float* c, int64_t T, int64_t N, int64_t K, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MISTRAL_B200_H_ */
