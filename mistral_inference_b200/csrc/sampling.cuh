// Device-side token selection and log-probabilities: the per-token tail of generate().
//
//   argmax_rows_kernel      greedy pick: torch.argmax(logits, -1) (generate.py:156), first index on ties
//   logprob_gather_kernel   log_softmax(logits, -1)[t, target[t]] (generate.py:101-117,134-135) without materialising [T, V]
//   sample_top_p_kernel     softmax(logits / temperature) -> nucleus (top-p) filter -> one draw (generate.py:151-170)
//
// All three are one CTA per row over fp32 logits [T, V] (the lm head's output).  Roofline: HBM/L2 -- V * 4 bytes per row and
// pass; argmax and logprob are single-pass (online log-sum-exp), top-p re-reads its row (L2 resident) during the threshold
// search.  Reductions are fixed-order (warp shuffles, then warp 0 over the per-warp partials): results are deterministic.
#pragma once
#include "common.cuh"

namespace mb200 {

constexpr int SP_THREADS = 1024;
constexpr int SP_WARPS = SP_THREADS / 32;

__device__ __forceinline__ unsigned long long argmax_key(float v, int idx) {
  // order-preserving map of the fp32 value in the high word, ~index in the low word: the maximum key is the largest value
  // and, among equal values, the smallest index (what torch.argmax returns)
  unsigned u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (unsigned)(0x7fffffff - idx);
}

// block-wide reductions over SP_THREADS threads; `scratch` holds SP_WARPS values; result broadcast to every thread
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  v = warp_sum(v);
  __syncthreads();  // scratch reuse
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = scratch[threadIdx.x & 31];  // SP_WARPS == 32: every lane reads one partial
  t = warp_sum(t);
  return t;
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = scratch[threadIdx.x & 31];
  t = warp_max(t);
  return t;
}

__global__ void __launch_bounds__(SP_THREADS) argmax_rows_kernel(const float* __restrict__ logits, long long* __restrict__ out, int V) {
  const float* row = logits + (int64_t)blockIdx.x * V;
  unsigned long long best = 0ull;
  for (int i = threadIdx.x; i < V; i += SP_THREADS) best = max(best, argmax_key(row[i], i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
  __shared__ unsigned long long sm[SP_WARPS];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x < 32) {
    best = sm[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (threadIdx.x == 0) out[blockIdx.x] = (long long)(0x7fffffff - (int)(best & 0xffffffffull));
  }
}

// out[t] = logits[t, target[t]] - max - log(sum(exp(logits[t, :] - max)))   (fp32, like torch.log_softmax on fp32 logits)
// rows with target[t] < 0 are skipped (out untouched).
__global__ void __launch_bounds__(SP_THREADS) logprob_gather_kernel(const float* __restrict__ logits, const long long* __restrict__ target,
                                                                    float* __restrict__ out, int V) {
  const long long tgt = target[blockIdx.x];
  if (tgt < 0) return;
  const float* row = logits + (int64_t)blockIdx.x * V;
  __shared__ float scratch[SP_WARPS];
  float m = -INFINITY;
  for (int i = threadIdx.x; i < V; i += SP_THREADS) m = fmaxf(m, row[i]);
  m = block_max(m, scratch);
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += SP_THREADS) s += expf(row[i] - m);
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) out[blockIdx.x] = (row[tgt] - m) - logf(s);
}

// Nucleus sampling, one draw per row.  The reference (generate.py:151-170):
//   probs = softmax(logits / temperature); sort descending; keep token j iff (mass of the tokens ranked before j) <= p;
//   renormalise; torch.multinomial(1).
// Without a sort: token i is kept iff S(p_i) <= top_p with S(q) = sum of probabilities strictly greater than q, i.e. the kept
// set is {i : p_i >= tau} for the smallest probability tau with S(tau) <= top_p; tau is found by bisection on the fp32 bit
// pattern (exact after 32 steps; probabilities are positive, so the patterns are ordered like the values).  The draw is the
// inverse CDF over the kept tokens in index order with the caller's uniform u[row] in [0, 1) -- the same distribution as
// multinomial over the sorted, renormalised vector (order is irrelevant).  Tokens with EQUAL probability at the cut are kept or
// dropped together, where the reference's unstable sort keeps an arbitrary subset of them.
__global__ void __launch_bounds__(SP_THREADS) sample_top_p_kernel(const float* __restrict__ logits, const float* __restrict__ uniform,
                                                                  long long* __restrict__ out, int V, float inv_temperature, float top_p) {
  const float* row = logits + (int64_t)blockIdx.x * V;
  __shared__ float scratch[SP_WARPS];
  __shared__ float warp_mass[SP_WARPS];
  float m = -INFINITY;
  for (int i = threadIdx.x; i < V; i += SP_THREADS) m = fmaxf(m, row[i] * inv_temperature);
  m = block_max(m, scratch);
  float z = 0.f;
  for (int i = threadIdx.x; i < V; i += SP_THREADS) z += expf(row[i] * inv_temperature - m);
  z = block_sum(z, scratch);
  const float inv_z = 1.0f / z;
  auto prob = [&](int i) { return expf(row[i] * inv_temperature - m) * inv_z; };
  // bisection on the bit pattern of tau in (0, 1]: invariant S(hi) <= top_p (S(1.0) = 0), S(lo) > top_p or lo = 0
  unsigned lo = 0u, hi = __float_as_uint(1.0f);
  while (hi - lo > 1u) {
    const unsigned mid = lo + (hi - lo) / 2u;
    const float q = __uint_as_float(mid);
    float s = 0.f;
    for (int i = threadIdx.x; i < V; i += SP_THREADS) {
      const float p = prob(i);
      s += p > q ? p : 0.f;
    }
    s = block_sum(s, scratch);
    if (s <= top_p)
      hi = mid;
    else
      lo = mid;
  }
  // kept set: p_i >= tau where tau = the smallest ACTUAL probability > lo's value ... any p in (value(lo), value(hi)] equals
  // value(hi) (adjacent floats), so "p >= value(hi)" is exact.  If even the largest probability alone exceeds top_p the first
  // token is still kept (its preceding mass is 0 <= p), which S(p_max) = 0 <= top_p guarantees here too.
  const float tau = __uint_as_float(hi);
  // mass of the kept set, then the draw: thread-contiguous chunks so that a prefix over threads is a prefix over indices
  const int per = (V + SP_THREADS - 1) / SP_THREADS;
  const int i0 = threadIdx.x * per, i1 = min(V, i0 + per);
  float mine = 0.f;
  for (int i = i0; i < i1; ++i) {
    const float p = prob(i);
    mine += p >= tau ? p : 0.f;
  }
  // inclusive scan over threads: within the warp, then over the warps
  float incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += t;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 31) warp_mass[threadIdx.x >> 5] = incl;
  __syncthreads();
  float before = 0.f, total = 0.f;
  for (int w = 0; w < SP_WARPS; ++w) {
    const float wm = warp_mass[w];
    if (w < (int)(threadIdx.x >> 5)) before += wm;
    total += wm;
  }
  // thread t owns the targets in [upper(t-1), upper(t)); the bounds come from the SAME numbers on both sides of every edge,
  // and a claim is resolved to the lowest thread, so there is exactly one winner even where rounding makes the prefix
  // non-monotone by an ulp
  __shared__ float upper_sm[SP_THREADS];
  __shared__ int winner_tid, winner;
  const float upper = before + incl;
  upper_sm[threadIdx.x] = upper;
  if (threadIdx.x == 0) {
    winner_tid = SP_THREADS;
    winner = -1;
  }
  __syncthreads();
  const float lower = threadIdx.x ? upper_sm[threadIdx.x - 1] : 0.f;
  total = upper_sm[SP_THREADS - 1];
  const float target = uniform[blockIdx.x] * total;
  if (mine > 0.f && target >= lower && target < upper) atomicMin(&winner_tid, (int)threadIdx.x);
  __syncthreads();
  if ((int)threadIdx.x == winner_tid) {
    float run = lower;
    int pick = -1;
    for (int i = i0; i < i1; ++i) {
      const float p = prob(i);
      if (p >= tau) {
        pick = i;  // last kept token so far: the fallback when rounding pushes the target past the chunk's sum
        run += p;
        if (target < run) break;
      }
    }
    winner = pick;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (winner < 0) {  // target == total after rounding: the last kept token overall
      for (int i = V - 1; i >= 0; --i)
        if (prob(i) >= tau) {
          winner = i;
          break;
        }
    }
    out[blockIdx.x] = winner;
  }
}

}  // namespace mb200
