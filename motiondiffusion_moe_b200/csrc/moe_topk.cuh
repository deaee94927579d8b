// softmax + top-2 of the MoE gate (reference: models/switch_moe.py:53-57), shared by the routing kernels (moe.cu, ep.cu) and
// the gate stage of the fused Linear + LayerNorm kernel (gemm_ln.cu).
#pragma once
#include "common.cuh"

namespace {

// softmax over E logits with ATen's softmax_warp_forward arithmetic (max-subtract, expf, butterfly
// xor-shuffle sum over next_pow2(E) lanes, IEEE divide), then top-2 with the tie order of
// torch.topk on CUDA: lowest indices are selected first; equal values are emitted higher index first.
// E is a compile-time power of two so that everything stays in registers.
template <int E>
__device__ __forceinline__ void softmax_top2(const float (&logits)[E], float (&probs)[E], int& i0, int& i1,
                                             float& v0, float& v1) {
  float mx = logits[0];
#pragma unroll
  for (int e = 1; e < E; ++e) mx = fmaxf(mx, logits[e]);
  float ex[E], red[E];
#pragma unroll
  for (int e = 0; e < E; ++e) { ex[e] = expf(logits[e] - mx); red[e] = ex[e]; }
#pragma unroll
  for (int off = E >> 1; off > 0; off >>= 1) {
    float nxt[E];
#pragma unroll
    for (int e = 0; e < E; ++e) nxt[e] = red[e] + red[e ^ off];
#pragma unroll
    for (int e = 0; e < E; ++e) red[e] = nxt[e];
  }
  const float sum = red[0];
#pragma unroll
  for (int e = 0; e < E; ++e) probs[e] = ex[e] / sum;
  int a = 0;
  float pa = probs[0];
#pragma unroll
  for (int e = 1; e < E; ++e)
    if (probs[e] > pa) { a = e; pa = probs[e]; }
  // b starts at a valid index (not -1): with non-finite probabilities no comparison succeeds and the indices must
  // still stay inside [0, E) - a numerical blow-up may not turn into an out-of-bounds scatter in permute / dispatch
  int b = (a == 0) ? 1 : 0;
  float pb = probs[b];
#pragma unroll
  for (int e = 0; e < E; ++e)
    if (e != a && e != ((a == 0) ? 1 : 0) && probs[e] > pb) { b = e; pb = probs[e]; }
  if (pa == pb) { i0 = b; i1 = a; v0 = pb; v1 = pa; }  // b > a here: tie => higher index first
  else { i0 = a; i1 = b; v0 = pa; v1 = pb; }
}

}  // namespace
