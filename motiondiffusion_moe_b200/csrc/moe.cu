// MoE routing path of MoEMultiBranchFFN / SwitchMoELayer
// (reference: models/multi_branch.py:52-61, models/switch_moe.py:44-111).
//
//   gate     per token, per branch: LN -> gate GEMV (D -> E) -> softmax (ATen arithmetic order) ->
//            top-2 with torch.topk's CUDA tie order; per-128-token-block histograms (warp-private
//            counters, no atomics => deterministic).
//   scan     one CTA: expert segment offsets padded to the 128-row GEMM tile, per-block bases, the
//            grouped-GEMM tile tables, the usage / importance counters.  No host synchronisation.
//   permute  expert-sorted copy of the LN'd rows (ballot/match ranks inside a block), row scales.
//   combine  gathers the NB*K expert rows of a token, sums them, and applies the FiLM of the
//            following StylizationBlock (its LN, scale/shift and SiLU).
// All four are HBM-bound; algorithmic bytes per token (D elements, s = sizeof activation):
//   gate 4D (read x) ; permute 4D + NB*K*D*s ; combine NB*K*D*s + D*s.
#include "common.cuh"
#include "rowmath.cuh"
#include "moe_topk.cuh"

namespace {

// A/B knobs (tools/build_variant.sh): MDM_PERM_SPLIT = CTAs per 128-token block of the permute, MDM_SCAN_FAST = single-round-trip scan
#ifndef MDM_PERM_SPLIT
#define MDM_PERM_SPLIT 4          // 1 = one CTA per block (A/B builds: tools/build_variant.sh)
#endif
#ifndef MDM_SCAN_FAST
#define MDM_SCAN_FAST 1
#endif
constexpr int TOK_PER_BLK = 128;
constexpr int MAX_G = 32;   // NB * E groups
constexpr int MAX_E = 16;

// Sum GP per-lane partial values across the warp: after the call lane l holds (in p[0]) the total of
// value index l >> (5 - log2(GP)).  GP-1 + (5 - log2 GP) shuffles instead of 5*GP.
template <int GP>
__device__ __forceinline__ void warp_reduce_scatter(float (&p)[GP], int lane) {
  int off = 16;
#pragma unroll
  for (int n = GP; n > 1; n >>= 1, off >>= 1) {
    const int half = n >> 1;
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? p[i] : p[i + half];
      const float keep = upper ? p[i + half] : p[i];
      p[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
#pragma unroll
  for (; off > 0; off >>= 1) p[0] += __shfl_xor_sync(0xffffffffu, p[0], off);
}

// Two tokens per warp iteration: every shared-memory read of a gate / LayerNorm weight vector feeds
// both tokens (the kernel was bound by shared-memory bandwidth with one), and the softmax / top-2
// of the TT*NB (token, branch) pairs is evaluated once, each by its own lane, instead of 32 times.
// The per-(token, expert) arithmetic order is unchanged (bit-identical logits, probabilities, counters).
// The 128-token block is the histogram granule of the C-ABI.  Two shapes (B200, D = 512, 16 groups):
//   GATE_WARPS = 16 (8 tokens per warp, one block per SM, the rows of the next token pair requested before the
//     current pair is evaluated) when there are fewer blocks than SMs: the parallelism has to come from
//     inside the block (N = 12 544: 32.5 -> 25.6 us);
//   GATE_WARPS = 8 (16 tokens per warp, two blocks per SM) otherwise (N = 25 088: 37 us; 16 warps: 49 us).
// FORCED (parity hook, mdm_moe_gate_forced): the expert INDICES come from `forced_idx` [N, NB, 2] (e.g. the routing of
// the fp32 reference run) instead of the top-2 search; the gate weights are still this kernel's own softmax
// probabilities of those experts, so LayerNorm, gate GEMV and softmax stay under test.
template <int VPT, int E, int NB, int GATE_WARPS, bool FORCED = false>
__global__ void __launch_bounds__(GATE_WARPS * 32, (GATE_WARPS == 8 && VPT <= 16 && NB * E <= 16 && !FORCED) ? 2 : 1)
moe_gate_kernel(const float* __restrict__ x, long N, int D, const float* __restrict__ ln_w,
                const float* __restrict__ ln_b, const float* __restrict__ gate_w,
                const float* __restrict__ gate_b, int* __restrict__ idx, float* __restrict__ vals,
                float* __restrict__ stats, int* __restrict__ blk_hist, float* __restrict__ blk_imp,
                const int* __restrict__ forced_idx = nullptr) {
  constexpr int G = NB * E;
  constexpr int LG = (G == 32) ? 5 : (G == 16) ? 4 : (G == 8) ? 3 : (G == 4) ? 2 : 1;
  constexpr int TT = 2;
  extern __shared__ float sm[];
  float* gw = sm;               // [G][D] gate weights
  float* lw = gw + G * D;       // [NB][D] LayerNorm weight
  float* lb = lw + NB * D;      // [NB][D] LayerNorm bias
  constexpr int TPW = TOK_PER_BLK / GATE_WARPS;   // tokens per warp
  constexpr bool PREFETCH = GATE_WARPS == 16;     // rows of the next token pair prefetched into registers
  // The 8-warp shape (two blocks per SM, 128 registers per thread) has no registers for that: its next token pair
  // arrives by cp.async in a per-warp shared-memory staging buffer (2 stages x 2 tokens x D floats) while the current
  // pair is evaluated, every lane copying exactly the 16-byte chunks it reads back.  Same arithmetic, same bits; the
  // global-load latency (the kernel ran at 1.1 TB/s of its 51 MB, 17 % warps active) leaves the critical path.
  constexpr bool ASYNC = GATE_WARPS == 8;
  float* my_stg = lb + NB * D + (threadIdx.x >> 5) * (2 * TT * D);
  __shared__ int w_all[GATE_WARPS][MAX_G], w_top1[GATE_WARPS][MAX_G];
  __shared__ float w_imp[GATE_WARPS][MAX_G];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long tok0 = (long)blockIdx.x * TOK_PER_BLK + warp * TPW;
  float nxt[TT][VPT];                              // rows of the next token pair, in flight
  auto fetch = [&](long tok) {
    if (tok < N) {
      load_row<VPT, float>(x + tok * D, lane, nxt[0]);
      load_row<VPT, float>(x + (tok + 1 < N ? tok + 1 : tok) * D, lane, nxt[1]);
    }
  };
  auto issue = [&](long tok, int stage) {
    if (tok < N) {
      const float* s0 = x + tok * D;
      const float* s1 = x + (tok + 1 < N ? tok + 1 : tok) * D;
      float* d0 = my_stg + (stage * TT) * D;
#pragma unroll
      for (int j = 0; j < VPT / 4; ++j) {
        const int o = (j * 32 + lane) * 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(d0 + o)), "l"(s0 + o));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(d0 + D + o)), "l"(s1 + o));
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");     // always: the group count stays uniform
  };
  pdl_enter();
  if (PREFETCH) fetch(tok0);
  if (ASYNC) issue(tok0, 0);
  int stage = 0;
  for (int i = threadIdx.x; i < G * D / 4; i += GATE_WARPS * 32)
    reinterpret_cast<float4*>(gw)[i] = __ldg(reinterpret_cast<const float4*>(gate_w) + i);
  for (int i = threadIdx.x; i < NB * D; i += GATE_WARPS * 32) { lw[i] = ln_w[i]; lb[i] = ln_b[i]; }
  __syncthreads();
  int cnt_all = 0, cnt_top1 = 0;  // lane g owns group g
  float imp = 0.f;
  const int my_t = (lane / NB) % TT, my_br = lane % NB;   // the (token, branch) pair this lane resolves
  for (int it = 0; it < TPW; it += TT) {
    const long tok = tok0 + it;
    if (tok >= N) break;
    const bool two = tok + 1 < N;
    float v[TT][VPT];
    if (ASYNC) {
      issue(it + TT < TPW ? tok + TT : N, stage ^ 1);         // next pair (an empty group past the end)
      asm volatile("cp.async.wait_group 1;" ::: "memory");    // the current pair has landed
#pragma unroll
      for (int t = 0; t < TT; ++t) load_row<VPT, float>(my_stg + (stage * TT + t) * D, lane, nxt[t]);
      stage ^= 1;
    } else if (!PREFETCH) {
      fetch(tok);
    }
#pragma unroll
    for (int t = 0; t < TT; ++t)
#pragma unroll
      for (int i = 0; i < VPT; ++i) v[t][i] = nxt[t][i];
    if (PREFETCH && it + TT < TPW) fetch(tok + TT);
    float mean[TT], rstd[TT];
#pragma unroll
    for (int t = 0; t < TT; ++t) row_stats<VPT>(v[t], D, mean[t], rstd[t]);
    if (lane == 0) {
      *reinterpret_cast<float2*>(stats + tok * 2) = make_float2(mean[0], rstd[0]);
      if (two) *reinterpret_cast<float2*>(stats + (tok + 1) * 2) = make_float2(mean[1], rstd[1]);
    }
    float part[TT][G];
#pragma unroll
    for (int br = 0; br < NB; ++br) {
      float hrow[TT][VPT];
#pragma unroll
      for (int j = 0; j < VPT / 4; ++j) {
        const float4 w4 = *reinterpret_cast<const float4*>(lw + br * D + (j * 32 + lane) * 4);
        const float4 b4 = *reinterpret_cast<const float4*>(lb + br * D + (j * 32 + lane) * 4);
#pragma unroll
        for (int t = 0; t < TT; ++t) {
          hrow[t][4 * j] = (v[t][4 * j] - mean[t]) * rstd[t] * w4.x + b4.x;
          hrow[t][4 * j + 1] = (v[t][4 * j + 1] - mean[t]) * rstd[t] * w4.y + b4.y;
          hrow[t][4 * j + 2] = (v[t][4 * j + 2] - mean[t]) * rstd[t] * w4.z + b4.z;
          hrow[t][4 * j + 3] = (v[t][4 * j + 3] - mean[t]) * rstd[t] * w4.w + b4.w;
        }
      }
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float* wr = gw + (br * E + e) * D;
        float a[TT];
#pragma unroll
        for (int t = 0; t < TT; ++t) a[t] = 0.f;
#pragma unroll
        for (int j = 0; j < VPT / 4; ++j) {
          const float4 w4 = *reinterpret_cast<const float4*>(wr + (j * 32 + lane) * 4);
#pragma unroll
          for (int t = 0; t < TT; ++t) {
            a[t] = fmaf(hrow[t][4 * j], w4.x, a[t]);
            a[t] = fmaf(hrow[t][4 * j + 1], w4.y, a[t]);
            a[t] = fmaf(hrow[t][4 * j + 2], w4.z, a[t]);
            a[t] = fmaf(hrow[t][4 * j + 3], w4.w, a[t]);
          }
        }
#pragma unroll
        for (int t = 0; t < TT; ++t) part[t][br * E + e] = a[t];
      }
    }
#pragma unroll
    for (int t = 0; t < TT; ++t) warp_reduce_scatter<G>(part[t], lane);  // lane (g << (5-LG)) holds group g
    float logits[E], probs[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int src = (my_br * E + e) << (5 - LG);
      const float d0 = __shfl_sync(0xffffffffu, part[0][0], src);
      const float d1 = __shfl_sync(0xffffffffu, part[1][0], src);
      logits[e] = (my_t ? d1 : d0) + __ldg(gate_b + my_br * E + e);
    }
    int i0, i1;
    float v0, v1;
    softmax_top2<E>(logits, probs, i0, i1, v0, v1);
    if (FORCED && lane < TT * NB && (my_t == 0 || two)) {
      const int2 f = *reinterpret_cast<const int2*>(forced_idx + ((tok + my_t) * NB + my_br) * 2);
      i0 = min(max(f.x, 0), E - 1); i1 = min(max(f.y, 0), E - 1);
#pragma unroll
      for (int e = 0; e < E; ++e) { if (e == i0) v0 = probs[e]; if (e == i1) v1 = probs[e]; }
    }
    if (lane < TT * NB && (my_t == 0 || two)) {
      const long o = ((tok + my_t) * NB + my_br) * 2;
      *reinterpret_cast<int2*>(idx + o) = make_int2(i0, i1);
      *reinterpret_cast<float2*>(vals + o) = make_float2(v0, v1);
    }
    const int mg0 = my_br * E + i0, mg1 = my_br * E + i1;
#pragma unroll
    for (int q = 0; q < TT * NB; ++q) {   // token-major, branch-minor: the accumulation order of v1
      const int g0 = __shfl_sync(0xffffffffu, mg0, q), g1 = __shfl_sync(0xffffffffu, mg1, q);
      const float a0 = __shfl_sync(0xffffffffu, v0, q), a1 = __shfl_sync(0xffffffffu, v1, q);
      if (q / NB == 0 || two) {
        if (lane == g0) { cnt_all++; cnt_top1++; imp += a0; }
        if (lane == g1) { cnt_all++; imp += a1; }
      }
    }
  }
  if (ASYNC) asm volatile("cp.async.wait_group 0;" ::: "memory");
  w_all[warp][lane] = cnt_all;
  w_top1[warp][lane] = cnt_top1;
  w_imp[warp][lane] = imp;
  __syncthreads();
  if (threadIdx.x < G) {
    int a = 0, t1 = 0;
    float im = 0.f;
    for (int w = 0; w < GATE_WARPS; ++w) { a += w_all[w][threadIdx.x]; t1 += w_top1[w][threadIdx.x]; im += w_imp[w][threadIdx.x]; }
    blk_hist[((long)blockIdx.x * 2) * G + threadIdx.x] = a;
    blk_hist[((long)blockIdx.x * 2 + 1) * G + threadIdx.x] = t1;
    blk_imp[(long)blockIdx.x * G + threadIdx.x] = im;
  }
}

// One warp per expert group: lane l owns a contiguous run of blocks (two passes: run totals, warp
// exclusive scan, per-block bases), then warp 0 does the segment scan over the groups.  (The first
// version walked the 196 blocks of every group with a single lane: 15-20 us per call.)
// STAGED: the block histograms / importance sums are first copied into shared memory with coalesced 16-byte loads.
// The lanes' own accesses are 4-byte reads 4 G bytes (x per) apart, i.e. one 32-byte sector each: ~350 KB of sector
// traffic into this one SM for 38 KB of data at N = 25 088 (12.4 us per call, 16 serialised calls per step).
template <bool STAGED>
__global__ void __launch_bounds__(1024)
moe_scan_kernel(const int* __restrict__ blk_hist_g, const float* __restrict__ blk_imp_g, int nblk, int G, int E,
                int F, int D, int* __restrict__ blk_base, int* __restrict__ seg_offsets,
                MTile* __restrict__ tiles_up, MTile* __restrict__ tiles_down, int* __restrict__ num_tiles,
                float* __restrict__ usage, float* __restrict__ importance) {
  __shared__ int total_s[MAX_G];
  extern __shared__ int4 scan_stage[];
  const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_enter();
  const int* blk_hist = blk_hist_g;
  const float* blk_imp = blk_imp_g;
  int hp = 2 * G, ip = G;                 // row pitches (ints / floats) of the two tables as read below
  if constexpr (STAGED) {
    // odd pitches + an odd number of blocks per lane: the lanes of a warp (runs of `per` rows) hit 32 different banks
    hp = 2 * G + 1; ip = G + 1;
    int* hs = reinterpret_cast<int*>(scan_stage);
    float* is = reinterpret_cast<float*>(hs + nblk * hp);
    const int nh = nblk * 2 * G / 4, ni = nblk * G / 4;          // int4 / float4 counts (G % 4 == 0: both divide)
    // (four + two 16-byte loads per thread in flight before the first store: one round trip for N <= 32 768)
    const int nthr = blockDim.x;
    for (int i0 = threadIdx.x; i0 < nh; i0 += 4 * nthr) {
      int4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i0 + u * nthr < nh) v[u] = reinterpret_cast<const int4*>(blk_hist_g)[i0 + u * nthr];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * nthr;
        if (i < nh) {
          const int e = 4 * i, row = e / (2 * G), col = e - row * 2 * G;
          int* d = hs + row * hp + col;
          d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w;
        }
      }
    }
    for (int i0 = threadIdx.x; i0 < ni; i0 += 2 * nthr) {
      float4 v[2];
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (i0 + u * nthr < ni) v[u] = reinterpret_cast<const float4*>(blk_imp_g)[i0 + u * nthr];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int i = i0 + u * nthr;
        if (i < ni) {
          const int e = 4 * i, row = e / G, col = e - row * G;
          float* d = is + row * ip + col;
          d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w;
        }
      }
    }
    __syncthreads();
    blk_hist = hs;
    blk_imp = is;
  }
  if (g < G) {
    int per = (nblk + 31) / 32;
    if (STAGED) per |= 1;
    const int b0 = lane * per, b1 = min(nblk, b0 + per);
    int tot = 0, top1 = 0;
    float imp = 0.f;
    // up to 8 blocks per lane (N <= 32 768 tokens): every load of the lane is issued at once and the histogram stays in
    // registers for the second pass - one L2 round trip instead of four dependent ones in this single-CTA kernel
    constexpr int PER_FAST = 8;
    int hreg[PER_FAST];
    const bool fast = MDM_SCAN_FAST && per <= PER_FAST;
    if (fast) {
      int t1[PER_FAST];
      float im[PER_FAST];
#pragma unroll
      for (int j = 0; j < PER_FAST; ++j) {
        const int b = b0 + j;
        const bool ok = j < per && b < b1;
        hreg[j] = ok ? blk_hist[(long)b * hp + g] : 0;
        t1[j] = ok ? blk_hist[(long)b * hp + G + g] : 0;
        im[j] = ok ? blk_imp[(long)b * ip + g] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < PER_FAST; ++j) {
        tot += hreg[j]; top1 += t1[j];
        if (j < per && b0 + j < b1) imp += im[j];      // same additions in the same order as the loop below
      }
    } else {
      for (int b = b0; b < b1; ++b) {
        tot += blk_hist[(long)b * hp + g];
        top1 += blk_hist[(long)b * hp + G + g];
        imp += blk_imp[(long)b * ip + g];
      }
    }
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    int run = incl - tot;
    if (fast) {
#pragma unroll
      for (int j = 0; j < PER_FAST; ++j) {
        const int b = b0 + j;
        if (j < per && b < b1) { blk_base[(long)b * G + g] = run; run += hreg[j]; }
      }
    } else {
      for (int b = b0; b < b1; ++b) {
        blk_base[(long)b * G + g] = run;
        run += blk_hist[(long)b * hp + g];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) top1 += __shfl_xor_sync(0xffffffffu, top1, o);
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    // importance: per-lane runs in block order, then the 32 run sums in lane order (a fixed order:
    // the counter is deterministic from launch to launch)
    float imp_tot = 0.f;
#pragma unroll
    for (int l = 0; l < 32; ++l) imp_tot += __shfl_sync(0xffffffffu, imp, l);
    if (lane == 0) {
      total_s[g] = total;
      if (usage) usage[g] += (float)top1;
      if (importance) importance[g] += imp_tot;
    }
  }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int gg = threadIdx.x;
  const int total = gg < G ? total_s[gg] : 0;
  const int padded = ((total + 127) / 128) * 128;
  int incl = padded;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (gg >= o) incl += n;
  }
  const int off = incl - padded;
  if (gg < G) {
    seg_offsets[gg] = off;
    const int ntile = padded / 128, tile0 = off / 128;
    for (int i = 0; i < ntile; ++i) {
      const int rows = min(128, total - i * 128);
      MTile u; u.a_row0 = off + i * 128; u.c_row0 = u.a_row0; u.w_row0 = gg * F; u.rows_valid = rows;
      MTile d = u; d.w_row0 = gg * D;
      tiles_up[tile0 + i] = u;
      tiles_down[tile0 + i] = d;
    }
  }
  const int all = __shfl_sync(0xffffffffu, incl, 31);
  if (gg == 0) { seg_offsets[G] = all; *num_tiles = all / 128; }
}

constexpr int PERM_WARPS = 16;
// A 128-token block of the gate's histogram is moved by PERM_SPLIT CTAs: each recomputes the block's ranks (512 index
// loads + one match per pair) and moves a quarter of its rows.  One CTA per block was 196 CTAs of 512 threads at
// N = 25 088, i.e. two waves with the second one third full, and 13 CTAs at 8 sequences per GPU.
constexpr int PERM_SPLIT = MDM_PERM_SPLIT;
template <int VPT, typename TO>
__global__ void __launch_bounds__(PERM_WARPS * 32, (VPT <= 16 && MDM_PERM_SPLIT > 1) ? 2 : 1)   // two CTAs per SM up to D = 512 (64 registers)
moe_permute_kernel(const float* __restrict__ x, long N, int D, int NB, int E, const float* __restrict__ ln_w,
                   const float* __restrict__ ln_b, const int* __restrict__ idx, const float* __restrict__ vals,
                   const float* __restrict__ stats, const int* __restrict__ blk_base,
                   const int* __restrict__ seg_offsets, TO* __restrict__ xp, int* __restrict__ perm,
                   float* __restrict__ rowscale) {
  // pairs of a block: p = token_local * NBK + slot, NBK = NB*2; 32 pairs per segment
  __shared__ int seg_cnt[16][MAX_G];
  __shared__ int pos_s[TOK_PER_BLK * 4];
  const int G = NB * E, NBK = NB * 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npairs = TOK_PER_BLK * NBK;           // <= 512
  const int nseg = npairs / 32;                   // <= 16
  pdl_enter();
  for (int i = threadIdx.x; i < 16 * MAX_G; i += PERM_WARPS * 32) (&seg_cnt[0][0])[i] = 0;
  __syncthreads();
  const int blk = blockIdx.x / PERM_SPLIT, part = blockIdx.x % PERM_SPLIT;
  const long tok_blk0 = (long)blk * TOK_PER_BLK;
  constexpr int PART_TOK = TOK_PER_BLK / PERM_SPLIT;   // tokens whose rows this CTA moves
  constexpr int SPW = 16 / PERM_WARPS;            // 32-pair segments per warp
  int my_g[SPW], my_rank[SPW];
  for (int r = 0; r < SPW; ++r) {
    const int seg = warp * SPW + r;
    my_g[r] = -1; my_rank[r] = 0;
    if (seg < nseg) {
      const int p = seg * 32 + lane;
      const long tok = tok_blk0 + p / NBK;
      const int slot = p % NBK;
      int g = -1;
      if (tok < N) g = (slot >> 1) * E + idx[tok * NBK + slot];
      const unsigned peers = __match_any_sync(0xffffffffu, g);
      my_g[r] = g;
      my_rank[r] = __popc(peers & ((1u << lane) - 1u));
      if (g >= 0 && my_rank[r] == 0) seg_cnt[seg][g] = __popc(peers);
    }
  }
  __syncthreads();
  for (int r = 0; r < SPW; ++r) {
    const int seg = warp * SPW + r;
    if (seg < nseg && my_g[r] >= 0) {
      const int g = my_g[r];
      int base = 0;
      for (int s = 0; s < seg; ++s) base += seg_cnt[s][g];
      const int p = seg * 32 + lane;
      const long tok = tok_blk0 + p / NBK;
      const int slot = p % NBK;
      const int pos = seg_offsets[g] + blk_base[(long)blk * G + g] + base + my_rank[r];
      pos_s[p] = pos;
      if ((p / NBK) / PART_TOK == part) {          // every pair is published once, by the CTA that moves its row
        perm[tok * NBK + slot] = pos;
        rowscale[pos] = vals[tok * NBK + slot] / (float)NB;
      }
    }
  }
  __syncthreads();
  constexpr int TPW = PART_TOK / PERM_WARPS, UNR = TPW < 4 ? TPW : 4;   // tokens per warp; rows in flight per warp
  static_assert(TPW >= 1 && TPW * PERM_WARPS * PERM_SPLIT == TOK_PER_BLK, "token split");
#pragma unroll 1
  for (int it = 0; it < TPW; it += UNR) {
    float v[UNR][VPT];
    float2 ms[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long tok = tok_blk0 + part * PART_TOK + warp * TPW + it + u;
      if (tok < N) {
        load_row<VPT, float>(x + tok * D, lane, v[u]);
        ms[u] = *reinterpret_cast<const float2*>(stats + tok * 2);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int tl = part * PART_TOK + warp * TPW + it + u;
      if (tok_blk0 + tl >= N) break;
      for (int br = 0; br < NB; ++br) {
        float hrow[VPT];
#pragma unroll
        for (int i = 0; i < VPT; ++i) hrow[i] = v[u][i];
        affine_row<VPT>(hrow, ms[u].x, ms[u].y, ln_w + br * D, ln_b + br * D, lane);
        store_row<VPT, TO>(xp + (long)pos_s[tl * NBK + br * 2] * D, lane, hrow);
        store_row<VPT, TO>(xp + (long)pos_s[tl * NBK + br * 2 + 1] * D, lane, hrow);
      }
    }
  }
}

template <int VPT, typename TI>
__global__ void __launch_bounds__(256)
moe_combine_film_kernel(const TI* __restrict__ yp, const int* __restrict__ perm, long N, int D, int NBK,
                        const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                        const float* __restrict__ film, int rows_per_seq, TI* __restrict__ out) {
  const long tok = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  pdl_enter();
  if (tok >= N) return;
  const int lane = threadIdx.x & 31;
  float acc[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) acc[i] = 0.f;
  for (int br = 0; br < NBK / 2; ++br) {
    float a[VPT], b[VPT];
    load_row<VPT, TI>(yp + (long)perm[tok * NBK + br * 2] * D, lane, a);
    load_row<VPT, TI>(yp + (long)perm[tok * NBK + br * 2 + 1] * D, lane, b);
#pragma unroll
    for (int i = 0; i < VPT; ++i) acc[i] += a[i] + b[i];
  }
  layernorm_row<VPT>(acc, ln_w, ln_b, lane, D);
  film_row<VPT>(acc, film + (tok / rows_per_seq) * 2 * D, lane, D);
#pragma unroll
  for (int i = 0; i < VPT; ++i) acc[i] = silu_out<TI>(acc[i]);
  store_row<VPT, TI>(out + tok * D, lane, acc);
}

template <int E>
__global__ void softmax_topk_kernel(const float* __restrict__ logits, long N, float* __restrict__ probs,
                                    int64_t* __restrict__ idx64, float* __restrict__ vals) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float l[E], p[E];
#pragma unroll
  for (int e = 0; e < E; ++e) l[e] = logits[i * E + e];
  int i0, i1;
  float v0, v1;
  softmax_top2<E>(l, p, i0, i1, v0, v1);
  if (probs) {
#pragma unroll
    for (int e = 0; e < E; ++e) probs[i * E + e] = p[e];
  }
  idx64[i * 2] = i0; idx64[i * 2 + 1] = i1;
  vals[i * 2] = v0; vals[i * 2 + 1] = v1;
}

}  // namespace

#define VPT_SWITCH(D, ...)               \
  switch (D) {                           \
    case 128: { constexpr int V = 4; __VA_ARGS__; break; }  \
    case 256: { constexpr int V = 8; __VA_ARGS__; break; }  \
    case 512: { constexpr int V = 16; __VA_ARGS__; break; } \
    case 1024: { constexpr int V = 32; __VA_ARGS__; break; }\
    default: return MDM_ERR_UNSUPPORTED; \
  }

template <int VPT, int E, int NB>
int launch_gate(const float* x, long N, int D, const float* ln_w, const float* ln_b, const float* gate_w,
                const float* gate_b, int* idx, float* vals, float* stats, int* blk_hist, float* blk_imp,
                cudaStream_t st, const int* forced_idx = nullptr) {
  const int nblk = (int)((N + TOK_PER_BLK - 1) / TOK_PER_BLK);
  const size_t smem = sizeof(float) * ((size_t)NB * E * D + 2 * (size_t)NB * D);
  const size_t smem8 = smem + sizeof(float) * 8 * 2 * 2 * (size_t)D;     // + cp.async staging: 8 warps x 2 stages x 2 tokens
  if (forced_idx) {
    if (smem8 > 48 * 1024 && cudaFuncSetAttribute(moe_gate_kernel<VPT, E, NB, 8, true>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8) != cudaSuccess)
      return MDM_ERR_CUDA;
    mdm_launch(moe_gate_kernel<VPT, E, NB, 8, true>, nblk, 256, smem8, st, x, N, D, ln_w, ln_b, gate_w, gate_b, idx, vals, stats,
                                                                   blk_hist, blk_imp, forced_idx);
    return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
  }
  if (smem8 > 48 * 1024 &&
      (cudaFuncSetAttribute(moe_gate_kernel<VPT, E, NB, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8) !=
           cudaSuccess ||
       cudaFuncSetAttribute(moe_gate_kernel<VPT, E, NB, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
           cudaSuccess))
    return MDM_ERR_CUDA;
  const int sms = mdm_num_sms();
  if (nblk <= sms)
    mdm_launch(moe_gate_kernel<VPT, E, NB, 16>, nblk, 512, smem, st, x, N, D, ln_w, ln_b, gate_w, gate_b, idx, vals, stats,
                                                             blk_hist, blk_imp, (const int*)nullptr);
  else
    mdm_launch(moe_gate_kernel<VPT, E, NB, 8>, nblk, 256, smem8, st, x, N, D, ln_w, ln_b, gate_w, gate_b, idx, vals, stats,
                                                             blk_hist, blk_imp, (const int*)nullptr);
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

template <int VPT>
int gate_dispatch(int NB, int E, const float* x, long N, int D, const float* ln_w, const float* ln_b,
                  const float* gate_w, const float* gate_b, int* idx, float* vals, float* stats, int* blk_hist,
                  float* blk_imp, cudaStream_t st, const int* forced_idx = nullptr) {
#define GATE_CASE(e, nb) \
  if (E == e && NB == nb) return launch_gate<VPT, e, nb>(x, N, D, ln_w, ln_b, gate_w, gate_b, idx, vals, stats, blk_hist, blk_imp, st, forced_idx);
  GATE_CASE(2, 1) GATE_CASE(4, 1) GATE_CASE(8, 1) GATE_CASE(16, 1)
  GATE_CASE(2, 2) GATE_CASE(4, 2) GATE_CASE(8, 2) GATE_CASE(16, 2)
#undef GATE_CASE
  return MDM_ERR_UNSUPPORTED;  // the expert count must be a power of two <= 16 (reference uses 4 / 8)
}

static int moe_gate_impl(const float* x, long N, int D, int NB, int E, int K, const float* ln_w,
                         const float* ln_b, const float* gate_w, const float* gate_b, int* idx,
                         float* vals, float* stats, int* blk_hist, float* blk_imp, const int* forced, void* stream) {
  if (!x || !ln_w || !ln_b || !gate_w || !gate_b || !idx || !vals || !stats || !blk_hist || !blk_imp)
    return MDM_ERR_ARG;
  if (K != 2 || E < 2 || E > MAX_E || NB * E > MAX_G || NB < 1) return MDM_ERR_UNSUPPORTED;
  if (N == 0) return MDM_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (D) {
    case 128: return gate_dispatch<4>(NB, E, x, N, D, ln_w, ln_b, gate_w, gate_b, idx, vals, stats, blk_hist, blk_imp, st, forced);
    case 256: return gate_dispatch<8>(NB, E, x, N, D, ln_w, ln_b, gate_w, gate_b, idx, vals, stats, blk_hist, blk_imp, st, forced);
    case 512: return gate_dispatch<16>(NB, E, x, N, D, ln_w, ln_b, gate_w, gate_b, idx, vals, stats, blk_hist, blk_imp, st, forced);
    case 1024: return gate_dispatch<32>(NB, E, x, N, D, ln_w, ln_b, gate_w, gate_b, idx, vals, stats, blk_hist, blk_imp, st, forced);
    default: return MDM_ERR_UNSUPPORTED;
  }
}

extern "C" MDM_API int mdm_moe_gate(const float* x, long N, int D, int NB, int E, int K, const float* ln_w,
                                    const float* ln_b, const float* gate_w, const float* gate_b, int* idx,
                                    float* vals, float* stats, int* blk_hist, float* blk_imp, void* stream) {
  return moe_gate_impl(x, N, D, NB, E, K, ln_w, ln_b, gate_w, gate_b, idx, vals, stats, blk_hist, blk_imp, nullptr, stream);
}

extern "C" MDM_API int mdm_moe_gate_forced(const float* x, long N, int D, int NB, int E, int K, const float* ln_w,
                                           const float* ln_b, const float* gate_w, const float* gate_b,
                                           const int* forced_idx, int* idx, float* vals, float* stats, int* blk_hist,
                                           float* blk_imp, void* stream) {
  if (!forced_idx) return MDM_ERR_ARG;
  return moe_gate_impl(x, N, D, NB, E, K, ln_w, ln_b, gate_w, gate_b, idx, vals, stats, blk_hist, blk_imp, forced_idx, stream);
}

extern "C" MDM_API int mdm_moe_scan(const int* blk_hist, const float* blk_imp, const int* idx, long N, int NB,
                                    int E, int K, int F, int D, int* blk_base, int* seg_offsets,
                                    void* tiles_up, void* tiles_down, int* num_tiles, float* usage,
                                    float* importance, void* stream) {
  (void)idx;
  if (!blk_hist || !blk_imp || !blk_base || !seg_offsets || !tiles_up || !tiles_down || !num_tiles)
    return MDM_ERR_ARG;
  if (K != 2 || NB * E > MAX_G) return MDM_ERR_UNSUPPORTED;
  const int nblk = (int)((N + TOK_PER_BLK - 1) / TOK_PER_BLK);
  const int G = NB * E;
  const size_t stage = (size_t)nblk * (3 * G + 2) * 4;
  const bool aligned = ((reinterpret_cast<uintptr_t>(blk_hist) | reinterpret_cast<uintptr_t>(blk_imp)) & 15) == 0 && (G & 3) == 0;
  static unsigned long long attr = 0;   // one bit per device ordinal
  const unsigned long long dev_bit = 1ull << mdm_cur_dev();
  bool staged = MDM_SCAN_FAST && aligned && stage <= 200 * 1024;
  if (staged && stage > 40 * 1024 && !(attr & dev_bit)) {
    if (cudaFuncSetAttribute(moe_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess) attr |= dev_bit;
    else { (void)cudaGetLastError(); staged = false; }
  }
  if (staged)
    mdm_launch(moe_scan_kernel<true>, 1, 32 * G, stage, reinterpret_cast<cudaStream_t>(stream), blk_hist, blk_imp, nblk, G, E, F, D, blk_base,
               seg_offsets, reinterpret_cast<MTile*>(tiles_up), reinterpret_cast<MTile*>(tiles_down), num_tiles, usage, importance);
  else
    mdm_launch(moe_scan_kernel<false>, 1, 32 * G, 0, reinterpret_cast<cudaStream_t>(stream), blk_hist, blk_imp, nblk, G, E, F, D, blk_base,
               seg_offsets, reinterpret_cast<MTile*>(tiles_up), reinterpret_cast<MTile*>(tiles_down), num_tiles, usage, importance);
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_moe_permute(const float* x, long N, int D, int NB, int E, int K, const float* ln_w,
                                       const float* ln_b, const int* idx, const float* vals, const float* stats,
                                       const int* blk_base, const int* seg_offsets, void* xp, int dt,
                                       int* perm, float* rowscale, void* stream) {
  if (!x || !idx || !vals || !stats || !blk_base || !seg_offsets || !xp || !perm || !rowscale) return MDM_ERR_ARG;
  if (K != 2 || NB < 1 || NB > 2 || NB * E > MAX_G) return MDM_ERR_UNSUPPORTED;
  if (N == 0) return MDM_OK;
  const int nblk = (int)((N + TOK_PER_BLK - 1) / TOK_PER_BLK);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VPT_SWITCH(D, {
    if (dt == MDM_F32)
      mdm_launch(moe_permute_kernel<V, float>, nblk * PERM_SPLIT, PERM_WARPS * 32, 0, st, x, N, D, NB, E, ln_w, ln_b, idx, vals, stats, blk_base,
                                                          seg_offsets, reinterpret_cast<float*>(xp), perm, rowscale);
    else
      mdm_launch(moe_permute_kernel<V, bf16>, nblk * PERM_SPLIT, PERM_WARPS * 32, 0, st, x, N, D, NB, E, ln_w, ln_b, idx, vals, stats, blk_base,
                                                         seg_offsets, reinterpret_cast<bf16*>(xp), perm, rowscale);
  });
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_moe_combine_film(const void* yp, int dt, const int* perm, long N, int D, int NBK,
                                            const float* ln_w, const float* ln_b, const float* film,
                                            int rows_per_seq, void* out, void* stream) {
  if (!yp || !perm || !ln_w || !ln_b || !film || !out || rows_per_seq <= 0) return MDM_ERR_ARG;
  if (NBK < 2 || (NBK & 1)) return MDM_ERR_UNSUPPORTED;
  if (N == 0) return MDM_OK;
  const unsigned grid = (unsigned)((N + 7) / 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VPT_SWITCH(D, {
    if (dt == MDM_F32)
      mdm_launch(moe_combine_film_kernel<V, float>, grid, 256, 0, st, reinterpret_cast<const float*>(yp), perm, N, D, NBK,
                                                               ln_w, ln_b, film, rows_per_seq,
                                                               reinterpret_cast<float*>(out));
    else
      mdm_launch(moe_combine_film_kernel<V, bf16>, grid, 256, 0, st, reinterpret_cast<const bf16*>(yp), perm, N, D, NBK,
                                                              ln_w, ln_b, film, rows_per_seq,
                                                              reinterpret_cast<bf16*>(out));
  });
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_softmax_topk(const float* logits, long N, int E, int K, float* probs, int64_t* idx64,
                                        float* vals, void* stream) {
  if (!logits || !idx64 || !vals) return MDM_ERR_ARG;
  if (K != 2) return MDM_ERR_UNSUPPORTED;
  if (N == 0) return MDM_OK;
  const unsigned grid = (unsigned)((N + 255) / 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (E) {
    case 2: softmax_topk_kernel<2><<<grid, 256, 0, st>>>(logits, N, probs, idx64, vals); break;
    case 4: softmax_topk_kernel<4><<<grid, 256, 0, st>>>(logits, N, probs, idx64, vals); break;
    case 8: softmax_topk_kernel<8><<<grid, 256, 0, st>>>(logits, N, probs, idx64, vals); break;
    case 16: softmax_topk_kernel<16><<<grid, 256, 0, st>>>(logits, N, probs, idx64, vals); break;
    default: return MDM_ERR_UNSUPPORTED;
  }
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}
