// Row-wise fused normalisation pipeline: one warp per row, everything in registers, 16-byte accesses.
// Covers every nn.LayerNorm, the F.normalize * sqrt(D) of PerformerSelfAttention and StylizationBlock's
// FiLM + SiLU (reference: models/stylization.py:29-30; models/fast_attention.py:142,169-172,210,225,248;
// models/multi_branch.py:55).  HBM-bound: algorithmic bytes = D * (in + sum of outputs) per row.
#include "common.cuh"
#include "rowmath.cuh"

namespace {

// Persistent grid-stride version: a warp keeps the (up to two) LayerNorm affine vectors in registers
// for all of its rows, prefetches its next row while it works on the current one, and in the
// bf16-operand mode evaluates SiLU on the MUFU units.  The first version re-read 6 parameter vectors
// through L1 for every row and spent more issue slots on expf / IEEE divides than on the row itself.
template <int VPT>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, int lane, float (&r)[VPT]) {
#pragma unroll
  for (int j = 0; j < VPT / 4; ++j) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p + (j * 32 + lane) * 4));
    r[4 * j] = t.x; r[4 * j + 1] = t.y; r[4 * j + 2] = t.z; r[4 * j + 3] = t.w;
  }
}
// mean / rstd with pairwise (tree) partial sums: same mathematics as row_stats(), 4 dependent adds
// instead of VPT.  Only used by the bf16-operand specialisations (fp32 mode keeps row_stats()).
template <int VPT>
__device__ __forceinline__ float tree_sum(const float (&a)[VPT]) {
  float t[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) t[i] = a[i];
#pragma unroll
  for (int n = VPT / 2; n >= 1; n >>= 1)
#pragma unroll
    for (int i = 0; i < n; ++i) t[i] += t[i + n];
  return t[0];
}
// Packed (FFMA2 / FMUL2 / FADD2) versions for the bf16-operand specialisations: the row kernels are issue-bound
// (~650 instructions per row for the five-stage pipeline), two elements per issue slot.
template <int VPT>
__device__ __forceinline__ float2 pair_sum(const float (&a)[VPT]) {       // two interleaved partial sums, tree order
  float2 t[VPT / 2];
#pragma unroll
  for (int i = 0; i < VPT / 2; ++i) t[i] = make_float2(a[2 * i], a[2 * i + 1]);
#pragma unroll
  for (int n = VPT / 4; n >= 1; n >>= 1)
#pragma unroll
    for (int i = 0; i < n; ++i) t[i] = add2(t[i], t[i + n]);
  return t[0];
}
template <int VPT>
__device__ __forceinline__ void ln_regs_packed(float (&v)[VPT], const float* __restrict__ w_s, const float* __restrict__ b_s,
                                               int lane, int D) {
  const float inv_d = 1.0f / (float)D;
  const float2 s2 = pair_sum<VPT>(v);
  const float mean = warp_sum(s2.x + s2.y) * inv_d;
  const float2 nm = make_float2(-mean, -mean);
  float2 q2 = make_float2(0.f, 0.f), q3 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < VPT / 2; i += 2) {
    const float2 d0 = add2(make_float2(v[2 * i], v[2 * i + 1]), nm), d1 = add2(make_float2(v[2 * i + 2], v[2 * i + 3]), nm);
    v[2 * i] = d0.x; v[2 * i + 1] = d0.y; v[2 * i + 2] = d1.x; v[2 * i + 3] = d1.y;
    q2 = fma2(d0, d0, q2);
    q3 = fma2(d1, d1, q3);
  }
  const float rstd = rsqrtf(warp_sum((q2.x + q3.x) + (q2.y + q3.y)) * inv_d + 1e-5f);
  const float2 r2 = make_float2(rstd, rstd);
#pragma unroll
  for (int j = 0; j < VPT / 4; ++j) {
    const float4 w4 = *reinterpret_cast<const float4*>(w_s + (j * 32 + lane) * 4);
    const float4 b4 = *reinterpret_cast<const float4*>(b_s + (j * 32 + lane) * 4);
    const float2 y0 = fma2(make_float2(v[4 * j], v[4 * j + 1]), mul2(make_float2(w4.x, w4.y), r2), make_float2(b4.x, b4.y));
    const float2 y1 = fma2(make_float2(v[4 * j + 2], v[4 * j + 3]), mul2(make_float2(w4.z, w4.w), r2), make_float2(b4.z, b4.w));
    v[4 * j] = y0.x; v[4 * j + 1] = y0.y; v[4 * j + 2] = y1.x; v[4 * j + 3] = y1.y;
  }
}
template <int VPT>
__device__ __forceinline__ void l2norm_packed(float (&v)[VPT], int D) {
  float2 q2 = make_float2(0.f, 0.f), q3 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < VPT / 2; i += 2) {
    const float2 a = make_float2(v[2 * i], v[2 * i + 1]), b = make_float2(v[2 * i + 2], v[2 * i + 3]);
    q2 = fma2(a, a, q2);
    q3 = fma2(b, b, q3);
  }
  const float denom = fmaxf(sqrtf(warp_sum((q2.x + q3.x) + (q2.y + q3.y))), 1e-12f);
  const float sc = sqrtf((float)D) / denom;
  const float2 s2 = make_float2(sc, sc);
#pragma unroll
  for (int i = 0; i < VPT / 2; ++i) {
    const float2 y = mul2(make_float2(v[2 * i], v[2 * i + 1]), s2);
    v[2 * i] = y.x; v[2 * i + 1] = y.y;
  }
}
template <int VPT>
__device__ __forceinline__ void film_packed(float (&v)[VPT], const float* __restrict__ film, int lane, int D) {
  const float2 one = make_float2(1.f, 1.f);
#pragma unroll
  for (int j = 0; j < VPT / 4; ++j) {
    const float4 sc = *reinterpret_cast<const float4*>(film + (j * 32 + lane) * 4);
    const float4 sh = *reinterpret_cast<const float4*>(film + D + (j * 32 + lane) * 4);
    const float2 y0 = fma2(make_float2(v[4 * j], v[4 * j + 1]), add2(make_float2(sc.x, sc.y), one), make_float2(sh.x, sh.y));
    const float2 y1 = fma2(make_float2(v[4 * j + 2], v[4 * j + 3]), add2(make_float2(sc.z, sc.w), one), make_float2(sh.z, sh.w));
    v[4 * j] = y0.x; v[4 * j + 1] = y0.y; v[4 * j + 2] = y1.x; v[4 * j + 3] = y1.y;
  }
}
template <int VPT>
__device__ __forceinline__ void silu_packed(float (&v)[VPT]) {
  const float2 nl2e = make_float2(-1.4426950408889634f, -1.4426950408889634f), one = make_float2(1.f, 1.f);
#pragma unroll
  for (int i = 0; i < VPT / 2; ++i) {
    const float2 x = make_float2(v[2 * i], v[2 * i + 1]);
    const float2 a = mul2(x, nl2e);
    float2 e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(a.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(a.y));
    e = add2(e, one);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(e.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(e.y));
    const float2 y = mul2(x, r);
    v[2 * i] = y.x; v[2 * i + 1] = y.y;
  }
}

template <int VPT, bool TREE>
__device__ __forceinline__ void ln_regs(float (&v)[VPT], const float* __restrict__ w_s, const float* __restrict__ b_s,
                                        int lane, int D) {
  if (TREE) {
    ln_regs_packed<VPT>(v, w_s, b_s, lane, D);
    return;
  }
  float mean, rstd;
  if (TREE) {
    const float inv_d = 1.0f / (float)D;
    mean = warp_sum(tree_sum<VPT>(v)) * inv_d;
    float d2[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) { const float d = v[i] - mean; d2[i] = d * d; }
    rstd = rsqrtf(warp_sum(tree_sum<VPT>(d2)) * inv_d + 1e-5f);
  } else {
    row_stats<VPT>(v, D, mean, rstd);
  }
#pragma unroll
  for (int j = 0; j < VPT / 4; ++j) {   // affine vectors live in shared memory (16-byte, conflict-free reads)
    const float4 w4 = *reinterpret_cast<const float4*>(w_s + (j * 32 + lane) * 4);
    const float4 b4 = *reinterpret_cast<const float4*>(b_s + (j * 32 + lane) * 4);
    v[4 * j] = (v[4 * j] - mean) * rstd * w4.x + b4.x;
    v[4 * j + 1] = (v[4 * j + 1] - mean) * rstd * w4.y + b4.y;
    v[4 * j + 2] = (v[4 * j + 2] - mean) * rstd * w4.z + b4.z;
    v[4 * j + 3] = (v[4 * j + 3] - mean) * rstd * w4.w + b4.w;
  }
}
__device__ __forceinline__ float silu_mufu(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-x * 1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}

// one row in its memory representation, spread over the warp like load_row()
template <int VPT, typename T> struct RawRow;
template <int VPT> struct RawRow<VPT, float> {
  float4 r[VPT / 4];
  __device__ __forceinline__ void load(const float* __restrict__ p, int lane) {
#pragma unroll
    for (int j = 0; j < VPT / 4; ++j) r[j] = *reinterpret_cast<const float4*>(p + (j * 32 + lane) * 4);
  }
  __device__ __forceinline__ void unpack(float (&v)[VPT]) const {
#pragma unroll
    for (int j = 0; j < VPT / 4; ++j) { v[4 * j] = r[j].x; v[4 * j + 1] = r[j].y; v[4 * j + 2] = r[j].z; v[4 * j + 3] = r[j].w; }
  }
};
template <int VPT> struct RawRow<VPT, bf16> {
  uint2 r[VPT / 4];
  __device__ __forceinline__ void load(const bf16* __restrict__ p, int lane) {
#pragma unroll
    for (int j = 0; j < VPT / 4; ++j) r[j] = *reinterpret_cast<const uint2*>(p + (j * 32 + lane) * 4);
  }
  __device__ __forceinline__ void unpack(float (&v)[VPT]) const {
#pragma unroll
    for (int j = 0; j < VPT / 4; ++j) {
      const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&r[j].x);
      const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&r[j].y);
      v[4 * j] = __low2float(a); v[4 * j + 1] = __high2float(a);
      v[4 * j + 2] = __low2float(b); v[4 * j + 3] = __high2float(b);
    }
  }
};

enum { F_LN1 = 1, F_L2 = 2, F_LN2 = 4, F_FILM = 8, F_SILU = 16, F_RUNTIME = -1 };

// FLAGS >= 0: the stage set is a compile-time constant (no speculative work: the runtime-flag version
// executed 584 instructions per row for a plain LayerNorm because the compiler if-converted the
// skipped stages); FLAGS == F_RUNTIME: any combination, decided per launch.
template <int VPT, typename TI, typename TO, bool FAST, int FLAGS>
__global__ void __launch_bounds__(256, (VPT >= 32 ? 1 : 3)) rowop_kernel(const RowOp op, long rows, int D) {
  __shared__ __align__(16) float prm[4][VPT * 32];   // ln1_w, ln1_b, ln2_w, ln2_b
  const int lane = threadIdx.x & 31;
  const long stride = (long)gridDim.x * 8;
  long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const bool do_ln1 = FLAGS >= 0 ? (FLAGS & F_LN1) != 0 : op.ln1_w != nullptr;
  const bool do_l2 = FLAGS >= 0 ? (FLAGS & F_L2) != 0 : op.l2norm != 0;
  const bool do_ln2 = FLAGS >= 0 ? (FLAGS & F_LN2) != 0 : op.ln2_w != nullptr;
  const bool do_film = FLAGS >= 0 ? (FLAGS & F_FILM) != 0 : op.film != nullptr;
  const bool do_silu = FLAGS >= 0 ? (FLAGS & F_SILU) != 0 : op.silu != 0;
  const TI* in = reinterpret_cast<const TI*>(op.in);
  pdl_enter();
  // The affine vectors are staged in shared memory rather than registers: the kernel is bound by
  // per-warp dependency latency (ncu: 45 % issue utilisation with 4 warps per scheduler), so the
  // registers buy more resident warps (3 CTAs per SM) instead.
  for (int i = threadIdx.x; i < VPT * 32; i += 256) {
    if (do_ln1) { prm[0][i] = op.ln1_w[i]; prm[1][i] = op.ln1_b[i]; }
    if (do_ln2) { prm[2][i] = op.ln2_w[i]; prm[3][i] = op.ln2_b[i]; }
  }
  __syncthreads();
  if (row >= rows) return;
  // Prefetch ring of PF rows per warp, kept in their raw (packed) form: the kernel is bound by load
  // latency, not bandwidth or issue (one row in flight per warp gave 17 us for a LayerNorm of 51 MB).
  constexpr int PF = sizeof(TI) == 2 ? 3 : 1;   // fp32 rows already run at 5 TB/s with one row ahead
  RawRow<VPT, TI> ring[PF];
#pragma unroll
  for (int j = 0; j < PF; ++j)
    if (row + j * stride < rows) ring[j].load(in + (row + j * stride) * D, lane);
  for (; row < rows; row += PF * stride) {
#pragma unroll
    for (int j = 0; j < PF; ++j) {
      const long r = row + j * stride;
      if (r >= rows) break;
      float v[VPT];
      ring[j].unpack(v);
      if (r + PF * stride < rows) ring[j].load(in + (r + PF * stride) * D, lane);
      if (op.out0_a) store_row<VPT, TO>(reinterpret_cast<TO*>(op.out0_a) + r * D, lane, v);
      if (do_ln1) ln_regs<VPT, FAST>(v, prm[0], prm[1], lane, D);
      if (do_l2) { if (FAST) l2norm_packed<VPT>(v, D); else l2norm_row<VPT>(v, D); }
      if (op.out1_f32) store_row<VPT, float>(op.out1_f32 + r * D, lane, v);
      if (op.out1_a) store_row<VPT, TO>(reinterpret_cast<TO*>(op.out1_a) + r * D, lane, v);
      if (do_ln2) ln_regs<VPT, FAST>(v, prm[2], prm[3], lane, D);
      if (do_film) {
        if (FAST) film_packed<VPT>(v, op.film + (r / op.rows_per_seq) * 2 * D, lane, D);
        else film_row<VPT>(v, op.film + (r / op.rows_per_seq) * 2 * D, lane, D);
      }
      if (do_silu) {
        if (FAST) silu_packed<VPT>(v);
        else {
#pragma unroll
          for (int i = 0; i < VPT; ++i) v[i] = silu_f(v[i]);
        }
      }
      if (op.out2_f32) store_row<VPT, float>(op.out2_f32 + r * D, lane, v);
      if (op.out2_a) store_row<VPT, TO>(reinterpret_cast<TO*>(op.out2_a) + r * D, lane, v);
    }
  }
}

template <int VPT>
int dispatch(const RowOp& op, long rows, int D, int out_dt, cudaStream_t st) {
  // three resident CTAs per SM (<= 85 registers), every warp strides over its share of the rows
  const long want = (rows + 7) / 8, cap = (VPT >= 32 ? 1L : 3L) * mdm_num_sms();
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  const int flags = (op.ln1_w ? F_LN1 : 0) | (op.l2norm ? F_L2 : 0) | (op.ln2_w ? F_LN2 : 0) | (op.film ? F_FILM : 0) |
                    (op.silu ? F_SILU : 0);
  cudaError_t err = cudaSuccess;
#define ROWOP(TI_, TO_, FAST_, FL_) err = mdm_launch(rowop_kernel<VPT, TI_, TO_, FAST_, FL_>, grid, 256, 0, st, op, rows, D)
  if (op.in_dt == MDM_F32 && out_dt == MDM_F32) {
    ROWOP(float, float, false, F_RUNTIME);
  } else if (op.in_dt == MDM_F32 && out_dt == MDM_BF16) {   // stage sets of MotionTransformer._layer
    if (flags == (F_LN1 | F_LN2)) ROWOP(float, bf16, true, F_LN1 | F_LN2);
    else if (flags == F_LN1) ROWOP(float, bf16, true, F_LN1);
    else if (flags == 0) ROWOP(float, bf16, true, 0);
    else ROWOP(float, bf16, true, F_RUNTIME);
  } else if (op.in_dt == MDM_BF16 && out_dt == MDM_BF16) {
    if (flags == (F_LN1 | F_L2 | F_LN2 | F_FILM | F_SILU)) ROWOP(bf16, bf16, true, F_LN1 | F_L2 | F_LN2 | F_FILM | F_SILU);
    else if (flags == (F_LN2 | F_FILM | F_SILU)) ROWOP(bf16, bf16, true, F_LN2 | F_FILM | F_SILU);
    else if (flags == F_LN1) ROWOP(bf16, bf16, true, F_LN1);
    else ROWOP(bf16, bf16, true, F_RUNTIME);
  } else if (op.in_dt == MDM_BF16 && out_dt == MDM_F32) {
    ROWOP(bf16, float, false, F_RUNTIME);
  } else {
    return MDM_ERR_ARG;
  }
#undef ROWOP
  return err == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}


// ---------------------------------------------------------------------------------------------
// Backward twin of the row pipeline (training step): given the gradient of the
// pipeline's final output, the gradient of its input plus per-CTA partial gradients of the LayerNorm affine vectors
// and of the FiLM (scale | shift) rows.  The forward intermediates are recomputed in registers from the input row
// (cheaper than storing five [N, D] tensors).  One warp per row, 32 rows of ONE sequence per CTA, so that the FiLM
// gradient of a sequence is a fixed-order sum of per-CTA partials (mdm_sum_partials): no atomics, deterministic.
//   a = LN1(x)      b = a * sqrt(D) / max(|a|, eps)      c = LN2(b)      d = c * (1 + sc) + sh      e = SiLU(d)
constexpr int BWD_ROWS = 128;   // rows of one sequence per CTA (16 per warp): 4x fewer partials to sum than with 32
// dmid (optional, fp32): a gradient that arrives at the out1 point (after LN1 / L2, before LN2), e.g. the residual-stream
// gradient of x1 = LN(pre) whose second consumer is LN2.  din may have its own type (TD) and may be accumulated into.
template <int VPT, typename TI, typename TG, typename TD>
__global__ void __launch_bounds__(256, 1)
rowop_bwd_kernel(const RowOp op, long rows, int D, int rows_per_seq, int chunks, int n_seq, const TG* __restrict__ dout,
                 const float* __restrict__ dmid, TD* din, int accumulate, float* __restrict__ dparam_part,
                 float* __restrict__ dfilm_part) {
  __shared__ float red[8][VPT * 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = blockIdx.x / chunks, ch = blockIdx.x - seq * chunks;
  const bool do_ln1 = op.ln1_w != nullptr, do_l2 = op.l2norm != 0, do_ln2 = op.ln2_w != nullptr, do_film = op.film != nullptr,
             do_silu = op.silu != 0;
  float w1[VPT], w2[VPT], sc[VPT];
  float gw1[VPT], gb1[VPT], gw2[VPT], gb2[VPT], gsc[VPT], gsh[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) { gw1[i] = gb1[i] = gw2[i] = gb2[i] = gsc[i] = gsh[i] = 0.f; w1[i] = w2[i] = 1.f; sc[i] = 0.f; }
  float b1[VPT], b2[VPT], sh[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) b1[i] = b2[i] = sh[i] = 0.f;
  if (do_ln1) { load_vec<VPT>(op.ln1_w, lane, w1); load_vec<VPT>(op.ln1_b, lane, b1); }
  if (do_ln2) { load_vec<VPT>(op.ln2_w, lane, w2); load_vec<VPT>(op.ln2_b, lane, b2); }
  if (do_film) { load_vec<VPT>(op.film + (long)seq * 2 * D, lane, sc); load_vec<VPT>(op.film + (long)seq * 2 * D + D, lane, sh); }
  const float inv_d = 1.0f / (float)D;
  const long seq_row0 = (long)seq * rows_per_seq;
  const int r_begin = ch * BWD_ROWS, r_end = min(rows_per_seq, r_begin + BWD_ROWS);
  for (int rl = r_begin + warp; rl < r_end; rl += 8) {
    const long r = seq_row0 + rl;
    if (r >= rows) break;
    float x[VPT], g[VPT];
    load_row<VPT, TI>(reinterpret_cast<const TI*>(op.in) + r * D, lane, x);
    load_row<VPT, TG>(dout + r * D, lane, g);
    // ---- forward recompute
    float xh1[VPT], r1 = 1.f;                 // LN1: normalised input, rstd
    if (do_ln1) {
      float mean;
      row_stats<VPT>(x, D, mean, r1);
#pragma unroll
      for (int i = 0; i < VPT; ++i) { xh1[i] = (x[i] - mean) * r1; x[i] = xh1[i] * w1[i] + b1[i]; }
    }
    float a[VPT], s_l2 = 1.f, n2 = 1.f;       // a = LN1 output (input of the L2 stage)
#pragma unroll
    for (int i = 0; i < VPT; ++i) a[i] = x[i];
    if (do_l2) {
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < VPT; ++i) q = fmaf(a[i], a[i], q);
      n2 = warp_sum(q);
      s_l2 = sqrtf((float)D) / fmaxf(sqrtf(n2), 1e-12f);
#pragma unroll
      for (int i = 0; i < VPT; ++i) x[i] = a[i] * s_l2;
    }
    float xh2[VPT], r2 = 1.f;
    if (do_ln2) {
      float mean;
      row_stats<VPT>(x, D, mean, r2);
#pragma unroll
      for (int i = 0; i < VPT; ++i) { xh2[i] = (x[i] - mean) * r2; x[i] = xh2[i] * w2[i] + b2[i]; }
    }
    // x = c (LN2 output); d = c * (1 + sc) + sh
    // ---- backward
    if (do_silu) {
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        const float d = do_film ? x[i] * (1.f + sc[i]) + sh[i] : x[i];
        const float sg = 1.f / (1.f + expf(-d));
        g[i] *= sg * (1.f + d * (1.f - sg));
      }
    }
    if (do_film) {
#pragma unroll
      for (int i = 0; i < VPT; ++i) { gsc[i] = fmaf(g[i], x[i], gsc[i]); gsh[i] += g[i]; g[i] *= 1.f + sc[i]; }
    }
    if (do_ln2) {
      float m1 = 0.f, m2 = 0.f;
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        gw2[i] = fmaf(g[i], xh2[i], gw2[i]); gb2[i] += g[i];
        g[i] *= w2[i];
        m1 += g[i]; m2 = fmaf(g[i], xh2[i], m2);
      }
      m1 = warp_sum(m1) * inv_d; m2 = warp_sum(m2) * inv_d;
#pragma unroll
      for (int i = 0; i < VPT; ++i) g[i] = r2 * (g[i] - m1 - xh2[i] * m2);
    }
    if (dmid) {
      float gm[VPT];
      load_row<VPT, float>(dmid + r * D, lane, gm);
#pragma unroll
      for (int i = 0; i < VPT; ++i) g[i] += gm[i];
    }
    if (do_l2) {                               // b = a * s,  s = sqrt(D) / |a|   (|a| > eps assumed, as in the forward)
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < VPT; ++i) dot = fmaf(a[i], g[i], dot);
      dot = warp_sum(dot) / fmaxf(n2, 1e-24f);
#pragma unroll
      for (int i = 0; i < VPT; ++i) g[i] = s_l2 * (g[i] - a[i] * dot);
    }
    if (do_ln1) {
      float m1 = 0.f, m2 = 0.f;
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        gw1[i] = fmaf(g[i], xh1[i], gw1[i]); gb1[i] += g[i];
        g[i] *= w1[i];
        m1 += g[i]; m2 = fmaf(g[i], xh1[i], m2);
      }
      m1 = warp_sum(m1) * inv_d; m2 = warp_sum(m2) * inv_d;
#pragma unroll
      for (int i = 0; i < VPT; ++i) g[i] = r1 * (g[i] - m1 - xh1[i] * m2);
    }
    if (accumulate) {
      float old[VPT];
      load_row<VPT, TD>(din + r * D, lane, old);
#pragma unroll
      for (int i = 0; i < VPT; ++i) g[i] += old[i];
    }
    store_row<VPT, TD>(din + r * D, lane, g);
  }
  // ---- per-CTA partial sums of the parameter gradients (fixed order over the 8 warps)
  auto reduce_to = [&](const float (&v)[VPT], float* dst) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < VPT / 4; ++j)
      *reinterpret_cast<float4*>(&red[warp][(j * 32 + lane) * 4]) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += 256) {
      float sacc = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) sacc += red[w][c];
      dst[c] = sacc;
    }
  };
  float* pp = dparam_part + (long)blockIdx.x * 4 * D;
  reduce_to(gw1, pp); reduce_to(gb1, pp + D); reduce_to(gw2, pp + 2 * D); reduce_to(gb2, pp + 3 * D);
  if (do_film) {
    float* fp = dfilm_part + ((long)ch * n_seq + seq) * 2 * D;
    reduce_to(gsc, fp); reduce_to(gsh, fp + D);
  }
}

}  // namespace

extern "C" MDM_API int mdm_rowop(const MdmRowOp* op, long rows, int D, int out_dt, void* stream) {
  if (!op || !op->in || rows < 0) return MDM_ERR_ARG;
  if (rows == 0) return MDM_OK;
  if (op->film && op->rows_per_seq <= 0) return MDM_ERR_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (D) {
    case 128: return dispatch<4>(*op, rows, D, out_dt, st);
    case 256: return dispatch<8>(*op, rows, D, out_dt, st);
    case 512: return dispatch<16>(*op, rows, D, out_dt, st);
    case 1024: return dispatch<32>(*op, rows, D, out_dt, st);
    default: return MDM_ERR_UNSUPPORTED;
  }
}

// C-ABI: see include/mdm_b200.h
template <int VPT>
static int launch_rowop_bwd(const MdmRowOp* op, long rows, int D, int grad_dt, const void* dout, const float* dmid, void* din,
                            int din_dt, int accumulate, float* dparam_part, float* dfilm_part, int rps, int n_seq, int chunks,
                            cudaStream_t st) {
  const unsigned grid = (unsigned)(n_seq * chunks);
#define BWD(TI_, TG_, TD_) rowop_bwd_kernel<VPT, TI_, TG_, TD_><<<grid, 256, 0, st>>>(*op, rows, D, rps, chunks, n_seq, \
      reinterpret_cast<const TG_*>(dout), dmid, reinterpret_cast<TD_*>(din), accumulate, dparam_part, dfilm_part)
  if (op->in_dt == MDM_BF16 && grad_dt == MDM_BF16 && din_dt == MDM_BF16) BWD(bf16, bf16, bf16);
  else if (op->in_dt == MDM_F32 && grad_dt == MDM_BF16 && din_dt == MDM_F32) BWD(float, bf16, float);
  else if (op->in_dt == MDM_F32 && grad_dt == MDM_F32 && din_dt == MDM_F32) BWD(float, float, float);
  else if (op->in_dt == MDM_BF16 && grad_dt == MDM_F32 && din_dt == MDM_F32) BWD(bf16, float, float);
  else if (op->in_dt == MDM_F32 && grad_dt == MDM_BF16 && din_dt == MDM_BF16) BWD(float, bf16, bf16);
  else return MDM_ERR_UNSUPPORTED;
#undef BWD
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_rowop_bwd2(const MdmRowOp* op, long rows, int D, int grad_dt, const void* dout, const void* dmid,
                                      int dmid_dt, void* din, float* dparam_part, float* dfilm_part, int din_flags,
                                      int* n_param_parts, int* n_film_chunks, void* stream) {
  if (!op || !op->in || !n_param_parts || !n_film_chunks) return MDM_ERR_ARG;
  if (op->film && op->rows_per_seq <= 0) return MDM_ERR_ARG;
  const int rps = op->film ? op->rows_per_seq : (int)(rows < 0x7fffffffL ? rows : 0x7fffffff);
  // without FiLM the rows need no per-sequence grouping: chunk the whole range
  const int n_seq = (int)((rows + rps - 1) / rps), chunks = (rps + BWD_ROWS - 1) / BWD_ROWS;
  *n_param_parts = n_seq * chunks;
  *n_film_chunks = chunks;
  if (!dout || !din || !dparam_part) return MDM_OK;       // size query
  if (op->film && !dfilm_part) return MDM_ERR_ARG;
  if (dmid && dmid_dt != MDM_F32) return MDM_ERR_UNSUPPORTED;
  if (rows == 0) return MDM_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int din_dt = din_flags & 1, acc = (din_flags >> 1) & 1;
  const float* dm = reinterpret_cast<const float*>(dmid);
  switch (D) {
    case 128: return launch_rowop_bwd<4>(op, rows, D, grad_dt, dout, dm, din, din_dt, acc, dparam_part, dfilm_part, rps, n_seq, chunks, st);
    case 256: return launch_rowop_bwd<8>(op, rows, D, grad_dt, dout, dm, din, din_dt, acc, dparam_part, dfilm_part, rps, n_seq, chunks, st);
    case 512: return launch_rowop_bwd<16>(op, rows, D, grad_dt, dout, dm, din, din_dt, acc, dparam_part, dfilm_part, rps, n_seq, chunks, st);
    case 1024: return launch_rowop_bwd<32>(op, rows, D, grad_dt, dout, dm, din, din_dt, acc, dparam_part, dfilm_part, rps, n_seq, chunks, st);
    default: return MDM_ERR_UNSUPPORTED;
  }
}

extern "C" MDM_API int mdm_rowop_bwd(const MdmRowOp* op, long rows, int D, int grad_dt, const void* dout, void* din,
                                     float* dparam_part, float* dfilm_part, int* n_param_parts, int* n_film_chunks,
                                     void* stream) {
  return mdm_rowop_bwd2(op, rows, D, grad_dt, dout, nullptr, MDM_F32, din, dparam_part, dfilm_part, grad_dt & 1, n_param_parts,
                        n_film_chunks, stream);
}
