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
template <int VPT, bool TREE>
__device__ __forceinline__ void ln_regs(float (&v)[VPT], const float* __restrict__ w_s, const float* __restrict__ b_s,
                                        int lane, int D) {
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
      if (do_l2) l2norm_row<VPT>(v, D);
      if (op.out1_f32) store_row<VPT, float>(op.out1_f32 + r * D, lane, v);
      if (op.out1_a) store_row<VPT, TO>(reinterpret_cast<TO*>(op.out1_a) + r * D, lane, v);
      if (do_ln2) ln_regs<VPT, FAST>(v, prm[2], prm[3], lane, D);
      if (do_film) film_row<VPT>(v, op.film + (r / op.rows_per_seq) * 2 * D, lane, D);
      if (do_silu) {
#pragma unroll
        for (int i = 0; i < VPT; ++i) v[i] = FAST ? silu_mufu(v[i]) : silu_f(v[i]);
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
#define ROWOP(TI_, TO_, FAST_, FL_) rowop_kernel<VPT, TI_, TO_, FAST_, FL_><<<grid, 256, 0, st>>>(op, rows, D)
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
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
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
