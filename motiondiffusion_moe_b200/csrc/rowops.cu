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
template <int VPT>
__device__ __forceinline__ void ln_regs(float (&v)[VPT], const float (&w)[VPT], const float (&b)[VPT], int D) {
  float mean, rstd;
  row_stats<VPT>(v, D, mean, rstd);
#pragma unroll
  for (int i = 0; i < VPT; ++i) v[i] = (v[i] - mean) * rstd * w[i] + b[i];
}
__device__ __forceinline__ float silu_mufu(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-x * 1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}

template <int VPT, typename TI, typename TO, bool FAST>
__global__ void __launch_bounds__(256, (VPT >= 32 ? 1 : 2)) rowop_kernel(const RowOp op, long rows, int D) {
  const int lane = threadIdx.x & 31;
  const long stride = (long)gridDim.x * 8;
  long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const TI* in = reinterpret_cast<const TI*>(op.in);
  float w1[VPT], b1[VPT], w2[VPT], b2[VPT];
  if (op.ln1_w) { load_vec<VPT>(op.ln1_w, lane, w1); load_vec<VPT>(op.ln1_b, lane, b1); }
  if (op.ln2_w) { load_vec<VPT>(op.ln2_w, lane, w2); load_vec<VPT>(op.ln2_b, lane, b2); }
  float nxt[VPT];
  load_row<VPT, TI>(in + row * D, lane, nxt);
  for (; row < rows; row += stride) {
    float v[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) v[i] = nxt[i];
    if (row + stride < rows) load_row<VPT, TI>(in + (row + stride) * D, lane, nxt);
    if (op.out0_a) store_row<VPT, TO>(reinterpret_cast<TO*>(op.out0_a) + row * D, lane, v);
    if (op.ln1_w) ln_regs<VPT>(v, w1, b1, D);
    if (op.l2norm) l2norm_row<VPT>(v, D);
    if (op.out1_f32) store_row<VPT, float>(op.out1_f32 + row * D, lane, v);
    if (op.out1_a) store_row<VPT, TO>(reinterpret_cast<TO*>(op.out1_a) + row * D, lane, v);
    if (op.ln2_w) ln_regs<VPT>(v, w2, b2, D);
    if (op.film) film_row<VPT>(v, op.film + (row / op.rows_per_seq) * 2 * D, lane, D);
    if (op.silu) {
#pragma unroll
      for (int i = 0; i < VPT; ++i) v[i] = FAST ? silu_mufu(v[i]) : silu_f(v[i]);
    }
    if (op.out2_f32) store_row<VPT, float>(op.out2_f32 + row * D, lane, v);
    if (op.out2_a) store_row<VPT, TO>(reinterpret_cast<TO*>(op.out2_a) + row * D, lane, v);
  }
}

template <int VPT>
int dispatch(const RowOp& op, long rows, int D, int out_dt, cudaStream_t st) {
  // two resident CTAs per SM (<= 128 registers), every warp strides over its share of the rows
  const long want = (rows + 7) / 8, cap = (VPT >= 32 ? 1L : 2L) * mdm_num_sms();
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  if (op.in_dt == MDM_F32 && out_dt == MDM_F32)
    rowop_kernel<VPT, float, float, false><<<grid, 256, 0, st>>>(op, rows, D);
  else if (op.in_dt == MDM_F32 && out_dt == MDM_BF16)
    rowop_kernel<VPT, float, bf16, true><<<grid, 256, 0, st>>>(op, rows, D);
  else if (op.in_dt == MDM_BF16 && out_dt == MDM_BF16)
    rowop_kernel<VPT, bf16, bf16, true><<<grid, 256, 0, st>>>(op, rows, D);
  else if (op.in_dt == MDM_BF16 && out_dt == MDM_F32)
    rowop_kernel<VPT, bf16, float, false><<<grid, 256, 0, st>>>(op, rows, D);
  else
    return MDM_ERR_ARG;
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
