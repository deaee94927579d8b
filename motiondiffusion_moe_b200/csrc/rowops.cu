// Row-wise fused normalisation pipeline: one warp per row, everything in registers, 16-byte accesses.
// Covers every nn.LayerNorm, the F.normalize * sqrt(D) of PerformerSelfAttention and StylizationBlock's
// FiLM + SiLU (reference: models/stylization.py:29-30; models/fast_attention.py:142,169-172,210,225,248;
// models/multi_branch.py:55).  HBM-bound: algorithmic bytes = D * (in + sum of outputs) per row.
#include "common.cuh"
#include "rowmath.cuh"

namespace {

template <int VPT, typename TI, typename TO>
__global__ void __launch_bounds__(256) rowop_kernel(const RowOp op, long rows, int D) {
  const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float v[VPT];
  load_row<VPT, TI>(reinterpret_cast<const TI*>(op.in) + row * D, lane, v);
  if (op.out0_a) store_row<VPT, TO>(reinterpret_cast<TO*>(op.out0_a) + row * D, lane, v);
  if (op.ln1_w) layernorm_row<VPT>(v, op.ln1_w, op.ln1_b, lane, D);
  if (op.l2norm) l2norm_row<VPT>(v, D);
  if (op.out1_f32) store_row<VPT, float>(op.out1_f32 + row * D, lane, v);
  if (op.out1_a) store_row<VPT, TO>(reinterpret_cast<TO*>(op.out1_a) + row * D, lane, v);
  if (op.ln2_w) layernorm_row<VPT>(v, op.ln2_w, op.ln2_b, lane, D);
  if (op.film) film_row<VPT>(v, op.film + (row / op.rows_per_seq) * 2 * D, lane, D);
  if (op.silu) {
#pragma unroll
    for (int i = 0; i < VPT; ++i) v[i] = silu_f(v[i]);
  }
  if (op.out2_f32) store_row<VPT, float>(op.out2_f32 + row * D, lane, v);
  if (op.out2_a) store_row<VPT, TO>(reinterpret_cast<TO*>(op.out2_a) + row * D, lane, v);
}

template <int VPT>
int dispatch(const RowOp& op, long rows, int D, int out_dt, cudaStream_t st) {
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (op.in_dt == MDM_F32 && out_dt == MDM_F32)
    rowop_kernel<VPT, float, float><<<grid, 256, 0, st>>>(op, rows, D);
  else if (op.in_dt == MDM_F32 && out_dt == MDM_BF16)
    rowop_kernel<VPT, float, bf16><<<grid, 256, 0, st>>>(op, rows, D);
  else if (op.in_dt == MDM_BF16 && out_dt == MDM_BF16)
    rowop_kernel<VPT, bf16, bf16><<<grid, 256, 0, st>>>(op, rows, D);
  else if (op.in_dt == MDM_BF16 && out_dt == MDM_F32)
    rowop_kernel<VPT, bf16, float><<<grid, 256, 0, st>>>(op, rows, D);
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
