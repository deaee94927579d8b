// fp32 GEMM on CUDA cores: the reference's fp32 precision mode (1e-5 parity target), same epilogue
// and grouped-tile contract as the tcgen05 kernel in gemm_tc.cu.
//   C[m, n] = epilogue( sum_k A[m, k] * W[n, k] ),  A: [rows, K] fp32, W: [N, K] fp32 (nn.Linear layout)
#include "common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, int lda, long a_rows, const float* __restrict__ W, int ldw,
                long w_rows, int M, int N, int K, const MTile* __restrict__ mtiles,
                const int* __restrict__ num_m_tiles_dev, const GemmEpi epi) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Ws[TK][TN + 4];
  const int mt = blockIdx.y >> 1;      // 128-row tile index
  const int half = blockIdx.y & 1;     // which 64-row half
  if (num_m_tiles_dev && mt >= *num_m_tiles_dev) return;
  long a_row0 = (long)mt * 128, c_row0 = a_row0;
  int w_row0 = 0, rows_valid = M - mt * 128;
  if (mtiles) {
    const MTile mi = mtiles[mt];
    a_row0 = mi.a_row0; c_row0 = mi.c_row0; w_row0 = mi.w_row0; rows_valid = mi.rows_valid;
  }
  if (half * TM >= rows_valid) return;
  a_row0 += half * TM; c_row0 += half * TM; rows_valid -= half * TM;
  const int n_base = blockIdx.x * TN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  const int lr = threadIdx.x >> 2;         // 0..63: tile row loaded by this thread
  const int lk = (threadIdx.x & 3) * 4;    // 0,4,8,12
  int kb = 0, ke = K;                      // per-tile contraction range (MdmGemmEpi.tile_k), exact here (no rounding to 64)
  if (epi.tile_k) { kb = epi.tile_k[2 * mt]; ke = min(K, kb + epi.tile_k[2 * mt + 1]); }
  for (int k0 = kb; k0 < ke; k0 += TK) {
    {
      const long ar = a_row0 + lr;
      const bool ok = lr < rows_valid && ar < a_rows;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + lk + j;
        As[lk + j][lr] = (ok && k < ke) ? A[ar * lda + k] : 0.f;
      }
      const long wr = (long)w_row0 + n_base + lr;
      const bool wok = (n_base + lr) < N && wr < w_rows;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + lk + j;
        Ws[lk + j][lr] = (wok && k < ke) ? W[wr * ldw + k] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ws[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
    if (r >= rows_valid) continue;
    const long m = c_row0 + r;
    const float rs = epi.rowscale ? epi.rowscale[m] : 1.f;
    const float rm = epi.rowmask ? epi.rowmask[m] : 1.f;
    const float scale = rs * rm * epi.alpha;
    const float* resid_row = nullptr;
    if (epi.resid) {
      const long rr = epi.resid_mod > 0 ? (m % epi.resid_mod) : m;
      resid_row = epi.resid + rr * epi.ld_resid;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n_base + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (epi.bias) v += epi.bias[w_row0 + n];
      v = apply_act(v, epi.act) * scale;
      float* out2 = reinterpret_cast<float*>(epi.out_bf16);  // operand-typed secondary output: fp32 here
      if (out2 && epi.bf16_pre_resid) out2[m * epi.ld_bf16 + n] = v;
      if (resid_row) v += epi.beta * resid_row[n];
      if (epi.out_f32) epi.out_f32[m * epi.ld_f32 + n] = v;
      if (out2 && !epi.bf16_pre_resid) out2[m * epi.ld_bf16 + n] = v;
    }
  }
}

}  // namespace

extern "C" MDM_API int mdm_gemm_f32(const float* A, int lda, long a_rows, const float* W, int ldw,
                                    long w_rows, int M, int N, int K, const void* mtiles,
                                    int num_m_tiles, const int* num_m_tiles_dev, const MdmGemmEpi* epi,
                                    void* stream) {
  if (epi && epi->tile_k && !mtiles) return MDM_ERR_ARG;
  if (!A || !W || !epi || M < 0 || N <= 0 || K <= 0) return MDM_ERR_ARG;
  if (!mtiles) num_m_tiles = (M + 127) / 128;
  if (num_m_tiles <= 0) return MDM_OK;
  dim3 grid((N + TN - 1) / TN, num_m_tiles * 2);
  gemm_f32_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      A, lda, a_rows, W, ldw, w_rows, M, N, K, reinterpret_cast<const MTile*>(mtiles), num_m_tiles_dev,
      *epi);
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}
