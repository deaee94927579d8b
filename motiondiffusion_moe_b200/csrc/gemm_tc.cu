// bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
//   C[m, n] = epilogue( sum_k A[m, k] * W[n, k] )      A: [rows, K] bf16 K-major, W: [N, K] bf16
//
// W is exactly an nn.Linear weight ([out, in]), so every Linear / expert FFN / 1x2 conv of the
// reference (models/transformer.py, models/switch_moe.py:19-25, models/fast_attention.py) maps onto
// this one kernel family.  Persistent, warp-specialised:
//   warp 0 lane 0 : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring)
//   warp 1 lane 0 : MMA issuer     (tcgen05.mma 128 x BN x 16, fp32 accumulate in TMEM, 2 stages)
//   warps 2..5    : epilogue       (tcgen05.ld -> bias/act/scale/residual -> global)
// Grouped mode (MoE experts, stacked FiLM MLPs): an MTile table maps each 128-row tile to its
// A rows, C rows and weight rows; the table and its length may be produced on the device.
#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024;  // +1024 align slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               int M, int N, int K, int num_m_tiles_host, const int* __restrict__ num_m_tiles_dev,
               const MTile* __restrict__ mtiles, const GemmEpi epi) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full[0], 1);
    mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], 4);
    mbar_init(&tmem_empty[1], 4);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int num_m_tiles = num_m_tiles_dev ? *num_m_tiles_dev : num_m_tiles_host;
  const int num_n_tiles = (N + BN - 1) / BN;
  const int num_tiles = num_m_tiles * num_n_tiles;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int mt = t / num_n_tiles, nt = t - mt * num_n_tiles;
      int a_row0 = mt * BM, w_row0 = 0;
      if (mtiles) {
        const MTile mi = mtiles[mt];
        a_row0 = mi.a_row0;
        w_row0 = mi.w_row0;
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
        tma_load_2d(&tmA, &full_bar[stage], sa, kb * BK, a_row0);
        tma_load_2d(&tmB, &full_bar[stage], sb, kb * BK, w_row0 + nt * BN);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
        const uint64_t adesc = make_sw128_kmajor_desc(sa);
        const uint64_t bdesc = make_sw128_kmajor_desc(sa + L::A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 16 bf16 = 32 bytes inside the 128B swizzle row: +2 in 16-byte units
          umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------ epilogue
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int mt = t / num_n_tiles, nt = t - mt * num_n_tiles;
      int c_row0 = mt * BM, w_row0 = 0, rows_valid = M - mt * BM;
      if (mtiles) {
        const MTile mi = mtiles[mt];
        c_row0 = mi.c_row0;
        w_row0 = mi.w_row0;
        rows_valid = mi.rows_valid;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int r = quad * 32 + lane;
      const bool row_ok = r < rows_valid;
      const long m = (long)c_row0 + r;
      const float rs = (epi.rowscale && row_ok) ? epi.rowscale[m] : 1.0f;
      const float rm = (epi.rowmask && row_ok) ? epi.rowmask[m] : 1.0f;
      const float scale = rs * rm * epi.alpha;
      const float* resid_row = nullptr;
      if (epi.resid && row_ok) {
        const long rr = epi.resid_mod > 0 ? (m % epi.resid_mod) : m;
        resid_row = epi.resid + rr * epi.ld_resid;
      }
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int n0 = nt * BN + c * 32;
        if (n0 >= N) break;
        uint32_t raw[32];
        tmem_ld32(t_addr + c * 32, raw);
        tmem_ld_wait();
        if (!row_ok) continue;
        const bool full = (n0 + 32 <= N);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
        if (epi.bias) {
          const float* bp = epi.bias + w_row0 + n0;
          if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < N) v[j] += __ldg(bp + j);
          }
        }
        if (epi.act != MDM_ACT_NONE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], epi.act);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= scale;

        if (epi.out_bf16 && epi.bf16_pre_resid) {
          bf16* op = reinterpret_cast<bf16*>(epi.out_bf16) + m * epi.ld_bf16 + n0;
          if (full && (epi.ld_bf16 & 7) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 pk;
              __nv_bfloat162 p0 = __floats2bfloat162_rn(v[j], v[j + 1]);
              __nv_bfloat162 p1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
              __nv_bfloat162 p2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
              __nv_bfloat162 p3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
              pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
              pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
              *reinterpret_cast<uint4*>(op + j) = pk;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < N) op[j] = __float2bfloat16_rn(v[j]);
          }
        }
        if (resid_row) {
          const float* rp = resid_row + n0;
          if (full && (epi.ld_resid & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 r4 = *reinterpret_cast<const float4*>(rp + j);
              v[j] += epi.beta * r4.x; v[j + 1] += epi.beta * r4.y;
              v[j + 2] += epi.beta * r4.z; v[j + 3] += epi.beta * r4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < N) v[j] += epi.beta * rp[j];
          }
        }
        if (epi.out_f32) {
          float* op = epi.out_f32 + m * epi.ld_f32 + n0;
          if (full && (epi.ld_f32 & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < N) op[j] = v[j];
          }
        }
        if (epi.out_bf16 && !epi.bf16_pre_resid) {
          bf16* op = reinterpret_cast<bf16*>(epi.out_bf16) + m * epi.ld_bf16 + n0;
          if (full && (epi.ld_bf16 & 7) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 pk;
              __nv_bfloat162 p0 = __floats2bfloat162_rn(v[j], v[j + 1]);
              __nv_bfloat162 p1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
              __nv_bfloat162 p2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
              __nv_bfloat162 p3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
              pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
              pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
              *reinterpret_cast<uint4*>(op + j) = pk;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < N) op[j] = __float2bfloat16_rn(v[j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2D bf16 tensor [rows, cols] with leading dimension ld (elements); box = [box_rows, 64 cols].
bool make_map(CUtensorMap* map, const void* ptr, long rows, long cols, long ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

int g_num_sms = 0;
int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return g_num_sms;
}

template <int BN, int STAGES>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, int num_m_tiles,
           const int* num_m_tiles_dev, const MTile* mtiles, const GemmEpi& epi, int max_ctas,
           cudaStream_t stream) {
  using L = SmemLayout<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             L::TOTAL) != cudaSuccess)
      return MDM_ERR_CUDA;
    attr_set = true;
  }
  const int num_n_tiles = (N + BN - 1) / BN;
  long tiles = (long)num_m_tiles * num_n_tiles;
  int grid = (int)(tiles < max_ctas ? tiles : max_ctas);
  if (num_m_tiles_dev) grid = max_ctas;
  if (grid < 1) grid = 1;
  gemm_tc_kernel<BN, STAGES><<<grid, NUM_THREADS, L::TOTAL, stream>>>(
      ta, tb, M, N, K, num_m_tiles, num_m_tiles_dev, mtiles, epi);
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

}  // namespace

// C-ABI: see include/mdm_b200.h
extern "C" MDM_API int mdm_gemm_bf16(const void* A, int lda, long a_rows, const void* W, int ldw, long w_rows,
                             int M, int N, int K, const void* mtiles, int num_m_tiles,
                             const int* num_m_tiles_dev, const GemmEpi* epi, int max_ctas,
                             void* stream) {
  if (!A || !W || !epi || M < 0 || N <= 0 || K <= 0) return MDM_ERR_ARG;
  if ((lda & 7) || (ldw & 7)) return MDM_ERR_ARG;  // TMA needs 16-byte row pitch
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15)) return MDM_ERR_ARG;
  if (!mtiles) num_m_tiles = (M + BM - 1) / BM;
  if (num_m_tiles == 0 && !num_m_tiles_dev) return MDM_OK;
  const int sms = num_sms();
  if (max_ctas <= 0 || max_ctas > sms) max_ctas = sms;
  const bool wide = N > 128;
  CUtensorMap ta, tb;
  if (!make_map(&ta, A, a_rows, K, lda, BM)) return MDM_ERR_CUDA;
  if (!make_map(&tb, W, w_rows, K, ldw, wide ? 256 : 128)) return MDM_ERR_CUDA;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const MTile* mt = reinterpret_cast<const MTile*>(mtiles);
  if (wide) return launch<256, 4>(ta, tb, M, N, K, num_m_tiles, num_m_tiles_dev, mt, *epi, max_ctas, st);
  return launch<128, 6>(ta, tb, M, N, K, num_m_tiles, num_m_tiles_dev, mt, *epi, max_ctas, st);
}
