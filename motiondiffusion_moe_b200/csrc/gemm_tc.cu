// bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
//   C[m, n] = epilogue( sum_k A[m, k] * W[n, k] )      A: [rows, K] bf16 K-major, W: [N, K] bf16
//
// W is exactly an nn.Linear weight ([out, in]), so every Linear / expert FFN / 1x2 conv of the
// reference (models/transformer.py, models/switch_moe.py:19-25, models/fast_attention.py) maps onto
// this one kernel family.  Persistent, warp-specialised:
//   warp 0 lane 0 : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring)
//   warp 1 lane 0 : MMA issuer     (tcgen05.mma 128 x BN x 16, fp32 accumulate in TMEM, 2 stages)
//   warps 4..11   : epilogue       (tcgen05.ld -> bias/act/scale -> (smem transpose) -> residual -> global)
// Roles are aligned to warpgroups so that setmaxnreg can move registers from warpgroup 0 (40 each)
// to the two epilogue warpgroups (232 each): the fp32 epilogue keeps the whole residual tile of its
// warp (128 registers) in flight while the MMAs of the tile are still running.
// Grouped mode (MoE experts, stacked FiLM MLPs): an MTile table maps each 128-row tile to its
// A rows, C rows and weight rows; the table and its length may be produced on the device.
#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int FIRST_EPI_WARP = 4;  // warpgroup 0 = {TMA, MMA, 2 idle warps}; warpgroups 1, 2 = epilogue
constexpr int NUM_THREADS = (FIRST_EPI_WARP + NUM_EPI_WARPS) * 32;
constexpr int STAGE_T_BYTES = 32 * 32 * 4;  // per-warp 32x32 fp32 transpose buffer

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFFSET = TR_OFFSET + NUM_EPI_WARPS * STAGE_T_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024;  // +1024 align slack
};

// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)) with erf from Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7) on
// the MUFU rcp / ex2 units: ~14 issue slots instead of erff's ~50, so that the epilogue of a K=512
// GEMM stays under its MMA time.  Max abs deviation from the exact erf GELU: 4.5e-7 (bf16 ulp at 1 is 7.8e-3).
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
  const float erf_abs = fmaf(-p, e, 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}
// GELU for bf16-only outputs: x * Phi(x) with Phi(x) = 0.5 + 0.5 tanh(x Q(x^2)), Q fitted (minimax over
// |x| <= 6, tools/fit_gelu.py) so that tanh(x Q(x^2)) == erf(x / sqrt 2): these are NOT the constants of
// the "tanh GELU" variant.  Max abs deviation from the exact-erf GELU: 5.4e-5 from the fit plus
// 2^-11 relative from MUFU.TANH, i.e. <= 1/8 of a bf16 rounding step of the result.  8 issue slots
// and one MUFU op per element instead of 15 and two (expert up-projection: 169 -> 142 us).
__device__ __forceinline__ float gelu_tanh_fit(float x) {
  const float s = fminf(x * x, 25.0f);
  float q = fmaf(-3.81889112e-04f, s, 3.72153111e-02f);
  q = fmaf(q, s, 7.97237410e-01f);
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(x * q));
  const float hx = 0.5f * x;
  return fmaf(hx, th, hx);
}
__device__ __forceinline__ float silu_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-x * 1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}
__device__ __forceinline__ float act_fast(float v, int act) {
  if (act == MDM_ACT_GELU) return gelu_fast(v);
  if (act == MDM_ACT_SILU) return silu_fast(v);
  if (act == MDM_ACT_EXPFEAT) return expf(fminf(fmaxf(v, -15.f), 15.f)) * 0.1f;
  return v;
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&p);
}
// Epilogue flavours.  Each is its own kernel instantiation so that the hot loop of a launch stays
// inside the instruction cache: with all variants (and their scalar tail code) in one body the fp32
// epilogue ran at ~20 cycles per issued instruction, "no instruction" being its largest stall (ncu).
//   EPI_BF16 : bf16 output only, 16-byte vector stores (N % 8 == 0, aligned)
//   EPI_F32  : fp32 output and/or residual (+ optional bf16 copy), vector accesses (N % 4 == 0, aligned)
//   EPI_ANY  : any shape / alignment (scalar tails); used for the few odd shapes (263 features ...)
enum { EPI_BF16 = 0, EPI_F32 = 1, EPI_ANY = 2 };

template <int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               int M, int N, int K, int num_m_tiles_host, const int* __restrict__ num_m_tiles_dev,
               const MTile* __restrict__ mtiles, const GemmEpi epi) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by offset (not by an integer round trip), so that the compiler still knows
  // these are shared-memory addresses and emits LDS/STS instead of generic LD/ST in the epilogue
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
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
    mbar_init(&tmem_empty[0], NUM_EPI_WARPS);
    mbar_init(&tmem_empty[1], NUM_EPI_WARPS);
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

  if (warp < FIRST_EPI_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
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
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ------------------------------------------------------------ epilogue (8 warps)
    // warp -> TMEM lane quadrant (warp & 3) x column parity: two warps share a quadrant and take
    // alternate column units.  Accumulators arrive with lane == row; bias, activation and row scale
    // are applied in that layout, then the 32-row block is transposed through a 4 KB XOR-swizzled
    // shared-memory buffer with 128-bit accesses so that every global access (residual load, fp32 /
    // bf16 store) is a full 128-byte row segment:
    //   read phase: lane -> (row = 4*i + lane/8, 16-byte chunk = lane%8), i = 0..7.
    const int quad = warp & 3;
    const int cpar = (warp - FIRST_EPI_WARP) >> 2;
    uint4* tr = reinterpret_cast<uint4*>(smem + L::TR_OFFSET) + (warp - FIRST_EPI_WARP) * 256;  // 32 rows x 8 chunks
    float4* trf = reinterpret_cast<float4*>(tr);
    const bool f32_path = EPI == EPI_F32 || (EPI == EPI_ANY && (epi.out_f32 != nullptr || epi.resid != nullptr));
    const int rsub = lane >> 3, ch = lane & 7;
    constexpr int NCH = BN / 64;   // 32-column chunks per warp and tile
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
      const int r = quad * 32 + lane;
      const bool row_ok = r < rows_valid;
      const long m = (long)c_row0 + r;
      const float rs = (epi.rowscale && row_ok) ? epi.rowscale[m] : 1.0f;
      const float rm = (epi.rowmask && row_ok) ? epi.rowmask[m] : 1.0f;
      const float scale = rs * rm * epi.alpha;
      const bool has_scale = epi.rowscale || epi.rowmask || epi.alpha != 1.0f;   // warp-uniform
      const int rmax = min(32, rows_valid - quad * 32);   // valid rows of this warp's 32-row block
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;
      // per-tile base pointers; in-tile offsets stay 32-bit
      const long blk_row0 = (long)c_row0 + quad * 32;
      float* of_blk = epi.out_f32 ? epi.out_f32 + blk_row0 * epi.ld_f32 : nullptr;
      bf16* ob_blk = epi.out_bf16 ? reinterpret_cast<bf16*>(epi.out_bf16) + blk_row0 * epi.ld_bf16 : nullptr;
      const float* rs_blk = (epi.resid && epi.resid_mod <= 0) ? epi.resid + blk_row0 * epi.ld_resid : epi.resid;
      const int rmod_base = epi.resid_mod > 0 ? (int)(blk_row0 % epi.resid_mod) : 0;

      // TMEM chunk c (32 columns from n0) -> registers, + bias, activation, row scale
      auto load_chunk = [&](int c, int n0, float (&v)[32]) {
        uint32_t raw[32];
        tmem_ld32(t_addr + c * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
        if (epi.bias) {
          const float* bp = epi.bias + w_row0 + n0;
          if (n0 + 32 <= N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          } else if (EPI != EPI_ANY) {   // N % 4 == 0 here: 4-column granules, static register indices
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (n0 + j + 4 <= N) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + j));
                v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < N) v[j] += __ldg(bp + j);
          }
        }
        if (epi.act == MDM_ACT_GELU) {
          if (f32_path) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_fit(v[j]);
          }
        } else if (epi.act == MDM_ACT_SILU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = silu_fast(v[j]);
        } else if (epi.act != MDM_ACT_NONE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = act_fast(v[j], epi.act);
        }
        if (has_scale) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= scale;
        }
      };

      if constexpr (EPI == EPI_F32) {
        // ---------------- fp32 / residual outputs, vector accesses only
        // The residual rows of all NCH chunks of this warp (read-phase layout) are requested before
        // the wait for the accumulator: NCH x 4 KB per warp in flight while the MMAs of the tile run.
        float4 res[NCH][8];
        if (epi.resid) {
#pragma unroll
          for (int k = 0; k < NCH; ++k) {
            const int n = nt * BN + (cpar + 2 * k) * 32 + ch * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = i * 4 + rsub;
              res[k][i] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (row < rmax && n < N) {
                const int rr_ = epi.resid_mod > 0 ? (rmod_base + row) % epi.resid_mod : row;
                const float* rp = rs_blk + (long)rr_ * epi.ld_resid + n;
                asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
                             : "=f"(res[k][i].x), "=f"(res[k][i].y), "=f"(res[k][i].z), "=f"(res[k][i].w) : "l"(rp));
              }
            }
          }
        }
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
          const int c = cpar + 2 * k;
          const int n0 = nt * BN + c * 32;
          if (n0 < N) {
            const int n = n0 + ch * 4;                        // this lane's 4 columns in the read phase
            float v[32];
            load_chunk(c, n0, v);
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8)
              trf[lane * 8 + (c8 ^ (lane & 7))] = make_float4(v[4 * c8], v[4 * c8 + 1], v[4 * c8 + 2], v[4 * c8 + 3]);
            __syncwarp();
            float4 x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = i * 4 + rsub;
              x[i] = trf[row * 8 + (ch ^ (row & 7))];
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = i * 4 + rsub;
              if (row < rmax && n < N) {
                if (ob_blk && epi.bf16_pre_resid) {
                  uint2 pk; pk.x = pack2(x[i].x, x[i].y); pk.y = pack2(x[i].z, x[i].w);
                  *reinterpret_cast<uint2*>(ob_blk + row * epi.ld_bf16 + n) = pk;
                }
                if (epi.resid) {
                  x[i].x = fmaf(epi.beta, res[k][i].x, x[i].x); x[i].y = fmaf(epi.beta, res[k][i].y, x[i].y);
                  x[i].z = fmaf(epi.beta, res[k][i].z, x[i].z); x[i].w = fmaf(epi.beta, res[k][i].w, x[i].w);
                }
                if (of_blk) *reinterpret_cast<float4*>(of_blk + row * epi.ld_f32 + n) = x[i];
                if (ob_blk && !epi.bf16_pre_resid) {
                  uint2 pk; pk.x = pack2(x[i].x, x[i].y); pk.y = pack2(x[i].z, x[i].w);
                  *reinterpret_cast<uint2*>(ob_blk + row * epi.ld_bf16 + n) = pk;
                }
              }
            }
          }
        }
      } else if constexpr (EPI == EPI_BF16) {
        // ---------------- bf16-only output: 64-column units (two TMEM chunks) staged as bf16
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int u = cpar; u < BN / 64; u += 2) {
          const int n0 = nt * BN + u * 64;
          if (n0 >= N) break;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            if (n0 + hh * 32 < N) {
              float v[32];
              load_chunk(u * 2 + hh, n0 + hh * 32, v);
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4) {
                uint4 pk;
                pk.x = pack2(v[8 * c4], v[8 * c4 + 1]); pk.y = pack2(v[8 * c4 + 2], v[8 * c4 + 3]);
                pk.z = pack2(v[8 * c4 + 4], v[8 * c4 + 5]); pk.w = pack2(v[8 * c4 + 6], v[8 * c4 + 7]);
                tr[lane * 8 + ((hh * 4 + c4) ^ (lane & 7))] = pk;
              }
            }
          }
          __syncwarp();
          const int n = n0 + ch * 8;   // this lane's 8 columns in the read phase
          uint4 w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = i * 4 + rsub;
            w[i] = tr[row * 8 + (ch ^ (row & 7))];
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = i * 4 + rsub;
            if (row < rmax && n < N) *reinterpret_cast<uint4*>(ob_blk + row * epi.ld_bf16 + n) = w[i];
          }
        }
      } else {
        // ---------------- any shape / alignment (scalar tails), rolled loops
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        if (f32_path) {
          const bool vec_ok = ((epi.ld_f32 & 3) == 0 || !epi.out_f32) && ((epi.ld_resid & 3) == 0 || !epi.resid) &&
                              ((epi.ld_bf16 & 3) == 0 || !epi.out_bf16);
#pragma unroll 1
          for (int c = cpar; c < BN / 32; c += 2) {
            const int n0 = nt * BN + c * 32;
            if (n0 >= N) break;
            const int n = n0 + ch * 4;
            const bool cvec = vec_ok && (n + 4 <= N);
            float v[32];
            load_chunk(c, n0, v);
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8)
              trf[lane * 8 + (c8 ^ (lane & 7))] = make_float4(v[4 * c8], v[4 * c8 + 1], v[4 * c8 + 2], v[4 * c8 + 3]);
            __syncwarp();
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
              const int row = i * 4 + rsub;
              const float4 x4 = trf[row * 8 + (ch ^ (row & 7))];
              if (row < rmax && n < N) {
                float x[4] = {x4.x, x4.y, x4.z, x4.w};
                const int rr_ = epi.resid_mod > 0 ? (rmod_base + row) % epi.resid_mod : row;
                const float* rp = epi.resid ? rs_blk + (long)rr_ * epi.ld_resid + n : nullptr;
                bf16* ob = ob_blk ? ob_blk + row * epi.ld_bf16 + n : nullptr;
                float* of = of_blk ? of_blk + row * epi.ld_f32 + n : nullptr;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if (n + j < N) {
                    if (ob && epi.bf16_pre_resid) ob[j] = __float2bfloat16_rn(x[j]);
                    if (rp) x[j] = fmaf(epi.beta, __ldg(rp + j), x[j]);
                    if (of) of[j] = x[j];
                    if (ob && !epi.bf16_pre_resid) ob[j] = __float2bfloat16_rn(x[j]);
                  }
                }
                (void)cvec;
              }
            }
            __syncwarp();
          }
        } else if (ob_blk) {
#pragma unroll 1
          for (int u = cpar; u < BN / 64; u += 2) {
            const int n0 = nt * BN + u * 64;
            if (n0 >= N) break;
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
              if (n0 + hh * 32 < N) {
                float v[32];
                load_chunk(u * 2 + hh, n0 + hh * 32, v);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                  uint4 pk;
                  pk.x = pack2(v[8 * c4], v[8 * c4 + 1]); pk.y = pack2(v[8 * c4 + 2], v[8 * c4 + 3]);
                  pk.z = pack2(v[8 * c4 + 4], v[8 * c4 + 5]); pk.w = pack2(v[8 * c4 + 6], v[8 * c4 + 7]);
                  tr[lane * 8 + ((hh * 4 + c4) ^ (lane & 7))] = pk;
                }
              }
            }
            __syncwarp();
            const int n = n0 + ch * 8;
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
              const int row = i * 4 + rsub;
              const uint4 w = tr[row * 8 + (ch ^ (row & 7))];
              if (row < rmax && n < N) {
                bf16* ob = ob_blk + row * epi.ld_bf16 + n;
                const bf16* e = reinterpret_cast<const bf16*>(&w);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (n + j < N) ob[j] = e[j];
              }
            }
            __syncwarp();
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

template <int BN, int STAGES, int EPI>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, int num_m_tiles,
           const int* num_m_tiles_dev, const MTile* mtiles, const GemmEpi& epi, int max_ctas,
           cudaStream_t stream) {
  using L = SmemLayout<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             L::TOTAL) != cudaSuccess)
      return MDM_ERR_CUDA;
    attr_set = true;
  }
  const int num_n_tiles = (N + BN - 1) / BN;
  long tiles = (long)num_m_tiles * num_n_tiles;
  int grid = (int)(tiles < max_ctas ? tiles : max_ctas);
  if (num_m_tiles_dev) grid = max_ctas;
  if (grid < 1) grid = 1;
  gemm_tc_kernel<BN, STAGES, EPI><<<grid, NUM_THREADS, L::TOTAL, stream>>>(
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
  // tile width: 256 columns unless N is small or the 256-wide tiling leaves the last wave mostly idle
  bool wide = N > 128;
  if (wide && !num_m_tiles_dev) {
    const long t256 = (long)num_m_tiles * ((N + 255) / 256), t128 = (long)num_m_tiles * ((N + 127) / 128);
    const double e256 = (double)t256 / (double)(((t256 + max_ctas - 1) / max_ctas) * max_ctas);
    const double e128 = (double)t128 / (double)(((t128 + max_ctas - 1) / max_ctas) * max_ctas);
    if (e128 > e256 + 0.12) wide = false;
  }
  CUtensorMap ta, tb;
  if (!make_map(&ta, A, a_rows, K, lda, BM)) return MDM_ERR_CUDA;
  if (!make_map(&tb, W, w_rows, K, ldw, wide ? 256 : 128)) return MDM_ERR_CUDA;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const MTile* mt = reinterpret_cast<const MTile*>(mtiles);
  // epilogue flavour: the vectorised kernels need aligned rows and N % 8 (bf16) / N % 4 (fp32)
  auto al = [](const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; };
  const bool f32_path = epi->out_f32 || epi->resid;
  int kind = EPI_ANY;
  if (!f32_path && epi->out_bf16 && (N & 7) == 0 && (epi->ld_bf16 & 7) == 0 && al(epi->out_bf16, 16)) kind = EPI_BF16;
  if (f32_path && (N & 3) == 0 && (!epi->out_f32 || ((epi->ld_f32 & 3) == 0 && al(epi->out_f32, 16))) &&
      (!epi->resid || ((epi->ld_resid & 3) == 0 && al(epi->resid, 16))) &&
      (!epi->out_bf16 || ((epi->ld_bf16 & 3) == 0 && al(epi->out_bf16, 8))))
    kind = EPI_F32;
#define MDM_LAUNCH(BN_, ST_, E_) launch<BN_, ST_, E_>(ta, tb, M, N, K, num_m_tiles, num_m_tiles_dev, mt, *epi, max_ctas, st)
  if (wide) return kind == EPI_BF16 ? MDM_LAUNCH(256, 4, EPI_BF16) : kind == EPI_F32 ? MDM_LAUNCH(256, 4, EPI_F32) : MDM_LAUNCH(256, 4, EPI_ANY);
  return kind == EPI_BF16 ? MDM_LAUNCH(128, 6, EPI_BF16) : kind == EPI_F32 ? MDM_LAUNCH(128, 6, EPI_F32) : MDM_LAUNCH(128, 6, EPI_ANY);
#undef MDM_LAUNCH
}
