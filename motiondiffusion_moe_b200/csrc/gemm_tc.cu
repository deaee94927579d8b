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

#include <stdlib.h>
#include "gemm_epilogue.cuh"
#include "rowmath.cuh"
#include "cluster.cuh"
#include "tensormap.cuh"

#ifdef MDM_GEMM_PROFILE
// bring-up instrumentation: per CTA {epilogue wait cycles, epilogue work cycles, MMA-thread cycles waiting
// for a free accumulator, MMA-thread cycles waiting for operands, total cycles, tiles}
__device__ unsigned long long g_gemm_prof[148 * 8];
extern "C" MDM_API int mdm_debug_read_gemm_prof(unsigned long long* host, int n) {
  return cudaMemcpyFromSymbol(host, g_gemm_prof, sizeof(unsigned long long) * n) == cudaSuccess ? 0 : 2;
}
extern "C" MDM_API int mdm_debug_read_epi_phase(unsigned long long* host, int reset) {
  unsigned long long z[8] = {0};
  if (cudaMemcpyFromSymbol(host, g_epi_phase, sizeof(z)) != cudaSuccess) return 2;
  if (reset && cudaMemcpyToSymbol(g_epi_phase, z, sizeof(z)) != cudaSuccess) return 2;
  return 0;
}
#define PROF_T() clock64()
#else
#define PROF_T() 0ll
#endif

namespace {

// rows of a grouped GEMM's output buffer (TMA-store bound): the caller passes the buffer height as M
inline long a_rows_out(const GemmEpi*, int M) { return M; }

// MN == true ("TN" contraction, the weight gradient dW = dY^T X of the training step): both operands are read MN-major
// straight from row-major [tokens, features] tensors - TMA boxes {64 features, BK tokens}, MN-major shared-memory
// descriptors, a_major / b_major set in the instruction descriptor - so the contraction over the TOKENS needs no
// transposed copies of dY and X.  a_row0 / w_row0 of a tile are then FEATURE offsets, k indexes tokens.
template <int BN, int STAGES, int EPI, int ACT, bool MN = false>
__global__ void __launch_bounds__(num_threads(EPI), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ CUtensorMap tmF, int M, int N, int K, int num_m_tiles_host,
               const int* __restrict__ num_m_tiles_dev, const MTile* __restrict__ mtiles, const GemmEpi epi) {
  using L = SmemLayout<BN, STAGES, EPI>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by offset (not by an integer round trip), so that the compiler still knows
  // these are shared-memory addresses and emits LDS/STS instead of generic LD/ST in the epilogue
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* res_bar = reinterpret_cast<uint64_t*>(tmem_ptr + 4);   // EPI_F32T: two residual barriers per epilogue warp

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
    mbar_init(&tmem_empty[0], epi_warps(EPI));
    mbar_init(&tmem_empty[1], epi_warps(EPI));
    if (EPI == EPI_F32T) {
#pragma unroll
      for (int i = 0; i < 16; ++i) mbar_init(&res_bar[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_enter();   // everything above overlapped the previous kernel's tail; global memory is touched only from here on

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
        int k0 = 0, nkb = num_kb;
        if (epi.tile_k) { k0 = epi.tile_k[2 * mt]; nkb = (epi.tile_k[2 * mt + 1] + BK - 1) / BK; }
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          if constexpr (MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(&tmA, &full_bar[stage], sa + j * (64 * BK * 2), a_row0 + 64 * j, k0 + kb * BK);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(&tmB, &full_bar[stage], sb + j * (64 * BK * 2), w_row0 + nt * BN + 64 * j, k0 + kb * BK);
          } else {
            tma_load_2d(&tmA, &full_bar[stage], sa, k0 + kb * BK, a_row0);
            tma_load_2d(&tmB, &full_bar[stage], sb, k0 + kb * BK, w_row0 + nt * BN);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ------------------------------------------------------------ MMA issuer
      constexpr uint32_t idesc = MN ? make_idesc_bf16_mn(BM, BN) : make_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      long long w_acc = 0, w_op = 0;
      const long long t_begin = PROF_T();
      (void)t_begin;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        long long c0 = PROF_T();
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        w_acc += PROF_T() - c0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const int nkb = epi.tile_k ? (epi.tile_k[2 * (t / num_n_tiles) + 1] + BK - 1) / BK : num_kb;
        for (int kb = 0; kb < nkb; ++kb) {
          c0 = PROF_T();
          mbar_wait(&full_bar[stage], phase);
          w_op += PROF_T() - c0;
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint64_t adesc = MN ? make_sw128_mnmajor_desc(sa, 64 * BK * 2) : make_sw128_kmajor_desc(sa);
          const uint64_t bdesc = MN ? make_sw128_mnmajor_desc(sa + L::A_BYTES, 64 * BK * 2) : make_sw128_kmajor_desc(sa + L::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: advance 16 bf16 = 32 bytes inside the 128B swizzle row: +2 in 16-byte units;
            // MN-major: advance 16 token rows of 128 bytes = 2048 bytes: +128
            constexpr int step = MN ? (UMMA_K * 128) >> 4 : 2;
            umma_bf16(d_tmem, adesc + step * k, bdesc + step * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
#ifdef MDM_GEMM_PROFILE
      g_gemm_prof[blockIdx.x * 8 + 2] = w_acc;
      g_gemm_prof[blockIdx.x * 8 + 3] = w_op;
      g_gemm_prof[blockIdx.x * 8 + 4] = PROF_T() - t_begin;
#endif
    }
  } else {
    if constexpr (EPI == EPI_BF16W) {
      // 640 threads x 96 registers are allocated at launch; warpgroup 0 returns 128 x 56, i.e. +14 per
      // epilogue thread: 104 is the largest multiple of 8 that can be granted (112 would block forever)
      asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    } else {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    }
    // ------------------------------------------------------------ epilogue (8 warps)
    // warp -> TMEM lane quadrant (warp & 3) x column parity: two warps share a quadrant and take
    // alternate column units.  Accumulators arrive with lane == row; bias, activation and row scale
    // are applied in that layout, then the 32-row block is transposed through a 4 KB XOR-swizzled
    // shared-memory buffer with 128-bit accesses so that every global access (residual load, fp32 /
    // bf16 store) is a full 128-byte row segment:
    //   read phase: lane -> (row = 4*i + lane/8, 16-byte chunk = lane%8), i = 0..7.
    const int quad = warp & 3;
    const int cpar = (warp - FIRST_EPI_WARP) >> 2;
    uint4* tr = reinterpret_cast<uint4*>(smem + L::TR_OFFSET) +
                (warp - FIRST_EPI_WARP) * (EPI == EPI_BF16W ? 128 : (EPI == EPI_F32T ? 512 : 256));
    uint32_t rphase = 0;
    const EpiTma tm{&tmC, &tmR, &tmF, res_bar + 2 * (warp - FIRST_EPI_WARP), &rphase};
    int acc = 0;
    uint32_t acc_phase = 0;
    long long e_wait = 0, e_work = 0;
    int e_tiles = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int mt = t / num_n_tiles, nt = t - mt * num_n_tiles;
      int c_row0 = mt * BM, w_row0 = 0, rows_valid = M - mt * BM;
      if (mtiles) {
        const MTile mi = mtiles[mt];
        c_row0 = mi.c_row0;
        w_row0 = mi.w_row0;
        rows_valid = mi.rows_valid;
      }
      const long long c0 = PROF_T();
      mbar_wait(&tmem_full[acc], acc_phase);     // (the epilogue waits again: returns at once)
      const long long c1 = PROF_T();
      e_wait += c1 - c0;
      ++e_tiles;
      epilogue_tile<BN, EPI, ACT>(epi, tm, N, nt, c_row0, w_row0, rows_valid,
                             tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN, &tmem_full[acc], acc_phase, tr,
                             quad, cpar, lane);
      tc_fence_before();
      __syncwarp();
      e_work += PROF_T() - c1;
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
#ifdef MDM_GEMM_PROFILE
    if (warp == FIRST_EPI_WARP && lane == 0) {
      g_gemm_prof[blockIdx.x * 8 + 0] = e_wait;
      g_gemm_prof[blockIdx.x * 8 + 1] = e_work;
      g_gemm_prof[blockIdx.x * 8 + 5] = e_tiles;
    }
#endif
  }

  if ((EPI == EPI_BF16W || EPI == EPI_F32T) && warp >= FIRST_EPI_WARP && lane == 0)
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // TMA stores of this warp have left shared memory
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// =============================================================================================
// CTA-pair variant (tcgen05 cta_group::2): a cluster of two CTAs computes one 256 x 256 tile.  CTA r
// stages rows [128 r, 128 r + 128) of A and rows [128 r, +128) of the 256 weight rows of the tile, so a
// k-block costs each SM 32 KB of L2 -> shared-memory traffic instead of 48 KB (the single-CTA kernel
// is bound by exactly that feed: 67 % tensor-pipe on the K = 1024 expert down-projection).  The leader
// CTA issues the MMAs for both; accumulator rows 128 r .. live in CTA r's TMEM and are drained by
// that CTA's own epilogue warps (same epilogue code as above).
//   full[s]       leader only, 1 arrival (leader's expect_tx of both CTAs' bytes) + TMA bytes of both
//   empty[s]      one per CTA, signalled by the leader's multicast tcgen05.commit
//   tmem_full[a]  one per CTA, multicast commit after the last k-block
//   tmem_empty[a] leader only, 2 x 8 arrivals (the peer's epilogue warps arrive remotely)
constexpr int BN2 = 256;        // tile width of the pair kernel
constexpr int HALF_N2 = 128;    // weight rows staged per CTA

template <int STAGES, int EPI>
struct SmemLayout2 {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = HALF_N2 * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFFSET = TR_OFFSET + tr_bytes(EPI);
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 16 * 8 + 1024;
};

template <int STAGES, int EPI, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(num_threads(EPI), 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                const __grid_constant__ CUtensorMap tmF, int M, int N, int K, int num_m_tiles_host,
                const int* __restrict__ num_m_tiles_dev, const MTile* __restrict__ mtiles, const GemmEpi epi) {
  using L = SmemLayout2<STAGES, EPI>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* res_bar = reinterpret_cast<uint64_t*>(tmem_ptr + 4);   // EPI_F32T: two residual barriers per epilogue warp

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

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
    mbar_init(&tmem_empty[0], 2 * epi_warps(EPI));
    mbar_init(&tmem_empty[1], 2 * epi_warps(EPI));
    if (EPI == EPI_F32T) {
#pragma unroll
      for (int i = 0; i < 16; ++i) mbar_init(&res_bar[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_ptr, 2 * BN2);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_enter();

  const int num_m_tiles = num_m_tiles_dev ? *num_m_tiles_dev : num_m_tiles_host;
  const int num_pairs = (num_m_tiles + 1) >> 1;
  const int num_n_tiles = (N + BN2 - 1) / BN2;
  const int num_work = num_pairs * num_n_tiles;
  const int num_kb = (K + BK - 1) / BK;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp < FIRST_EPI_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0 && lane == 0) {
      // ------------------------------------------------------------ TMA producer (both CTAs)
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cluster_id; w < num_work; w += num_clusters) {
        const int pm = w / num_n_tiles, nt = w - pm * num_n_tiles;
        const int mt = 2 * pm + (int)rank;
        int a_row0 = mt * BM, w_row0 = 0;
        if (mtiles) {
          const MTile m0 = mtiles[2 * pm];            // both tiles of a pair share the weight rows
          w_row0 = m0.w_row0;
          a_row0 = mt < num_m_tiles ? mtiles[mt].a_row0 : m0.a_row0 + BM;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);   // the leader's barrier
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
          tma_load_2d_2sm(&tmA, fb, sa, kb * BK, a_row0);
          tma_load_2d_2sm(&tmB, fb, sb, kb * BK, w_row0 + nt * BN2 + (int)rank * HALF_N2);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 && lane == 0 && leader) {
      // ------------------------------------------------------------ MMA issuer (leader CTA only)
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN2);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      long long w_acc = 0, w_op = 0;
      const long long t_begin = PROF_T();
      (void)t_begin;
      for (int w = cluster_id; w < num_work; w += num_clusters) {
        long long c0 = PROF_T();
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        w_acc += PROF_T() - c0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN2;
        for (int kb = 0; kb < num_kb; ++kb) {
          c0 = PROF_T();
          mbar_wait(&full_bar[stage], phase);
          w_op += PROF_T() - c0;
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint64_t adesc = make_sw128_kmajor_desc(sa);
          const uint64_t bdesc = make_sw128_kmajor_desc(sa + L::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit_2sm(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
#ifdef MDM_GEMM_PROFILE
      g_gemm_prof[blockIdx.x * 8 + 2] = w_acc;
      g_gemm_prof[blockIdx.x * 8 + 3] = w_op;
      g_gemm_prof[blockIdx.x * 8 + 4] = PROF_T() - t_begin;
#endif
    }
  } else {
    if constexpr (EPI == EPI_BF16W) {
      // 640 threads x 96 registers are allocated at launch; warpgroup 0 returns 128 x 56, i.e. +14 per
      // epilogue thread: 104 is the largest multiple of 8 that can be granted (112 would block forever)
      asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    } else {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    }
    // ------------------------------------------------------------ epilogue (8 warps per CTA, own 128 rows)
    const int quad = warp & 3;
    const int cpar = (warp - FIRST_EPI_WARP) >> 2;
    uint4* tr = reinterpret_cast<uint4*>(smem + L::TR_OFFSET) +
                (warp - FIRST_EPI_WARP) * (EPI == EPI_BF16W ? 128 : (EPI == EPI_F32T ? 512 : 256));
    uint32_t rphase = 0;
    const EpiTma tm{&tmC, &tmR, &tmF, res_bar + 2 * (warp - FIRST_EPI_WARP), &rphase};
    int acc = 0;
    uint32_t acc_phase = 0;
    long long e_wait = 0, e_work = 0;
    int e_tiles = 0;
    for (int w = cluster_id; w < num_work; w += num_clusters) {
      const int pm = w / num_n_tiles, nt = w - pm * num_n_tiles;
      const int mt = 2 * pm + (int)rank;
      int c_row0 = mt * BM, w_row0 = 0, rows_valid = M - mt * BM;
      if (mtiles) {
        w_row0 = mtiles[2 * pm].w_row0;
        if (mt < num_m_tiles) {
          const MTile mi = mtiles[mt];
          c_row0 = mi.c_row0;
          rows_valid = mi.rows_valid;
        } else {
          rows_valid = 0;
        }
      }
      if (mt >= num_m_tiles) rows_valid = 0;
      const long long c0 = PROF_T();
      mbar_wait(&tmem_full[acc], acc_phase);
      const long long c1 = PROF_T();
      e_wait += c1 - c0;
      ++e_tiles;
      epilogue_tile<BN2, EPI, ACT>(epi, tm, N, nt, c_row0, w_row0, rows_valid,
                              tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN2, &tmem_full[acc], acc_phase, tr,
                              quad, cpar, lane);
      tc_fence_before();
      __syncwarp();
      e_work += PROF_T() - c1;
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
#ifdef MDM_GEMM_PROFILE
    if (warp == FIRST_EPI_WARP && lane == 0) {
      g_gemm_prof[blockIdx.x * 8 + 0] = e_wait;
      g_gemm_prof[blockIdx.x * 8 + 1] = e_work;
      g_gemm_prof[blockIdx.x * 8 + 5] = e_tiles;
    }
#endif
  }

  if ((EPI == EPI_BF16W || EPI == EPI_F32T) && warp >= FIRST_EPI_WARP && lane == 0)
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * BN2);
  }
}

// ------------------------------------------------------------------ host side
template <int BN, int STAGES, int EPI, int ACT, bool MN = false>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tr,
           const CUtensorMap& tf, int M, int N, int K, int num_m_tiles,
           const int* num_m_tiles_dev, const MTile* mtiles, const GemmEpi& epi, int max_ctas,
           cudaStream_t stream) {
  using L = SmemLayout<BN, STAGES, EPI>;
  static_assert(L::TOTAL <= 227 * 1024, "shared memory budget");
  static unsigned long long attr_set = 0;   // one bit per device ordinal: the attribute is per (function, device)
  const unsigned long long dev_bit = 1ull << mdm_cur_dev();
  if (!(attr_set & dev_bit)) {
    if (cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, EPI, ACT, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             L::TOTAL) != cudaSuccess)
      return MDM_ERR_CUDA;
    attr_set |= dev_bit;
  }
  const int num_n_tiles = (N + BN - 1) / BN;
  long tiles = (long)num_m_tiles * num_n_tiles;
  int grid = (int)(tiles < max_ctas ? tiles : max_ctas);
  if (num_m_tiles_dev) grid = max_ctas;
  if (grid < 1) grid = 1;
  return mdm_launch(gemm_tc_kernel<BN, STAGES, EPI, ACT, MN>, grid, num_threads(EPI), L::TOTAL, stream,
                    ta, tb, tc, tr, tf, M, N, K, num_m_tiles, num_m_tiles_dev, mtiles, epi) == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

template <int STAGES, int EPI, int ACT>
int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tr,
            const CUtensorMap& tf, int M, int N, int K, int num_m_tiles,
            const int* num_m_tiles_dev, const MTile* mtiles, const GemmEpi& epi, int max_ctas, cudaStream_t stream) {
  using L = SmemLayout2<STAGES, EPI>;
  static_assert(L::TOTAL <= 227 * 1024, "shared memory budget");
  static unsigned long long attr_set = 0;   // one bit per device ordinal: the attribute is per (function, device)
  const unsigned long long dev_bit = 1ull << mdm_cur_dev();
  if (!(attr_set & dev_bit)) {
    if (cudaFuncSetAttribute(gemm_tc2_kernel<STAGES, EPI, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) !=
        cudaSuccess)
      return MDM_ERR_CUDA;
    attr_set |= dev_bit;
  }
  const long work = (long)((num_m_tiles + 1) / 2) * ((N + BN2 - 1) / BN2);
  long clusters = max_ctas / 2;
  if (!num_m_tiles_dev && work < clusters) clusters = work;
  if (clusters < 1) clusters = 1;
  return mdm_launch(gemm_tc2_kernel<STAGES, EPI, ACT>, (unsigned)(2 * clusters), num_threads(EPI), L::TOTAL, stream,
                    ta, tb, tc, tr, tf, M, N, K, num_m_tiles, num_m_tiles_dev, mtiles, epi) == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}


// =============================================================================================
// Row pipeline fused into the GEMM that consumes it (K = D = 512): the A operand is never in global memory.
//   out[m, :] = beta * resid[m, :] + alpha * ( rowop(in[m, :]) . W^T + bias )
// rowop = the stage set of mdm_rowop (LayerNorm -> L2 norm * sqrt(D) -> LayerNorm -> FiLM -> SiLU, each optional).
// Per 128-row tile the eight epilogue warps first BUILD the A operand: 16 rows each, row pipeline in registers
// (warp per row, two rows interleaved), result written as bf16 straight into the eight 128B-swizzled K-major
// k-block tiles the tensor core reads (128 KB).  Meanwhile warp 0 streams the weight tiles by TMA through a 3-stage
// ring and warp 1 issues the tcgen05.mma chains for the N / 256 accumulators (TMEM columns 0..N-1).  When the
// accumulators are complete the A region is dead and becomes the staging area of the fp32 + residual epilogue
// (EPI_F32T: residual in / sum out by TMA).  Replaces rowop + GEMM pairs whose intermediate has one consumer:
// the StylizationBlock chains of the two Performer blocks and of the linear cross-attention (a2 -> s_out / ca_out).
constexpr int RD = 512;                 // row width == K
constexpr int RKB = RD / BK;            // k-blocks of the A operand
constexpr int RSTAGES = 3;              // weight ring
constexpr int RBN = 256;
struct RowGemmSmem {
  static constexpr int A_OFF = 0;                               // RKB x [128 rows x 128 B]; later epilogue staging
  static constexpr int B_OFF = RKB * BM * BK * 2;               // RSTAGES x [256 rows x 128 B]
  static constexpr int BAR_OFF = B_OFF + RSTAGES * RBN * BK * 2;
  // full / empty ring, a_ready, acc_full, tmem_empty, 16 residual barriers, TMEM pointer; + alignment slack
  static constexpr int TOTAL = BAR_OFF + (2 * RSTAGES + 3 + 16) * 8 + 16 + 1024;
};
static_assert(RowGemmSmem::TOTAL <= 227 * 1024, "shared memory budget");
static_assert(tr_bytes(EPI_F32T) <= RowGemmSmem::B_OFF, "epilogue staging fits the dead A operand");

enum { RF_LN1 = 1, RF_L2 = 2, RF_LN2 = 4, RF_FILM = 8, RF_SILU = 16 };

__device__ __forceinline__ float silu_mufu_g(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-x * 1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}

template <typename TI, int FLAGS>
__global__ void __launch_bounds__(num_threads(EPI_F32T), 1)
gemm_rowop_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                  const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmF,
                  const TI* __restrict__ in, const float* __restrict__ ln1_w, const float* __restrict__ ln1_b,
                  const float* __restrict__ ln2_w, const float* __restrict__ ln2_b, const float* __restrict__ film,
                  int rows_per_seq, int M, int N, const GemmEpi epi) {
  using L = RowGemmSmem;
  constexpr int VPT = RD / 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* As = smem + L::A_OFF;
  uint8_t* Bs = smem + L::B_OFF;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + RSTAGES;
  uint64_t* a_ready = empty_bar + RSTAGES;
  uint64_t* acc_full = a_ready + 1;
  uint64_t* tmem_empty = acc_full + 1;
  uint64_t* res_bar = tmem_empty + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_bar + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NT = N / RBN;
  const int num_m_tiles = (M + BM - 1) / BM;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < RSTAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(a_ready, 8);
    mbar_init(acc_full, 1);
    mbar_init(tmem_empty, 8);
#pragma unroll
    for (int i = 0; i < 16; ++i) mbar_init(&res_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_enter();

  if (warp < FIRST_EPI_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0 && lane == 0) {
      // ------------------------------------------------------------ weight tiles by TMA
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_m_tiles; t += gridDim.x) {
        for (int nt = 0; nt < NT; ++nt) {
          for (int kb = 0; kb < RKB; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], RBN * BK * 2);
            tma_load_2d(&tmB, &full_bar[stage], Bs + stage * (RBN * BK * 2), kb * BK, nt * RBN);
            if (++stage == RSTAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ------------------------------------------------------------ MMA issuer
      constexpr uint32_t idesc = make_idesc_bf16(BM, RBN);
      int stage = 0;
      uint32_t phase = 0, it = 0;
      const uint32_t a_addr = smem_u32(As), b_addr = smem_u32(Bs);
      long long w_a = 0, w_op = 0;
      const long long t_begin = PROF_T();
      (void)t_begin;
      for (int t = blockIdx.x; t < num_m_tiles; t += gridDim.x, ++it) {
        long long c0 = PROF_T();
        mbar_wait(tmem_empty, (it & 1) ^ 1);        // the previous tile's epilogue has drained the accumulators
        mbar_wait(a_ready, it & 1);                 // the A operand of this tile is in shared memory
        w_a += PROF_T() - c0;
        tc_fence_after();
        for (int nt = 0; nt < NT; ++nt) {
          for (int kb = 0; kb < RKB; ++kb) {
            c0 = PROF_T();
            mbar_wait(&full_bar[stage], phase);
            w_op += PROF_T() - c0;
            tc_fence_after();
            const uint64_t adesc = make_sw128_kmajor_desc(a_addr + kb * (BM * BK * 2));
            const uint64_t bdesc = make_sw128_kmajor_desc(b_addr + stage * (RBN * BK * 2));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_bf16(tmem_base + nt * RBN, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            umma_commit(&empty_bar[stage]);
            if (++stage == RSTAGES) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit(acc_full);
      }
#ifdef MDM_GEMM_PROFILE
      g_gemm_prof[blockIdx.x * 8 + 2] = w_a;
      g_gemm_prof[blockIdx.x * 8 + 3] = w_op;
      g_gemm_prof[blockIdx.x * 8 + 4] = PROF_T() - t_begin;
#endif
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ------------------------------------------------------------ A builder + epilogue (8 warps)
    const int widx = warp - FIRST_EPI_WARP;
    const int quad = warp & 3, cpar = widx >> 2;
    uint4* tr = reinterpret_cast<uint4*>(As) + widx * 512;
    uint32_t rphase = 0;
    const EpiTma tm{&tmC, &tmR, &tmF, res_bar + 2 * widx, &rphase};
    float w1[VPT], b1[VPT], w2[VPT], b2[VPT];
#pragma unroll
    for (int j = 0; j < VPT / 4; ++j) {
      const int c = (j * 32 + lane) * 4;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 a = (FLAGS & RF_LN1) ? __ldg(reinterpret_cast<const float4*>(ln1_w + c)) : z;
      const float4 bq = (FLAGS & RF_LN1) ? __ldg(reinterpret_cast<const float4*>(ln1_b + c)) : z;
      const float4 cc = (FLAGS & RF_LN2) ? __ldg(reinterpret_cast<const float4*>(ln2_w + c)) : z;
      const float4 d = (FLAGS & RF_LN2) ? __ldg(reinterpret_cast<const float4*>(ln2_b + c)) : z;
      w1[4 * j] = a.x; w1[4 * j + 1] = a.y; w1[4 * j + 2] = a.z; w1[4 * j + 3] = a.w;
      b1[4 * j] = bq.x; b1[4 * j + 1] = bq.y; b1[4 * j + 2] = bq.z; b1[4 * j + 3] = bq.w;
      w2[4 * j] = cc.x; w2[4 * j + 1] = cc.y; w2[4 * j + 2] = cc.z; w2[4 * j + 3] = cc.w;
      b2[4 * j] = d.x; b2[4 * j + 1] = d.y; b2[4 * j + 2] = d.z; b2[4 * j + 3] = d.w;
    }
    // the row pipeline on one row held as v[VPT] (lane owns columns (j*32 + lane)*4 + c)
    auto pipeline = [&](float (&v)[VPT], long m) {
      if (FLAGS & RF_LN1) {
        float mean, rstd;
        row_stats<VPT>(v, RD, mean, rstd);
#pragma unroll
        for (int i = 0; i < VPT; ++i) v[i] = (v[i] - mean) * rstd * w1[i] + b1[i];
      }
      if (FLAGS & RF_L2) l2norm_row<VPT>(v, RD);
      if (FLAGS & RF_LN2) {
        float mean, rstd;
        row_stats<VPT>(v, RD, mean, rstd);
#pragma unroll
        for (int i = 0; i < VPT; ++i) v[i] = (v[i] - mean) * rstd * w2[i] + b2[i];
      }
      if (FLAGS & RF_FILM) film_row<VPT>(v, film + (m / rows_per_seq) * 2 * RD, lane, RD);
      if (FLAGS & RF_SILU) {
#pragma unroll
        for (int i = 0; i < VPT; ++i) v[i] = silu_mufu_g(v[i]);
      }
    };
    auto store_a = [&](const float (&v)[VPT], int r) {     // bf16 row r -> the swizzled k-block tiles
#pragma unroll
      for (int j = 0; j < VPT / 4; ++j) {
        const int c = (j * 32 + lane) * 4;
        uint2 pk;
        pk.x = pack2(v[4 * j], v[4 * j + 1]);
        pk.y = pack2(v[4 * j + 2], v[4 * j + 3]);
        *reinterpret_cast<uint2*>(As + (c >> 6) * (BM * BK * 2) + r * 128 + (((((c & 63) >> 3) ^ r) & 7) << 4) + ((c & 7) << 1)) = pk;
      }
    };
    uint32_t it = 0;
    long long p_build = 0, p_wait = 0, p_epi = 0;
    for (int t = blockIdx.x; t < num_m_tiles; t += gridDim.x, ++it) {
      const long long q0 = PROF_T();
      if (it > 0) {      // the staging tiles of the previous epilogue live in the A region: every store must have read them
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      // ---- build A: rows widx*16 .. +15, two at a time
      const long row0 = (long)t * BM + widx * 16;
#pragma unroll 1
      for (int i = 0; i < 8; ++i) {
        const long ma = row0 + i, mb = row0 + i + 8;
        float va[VPT], vb[VPT];
        if (ma < M) load_row<VPT, TI>(in + ma * RD, lane, va);
        else {
#pragma unroll
          for (int e = 0; e < VPT; ++e) va[e] = 0.f;
        }
        if (mb < M) load_row<VPT, TI>(in + mb * RD, lane, vb);
        else {
#pragma unroll
          for (int e = 0; e < VPT; ++e) vb[e] = 0.f;
        }
        pipeline(va, ma < M ? ma : 0);
        pipeline(vb, mb < M ? mb : 0);
        store_a(va, widx * 16 + i);
        store_a(vb, widx * 16 + i + 8);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
      const long long q1 = PROF_T();
      // ---- epilogue once the accumulators are complete (the MMAs have finished reading A)
      mbar_wait(acc_full, it & 1);
      tc_fence_after();
      const long long q2 = PROF_T();
      const int c_row0 = t * BM, rows_valid = M - t * BM;
      for (int nt = 0; nt < NT; ++nt)
        epilogue_tile<RBN, EPI_F32T, MDM_ACT_NONE>(epi, tm, N, nt, c_row0, 0, rows_valid,
                                                   tmem_base + ((uint32_t)(quad * 32) << 16) + nt * RBN, acc_full, it & 1, tr,
                                                   quad, cpar, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
      p_build += q1 - q0; p_wait += q2 - q1; p_epi += PROF_T() - q2;
    }
#ifdef MDM_GEMM_PROFILE
    if (warp == FIRST_EPI_WARP && lane == 0) {
      g_gemm_prof[blockIdx.x * 8 + 0] = p_build;
      g_gemm_prof[blockIdx.x * 8 + 1] = p_wait;
      g_gemm_prof[blockIdx.x * 8 + 6] = p_epi;
      g_gemm_prof[blockIdx.x * 8 + 5] = it;
    }
#endif
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <typename TI, int FLAGS>
int launch_rowop_gemm(const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tr, const CUtensorMap& tf,
                      const MdmRowOp& op, int M, int N, const GemmEpi& epi, cudaStream_t st) {
  static unsigned long long attr = 0;   // one bit per device ordinal: the attribute is per (function, device)
  const unsigned long long dev_bit = 1ull << mdm_cur_dev();
  if (!(attr & dev_bit)) {
    if (cudaFuncSetAttribute(gemm_rowop_kernel<TI, FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             RowGemmSmem::TOTAL) != cudaSuccess)
      return MDM_ERR_CUDA;
    attr |= dev_bit;
  }
  const int tiles = (M + BM - 1) / BM, sms = num_sms();
  return mdm_launch(gemm_rowop_kernel<TI, FLAGS>, tiles < sms ? tiles : sms, num_threads(EPI_F32T), RowGemmSmem::TOTAL, st,
                    tb, tc, tr, tf, reinterpret_cast<const TI*>(op.in), op.ln1_w, op.ln1_b, op.ln2_w, op.ln2_b, op.film,
                    op.rows_per_seq, M, N, epi) == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
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
  if (epi->mn_major) {
    // "TN" contraction over the rows of A and W (tokens): C[M features of A, N features of W] = A^T W, fp32 output only
    // (partial products of the weight gradients); tiles address FEATURE offsets, tile_k token ranges
    if (!epi->out_f32 || epi->out_bf16 || epi->resid || epi->bias || epi->rowscale || epi->rowmask || epi->act != MDM_ACT_NONE ||
        (N & 3) || (epi->ld_f32 & 3) || (reinterpret_cast<uintptr_t>(epi->out_f32) & 15))
      return MDM_ERR_UNSUPPORTED;
    CUtensorMap ta, tb;
    if (!make_map(&ta, A, a_rows, lda, lda, BK) || !make_map(&tb, W, w_rows, ldw, ldw, BK)) return MDM_ERR_CUDA;
    const MTile* mt = reinterpret_cast<const MTile*>(mtiles);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (N > 128) return launch<256, 4, EPI_F32, MDM_ACT_NONE, true>(ta, tb, ta, ta, ta, M, N, K, num_m_tiles, num_m_tiles_dev, mt, *epi, max_ctas, st);
    return launch<128, 6, EPI_F32, MDM_ACT_NONE, true>(ta, tb, ta, ta, ta, M, N, K, num_m_tiles, num_m_tiles_dev, mt, *epi, max_ctas, st);
  }
  // tile width: 256 columns unless N is small or the 256-wide tiling leaves the last wave mostly idle
  bool wide = N > 128;
  if (wide && !num_m_tiles_dev) {
    const long t256 = (long)num_m_tiles * ((N + 255) / 256), t128 = (long)num_m_tiles * ((N + 127) / 128);
    const double e256 = (double)t256 / (double)(((t256 + max_ctas - 1) / max_ctas) * max_ctas);
    const double e128 = (double)t128 / (double)(((t128 + max_ctas - 1) / max_ctas) * max_ctas);
    if (e128 > e256 + 0.12) wide = false;
  }
  // CTA pairs (cta_group::2, 256 x 256 tiles) for wide GEMMs with at least two row tiles; a grouped
  // GEMM must declare that its tile table is pair-aligned (tiles 2i, 2i+1 share their weight rows)
  // Measured on B200 (tools/op_bench.py, tools/gemm_prof.py): K = 512 GEMMs are bound by their
  // epilogue, where the pair kernel is ~8 % slower; from K = 1024 up (operand-feed bound) it wins
  // (N x 512 x 2048: 56.8 -> 53.8 us).  MDM_GEMM_PAIR=0/1 forces it off / on for A/B runs.
  static const int pair_env = [] { const char* e = getenv("MDM_GEMM_PAIR"); return e ? atoi(e) : -1; }();
  if (epi->tile_k && !mtiles) return MDM_ERR_ARG;
  const bool pair_fit = wide && (num_m_tiles_dev || num_m_tiles >= 2) && (!mtiles || epi->pair_tiles) && max_ctas >= 2 &&
                        !epi->tile_k;
  const bool pair = pair_fit && (pair_env == 1 || (pair_env < 0 && K >= 1024));
  CUtensorMap ta, tb;
  if (!make_map(&ta, A, a_rows, K, lda, BM)) return MDM_ERR_CUDA;
  if (!make_map(&tb, W, w_rows, K, ldw, (wide && !pair) ? 256 : 128)) return MDM_ERR_CUDA;
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
  // one instantiation per (epilogue flavour, activation): bf16 {none, gelu, silu}, fp32 {none, gelu};
  // anything else goes through the any-shape kernel with the activation decided at run time
  const int act = epi->act;
#define MDM_GO(E_, A_)                                                                                            \
  do {                                                                                                            \
    if (pair) return launch2<6, E_, A_>(ta, tb, tc, tc, tc, M, N, K, num_m_tiles, num_m_tiles_dev, mt, *epi, max_ctas, st);   \
    if (wide) return launch<256, 4, E_, A_>(ta, tb, tc, tc, tc, M, N, K, num_m_tiles, num_m_tiles_dev, mt, *epi, max_ctas, st); \
    return launch<128, 6, E_, A_>(ta, tb, tc, tc, tc, M, N, K, num_m_tiles, num_m_tiles_dev, mt, *epi, max_ctas, st);          \
  } while (0)
  static const int epiw_env = [] { const char* e = getenv("MDM_GEMM_EPIW"); return e ? atoi(e) : 1; }();
  CUtensorMap tc = ta;    // only the TMA-store flavour reads it
  // rows of the output the kernel may write: the logical M (plain GEMM) or the whole buffer (grouped)
  const long c_rows = mtiles ? a_rows_out(epi, M) : M;
  const bool tma_out = kind == EPI_BF16 && epiw_env && (wide || pair) && (epi->ld_bf16 & 7) == 0 &&
                       make_out_map(&tc, epi->out_bf16, c_rows, N, epi->ld_bf16);
  if (tma_out) {   // 256-wide tiles: 16 epilogue warps, TMA stores
#define MDM_GOW(A_)                                                                                                  \
  do {                                                                                                               \
    if (pair) return launch2<6, EPI_BF16W, A_>(ta, tb, tc, tc, tc, M, N, K, num_m_tiles, num_m_tiles_dev, mt, *epi, max_ctas, st); \
    return launch<256, 4, EPI_BF16W, A_>(ta, tb, tc, tc, tc, M, N, K, num_m_tiles, num_m_tiles_dev, mt, *epi, max_ctas, st);      \
  } while (0)
    if (act == MDM_ACT_NONE) MDM_GOW(MDM_ACT_NONE);
    if (act == MDM_ACT_GELU) MDM_GOW(MDM_ACT_GELU);
    if (act == MDM_ACT_SILU) MDM_GOW(MDM_ACT_SILU);
#undef MDM_GOW
  }
  // fp32 output + fp32 residual of a plain (not grouped) GEMM: residual in / sum out by TMA (B200, N x 512 x 512:
  // 35.7 -> 26.5 us, with GELU 44.3 -> 30.3 us, with a bf16 copy 39.3 -> 31.6 us; these GEMMs move 128 MB for
  // 13 GFLOP, i.e. they are HBM-bound).  Not for the CTA-pair kernel: K >= 1024 is operand-feed bound and the
  // staging tiles cost it two pipeline stages (53.4 -> 54.5 us).
  // MDM_GEMM_F32T: 0 = off, 1 = on with the usual tile-width choice (default), 2 = always 128-wide tiles.
  static const int f32t_env = [] { const char* e = getenv("MDM_GEMM_F32T"); return e ? atoi(e) : 1; }();
  if (kind == EPI_F32 && f32t_env && !pair && !mtiles && epi->out_f32 && epi->resid && epi->resid_mod <= 0 &&
      (!epi->out_bf16 || ((epi->ld_bf16 & 7) == 0 && al(epi->out_bf16, 16))) &&
      (act == MDM_ACT_NONE || act == MDM_ACT_GELU)) {
    CUtensorMap tr, tf;
    bool ok = make_f32_map(&tr, epi->resid, M, N, epi->ld_resid) && make_f32_map(&tf, epi->out_f32, M, N, epi->ld_f32);
    if (ok && epi->out_bf16) ok = make_out_map(&tc, epi->out_bf16, M, N, epi->ld_bf16);
    if (ok) {
      const bool wide_t = wide && f32t_env != 2;
      if (wide_t != wide)     // the weight box follows the tile width
        if (!make_map(&tb, W, w_rows, K, ldw, wide_t ? 256 : 128)) return MDM_ERR_CUDA;
#define MDM_GOT(A_)                                                                                                   \
  do {                                                                                                                \
    if (wide_t) return launch<256, 3, EPI_F32T, A_>(ta, tb, tc, tr, tf, M, N, K, num_m_tiles, num_m_tiles_dev, mt, *epi, max_ctas, st); \
    return launch<128, 4, EPI_F32T, A_>(ta, tb, tc, tr, tf, M, N, K, num_m_tiles, num_m_tiles_dev, mt, *epi, max_ctas, st); \
  } while (0)
      if (act == MDM_ACT_NONE) MDM_GOT(MDM_ACT_NONE);
      MDM_GOT(MDM_ACT_GELU);
#undef MDM_GOT
    }
  }
  if (kind == EPI_BF16 && act == MDM_ACT_NONE) MDM_GO(EPI_BF16, MDM_ACT_NONE);
  if (kind == EPI_BF16 && act == MDM_ACT_GELU) MDM_GO(EPI_BF16, MDM_ACT_GELU);
  if (kind == EPI_BF16 && act == MDM_ACT_SILU) MDM_GO(EPI_BF16, MDM_ACT_SILU);
  if (kind == EPI_F32 && act == MDM_ACT_NONE) MDM_GO(EPI_F32, MDM_ACT_NONE);
  if (kind == EPI_F32 && act == MDM_ACT_GELU) MDM_GO(EPI_F32, MDM_ACT_GELU);
  MDM_GO(EPI_ANY, -1);
#undef MDM_GO
}

// C-ABI: see include/mdm_b200.h
extern "C" MDM_API int mdm_gemm_rowop(const MdmRowOp* op, long rows, int D, const void* W, int ldw, long w_rows, int N,
                                      const GemmEpi* epi, void* stream) {
  if (!op || !op->in || !W || !epi) return MDM_ERR_ARG;
  if (D != RD || N % RBN != 0 || N < RBN || N > 512 || rows <= 0 || rows > 0x7fffffffL) return MDM_ERR_UNSUPPORTED;
  if (!epi->out_f32 || !epi->resid || epi->resid_mod > 0 || epi->act != MDM_ACT_NONE || epi->rowscale || epi->rowmask ||
      epi->out_bf16)
    return MDM_ERR_UNSUPPORTED;
  if ((ldw & 7) || (epi->ld_f32 & 3) || (epi->ld_resid & 3)) return MDM_ERR_UNSUPPORTED;
  auto al = [](const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; };
  if (!al(W, 16) || !al(op->in, 16) || !al(epi->out_f32, 16) || !al(epi->resid, 16)) return MDM_ERR_UNSUPPORTED;
  if ((op->film != nullptr) && op->rows_per_seq <= 0) return MDM_ERR_ARG;
  const int M = (int)rows;
  CUtensorMap tb, tr, tf;
  if (!make_map(&tb, W, w_rows, D, ldw, RBN)) return MDM_ERR_CUDA;
  if (!make_f32_map(&tr, epi->resid, M, N, epi->ld_resid) || !make_f32_map(&tf, epi->out_f32, M, N, epi->ld_f32))
    return MDM_ERR_CUDA;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int flags = (op->ln1_w ? RF_LN1 : 0) | (op->l2norm ? RF_L2 : 0) | (op->ln2_w ? RF_LN2 : 0) | (op->film ? RF_FILM : 0) |
                    (op->silu ? RF_SILU : 0);
  // the stage sets of MotionTransformer._layer that feed a residual-stream GEMM
  if (op->in_dt == MDM_BF16 && flags == (RF_LN1 | RF_L2 | RF_LN2 | RF_FILM | RF_SILU))
    return launch_rowop_gemm<bf16, RF_LN1 | RF_L2 | RF_LN2 | RF_FILM | RF_SILU>(tb, tb, tr, tf, *op, M, N, *epi, st);
  if (op->in_dt == MDM_BF16 && flags == (RF_LN2 | RF_FILM | RF_SILU))
    return launch_rowop_gemm<bf16, RF_LN2 | RF_FILM | RF_SILU>(tb, tb, tr, tf, *op, M, N, *epi, st);
  return MDM_ERR_UNSUPPORTED;
}
