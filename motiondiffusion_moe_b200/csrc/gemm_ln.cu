// Full-row bf16 GEMM with the row pipeline (LayerNorm / L2 norm / FiLM / SiLU) in its epilogue: north_star (3),
// "StylizationBlock timestep-FiLM and LayerNorm fused into the adjacent GEMM epilogues".
//
//   y_pre = act(A[m, :] . W[512, K]^T + bias) * alpha          y = y_pre + beta * resid[m, :]
//   s     = LN source: y, or y_pre (MdmGemmEpi.bf16_pre_resid)
//   u     = L2norm?( LN1(s) )                                   -> out1 (fp32 or bf16)
//   z     = SiLU?( FiLM?( LN2(u) ) )                            -> out2 (bf16)
//
// Every LayerNorm of MoEExtendedDecoderLayer except the MoE gate's follows a Linear whose output IS the whole row
// (N = latent_dim = 512: models/fast_attention.py:142,166-176,210,225; models/stylization.py:27-30;
// models/transformer.py:55-64), so the tile is 128 rows x 512 columns = the whole TMEM (128 lanes x 512 columns of
// fp32): a row lives in ONE TMEM lane, i.e. in one thread of the epilogue.  Row statistics are therefore plain
// per-thread sums over tcgen05.ld chunks (no shuffles; the two warps that share a lane quadrant exchange two floats per
// row through shared memory), and TMEM doubles as the row buffer between the passes (tcgen05.st writes the finished
// row back over the accumulator):
//   pass A  acc -> +bias, activation, alpha, + beta * residual (TMA in) -> y out (TMA) ; sum / sum of squares of s; s -> TMEM
//   pass B  s -> u = LN1(s) -> out1 (TMA) ; sum / sum of squares of u  (they give |u| for the L2 norm AND the LN2
//           statistics of u * sqrt(D)/|u| in closed form)
//   pass C  s -> u -> z -> out2 (TMA)
// The MMA side is the CTA-pair scheme of gemm_tc2_kernel (cluster of two, tcgen05 cta_group::2, M = 256): each CTA
// stages its own 128 rows of A and HALF of the weight rows of each 256-column half, so a k-block costs an SM 48 KB of
// L2 -> shared-memory traffic for 128 x 512 x 64 MACs (the single-CTA 128 x 256 tile pays the same 48 KB for half of
// that, and is bound by exactly this feed).  One accumulator per CTA: MMA and epilogue of a tile do not overlap, the
// TMA ring (3 stages) refills during the epilogue.
// What it replaces at batch 64 (B200, tools/op_bench.py): p3 GEMM 17.6 us + five-stage rowop 31 us; fp32 + residual
// GEMM 27 us + LayerNorm rowop 15 us.
#include <stdlib.h>
#include "gemm_epilogue.cuh"
#include "cluster.cuh"
#include "tensormap.cuh"

namespace {

constexpr int LN_N = 512;
constexpr int LN_STAGES = 3;
constexpr int LN_EPI_WARPS = 8;
constexpr int LN_THREADS = (FIRST_EPI_WARP + LN_EPI_WARPS) * 32;

struct LnSmem {
  static constexpr int A_BYTES = BM * BK * 2;                      // this CTA's 128 rows of A
  static constexpr int BH_BYTES = 128 * BK * 2;                    // this CTA's 128 weight rows of one 256-column half
  static constexpr int STAGE_BYTES = A_BYTES + 2 * BH_BYTES;       // 48 KB
  static constexpr int TILE_OFF = LN_STAGES * STAGE_BYTES;         // per epilogue warp 8 KB: 2 fp32 / 4 bf16 staging tiles
  static constexpr int PRM_OFF = TILE_OFF + LN_EPI_WARPS * 8192;   // bias, ln1_w, ln1_b, ln2_w, ln2_b
  static constexpr int RED_OFF = PRM_OFF + 5 * LN_N * 4;           // float2 [2 column halves][128 rows]
  static constexpr int BAR_OFF = RED_OFF + 2 * 128 * 8;
  // full / empty ring, acc_full, acc_empty, 2 residual barriers per epilogue warp, TMEM pointer; + alignment slack
  static constexpr int TOTAL = BAR_OFF + (2 * LN_STAGES + 2 + 2 * LN_EPI_WARPS) * 8 + 16 + 1024;
};
static_assert(LnSmem::TOTAL <= 227 * 1024, "shared memory budget");

enum {
  LF_RESID = 1, LF_OUT_Y = 2, LF_COPY_S = 4, LF_LN_PRE = 8, LF_L2 = 16, LF_OUT1_F32 = 32, LF_OUT1_A = 64, LF_LN2 = 128,
  LF_FILM = 256, LF_SILU = 512, LF_OUT2 = 1024
};

struct LnArgs {
  const float* bias;
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  const float* film;
  float alpha, beta;
  int rows_per_seq;
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void quad_sync(int quad) {   // the two epilogue warps of one TMEM lane quadrant
  asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(src)) : "memory");
}

template <int FLAGS, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LN_THREADS, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmY,
               const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmO1F,
               const __grid_constant__ CUtensorMap tmO1A, const __grid_constant__ CUtensorMap tmO2, int M, int K,
               const LnArgs a) {
  using L = LnSmem;
  constexpr bool RESID = (FLAGS & LF_RESID) != 0, OUT_Y = (FLAGS & LF_OUT_Y) != 0, COPY_S = (FLAGS & LF_COPY_S) != 0;
  constexpr bool LN_PRE = (FLAGS & LF_LN_PRE) != 0, L2N = (FLAGS & LF_L2) != 0, OUT1_F32 = (FLAGS & LF_OUT1_F32) != 0;
  constexpr bool OUT1_A = (FLAGS & LF_OUT1_A) != 0, LN2 = (FLAGS & LF_LN2) != 0, FILM = (FLAGS & LF_FILM) != 0;
  constexpr bool SILU = (FLAGS & LF_SILU) != 0, OUT2 = (FLAGS & LF_OUT2) != 0;
  constexpr bool PASS_C = OUT2;                          // z is only computed when it has a destination
  constexpr bool STATS2 = PASS_C && (L2N || LN2);        // pass B accumulates the statistics of u
  constexpr bool PASS_B = OUT1_F32 || OUT1_A || STATS2 || (COPY_S && !PASS_C);
  static_assert(!(OUT1_F32 && OUT1_A), "out1 has one staging area: fp32 or bf16");
  static_assert(!(OUT1_F32 && COPY_S && !PASS_C), "the bf16 copy of s shares the staging area of out1");
  static_assert(!L2N || !(OUT1_F32 || OUT1_A), "out1 after the L2 norm is not built (no caller)");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* prm = reinterpret_cast<float*>(smem + L::PRM_OFF);
  float2* red = reinterpret_cast<float2*>(smem + L::RED_OFF);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + LN_STAGES;
  uint64_t* acc_full = empty_bar + LN_STAGES;
  uint64_t* acc_empty = acc_full + 1;
  uint64_t* res_bar = acc_empty + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_bar + 2 * LN_EPI_WARPS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < LN_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 2 * LN_EPI_WARPS);
#pragma unroll
    for (int i = 0; i < 2 * LN_EPI_WARPS; ++i) mbar_init(&res_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_ptr, LN_N);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_enter();   // first global access below

  const int num_m_tiles = (M + BM - 1) / BM;
  const int num_pairs = (num_m_tiles + 1) >> 1;
  const int num_kb = K / BK;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp < FIRST_EPI_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0 && lane == 0) {
      // ------------------------------------------------------------ TMA producer (both CTAs)
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cluster_id; w < num_pairs; w += num_clusters) {
        const int a_row0 = (2 * w + (int)rank) * BM;       // past M for the odd tile out: zero-filled by the TMA
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);   // the leader's barrier
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
          tma_load_2d_2sm(&tmA, fb, sa, kb * BK, a_row0);
          tma_load_2d_2sm(&tmB, fb, sa + L::A_BYTES, kb * BK, (int)rank * 128);
          tma_load_2d_2sm(&tmB, fb, sa + L::A_BYTES + L::BH_BYTES, kb * BK, 256 + (int)rank * 128);
          if (++stage == LN_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 && lane == 0 && leader) {
      // ------------------------------------------------------------ MMA issuer (leader CTA only)
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, 256);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int w = cluster_id; w < num_pairs; w += num_clusters) {
        mbar_wait(acc_empty, acc_phase ^ 1);               // the epilogues of both CTAs are done with TMEM
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint64_t adesc = make_sw128_kmajor_desc(sa);
          const uint64_t bdesc0 = make_sw128_kmajor_desc(sa + L::A_BYTES);
          const uint64_t bdesc1 = make_sw128_kmajor_desc(sa + L::A_BYTES + L::BH_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            umma_bf16_2sm(tmem_base, adesc + 2 * k, bdesc0 + 2 * k, idesc, (kb | k) != 0);
            umma_bf16_2sm(tmem_base + 256, adesc + 2 * k, bdesc1 + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit_2sm(&empty_bar[stage]);
          if (++stage == LN_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(acc_full);
        acc_phase ^= 1;
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ------------------------------------------------------------ epilogue (8 warps per CTA, own 128 rows)
    const int widx = warp - FIRST_EPI_WARP;
    const int quad = warp & 3, cpar = widx >> 2;
    uint8_t* wtile = smem + L::TILE_OFF + widx * 8192;
    float4* ftile = reinterpret_cast<float4*>(wtile);          // two 4 KB fp32 tiles (32 rows x 128 B, SWIZZLE_128B)
    uint4* btile = reinterpret_cast<uint4*>(wtile);            // four 2 KB bf16 tiles (32 rows x 64 B, SWIZZLE_64B)
    uint64_t* rbar = res_bar + 2 * widx;
    uint32_t rphase = 0;
    const float* bias_s = prm;
    const float* w1_s = prm + LN_N;
    const float* b1_s = prm + 2 * LN_N;
    const float* w2_s = prm + 3 * LN_N;
    const float* b2_s = prm + 4 * LN_N;
    {
      const int t = threadIdx.x - FIRST_EPI_WARP * 32;         // 0..255: two columns each
      for (int i = t; i < LN_N; i += LN_EPI_WARPS * 32) {
        prm[i] = a.bias ? a.bias[i] : 0.f;
        prm[LN_N + i] = a.ln1_w[i];
        prm[2 * LN_N + i] = a.ln1_b[i];
        if (LN2) { prm[3 * LN_N + i] = a.ln2_w[i]; prm[4 * LN_N + i] = a.ln2_b[i]; }
      }
      asm volatile("bar.sync 9, 256;" ::: "memory");            // the epilogue warps only
    }
    const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const float inv_n = 1.0f / (float)LN_N;
    uint32_t acc_phase = 0;
    for (int w = cluster_id; w < num_pairs; w += num_clusters) {
      const int row0 = (2 * w + (int)rank) * BM + quad * 32;    // first row of this warp's 32-row block
      const int r = row0 + lane;
      auto issue_res = [&](int k) {       // lane 0: residual chunk k -> fp32 tile k & 1
        mbar_expect_tx(&rbar[k & 1], 4096);
        tma_load_2d(&tmR, &rbar[k & 1], ftile + (k & 1) * 256, (cpar + 2 * k) * 32, row0);
      };
      if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // stores of the previous tile have left
        if (RESID) { issue_res(0); issue_res(1); }
      }
      __syncwarp();
      mbar_wait(acc_full, acc_phase);
      tc_fence_after();
      acc_phase ^= 1;

      // ---------------------------------------------------------------- pass A
      float sum = 0.f, sq = 0.f;
#pragma unroll 1
      for (int k = 0; k < 8; ++k) {
        const int n0 = (cpar + 2 * k) * 32;
        uint32_t raw[32];
        tmem_ld32(t_addr + n0, raw);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + n0 + 4 * j);
          v[4 * j] = __uint_as_float(raw[4 * j]) + b4.x;
          v[4 * j + 1] = __uint_as_float(raw[4 * j + 1]) + b4.y;
          v[4 * j + 2] = __uint_as_float(raw[4 * j + 2]) + b4.z;
          v[4 * j + 3] = __uint_as_float(raw[4 * j + 3]) + b4.w;
        }
        if (ACT == MDM_ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 g = gelu_tanh_fit2(make_float2(v[j], v[j + 1]));
            v[j] = g.x; v[j + 1] = g.y;
          }
        }
        if (a.alpha != 1.0f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= a.alpha;
        }
        if (LN_PRE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) { sum += v[j]; sq = fmaf(v[j], v[j], sq); raw[j] = __float_as_uint(v[j]); }
          tmem_st32(t_addr + n0, raw);
        }
        float4* t4 = ftile + (k & 1) * 256 + lane * 8;
        if (RESID) {
          mbar_wait(&rbar[k & 1], (rphase >> (k & 1)) & 1u);    // residual chunk k has landed
          rphase ^= 1u << (k & 1);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int slot = j ^ (lane & 7);
            const float4 r4 = t4[slot];
            v[4 * j] = fmaf(a.beta, r4.x, v[4 * j]);
            v[4 * j + 1] = fmaf(a.beta, r4.y, v[4 * j + 1]);
            v[4 * j + 2] = fmaf(a.beta, r4.z, v[4 * j + 2]);
            v[4 * j + 3] = fmaf(a.beta, r4.w, v[4 * j + 3]);
            if (OUT_Y) t4[slot] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        } else if (OUT_Y) {
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // tile k & 1 is free again
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) t4[j ^ (lane & 7)] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if (!LN_PRE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) { sum += v[j]; sq = fmaf(v[j], v[j], sq); raw[j] = __float_as_uint(v[j]); }
          tmem_st32(t_addr + n0, raw);
        }
        if (OUT_Y) fence_proxy_async();
        if (OUT_Y || RESID) __syncwarp();                       // every lane is done with the tile
        if (lane == 0) {
          if (OUT_Y) {
            tma_store_2d(&tmY, ftile + (k & 1) * 256, n0, row0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (RESID && k + 2 < 8) {
            if (OUT_Y) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the store has read the tile
            issue_res(k + 2);
          }
        }
      }
      tmem_st_wait();
      // row statistics of s: this warp's 256 columns + the partner warp's
      red[cpar * 128 + quad * 32 + lane] = make_float2(sum, sq);
      quad_sync(quad);
      {
        const float2 o = red[(cpar ^ 1) * 128 + quad * 32 + lane];
        sum += o.x;
        sq += o.y;
      }
      quad_sync(quad);
      const float mean1 = sum * inv_n;
      const float rstd1 = rsqrtf(fmaxf(fmaf(sq, inv_n, -mean1 * mean1), 0.f) + 1e-5f);

      // ---------------------------------------------------------------- pass B
      float su = 0.f, squ = 0.f;
      if (PASS_B) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll 1
        for (int k = 0; k < 8; ++k) {
          const int n0 = (cpar + 2 * k) * 32;
          uint32_t raw[32];
          tmem_ld32(t_addr + n0, raw);
          tmem_ld_wait();
          float u[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 w4 = *reinterpret_cast<const float4*>(w1_s + n0 + 4 * j);
            const float4 b4 = *reinterpret_cast<const float4*>(b1_s + n0 + 4 * j);
            u[4 * j] = fmaf((__uint_as_float(raw[4 * j]) - mean1) * rstd1, w4.x, b4.x);
            u[4 * j + 1] = fmaf((__uint_as_float(raw[4 * j + 1]) - mean1) * rstd1, w4.y, b4.y);
            u[4 * j + 2] = fmaf((__uint_as_float(raw[4 * j + 2]) - mean1) * rstd1, w4.z, b4.z);
            u[4 * j + 3] = fmaf((__uint_as_float(raw[4 * j + 3]) - mean1) * rstd1, w4.w, b4.w);
          }
          if (STATS2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { su += u[j]; squ = fmaf(u[j], u[j], squ); }
          }
          if (OUT1_F32 || OUT1_A || (COPY_S && !PASS_C)) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // buffers of chunk k - 2
            __syncwarp();
            if (OUT1_F32) {
              float4* t4 = ftile + (k & 1) * 256 + lane * 8;
#pragma unroll
              for (int j = 0; j < 8; ++j) t4[j ^ (lane & 7)] = make_float4(u[4 * j], u[4 * j + 1], u[4 * j + 2], u[4 * j + 3]);
            }
            if (OUT1_A) {
              uint4* bt = btile + ((k & 1) * 2) * 128 + lane * 4;
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4) {
                uint4 pk;
                pk.x = pack2(u[8 * c4], u[8 * c4 + 1]); pk.y = pack2(u[8 * c4 + 2], u[8 * c4 + 3]);
                pk.z = pack2(u[8 * c4 + 4], u[8 * c4 + 5]); pk.w = pack2(u[8 * c4 + 6], u[8 * c4 + 7]);
                bt[c4 ^ ((lane >> 1) & 3)] = pk;
              }
            }
            if (COPY_S && !PASS_C) {
              uint4* bt = btile + ((k & 1) * 2 + 1) * 128 + lane * 4;
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4) {
                uint4 pk;
                pk.x = pack2(__uint_as_float(raw[8 * c4]), __uint_as_float(raw[8 * c4 + 1]));
                pk.y = pack2(__uint_as_float(raw[8 * c4 + 2]), __uint_as_float(raw[8 * c4 + 3]));
                pk.z = pack2(__uint_as_float(raw[8 * c4 + 4]), __uint_as_float(raw[8 * c4 + 5]));
                pk.w = pack2(__uint_as_float(raw[8 * c4 + 6]), __uint_as_float(raw[8 * c4 + 7]));
                bt[c4 ^ ((lane >> 1) & 3)] = pk;
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              if (OUT1_F32) tma_store_2d(&tmO1F, ftile + (k & 1) * 256, n0, row0);
              if (OUT1_A) tma_store_2d(&tmO1A, btile + ((k & 1) * 2) * 128, n0, row0);
              if (COPY_S && !PASS_C) tma_store_2d(&tmS, btile + ((k & 1) * 2 + 1) * 128, n0, row0);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
        }
      }

      // ---------------------------------------------------------------- pass C
      if (PASS_C) {
        float mean_u = 0.f, g = 1.0f;
        if (STATS2) {
          red[cpar * 128 + quad * 32 + lane] = make_float2(su, squ);
          quad_sync(quad);
          const float2 o = red[(cpar ^ 1) * 128 + quad * 32 + lane];
          su += o.x;
          squ += o.y;
          quad_sync(quad);
          mean_u = su * inv_n;
          const float var_u = fmaxf(fmaf(squ, inv_n, -mean_u * mean_u), 0.f);
          // F.normalize(u) * sqrt(D): b = u * sc; LayerNorm(b) has mean sc * mean_u and variance sc^2 * var_u
          const float sc = L2N ? sqrtf((float)LN_N) / fmaxf(sqrtf(squ), 1e-12f) : 1.0f;
          g = LN2 ? sc * rsqrtf(sc * sc * var_u + 1e-5f) : sc;
          if (!LN2) mean_u = 0.f;
        }
        const int rc = min(r, M - 1);                            // rows past M only need a valid FiLM address
        const float* fp = FILM ? a.film + (long)(rc / a.rows_per_seq) * (2 * LN_N) : nullptr;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll 1
        for (int k = 0; k < 8; ++k) {
          const int n0 = (cpar + 2 * k) * 32;
          uint32_t raw[32];
          tmem_ld32(t_addr + n0, raw);
          tmem_ld_wait();
          float z[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 w4 = *reinterpret_cast<const float4*>(w1_s + n0 + 4 * j);
            const float4 b4 = *reinterpret_cast<const float4*>(b1_s + n0 + 4 * j);
            z[4 * j] = fmaf((__uint_as_float(raw[4 * j]) - mean1) * rstd1, w4.x, b4.x);
            z[4 * j + 1] = fmaf((__uint_as_float(raw[4 * j + 1]) - mean1) * rstd1, w4.y, b4.y);
            z[4 * j + 2] = fmaf((__uint_as_float(raw[4 * j + 2]) - mean1) * rstd1, w4.z, b4.z);
            z[4 * j + 3] = fmaf((__uint_as_float(raw[4 * j + 3]) - mean1) * rstd1, w4.w, b4.w);
          }
          if (LN2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 w4 = *reinterpret_cast<const float4*>(w2_s + n0 + 4 * j);
              const float4 b4 = *reinterpret_cast<const float4*>(b2_s + n0 + 4 * j);
              z[4 * j] = fmaf((z[4 * j] - mean_u) * g, w4.x, b4.x);
              z[4 * j + 1] = fmaf((z[4 * j + 1] - mean_u) * g, w4.y, b4.y);
              z[4 * j + 2] = fmaf((z[4 * j + 2] - mean_u) * g, w4.z, b4.z);
              z[4 * j + 3] = fmaf((z[4 * j + 3] - mean_u) * g, w4.w, b4.w);
            }
          } else if (L2N) {
#pragma unroll
            for (int j = 0; j < 32; ++j) z[j] *= g;
          }
          if (FILM) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 sc4 = __ldg(reinterpret_cast<const float4*>(fp + n0 + 4 * j));
              const float4 sh4 = __ldg(reinterpret_cast<const float4*>(fp + LN_N + n0 + 4 * j));
              z[4 * j] = fmaf(z[4 * j], 1.f + sc4.x, sh4.x);
              z[4 * j + 1] = fmaf(z[4 * j + 1], 1.f + sc4.y, sh4.y);
              z[4 * j + 2] = fmaf(z[4 * j + 2], 1.f + sc4.z, sh4.z);
              z[4 * j + 3] = fmaf(z[4 * j + 3], 1.f + sc4.w, sh4.w);
            }
          }
          if (SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) z[j] = silu_fast(z[j]);
          }
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");     // buffers of chunk k - 2
          __syncwarp();
          {
            uint4* bt = btile + ((k & 1) * 2) * 128 + lane * 4;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              uint4 pk;
              pk.x = pack2(z[8 * c4], z[8 * c4 + 1]); pk.y = pack2(z[8 * c4 + 2], z[8 * c4 + 3]);
              pk.z = pack2(z[8 * c4 + 4], z[8 * c4 + 5]); pk.w = pack2(z[8 * c4 + 6], z[8 * c4 + 7]);
              bt[c4 ^ ((lane >> 1) & 3)] = pk;
            }
          }
          if (COPY_S) {
            uint4* bt = btile + ((k & 1) * 2 + 1) * 128 + lane * 4;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              uint4 pk;
              pk.x = pack2(__uint_as_float(raw[8 * c4]), __uint_as_float(raw[8 * c4 + 1]));
              pk.y = pack2(__uint_as_float(raw[8 * c4 + 2]), __uint_as_float(raw[8 * c4 + 3]));
              pk.z = pack2(__uint_as_float(raw[8 * c4 + 4]), __uint_as_float(raw[8 * c4 + 5]));
              pk.w = pack2(__uint_as_float(raw[8 * c4 + 6]), __uint_as_float(raw[8 * c4 + 7]));
              bt[c4 ^ ((lane >> 1) & 3)] = pk;
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmO2, btile + ((k & 1) * 2) * 128, n0, row0);
            if (COPY_S) tma_store_2d(&tmS, btile + ((k & 1) * 2 + 1) * 128, n0, row0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      // this warp is done with TMEM: the next tile's MMAs may overwrite its columns
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(acc_empty), 0));
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // TMA stores of this warp have left shared memory
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, LN_N);
  }
}

template <int FLAGS, int ACT>
int launch_ln(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tr, const CUtensorMap& ty, const CUtensorMap& ts,
              const CUtensorMap& t1f, const CUtensorMap& t1a, const CUtensorMap& t2, int M, int K, const LnArgs& a,
              cudaStream_t st) {
  static unsigned long long attr_set = 0;   // one bit per device ordinal
  const unsigned long long dev_bit = 1ull << mdm_cur_dev();
  if (!(attr_set & dev_bit)) {
    if (cudaFuncSetAttribute(gemm_ln_kernel<FLAGS, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, LnSmem::TOTAL) !=
        cudaSuccess)
      return MDM_ERR_CUDA;
    attr_set |= dev_bit;
  }
  const int pairs = ((M + BM - 1) / BM + 1) / 2;
  int clusters = num_sms() / 2;
  if (pairs < clusters) clusters = pairs;
  if (clusters < 1) clusters = 1;
  return mdm_launch(gemm_ln_kernel<FLAGS, ACT>, (unsigned)(2 * clusters), LN_THREADS, LnSmem::TOTAL, st, ta, tb, tr, ty, ts,
                    t1f, t1a, t2, M, K, a) == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

}  // namespace

// C-ABI: see include/mdm_b200.h
extern "C" MDM_API int mdm_gemm_ln(const void* A, int lda, long a_rows, const void* W, int ldw, long w_rows, int M, int N, int K,
                                   const MdmGemmEpi* epi, const MdmRowOp* op, void* stream) {
  if (!A || !W || !epi || !op || M <= 0 || K <= 0) return MDM_ERR_ARG;
  if (N != LN_N || (K % BK) != 0 || w_rows < LN_N) return MDM_ERR_UNSUPPORTED;
  if ((lda & 7) || (ldw & 7)) return MDM_ERR_ARG;
  auto al = [](const void* p, uintptr_t n) { return (reinterpret_cast<uintptr_t>(p) & (n - 1)) == 0; };
  if (!al(A, 16) || !al(W, 16)) return MDM_ERR_ARG;
  if (epi->rowscale || epi->rowmask || epi->tile_k || epi->mn_major || epi->resid_mod > 0) return MDM_ERR_UNSUPPORTED;
  if (epi->act != MDM_ACT_NONE && epi->act != MDM_ACT_GELU) return MDM_ERR_UNSUPPORTED;
  if (!op->ln1_w || !op->ln1_b || op->out2_f32 || op->out0_a) return MDM_ERR_UNSUPPORTED;
  if ((op->ln2_w != nullptr) != (op->ln2_b != nullptr)) return MDM_ERR_ARG;
  if (op->film && op->rows_per_seq <= 0) return MDM_ERR_ARG;
  if (!op->out2_a && (op->ln2_w || op->film || op->silu || op->l2norm)) return MDM_ERR_ARG;   // stages without a destination
  int flags = 0;
  if (epi->resid) flags |= LF_RESID;
  if (epi->out_f32) flags |= LF_OUT_Y;
  if (epi->out_bf16) flags |= LF_COPY_S;
  if (epi->bf16_pre_resid && epi->resid) flags |= LF_LN_PRE;
  if (op->l2norm) flags |= LF_L2;
  if (op->out1_f32) flags |= LF_OUT1_F32;
  if (op->out1_a) flags |= LF_OUT1_A;
  if (op->ln2_w) flags |= LF_LN2;
  if (op->film) flags |= LF_FILM;
  if (op->silu) flags |= LF_SILU;
  if (op->out2_a) flags |= LF_OUT2;
  CUtensorMap ta, tb;
  if (!make_map(&ta, A, a_rows, K, lda, BM) || !make_map(&tb, W, w_rows, K, ldw, 128)) return MDM_ERR_CUDA;
  CUtensorMap tr = ta, ty = ta, ts = ta, t1f = ta, t1a = ta, t2 = ta;
  bool ok = true;
  if (epi->resid) ok = ok && (epi->ld_resid & 3) == 0 && al(epi->resid, 16) && make_f32_map(&tr, epi->resid, M, N, epi->ld_resid);
  if (epi->out_f32) ok = ok && (epi->ld_f32 & 3) == 0 && al(epi->out_f32, 16) && make_f32_map(&ty, epi->out_f32, M, N, epi->ld_f32);
  if (epi->out_bf16) ok = ok && (epi->ld_bf16 & 7) == 0 && al(epi->out_bf16, 16) && make_out_map(&ts, epi->out_bf16, M, N, epi->ld_bf16);
  if (op->out1_f32) ok = ok && al(op->out1_f32, 16) && make_f32_map(&t1f, op->out1_f32, M, N, N);
  if (op->out1_a) ok = ok && al(op->out1_a, 16) && make_out_map(&t1a, op->out1_a, M, N, N);
  if (op->out2_a) ok = ok && al(op->out2_a, 16) && make_out_map(&t2, op->out2_a, M, N, N);
  if (!ok) return MDM_ERR_UNSUPPORTED;
  LnArgs a;
  a.bias = epi->bias;
  a.ln1_w = op->ln1_w; a.ln1_b = op->ln1_b; a.ln2_w = op->ln2_w; a.ln2_b = op->ln2_b;
  a.film = op->film;
  a.alpha = epi->alpha;
  a.beta = epi->beta;
  a.rows_per_seq = op->rows_per_seq > 0 ? op->rows_per_seq : 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int act = epi->act;
  // the stage sets of MotionTransformer._layer (one instantiation each: the epilogue stays inside the instruction cache)
#define MDM_LN(F_, A_) if (flags == (F_) && act == (A_)) return launch_ln<(F_), (A_)>(ta, tb, tr, ty, ts, t1f, t1a, t2, M, K, a, st)
  // Performer projection -> post LN -> L2 norm -> StylizationBlock LN, FiLM, SiLU (fast_attention.py:166-176, stylization.py:27-30)
  MDM_LN(LF_L2 | LF_LN2 | LF_FILM | LF_SILU | LF_OUT2, MDM_ACT_NONE);
  // StylizationBlock output Linear + residual -> the next block's pre-norm (fast_attention.py:142)
  MDM_LN(LF_RESID | LF_OUT_Y | LF_OUT1_A, MDM_ACT_NONE);
  // DualSelfAttentionBlock skip Linear + GELU + residual -> post norm (kept in fp32) -> cross-attention norm (fast_attention.py:221-226,248)
  MDM_LN(LF_RESID | LF_OUT1_F32 | LF_LN2 | LF_OUT2, MDM_ACT_GELU);
  // MemoryEfficientCrossAttentionBlock output projection: residual sum out, LayerNorm of the projection itself (fast_attention.py:322-326)
  MDM_LN(LF_RESID | LF_OUT_Y | LF_LN_PRE | LF_OUT1_A, MDM_ACT_NONE);
  // feed-forward output Linear + residual = the layer output -> the next layer's two pre-norms, + its bf16 copy (fast_attention.py:210-215)
  MDM_LN(LF_RESID | LF_OUT_Y | LF_COPY_S | LF_OUT1_F32 | LF_LN2 | LF_OUT2, MDM_ACT_NONE);
#undef MDM_LN
  return MDM_ERR_UNSUPPORTED;
}
