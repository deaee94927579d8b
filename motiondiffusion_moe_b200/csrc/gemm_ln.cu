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
// models/transformer.py:55-64).  A cluster of two CTAs owns a 128-row block: CTA r computes columns [256 r, 256 r + 256)
// of those rows (its own tcgen05 cta_group::1 GEMM: A tile + its half of the weight rows by TMA, 3-stage ring), so a
// row's 512 values live in ONE TMEM lane of each of the two SMs, i.e. in one epilogue thread per SM.  Row statistics
// are plain per-thread sums over tcgen05.ld chunks (no shuffles); the two warps of a CTA that share a lane quadrant
// exchange their partial sums through shared memory, the two CTAs through distributed shared memory + an mbarrier.
// TMEM doubles as the row buffer between the passes (tcgen05.st writes the finished row back over the accumulator):
//   pass A  acc -> +bias, activation, alpha, + beta * residual (TMA in) -> y out (TMA) ; sum / sum of squares of s; s -> TMEM
//   pass B  s -> u = LN1(s) -> out1 (TMA) ; sum / sum of squares of u  (they give |u| for the L2 norm AND the LN2
//           statistics of u * sqrt(D)/|u| in closed form)
//   pass C  s -> u -> z -> out2 (TMA)
// 256 columns per CTA = half of TMEM: two accumulator buffers, the MMAs of the next row block run under the three
// epilogue passes of the current one.  (First version, commit 1c0bf6d: 128 x 512 tiles = all of TMEM per CTA,
// cta_group::2 pairs; epilogue and MMA serialised and 196 tiles on 148 SMs left the second wave two thirds empty:
// 42 us where this layout runs 196 half-size tile pairs in 2.65 waves of 74 clusters.)
// What it replaces at batch 64 (B200, tools/op_bench.py): p3 GEMM 17.6 us + five-stage rowop 31 us; fp32 + residual
// GEMM 27 us + LayerNorm rowop 15 us.
#include <stdlib.h>
#include "gemm_epilogue.cuh"
#include "cluster.cuh"
#include "tensormap.cuh"
#include "moe_topk.cuh"

namespace {

constexpr int LN_N = 512;            // row width
constexpr int LN_NC = 256;           // columns per CTA
constexpr int LN_STAGES = 3;
constexpr int LN_EPI_WARPS = 8;
constexpr int LN_THREADS = (FIRST_EPI_WARP + LN_EPI_WARPS) * 32;
constexpr int LN_CH = 4;             // 32-column chunks per epilogue warp and pass

constexpr int LN_G = 16;             // gate stage: 2 branches x 8 experts (MoEMultiBranchFFN of the default model)
template <bool GATE>
struct LnSmemT {
  // the gate stage keeps its 16 weight rows (16 KB per CTA) and the logit exchange in shared memory and pays for them
  // with one operand stage: with K = 512 the MMAs of the next block hide under the epilogue either way
  static constexpr int STAGES = GATE ? 2 : LN_STAGES;
  static constexpr int A_BYTES = BM * BK * 2;                      // 128 rows of A
  static constexpr int B_BYTES = LN_NC * BK * 2;                   // this CTA's 256 weight rows
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;            // 48 KB
  static constexpr int TILE_OFF = STAGES * STAGE_BYTES;            // per epilogue warp 8 KB: 2 fp32 / 4 bf16 staging tiles
  // this CTA's columns of: bias, ln1_w, ln1_b, ln2_w, ln2_b | gate: bias, ln_w[2], ln_b[2], gate_w[16]
  static constexpr int PRM_OFF = TILE_OFF + LN_EPI_WARPS * 8192;
  static constexpr int PRM_FLOATS = GATE ? (1 + 4 + LN_G) * LN_NC : 5 * LN_NC;
  static constexpr int RED_OFF = PRM_OFF + PRM_FLOATS * 4;         // float2 [2 column halves][128 rows]: inside the CTA
  static constexpr int XRED_OFF = RED_OFF + 2 * 128 * 8;           // float2 [2 block parities][2 exchanges][128 rows]: from the peer CTA
  // gate: float [128 rows][16] logit partials of the partner warp | float [2 parities][128 rows][8] from the peer CTA |
  // {int2 idx, float2 val} [2 parities][128 rows] results of this CTA's branch
  static constexpr int GATE_OFF = XRED_OFF + 4 * 128 * 8;
  static constexpr int GATE_BYTES = GATE ? 128 * LN_G * 4 + 2 * 128 * 8 * 4 + 2 * 128 * 16 : 0;
  static constexpr int BAR_OFF = GATE_OFF + GATE_BYTES;
  // full / empty ring, tmem_full[2], tmem_empty[2], 2 residual barriers per epilogue warp, 2 x 2 x 4 statistics barriers,
  // 2 x 4 gate barriers, TMEM pointer; + alignment slack
  static constexpr int TOTAL = BAR_OFF + (2 * LN_STAGES + 4 + 2 * LN_EPI_WARPS + 16 + 8) * 8 + 16 + 1024;
};
static_assert(LnSmemT<true>::TOTAL <= 227 * 1024 && LnSmemT<false>::TOTAL <= 227 * 1024, "shared memory budget");

enum {
  LF_RESID = 1, LF_OUT_Y = 2, LF_COPY_S = 4, LF_LN_PRE = 8, LF_L2 = 16, LF_OUT1_F32 = 32, LF_OUT1_A = 64, LF_LN2 = 128,
  LF_FILM = 256, LF_SILU = 512, LF_OUT2 = 1024, LF_GATE = 2048
};

struct LnArgs {
  const float* bias;
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  const float* film;
  float alpha, beta;
  int rows_per_seq;
  // gate stage (LF_GATE): LayerNorm affine of the two branches [2][512], gate weights [16][512] and bias [16];
  // outputs of mdm_moe_gate (include/mdm_b200.h): idx / vals [M][2][2], stats [M][2], per-128-row-block histograms
  const float *gate_ln_w, *gate_ln_b, *gate_w, *gate_b;
  int* idx;
  float* vals;
  float* stats;
  int* blk_hist;
  float* blk_imp;
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
// Partial row statistics for the peer CTA: an asynchronous store into ITS shared memory that signals ITS mbarrier
// with the byte count (st.async ... mbarrier::complete_tx::bytes): the sender does not wait for anything, and the
// receiver's ordinary mbarrier wait orders the data (the same mechanism as a TMA multicast from the peer).  The first
// version used st.shared::cluster + mbarrier.arrive.release.cluster / try_wait.acquire.cluster: a MEMBAR + ERRBAR on the
// sending warp and an L1 invalidation (CCTL.IVALL) on every waiting warp, 17 % of the epilogue's stall samples.
__device__ __forceinline__ void st_async_f2(uint32_t cluster_addr, float2 v, uint32_t cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
               ::"r"(cluster_addr), "f"(v.x), "f"(v.y), "r"(cluster_bar) : "memory");
}

__device__ __forceinline__ void st_async_f4(uint32_t cluster_addr, float4 v, uint32_t cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(cluster_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(cluster_bar) : "memory");
}

// ---- per-chunk (32 columns of one row per thread) helpers, packed fp32 pairs (FFMA2 / FADD2: one issue slot per two
// elements; the epilogue is issue / latency bound, not bandwidth bound)
__device__ __forceinline__ void stats_acc(const float (&v)[32], float2 (&s)[2], float2 (&q)[2]) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float2 p0 = make_float2(v[j], v[j + 1]), p1 = make_float2(v[j + 2], v[j + 3]);
    s[0] = add2(s[0], p0);
    s[1] = add2(s[1], p1);
    q[0] = fma2(p0, p0, q[0]);
    q[1] = fma2(p1, p1, q[1]);
  }
}
// v = (v * rstd + nmr) * w + b, nmr = -mean * rstd; w / b: shared-memory vectors (warp-uniform address: broadcast reads)
__device__ __forceinline__ void ln_apply(float (&v)[32], const float* __restrict__ w_s, const float* __restrict__ b_s, float rstd,
                                         float nmr) {
  const float2 r2 = make_float2(rstd, rstd), m2 = make_float2(nmr, nmr);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 w4 = *reinterpret_cast<const float4*>(w_s + 4 * j);
    const float4 b4 = *reinterpret_cast<const float4*>(b_s + 4 * j);
    const float2 u0 = fma2(fma2(make_float2(v[4 * j], v[4 * j + 1]), r2, m2), make_float2(w4.x, w4.y), make_float2(b4.x, b4.y));
    const float2 u1 = fma2(fma2(make_float2(v[4 * j + 2], v[4 * j + 3]), r2, m2), make_float2(w4.z, w4.w), make_float2(b4.z, b4.w));
    v[4 * j] = u0.x; v[4 * j + 1] = u0.y; v[4 * j + 2] = u1.x; v[4 * j + 3] = u1.y;
  }
}
// SiLU(x) = x * sigmoid(x) = h + h * tanh(h), h = x / 2: one MUFU op per element (ex2 + rcp: two, and the pass that
// applies it is bound by the 16 MUFU lanes of the SM).  MUFU.TANH: 2^-11 relative, i.e. <= 2.5e-4 of sigmoid: a
// sixteenth of a bf16 rounding step of the result.
__device__ __forceinline__ void silu32(float (&v)[32]) {
  const float2 half = make_float2(0.5f, 0.5f);
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const float2 h = mul2(make_float2(v[j], v[j + 1]), half);
    float2 t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(h.x));
    asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(h.y));
    const float2 y = fma2(h, t, h);
    v[j] = y.x; v[j + 1] = y.y;
  }
}
// row-per-lane staging tiles in the layouts the TMA store expects
__device__ __forceinline__ void tile_f32(float4* t, int lane, const float (&v)[32]) {       // 32 x 128 B, SWIZZLE_128B
  float4* t4 = t + lane * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) t4[j ^ (lane & 7)] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void tile_bf16(uint4* t, int lane, const float (&v)[32]) {        // 32 x 64 B, SWIZZLE_64B
  uint4* bt = t + lane * 4;
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) {
    uint4 pk;
    pk.x = pack2(v[8 * c4], v[8 * c4 + 1]); pk.y = pack2(v[8 * c4 + 2], v[8 * c4 + 3]);
    pk.z = pack2(v[8 * c4 + 4], v[8 * c4 + 5]); pk.w = pack2(v[8 * c4 + 6], v[8 * c4 + 7]);
    bt[c4 ^ ((lane >> 1) & 3)] = pk;
  }
}
__device__ __forceinline__ void load8(const float* p, float4 (&q)[8]) {     // 128 contiguous bytes of this lane's row
#pragma unroll
  for (int j = 0; j < 8; ++j)
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q[j].x), "=f"(q[j].y), "=f"(q[j].z), "=f"(q[j].w) : "l"(p + 4 * j));
}

template <int FLAGS, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LN_THREADS, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmY,
               const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmO1F,
               const __grid_constant__ CUtensorMap tmO1A, const __grid_constant__ CUtensorMap tmO2, int M, int K,
               const LnArgs a) {
  constexpr bool GATE = (FLAGS & LF_GATE) != 0;
  using L = LnSmemT<GATE>;
  constexpr int NSTG = L::STAGES;
  constexpr bool RESID = (FLAGS & LF_RESID) != 0, OUT_Y = (FLAGS & LF_OUT_Y) != 0, COPY_S = (FLAGS & LF_COPY_S) != 0;
  constexpr bool LN_PRE = (FLAGS & LF_LN_PRE) != 0, L2N = (FLAGS & LF_L2) != 0, OUT1_F32 = (FLAGS & LF_OUT1_F32) != 0;
  constexpr bool OUT1_A = (FLAGS & LF_OUT1_A) != 0, LN2 = (FLAGS & LF_LN2) != 0, FILM = (FLAGS & LF_FILM) != 0;
  constexpr bool SILU = (FLAGS & LF_SILU) != 0, OUT2 = (FLAGS & LF_OUT2) != 0;
  constexpr bool PASS_C = OUT2;                          // z is only computed when it has a destination
  constexpr bool STATS2 = PASS_C && (L2N || LN2);        // pass B accumulates the statistics of u
  constexpr bool PASS_B = OUT1_F32 || OUT1_A || STATS2 || (COPY_S && !PASS_C);
  static_assert(!GATE || !(PASS_B || PASS_C || LN_PRE), "the gate stage is its own second pass");
  static_assert(!(OUT1_F32 && OUT1_A), "out1 has one staging area: fp32 or bf16");
  static_assert(!(OUT1_F32 && COPY_S && !PASS_C), "the bf16 copy of s shares the staging area of out1");
  static_assert(!L2N || !(OUT1_F32 || OUT1_A), "out1 after the L2 norm is not built (no caller)");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* prm = reinterpret_cast<float*>(smem + L::PRM_OFF);
  float2* red = reinterpret_cast<float2*>(smem + L::RED_OFF);
  float2* xred = reinterpret_cast<float2*>(smem + L::XRED_OFF);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + LN_STAGES;     // (LN_STAGES slots reserved; NSTG of them used)
  uint64_t* tmem_full = empty_bar + LN_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_bar = tmem_empty + 2;
  // [2 block parities][2 exchanges][4 quadrants]: one local arrival that expects the 256 bytes the peer warp sends.
  // One barrier per block PARITY: a
  // barrier then completes a phase every second block, and the peer cannot be two blocks ahead (it needs this CTA's
  // totals of the block in between), so a waiter can never be lapped.
  uint64_t* xbar = res_bar + 2 * LN_EPI_WARPS;
  uint64_t* gbar = xbar + 16;                            // gate stage: [2 block parities][4 quadrants], same scheme
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(gbar + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < NSTG; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full[0], 1);
    mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], LN_EPI_WARPS);
    mbar_init(&tmem_empty[1], LN_EPI_WARPS);
#pragma unroll
    for (int i = 0; i < 2 * LN_EPI_WARPS; ++i) mbar_init(&res_bar[i], 1);
#pragma unroll
    for (int i = 0; i < 16; ++i) mbar_init(&xbar[i], 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) mbar_init(&gbar[i], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 2 * LN_NC);
  tc_fence_before();
  cluster_sync_all();          // the peer's exchange barriers are initialised before anything can arrive on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_enter();                 // first global access below

  const int num_blocks = (M + BM - 1) / BM;              // 128-row blocks, one per cluster and iteration
  const int num_kb = K / BK;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int ncol0 = (int)rank * LN_NC;                   // this CTA's first column / weight row

  if (warp < FIRST_EPI_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0 && lane == 0) {
      // ------------------------------------------------------------ TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cluster_id; w < num_blocks; w += num_clusters) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          tma_load_2d(&tmA, &full_bar[stage], sa, kb * BK, w * BM);
          tma_load_2d(&tmB, &full_bar[stage], sa + L::A_BYTES, kb * BK, ncol0);
          if (++stage == NSTG) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ------------------------------------------------------------ MMA issuer
      constexpr uint32_t idesc = make_idesc_bf16(BM, LN_NC);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int w = cluster_id; w < num_blocks; w += num_clusters) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * LN_NC;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint64_t adesc = make_sw128_kmajor_desc(sa);
          const uint64_t bdesc = make_sw128_kmajor_desc(sa + L::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (++stage == NSTG) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ------------------------------------------------------------ epilogue (8 warps: lane quadrant x column half)
    const int widx = warp - FIRST_EPI_WARP;
    const int quad = warp & 3, ch = widx >> 2;
    uint8_t* wtile = smem + L::TILE_OFF + widx * 8192;
    float4* ftile = reinterpret_cast<float4*>(wtile);          // two 4 KB fp32 tiles (32 rows x 128 B, SWIZZLE_128B)
    uint4* btile = reinterpret_cast<uint4*>(wtile);            // four 2 KB bf16 tiles (32 rows x 64 B, SWIZZLE_64B)
    uint64_t* rbar = res_bar + 2 * widx;
    uint32_t rphase = 0;
    const float* bias_s = prm + ch * 128;                      // this warp's 128 columns of the CTA's 256
    const float* w1_s = prm + LN_NC + ch * 128;
    const float* b1_s = prm + 2 * LN_NC + ch * 128;
    const float* w2_s = prm + 3 * LN_NC + ch * 128;
    const float* b2_s = prm + 4 * LN_NC + ch * 128;
    // gate stage: ln_w[br] at prm + (1 + br) * 256, ln_b[br] at prm + (3 + br) * 256, gate_w[g] at prm + (5 + g) * 256
    const float* gprm = prm + ch * 128;
    {
      const int t = threadIdx.x - FIRST_EPI_WARP * 32;         // 0..255: one column each
      prm[t] = a.bias ? a.bias[ncol0 + t] : 0.f;
      if (GATE) {
#pragma unroll
        for (int br = 0; br < 2; ++br) {
          prm[(1 + br) * LN_NC + t] = a.gate_ln_w[br * LN_N + ncol0 + t];
          prm[(3 + br) * LN_NC + t] = a.gate_ln_b[br * LN_N + ncol0 + t];
        }
#pragma unroll
        for (int g = 0; g < LN_G; ++g) prm[(5 + g) * LN_NC + t] = a.gate_w[g * LN_N + ncol0 + t];
      } else {
        prm[LN_NC + t] = a.ln1_w[ncol0 + t];
        prm[2 * LN_NC + t] = a.ln1_b[ncol0 + t];
        if (LN2) { prm[3 * LN_NC + t] = a.ln2_w[ncol0 + t]; prm[4 * LN_NC + t] = a.ln2_b[ncol0 + t]; }
      }
      asm volatile("bar.sync 9, 256;" ::: "memory");            // the epilogue warps only
    }
    const int gcol0 = ncol0 + ch * 128;                        // global column of this warp's chunk 0
    const float inv_n = 1.0f / (float)LN_N;
    int acc = 0;
    uint32_t acc_phase = 0, it = 0;
    // row statistics of the whole 512-column row: own 128 columns + the partner warp's (shared memory) + the peer
    // CTA's 256 (distributed shared memory).  e: exchange index (0 after pass A, 1 after pass B).
    auto row_total = [&](float& s0, float& s1, int e) {
      const int row = quad * 32 + lane;
      red[ch * 128 + row] = make_float2(s0, s1);
      quad_sync(quad);
      const float2 o = red[(ch ^ 1) * 128 + row];
      s0 += o.x;
      s1 += o.y;
      quad_sync(quad);
      const int par = (int)(it & 1);
      const int slot = ((par * 2 + e) * 128) + row;
      uint64_t* xb = &xbar[(par * 2 + e) * 4 + quad];
      if (ch == 0) {                                           // one warp per quadrant sends the CTA's total to the peer
        st_async_f2(mapa_u32(smem_u32(&xred[slot]), rank ^ 1), make_float2(s0, s1), mapa_u32(smem_u32(xb), rank ^ 1));
        if (lane == 0) mbar_expect_tx(xb, 32 * 8);             // own barrier: the peer warp's 32 float2
      }
      mbar_wait(xb, (it >> 1) & 1);
      const float2 p = xred[slot];
      s0 += p.x;
      s1 += p.y;
    };
    for (int w = cluster_id; w < num_blocks; w += num_clusters, ++it) {
      const int row0 = w * BM + quad * 32;                      // first row of this warp's 32-row block
      const int r = row0 + lane;
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * LN_NC + ch * 128;
      auto issue_res = [&](int k) {       // lane 0: residual chunk k -> fp32 tile k & 1
        mbar_expect_tx(&rbar[k & 1], 4096);
        tma_load_2d(&tmR, &rbar[k & 1], ftile + (k & 1) * 256, gcol0 + 32 * k, row0);
      };
      if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // stores of the previous block have left
        if (RESID) { issue_res(0); issue_res(1); }
      }
      __syncwarp();
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();

      // ---------------------------------------------------------------- pass A
      // chunk k: the TMEM load of chunk k + 1 is in flight while chunk k is finished
      float2 sA[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, qA[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      uint32_t ra[32], rb[32];
      auto pass_a = [&](int k, uint32_t (&raw)[32], uint32_t (&nraw)[32]) {
        const int n0 = 32 * k;
        tmem_ld_wait();
        if (k + 1 < LN_CH) tmem_ld32(t_addr + n0 + 32, nraw);
        float v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + n0 + 4 * j);
          const float2 x0 = add2(make_float2(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1])), make_float2(b4.x, b4.y));
          const float2 x1 = add2(make_float2(__uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3])), make_float2(b4.z, b4.w));
          v[4 * j] = x0.x; v[4 * j + 1] = x0.y; v[4 * j + 2] = x1.x; v[4 * j + 3] = x1.y;
        }
        if (ACT == MDM_ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 g = gelu_tanh_fit2(make_float2(v[j], v[j + 1]));
            v[j] = g.x; v[j + 1] = g.y;
          }
        }
        if (a.alpha != 1.0f) {
          const float2 a2 = make_float2(a.alpha, a.alpha);
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 x = mul2(make_float2(v[j], v[j + 1]), a2);
            v[j] = x.x; v[j + 1] = x.y;
          }
        }
        if (LN_PRE) {
          stats_acc(v, sA, qA);
#pragma unroll
          for (int j = 0; j < 32; ++j) raw[j] = __float_as_uint(v[j]);
          tmem_st32(t_addr + n0, raw);
        }
        float4* t4 = ftile + (k & 1) * 256 + lane * 8;
        if (RESID) {
          mbar_wait(&rbar[k & 1], (rphase >> (k & 1)) & 1u);    // residual chunk k has landed
          rphase ^= 1u << (k & 1);
          const float2 b2 = make_float2(a.beta, a.beta);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int slot = j ^ (lane & 7);
            const float4 r4 = t4[slot];
            const float2 y0 = fma2(b2, make_float2(r4.x, r4.y), make_float2(v[4 * j], v[4 * j + 1]));
            const float2 y1 = fma2(b2, make_float2(r4.z, r4.w), make_float2(v[4 * j + 2], v[4 * j + 3]));
            v[4 * j] = y0.x; v[4 * j + 1] = y0.y; v[4 * j + 2] = y1.x; v[4 * j + 3] = y1.y;
            if (OUT_Y) t4[slot] = make_float4(y0.x, y0.y, y1.x, y1.y);
          }
        } else if (OUT_Y) {
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // tile k & 1 is free again
          __syncwarp();
          tile_f32(ftile + (k & 1) * 256, lane, v);
        }
        if (!LN_PRE) {
          stats_acc(v, sA, qA);
#pragma unroll
          for (int j = 0; j < 32; ++j) raw[j] = __float_as_uint(v[j]);
          tmem_st32(t_addr + n0, raw);
        }
        if (OUT_Y) fence_proxy_async();
        if (OUT_Y || RESID) __syncwarp();                       // every lane is done with the tile
        if (lane == 0) {
          if (OUT_Y) {
            tma_store_2d(&tmY, ftile + (k & 1) * 256, gcol0 + n0, row0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (RESID && k + 2 < LN_CH) {
            if (OUT_Y) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the store has read the tile
            issue_res(k + 2);
          }
        }
      };
      tmem_ld32(t_addr, ra);
#pragma unroll 1
      for (int k = 0; k < LN_CH; k += 2) {
        pass_a(k, ra, rb);
        pass_a(k + 1, rb, ra);
      }
      tmem_st_wait();
      float sum = (sA[0].x + sA[0].y) + (sA[1].x + sA[1].y), sq = (qA[0].x + qA[0].y) + (qA[1].x + qA[1].y);
      row_total(sum, sq, 0);
      const float mean1 = sum * inv_n;
      const float rstd1 = rsqrtf(fmaxf(fmaf(sq, inv_n, -mean1 * mean1), 0.f) + 1e-5f);
      const float nmr1 = -mean1 * rstd1;

      // ---------------------------------------------------------------- gate stage (SwitchMoELayer gate of both branches)
      // LayerNorm(y) per branch (shared statistics) -> 16 dot products per row, summed per thread over its 128 columns;
      // partner warp through shared memory, then CTA r finalises branch r: it receives the peer's partial of its 8
      // logits (st.async), adds the bias, softmax + top-2 per row (one thread per row: no shuffles), writes idx / vals
      // / stats, and one warp counts the 128 rows of the block in token order (deterministic histograms).
      if (GATE) {
        float2 lg[LN_G];
#pragma unroll
        for (int g = 0; g < LN_G; ++g) lg[g] = make_float2(0.f, 0.f);
        auto pass_g = [&](int k, uint32_t (&raw)[32], uint32_t (&nraw)[32]) {
          const int n0 = 32 * k;
          tmem_ld_wait();
          if (k + 1 < LN_CH) tmem_ld32(t_addr + n0 + 32, nraw);
#pragma unroll
          for (int br = 0; br < 2; ++br) {
            float h[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) h[j] = __uint_as_float(raw[j]);
            ln_apply(h, gprm + (1 + br) * LN_NC + n0, gprm + (3 + br) * LN_NC + n0, rstd1, nmr1);
            // column-outer, expert-inner: eight independent accumulator chains per h value (the expert-outer order is
            // one 16-long dependent FFMA2 chain per expert)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 h0 = make_float2(h[4 * j], h[4 * j + 1]), h1 = make_float2(h[4 * j + 2], h[4 * j + 3]);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float4 w4 = *reinterpret_cast<const float4*>(gprm + (5 + br * 8 + e) * LN_NC + n0 + 4 * j);
                lg[br * 8 + e] = fma2(h1, make_float2(w4.z, w4.w), fma2(h0, make_float2(w4.x, w4.y), lg[br * 8 + e]));
              }
            }
          }
        };
        tmem_ld32(t_addr, ra);
#pragma unroll 1
        for (int k = 0; k < LN_CH; k += 2) {
          pass_g(k, ra, rb);
          pass_g(k + 1, rb, ra);
        }
        float lp[LN_G];
#pragma unroll
        for (int g = 0; g < LN_G; ++g) lp[g] = lg[g].x + lg[g].y;
        const int row = quad * 32 + lane;
        float4* lred = reinterpret_cast<float4*>(smem + L::GATE_OFF);                 // [128 rows][4 float4]
        const int par = (int)(it & 1);
        float4* xg = lred + 128 * 4 + par * 128 * 2;                                   // [2 parities][128 rows][2 float4]
        uint4* resb = reinterpret_cast<uint4*>(lred + 128 * 4 + 2 * 128 * 2) + par * 128;   // [2 parities][128 rows]
        if (ch == 1) {
#pragma unroll
          for (int q = 0; q < 4; ++q) lred[row * 4 + q] = make_float4(lp[4 * q], lp[4 * q + 1], lp[4 * q + 2], lp[4 * q + 3]);
        }
        quad_sync(quad);
        if (ch == 0) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 o = lred[row * 4 + q];
            lp[4 * q] += o.x; lp[4 * q + 1] += o.y; lp[4 * q + 2] += o.z; lp[4 * q + 3] += o.w;
          }
        }
        quad_sync(quad);
        if (ch == 0) {
          const int mine = (int)rank;                            // CTA r finalises branch r
          float own[8], oth[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) { own[e] = mine ? lp[8 + e] : lp[e]; oth[e] = mine ? lp[e] : lp[8 + e]; }
          uint64_t* gb = &gbar[par * 4 + quad];
          const uint32_t peer_bar = mapa_u32(smem_u32(gb), rank ^ 1);
          st_async_f4(mapa_u32(smem_u32(&xg[row * 2]), rank ^ 1), make_float4(oth[0], oth[1], oth[2], oth[3]), peer_bar);
          st_async_f4(mapa_u32(smem_u32(&xg[row * 2 + 1]), rank ^ 1), make_float4(oth[4], oth[5], oth[6], oth[7]), peer_bar);
          if (lane == 0) mbar_expect_tx(gb, 32 * 32);              // the peer warp's 32 x 8 floats
          mbar_wait(gb, (it >> 1) & 1);
          const float4 p0 = xg[row * 2], p1 = xg[row * 2 + 1];
          float logits[8], probs[8];
          logits[0] = own[0] + p0.x; logits[1] = own[1] + p0.y; logits[2] = own[2] + p0.z; logits[3] = own[3] + p0.w;
          logits[4] = own[4] + p1.x; logits[5] = own[5] + p1.y; logits[6] = own[6] + p1.z; logits[7] = own[7] + p1.w;
#pragma unroll
          for (int e = 0; e < 8; ++e) logits[e] += __ldg(a.gate_b + mine * 8 + e);
          int i0, i1;
          float v0, v1;
          softmax_top2<8>(logits, probs, i0, i1, v0, v1);
          if (r < M) {
            const long o = ((long)r * 2 + mine) * 2;
            *reinterpret_cast<int2*>(a.idx + o) = make_int2(i0, i1);
            *reinterpret_cast<float2*>(a.vals + o) = make_float2(v0, v1);
            if (mine == 0) *reinterpret_cast<float2*>(a.stats + (long)r * 2) = make_float2(mean1, rstd1);
          }
          resb[row] = make_uint4((uint32_t)i0, (uint32_t)i1, __float_as_uint(v0), __float_as_uint(v1));
          asm volatile("bar.sync 10, 128;" ::: "memory");          // the four finalising warps of this CTA
          if (quad == 0 && lane < 8) {
            // group (branch mine, expert lane): rows in token order, slot 0 before slot 1 - the accumulation order of
            // the reference's per-expert loops (switch_moe.py:72-92); fixed order = deterministic importance sums
            const int rows_here = min(BM, M - w * BM);
            int c_all = 0, c_top1 = 0;
            float imp = 0.f;
            for (int rr = 0; rr < rows_here; ++rr) {
              const uint4 q = resb[rr];
              if ((int)q.x == lane) { ++c_all; ++c_top1; imp += __uint_as_float(q.z); }
              if ((int)q.y == lane) { ++c_all; imp += __uint_as_float(q.w); }
            }
            const int g = mine * 8 + lane;
            a.blk_hist[((long)w * 2) * LN_G + g] = c_all;
            a.blk_hist[((long)w * 2 + 1) * LN_G + g] = c_top1;
            a.blk_imp[(long)w * LN_G + g] = imp;
          }
        }
      }

      // FiLM (scale | shift) of this lane's row, chunk 0: requested a whole pass before it is needed; chunk k + 1 is
      // requested as soon as chunk k's values are consumed
      const int rc = min(r, M - 1);                              // rows past M only need a valid FiLM address
      const float* fp = FILM ? a.film + (long)(rc / a.rows_per_seq) * (2 * LN_N) + gcol0 : nullptr;
      float4 fsc[8], fsh[8];
      if (FILM) { load8(fp, fsc); load8(fp + LN_N, fsh); }

      // ---------------------------------------------------------------- pass B
      float2 sB[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, qB[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      if (PASS_B) {
        constexpr bool STORE_B = OUT1_F32 || OUT1_A || (COPY_S && !PASS_C);
        if (STORE_B) {
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
        }
        auto pass_b = [&](int k, uint32_t (&raw)[32], uint32_t (&nraw)[32]) {
          const int n0 = 32 * k;
          tmem_ld_wait();
          if (k + 1 < LN_CH) tmem_ld32(t_addr + n0 + 32, nraw);
          float u[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) u[j] = __uint_as_float(raw[j]);
          if (STORE_B) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // buffers of chunk k - 2
            __syncwarp();
          }
          if (COPY_S && !PASS_C) tile_bf16(btile + ((k & 1) * 2 + 1) * 128, lane, u);       // bf16 copy of s itself
          ln_apply(u, w1_s + n0, b1_s + n0, rstd1, nmr1);
          if (STATS2) stats_acc(u, sB, qB);
          if (STORE_B) {
            if (OUT1_F32) tile_f32(ftile + (k & 1) * 256, lane, u);
            if (OUT1_A) tile_bf16(btile + ((k & 1) * 2) * 128, lane, u);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              if (OUT1_F32) tma_store_2d(&tmO1F, ftile + (k & 1) * 256, gcol0 + n0, row0);
              if (OUT1_A) tma_store_2d(&tmO1A, btile + ((k & 1) * 2) * 128, gcol0 + n0, row0);
              if (COPY_S && !PASS_C) tma_store_2d(&tmS, btile + ((k & 1) * 2 + 1) * 128, gcol0 + n0, row0);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
        };
        tmem_ld32(t_addr, ra);
#pragma unroll 1
        for (int k = 0; k < LN_CH; k += 2) {
          pass_b(k, ra, rb);
          pass_b(k + 1, rb, ra);
        }
      }

      // ---------------------------------------------------------------- pass C
      if (PASS_C) {
        float nmu = 0.f, g = 1.0f;                               // z = (u * g + nmu) * w2 + b2
        if (STATS2) {
          float su = (sB[0].x + sB[0].y) + (sB[1].x + sB[1].y), squ = (qB[0].x + qB[0].y) + (qB[1].x + qB[1].y);
          row_total(su, squ, 1);
          const float mean_u = su * inv_n;
          const float var_u = fmaxf(fmaf(squ, inv_n, -mean_u * mean_u), 0.f);
          // F.normalize(u) * sqrt(D): b = u * sc; LayerNorm(b) has mean sc * mean_u and variance sc^2 * var_u
          const float sc = L2N ? sqrtf((float)LN_N) / fmaxf(sqrtf(squ), 1e-12f) : 1.0f;
          g = LN2 ? sc * rsqrtf(sc * sc * var_u + 1e-5f) : sc;
          nmu = LN2 ? -mean_u * g : 0.f;
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        auto pass_c = [&](int k, uint32_t (&raw)[32], uint32_t (&nraw)[32]) {
          const int n0 = 32 * k;
          tmem_ld_wait();
          if (k + 1 < LN_CH) tmem_ld32(t_addr + n0 + 32, nraw);
          float z[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) z[j] = __uint_as_float(raw[j]);
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");     // buffers of chunk k - 2
          __syncwarp();
          if (COPY_S) tile_bf16(btile + ((k & 1) * 2 + 1) * 128, lane, z);
          ln_apply(z, w1_s + n0, b1_s + n0, rstd1, nmr1);
          if (LN2) {
            ln_apply(z, w2_s + n0, b2_s + n0, g, nmu);
          } else if (L2N) {
#pragma unroll
            for (int j = 0; j < 32; ++j) z[j] *= g;
          }
          if (FILM) {
            const float2 one = make_float2(1.f, 1.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 y0 = fma2(make_float2(z[4 * j], z[4 * j + 1]), add2(make_float2(fsc[j].x, fsc[j].y), one), make_float2(fsh[j].x, fsh[j].y));
              const float2 y1 = fma2(make_float2(z[4 * j + 2], z[4 * j + 3]), add2(make_float2(fsc[j].z, fsc[j].w), one), make_float2(fsh[j].z, fsh[j].w));
              z[4 * j] = y0.x; z[4 * j + 1] = y0.y; z[4 * j + 2] = y1.x; z[4 * j + 3] = y1.y;
            }
            if (k + 1 < LN_CH) { load8(fp + n0 + 32, fsc); load8(fp + LN_N + n0 + 32, fsh); }
          }
          if (SILU) silu32(z);
          tile_bf16(btile + ((k & 1) * 2) * 128, lane, z);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmO2, btile + ((k & 1) * 2) * 128, gcol0 + n0, row0);
            if (COPY_S) tma_store_2d(&tmS, btile + ((k & 1) * 2 + 1) * 128, gcol0 + n0, row0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        };
        tmem_ld32(t_addr, ra);
#pragma unroll 1
        for (int k = 0; k < LN_CH; k += 2) {
          pass_c(k, ra, rb);
          pass_c(k + 1, rb, ra);
        }
      }
      // this warp is done with the accumulator buffer: the MMAs of the block after next may overwrite it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // TMA stores of this warp have left shared memory
  }

  tc_fence_before();
  cluster_sync_all();          // no CTA exits while its peer may still write into its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * LN_NC);
  }
}

template <int FLAGS, int ACT>
int launch_ln(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tr, const CUtensorMap& ty, const CUtensorMap& ts,
              const CUtensorMap& t1f, const CUtensorMap& t1a, const CUtensorMap& t2, int M, int K, const LnArgs& a,
              cudaStream_t st) {
  static unsigned long long attr_set = 0;   // one bit per device ordinal
  const unsigned long long dev_bit = 1ull << mdm_cur_dev();
  if (!(attr_set & dev_bit)) {
    if (cudaFuncSetAttribute(gemm_ln_kernel<FLAGS, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, LnSmemT<(FLAGS & LF_GATE) != 0>::TOTAL) !=
        cudaSuccess)
      return MDM_ERR_CUDA;
    attr_set |= dev_bit;
  }
  const int blocks = (M + BM - 1) / BM;
  int clusters = num_sms() / 2;
  if (blocks < clusters) clusters = blocks;
  if (clusters < 1) clusters = 1;
  return mdm_launch(gemm_ln_kernel<FLAGS, ACT>, (unsigned)(2 * clusters), LN_THREADS, LnSmemT<(FLAGS & LF_GATE) != 0>::TOTAL, st, ta, tb, tr, ty, ts,
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
  if (!make_map(&ta, A, a_rows, K, lda, BM) || !make_map(&tb, W, w_rows, K, ldw, LN_NC)) return MDM_ERR_CUDA;
  CUtensorMap tr = ta, ty = ta, ts = ta, t1f = ta, t1a = ta, t2 = ta;
  bool ok = true;
  if (epi->resid) ok = ok && (epi->ld_resid & 3) == 0 && al(epi->resid, 16) && make_f32_map(&tr, epi->resid, M, N, epi->ld_resid);
  if (epi->out_f32) ok = ok && (epi->ld_f32 & 3) == 0 && al(epi->out_f32, 16) && make_f32_map(&ty, epi->out_f32, M, N, epi->ld_f32);
  if (epi->out_bf16) ok = ok && (epi->ld_bf16 & 7) == 0 && al(epi->out_bf16, 16) && make_out_map(&ts, epi->out_bf16, M, N, epi->ld_bf16);
  if (op->out1_f32) ok = ok && al(op->out1_f32, 16) && make_f32_map(&t1f, op->out1_f32, M, N, N);
  if (op->out1_a) ok = ok && al(op->out1_a, 16) && make_out_map(&t1a, op->out1_a, M, N, N);
  if (op->out2_a) ok = ok && al(op->out2_a, 16) && make_out_map(&t2, op->out2_a, M, N, N);
  if (!ok) return MDM_ERR_UNSUPPORTED;
  LnArgs a = {};
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

// C-ABI: see include/mdm_b200.h
extern "C" MDM_API int mdm_gemm_gate(const void* A, int lda, long a_rows, const void* W, int ldw, long w_rows, int M, int N, int K,
                                     const MdmGemmEpi* epi, int NB, int E, const float* ln_w, const float* ln_b,
                                     const float* gate_w, const float* gate_b, int* idx, float* vals, float* stats,
                                     int* blk_hist, float* blk_imp, void* stream) {
  if (!A || !W || !epi || !ln_w || !ln_b || !gate_w || !gate_b || !idx || !vals || !stats || !blk_hist || !blk_imp || M <= 0 ||
      K <= 0)
    return MDM_ERR_ARG;
  if (N != LN_N || (K % BK) != 0 || w_rows < LN_N || NB != 2 || NB * E != LN_G) return MDM_ERR_UNSUPPORTED;
  if ((lda & 7) || (ldw & 7)) return MDM_ERR_ARG;
  auto al = [](const void* p, uintptr_t n) { return (reinterpret_cast<uintptr_t>(p) & (n - 1)) == 0; };
  if (!al(A, 16) || !al(W, 16)) return MDM_ERR_ARG;
  if (epi->rowscale || epi->rowmask || epi->tile_k || epi->mn_major || epi->resid_mod > 0 || epi->act != MDM_ACT_NONE ||
      epi->out_bf16 || !epi->resid || !epi->out_f32)
    return MDM_ERR_UNSUPPORTED;
  if ((epi->ld_resid & 3) || (epi->ld_f32 & 3) || !al(epi->resid, 16) || !al(epi->out_f32, 16) || !al(idx, 8) || !al(vals, 8) ||
      !al(stats, 8))
    return MDM_ERR_UNSUPPORTED;
  CUtensorMap ta, tb, tr, ty;
  if (!make_map(&ta, A, a_rows, K, lda, BM) || !make_map(&tb, W, w_rows, K, ldw, LN_NC)) return MDM_ERR_CUDA;
  if (!make_f32_map(&tr, epi->resid, M, N, epi->ld_resid) || !make_f32_map(&ty, epi->out_f32, M, N, epi->ld_f32)) return MDM_ERR_CUDA;
  LnArgs a = {};
  a.bias = epi->bias;
  a.alpha = epi->alpha;
  a.beta = epi->beta;
  a.rows_per_seq = 1;
  a.gate_ln_w = ln_w; a.gate_ln_b = ln_b; a.gate_w = gate_w; a.gate_b = gate_b;
  a.idx = idx; a.vals = vals; a.stats = stats; a.blk_hist = blk_hist; a.blk_imp = blk_imp;
  return launch_ln<LF_RESID | LF_OUT_Y | LF_GATE, MDM_ACT_NONE>(ta, tb, tr, ty, ta, ta, ta, ta, M, K, a,
                                                                reinterpret_cast<cudaStream_t>(stream));
}
