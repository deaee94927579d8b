// Shared pieces of the tcgen05 GEMM kernels (gemm_tc.cu: one CTA per 128-row tile; gemm_tc2.cu: CTA pairs,
// cta_group::2): tile constants, fast activations and the per-tile epilogue of one epilogue warp.
#pragma once
#include "common.cuh"

#ifdef MDM_GEMM_PROFILE
__device__ unsigned long long g_epi_phase[8];   // cycles summed over (warp 4 lane 0 of every CTA): see tools/gemm_prof.py
#define EPI_MARK(k) do { if (prof_on) { const long long now_ = clock64(); atomicAdd(&g_epi_phase[k], (unsigned long long)(now_ - tprev_)); tprev_ = now_; } } while (0)
#else
#define EPI_MARK(k) do { } while (0)
#endif

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int FIRST_EPI_WARP = 4;  // warpgroup 0 = {TMA, MMA, 2 idle warps}; the following warpgroups = epilogue
constexpr int TR_BYTES = 32 * 1024;  // transpose buffers of all epilogue warps: 8 x 4 KB or 16 x 2 KB
// Epilogue flavours (each its own kernel instantiation, see below).  EPI_BF16W is EPI_BF16 with 16
// epilogue warps (four per TMEM lane quadrant, one 64-column unit each, 104 registers): the 8-warp
// epilogue is latency-bound at two warps per scheduler (ncu: 6.9 cycles per issued instruction).
// EPI_F32T is EPI_F32 with the residual tile arriving by TMA (prefetched one chunk ahead into a
// 128B-swizzled staging tile), the sum written back row-per-lane into the same tile and leaving through a
// TMA store: no LSU residual loads (32 x LDG.128 per thread and tile), no transposing pass, no STG.
enum { EPI_BF16 = 0, EPI_F32 = 1, EPI_ANY = 2, EPI_BF16W = 3, EPI_F32T = 4 };
__host__ __device__ constexpr int epi_warps(int epi) { return epi == EPI_BF16W ? 16 : 8; }
__host__ __device__ constexpr int num_threads(int epi) { return (FIRST_EPI_WARP + epi_warps(epi)) * 32; }
// shared memory of the epilogue staging tiles: 8 x 4 KB / 16 x 2 KB, or for EPI_F32T per warp two 4 KB fp32
// tiles (residual in / sum out, double-buffered) + one 2 KB bf16 tile (secondary output)
__host__ __device__ constexpr int tr_bytes(int epi) { return epi == EPI_F32T ? 8 * (8192 + 2048) : TR_BYTES; }

struct EpiTma {
  const CUtensorMap* tmC;   // bf16 output (EPI_BF16W; secondary output of EPI_F32T)
  const CUtensorMap* tmR;   // fp32 residual (EPI_F32T)
  const CUtensorMap* tmF;   // fp32 output (EPI_F32T)
  uint64_t* rbar;           // EPI_F32T: the two residual-arrival barriers of this warp
  uint32_t* rphase;         // EPI_F32T: their parity bits (bit 0 / 1), kept across tiles
};

template <int BN, int STAGES, int EPI>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFFSET = TR_OFFSET + tr_bytes(EPI);
  // barriers: full/empty ring, 2 + 2 accumulator barriers, TMEM pointer, 16 residual barriers; +1024 align slack
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 16 * 8 + 1024;
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
// gelu_tanh_fit on two elements: the same operations in the same order (bit-identical results)
__device__ __forceinline__ float2 gelu_tanh_fit2(float2 x) {
  float2 s = mul2(x, x);
  s.x = fminf(s.x, 25.0f);
  s.y = fminf(s.y, 25.0f);
  float2 q = fma2(make_float2(-3.81889112e-04f, -3.81889112e-04f), s, make_float2(3.72153111e-02f, 3.72153111e-02f));
  q = fma2(q, s, make_float2(7.97237410e-01f, 7.97237410e-01f));
  const float2 xq = mul2(x, q);
  float2 th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th.x) : "f"(xq.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(th.y) : "f"(xq.y));
  const float2 hx = mul2(x, make_float2(0.5f, 0.5f));
  return fma2(hx, th, hx);
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

// The epilogue of ONE warp for ONE 128 x BN accumulator tile: TMEM lane quadrant `quad` (rows quad*32 ..
// +31 of the tile), column chunks of parity `cpar`.  t_addr = TMEM address of (lane quad*32, first column
// of the accumulator buffer); acc_bar / acc_phase: the "accumulator complete" barrier to wait on.
// ACT >= 0: the activation is a compile-time constant (one instantiation per activation keeps the hot
// loop small: with every activation inlined a 32-column chunk was 826 SASS instructions of which ~100
// execute, and four unrolled chunks overflowed the instruction cache); ACT < 0: epi.act at run time.
template <int BN, int EPI, int ACT>
__device__ __forceinline__ void epilogue_tile(const GemmEpi& epi, const EpiTma& tm, int N, int nt, int c_row0,
                                              int w_row0, int rows_valid, uint32_t t_addr, uint64_t* acc_bar,
                                              uint32_t acc_phase, uint4* tr, int quad, int cpar, int lane) {
  const CUtensorMap* tmC = tm.tmC;
#ifdef MDM_GEMM_PROFILE
  const bool prof_on = (threadIdx.x == FIRST_EPI_WARP * 32);
  long long tprev_ = clock64();
#endif
  float4* trf = reinterpret_cast<float4*>(tr);
  const bool f32_path = EPI == EPI_F32 || EPI == EPI_F32T || (EPI == EPI_ANY && (epi.out_f32 != nullptr || epi.resid != nullptr));
  const int rsub = lane >> 3, ch = lane & 7;
  constexpr int NCH = EPI == EPI_BF16W ? 2 : BN / 64;   // 32-column chunks per warp and tile
  const int r = quad * 32 + lane;
  const bool row_ok = r < rows_valid;
  const long m = (long)c_row0 + r;
  const float rs = (epi.rowscale && row_ok) ? epi.rowscale[m] : 1.0f;
  const float rm = (epi.rowmask && row_ok) ? epi.rowmask[m] : 1.0f;
  const float scale = rs * rm * epi.alpha;
  const bool has_scale = epi.rowscale || epi.rowmask || epi.alpha != 1.0f;   // warp-uniform
  const int rmax = min(32, rows_valid - quad * 32);   // valid rows of this warp's 32-row block
    // per-tile base pointers; in-tile offsets stay 32-bit
  const long blk_row0 = (long)c_row0 + quad * 32;
  float* of_blk = epi.out_f32 ? epi.out_f32 + blk_row0 * epi.ld_f32 : nullptr;
  bf16* ob_blk = epi.out_bf16 ? reinterpret_cast<bf16*>(epi.out_bf16) + blk_row0 * epi.ld_bf16 : nullptr;
  const float* rs_blk = (epi.resid && epi.resid_mod <= 0) ? epi.resid + blk_row0 * epi.ld_resid : epi.resid;
  const int rmod_base = epi.resid_mod > 0 ? (int)(blk_row0 % epi.resid_mod) : 0;

  // Bias of the NCH chunks of this warp, one column per lane, requested before the accumulator wait and
  // broadcast with shuffles when a chunk is finished.  (The chunks' bias used to be re-read with 8
  // LDG.128 right after every TMEM load; with 227 KB of the SM configured as shared memory there is
  // practically no L1, so those were L2 round trips on the critical path: ~900 cycles per chunk.)
  // chunk k of this warp -> TMEM chunk index: fp32 path cpar + 2k ; bf16 path 2 (cpar + 2 (k / 2)) + k % 2
  auto chunk_of = [&](int k) {
    return EPI == EPI_BF16W ? 2 * cpar + k : (f32_path ? cpar + 2 * k : 2 * (cpar + 2 * (k >> 1)) + (k & 1));
  };
  float bias_r[NCH];
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int n = nt * BN + chunk_of(k) * 32 + lane;
    bias_r[k] = (epi.bias && n < N) ? __ldg(epi.bias + w_row0 + n) : 0.f;
  }

  // bias (lane-distributed register bk), activation and row scale on the 32 accumulator columns in raw
  auto finish_chunk = [&](const uint32_t (&raw)[32], float bk, float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
    if (epi.bias) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float2 r = add2(make_float2(v[j], v[j + 1]),
                              make_float2(__shfl_sync(0xffffffffu, bk, j), __shfl_sync(0xffffffffu, bk, j + 1)));
        v[j] = r.x; v[j + 1] = r.y;
      }
    }
    const int act = ACT >= 0 ? ACT : epi.act;
    if (act == MDM_ACT_GELU) {
      if (f32_path && EPI != EPI_F32T) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float2 r = gelu_tanh_fit2(make_float2(v[j], v[j + 1]));
          v[j] = r.x; v[j + 1] = r.y;
        }
      }
    } else if (act == MDM_ACT_SILU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = silu_fast(v[j]);
    } else if (act != MDM_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = act_fast(v[j], act);
    }
    if (has_scale) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float2 r = mul2(make_float2(v[j], v[j + 1]), make_float2(scale, scale));
        v[j] = r.x; v[j + 1] = r.y;
      }
    }
  };
  // TMEM chunk c -> registers, then finish_chunk (one chunk at a time)
  auto load_chunk = [&](int c, float bk, float (&v)[32]) {
    uint32_t raw[32];
    EPI_MARK(7);
    tmem_ld32(t_addr + c * 32, raw);
    tmem_ld_wait();
    EPI_MARK(0);
    finish_chunk(raw, bk, v);
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
    EPI_MARK(4);
    mbar_wait(acc_bar, acc_phase);
    tc_fence_after();
    EPI_MARK(5);
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int c = cpar + 2 * k;
      const int n0 = nt * BN + c * 32;
      if (n0 < N) {
        const int n = n0 + ch * 4;                        // this lane's 4 columns in the read phase
        float v[32];
        load_chunk(c, bias_r[k], v);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8)
          trf[lane * 8 + (c8 ^ (lane & 7))] = make_float4(v[4 * c8], v[4 * c8 + 1], v[4 * c8 + 2], v[4 * c8 + 3]);
        EPI_MARK(1);
        __syncwarp();
        float4 x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = i * 4 + rsub;
          x[i] = trf[row * 8 + (ch ^ (row & 7))];
        }
        __syncwarp();
        EPI_MARK(2);
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
        EPI_MARK(3);
      }
    }
  } else if constexpr (EPI == EPI_F32T) {
    // ---------------- fp32 output + fp32 residual (+ optional bf16 copy), all global traffic by TMA.
    // tr: this warp's two 4 KB fp32 tiles (32 rows x 128 B, SWIZZLE_128B); the 2 KB bf16 tiles of all warps
    // follow the 8 x 8 KB fp32 area.  Chunk k uses tile k & 1: residual in, (acc + beta * residual) out.
    // The residuals of chunks 0 and 1 are requested before the accumulator wait, chunk k + 2 as soon as
    // the store of chunk k has finished reading its tile.
    float4* ftile = reinterpret_cast<float4*>(tr);
    const int widx = quad + 4 * cpar;
    uint4* btile = tr + (8 * 8192 - widx * 8192) / 16 + widx * (2048 / 16);
    const int row0 = c_row0 + quad * 32;
    (void)ob_blk; (void)of_blk; (void)rs_blk; (void)rmod_base; (void)rmax; (void)trf; (void)rsub; (void)ch;
    const bool want_b = epi.out_bf16 != nullptr;
    auto col_of = [&](int k) { return nt * BN + (cpar + 2 * k) * 32; };
    auto issue_res = [&](int k) {       // lane 0: residual chunk k -> tile k & 1
      mbar_expect_tx(&tm.rbar[k & 1], 4096);
      tma_load_2d(tm.tmR, &tm.rbar[k & 1], ftile + (k & 1) * 256, col_of(k), row0);
    };
    if (lane == 0 && col_of(0) < N) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // stores of the previous tile have left
      issue_res(0);
      if (NCH > 1 && col_of(1) < N) issue_res(1);
    }
    EPI_MARK(4);
    mbar_wait(acc_bar, acc_phase);
    tc_fence_after();
    EPI_MARK(5);
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int c = cpar + 2 * k;
      const int n0 = col_of(k);
      if (n0 < N) {
        float v[32];
        load_chunk(c, bias_r[k], v);
        mbar_wait(&tm.rbar[k & 1], (*tm.rphase >> (k & 1)) & 1u);          // residual chunk k has landed
        *tm.rphase ^= 1u << (k & 1);
        EPI_MARK(2);
        float4* t4 = ftile + (k & 1) * 256 + lane * 8;
        if (want_b) {     // the previous bf16 store of this warp must have finished reading its tile
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int slot = j ^ (lane & 7);
          const float4 r4 = t4[slot];
          const float4 x = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          float4 y;
          y.x = fmaf(epi.beta, r4.x, x.x); y.y = fmaf(epi.beta, r4.y, x.y);
          y.z = fmaf(epi.beta, r4.z, x.z); y.w = fmaf(epi.beta, r4.w, x.w);
          t4[slot] = y;
          if (want_b) {     // bf16 copy with or without the residual term, SWIZZLE_64B tile (8-byte pieces)
            const float4 s4 = epi.bf16_pre_resid ? x : y;
            uint2 pk; pk.x = pack2(s4.x, s4.y); pk.y = pack2(s4.z, s4.w);
            const int c16 = (j >> 1) ^ ((lane >> 1) & 3);
            reinterpret_cast<uint2*>(btile + lane * 4 + c16)[j & 1] = pk;
          }
        }
        EPI_MARK(1);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                       ::"l"(reinterpret_cast<uint64_t>(tm.tmF)), "r"(n0), "r"(row0), "r"(smem_u32(ftile + (k & 1) * 256))
                       : "memory");
          if (want_b)
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                         ::"l"(reinterpret_cast<uint64_t>(tmC)), "r"(n0), "r"(row0), "r"(smem_u32(btile))
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          if (k + 2 < NCH && col_of(k + 2) < N) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // tile k & 1 is free again
            issue_res(k + 2);
          }
        }
        EPI_MARK(3);
      }
    }
  } else if constexpr (EPI == EPI_BF16W) {
    // ---------------- bf16-only output, 16 warps, TMA store: this warp owns the 64-column unit `cpar`
    // (0..3) of its lane quadrant.  Each 32-column half is written row-per-lane into a 2 KB staging tile
    // in the SWIZZLE_64B layout (16-byte slot ^ ((row >> 1) & 3): conflict-free) and leaves through one
    // cp.async.bulk.tensor store: no transposing read phase, and no LSU global stores, which measurably
    // slowed down the shared-memory traffic of the whole SM (tools/gemm_prof.py: qkv tile period 6 980
    // -> 5 044 cycles with the STG.128s removed).  Rows / columns outside the output are clipped by the
    // tensor map; rows >= rows_valid of a grouped tile fall into the segment's own padding.
    EPI_MARK(4);
    mbar_wait(acc_bar, acc_phase);
    tc_fence_after();
    EPI_MARK(5);
    const int n0 = nt * BN + cpar * 64;
    if (n0 < N) {
      uint32_t raw[2][32];
      EPI_MARK(7);
      tmem_ld32(t_addr + (cpar * 2) * 32, raw[0]);
      tmem_ld32(t_addr + (cpar * 2 + 1) * 32, raw[1]);
      tmem_ld_wait();
      EPI_MARK(0);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int nh = n0 + hh * 32;
        if (nh < N) {
          float v[32];
          finish_chunk(raw[hh], bias_r[hh], v);
          // the previous store of this warp must have finished READING the staging tile
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            uint4 pk;
            pk.x = pack2(v[8 * c4], v[8 * c4 + 1]); pk.y = pack2(v[8 * c4 + 2], v[8 * c4 + 3]);
            pk.z = pack2(v[8 * c4 + 4], v[8 * c4 + 5]); pk.w = pack2(v[8 * c4 + 6], v[8 * c4 + 7]);
            tr[lane * 4 + (c4 ^ ((lane >> 1) & 3))] = pk;
          }
          EPI_MARK(1);
          fence_proxy_async();          // generic-proxy writes -> visible to the TMA (async proxy)
          __syncwarp();
          if (lane == 0) {
            asm volatile(
                "cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                ::"l"(reinterpret_cast<uint64_t>(tmC)), "r"(nh), "r"(c_row0 + quad * 32), "r"(smem_u32(tr))
                : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          EPI_MARK(3);
        }
      }
    }
  } else if constexpr (EPI == EPI_BF16) {
    // ---------------- bf16-only output: 64-column units (two TMEM chunks) staged as bf16
    EPI_MARK(4);
    mbar_wait(acc_bar, acc_phase);
    tc_fence_after();
    EPI_MARK(5);
#pragma unroll
    for (int uu = 0; uu < NCH / 2; ++uu) {
      const int u = cpar + 2 * uu;
      const int n0 = nt * BN + u * 64;
      if (n0 >= N) break;
      // both 32-column halves of the unit are requested from TMEM before either is processed
      uint32_t raw[2][32];
      EPI_MARK(7);
      tmem_ld32(t_addr + (u * 2) * 32, raw[0]);
      tmem_ld32(t_addr + (u * 2 + 1) * 32, raw[1]);
      tmem_ld_wait();
      EPI_MARK(0);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        if (n0 + hh * 32 < N) {
          float v[32];
          finish_chunk(raw[hh], bias_r[2 * uu + hh], v);
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            uint4 pk;
            pk.x = pack2(v[8 * c4], v[8 * c4 + 1]); pk.y = pack2(v[8 * c4 + 2], v[8 * c4 + 3]);
            pk.z = pack2(v[8 * c4 + 4], v[8 * c4 + 5]); pk.w = pack2(v[8 * c4 + 6], v[8 * c4 + 7]);
            tr[lane * 8 + ((hh * 4 + c4) ^ (lane & 7))] = pk;
          }
        }
      }
      EPI_MARK(1);
      __syncwarp();
      const int n = n0 + ch * 8;   // this lane's 8 columns in the read phase
      uint4 w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + rsub;
        w[i] = tr[row * 8 + (ch ^ (row & 7))];
      }
      __syncwarp();
      EPI_MARK(2);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + rsub;
        if (row < rmax && n < N) *reinterpret_cast<uint4*>(ob_blk + row * epi.ld_bf16 + n) = w[i];
      }
      EPI_MARK(3);
    }
  } else {
    // ---------------- any shape / alignment (scalar tails), rolled loops
    EPI_MARK(4);
    mbar_wait(acc_bar, acc_phase);
    tc_fence_after();
    EPI_MARK(5);
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
        load_chunk(c, (epi.bias && n0 + lane < N) ? __ldg(epi.bias + w_row0 + n0 + lane) : 0.f, v);
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
            const int nb = n0 + hh * 32 + lane;
            load_chunk(u * 2 + hh, (epi.bias && nb < N) ? __ldg(epi.bias + w_row0 + nb) : 0.f, v);
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
}

}  // namespace
