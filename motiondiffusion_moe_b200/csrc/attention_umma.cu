// FastAttention.forward (reference: models/fast_attention.py:29-92) on the 5th-generation tensor cores:
// one CTA per (sequence, head), hd = M = 128, T <= 256.  The four products of the op run as tcgen05.mma
// with fp32 accumulators in TMEM; every operand is built in shared memory in the canonical K-major,
// 128B-swizzled layout by the CUDA cores (LayerNorm / L2 norm / feature maps are row-wise passes whose
// output IS the next operand), so nothing but q, k, v (in) and the result (out) touches global memory.
//
//   P0  q, k rows: 0.1 x -> LayerNorm(hd) -> L2 norm -> bf16         -> Qs [t][d], Ks [t][d]   (A / B operands)
//       v rows:    0.1 x -> LayerNorm(hd) -> bf16, TRANSPOSED          -> VT [l][t]            (B operand of kv)
//       P^T bf16 (packed once per model)                               -> Pt [m][d]
//   P1  K'^T = Pt . Ks^T   [m][t]  (M 128, N TP, K 128)    Q'' = Qs . Pt^T  [t][m]  (TP/128 tiles of M 128, N 128)
//   P2  K'^T epilogue: exp(clamp(.)) * 0.1, key mask (t >= len -> 0), bf16 -> KT [m][t] over Ks (A operand of kv)
//   P3  kv = KT . VT^T     [m][l]  (K = frames, only the 64-frame chunks below len)
//       Q' epilogue (overlaps the kv MMAs): feature map -> bf16 -> Qs in place (A operand of the apply product);
//       den[t] = max(sum_m q'[t][m] k'[t][m], 1e-6)            (fast_attention.py:83-86)
//   P4  kv epilogue: x 0.1 -> bf16, transposed -> kvT [l][m] over Pt (B operand of the apply product)
//   P5  out = Q' . kvT^T   [t][l]
//   P6  out epilogue: x 0.1 / den -> LayerNorm(hd) (lane == row: the row statistics need no shuffles) -> bf16 -> global
//
// bf16 rounding points are those of the mma.sync kernels in attention_tc.cu (normalised q / k / v, q', k', kv).
//
// hd = M = 64 (HPC = 2): one CTA takes TWO adjacent heads, i.e. the same 128 contiguous columns of q / k / v.  P^T is
// passed as the block-diagonal [128 x 128] matrix diag(P^T, P^T), so K'^T and Q'' come out per head from the same
// 128-wide products; the row statistics (LayerNorm / L2 norm over 64, the denominators) are taken per 64-column half;
// the cross-head blocks of kv are zeroed when kv^T is written, which makes the apply product block-diagonal as well.
// TMEM: K'^T in columns [0, TP), later kv in [0, 128); Q'' / out tiles in [256, 256 + TP).
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "cluster.cuh"
#include "tensormap.cuh"

#ifdef MDM_ATTN_PROFILE
__device__ unsigned long long g_fau_phase[16];   // cycles per phase summed over CTAs (thread 0): tools/fa_prof.py
extern "C" MDM_API int mdm_debug_read_fau_phase(unsigned long long* host, int reset) {
  unsigned long long z[16] = {0};
  if (cudaMemcpyFromSymbol(host, g_fau_phase, sizeof(z)) != cudaSuccess) return 2;
  if (reset && cudaMemcpyToSymbol(g_fau_phase, z, sizeof(z)) != cudaSuccess) return 2;
  return 0;
}
#define FAU_MARK(k) do { if (threadIdx.x == 0) { const long long n_ = clock64(); atomicAdd(&g_fau_phase[k], (unsigned long long)(n_ - fau_t_)); fau_t_ = n_; } } while (0)
#define FAU_INIT() long long fau_t_ = clock64()
#else
#define FAU_MARK(k) do { } while (0)
#define FAU_INIT() do { } while (0)
#endif

namespace {

constexpr int HD = 128;
constexpr int NTHR = 256;

__device__ __forceinline__ uint32_t pack2u(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float expfeat_u(float x) {
  const float c = fminf(fmaxf(x, -15.f), 15.f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(c, 1.4426950408889634f, -3.3219280948873623f)));
  return e;
}
// byte offset of element (row r, column c) in a K-major SWIZZLE_128B tile whose rows are 64 bf16 (128 B)
__device__ __forceinline__ uint32_t sw_off(int r, int c) {
  return (uint32_t)(r * 128 + ((((c >> 3) ^ r) & 7) << 4) + ((c & 7) << 1));
}

// TMA 3D tile load: coordinates (c0 = innermost element, c1 = row inside the sequence, c2 = sequence)
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// bf16 tensor [B][T][cols] (row pitch ld elements) as a 3D map with boxes of {64 columns, box_rows frames, 1 sequence},
// SWIZZLE_128B: a box lands as one K-half of an operand tile ([box_rows] x 128 B, 16-byte chunk ^ (row & 7)), and the
// frames beyond T are zero-filled by the TMA unit (a 2D map over [B * T] rows would read the next sequence instead).
bool make_seq_map(CUtensorMap* map, const void* ptr, int B, int T, long cols, long ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)T * (cuuint64_t)ld * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

__device__ __forceinline__ float2 bf2_to_f2(uint32_t u) {      // packed bf16 pair -> two fp32 (exact)
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
template <int HPC = 1>
__device__ __forceinline__ float group8_sum(float v, unsigned gmask) {
  v += __shfl_xor_sync(gmask, v, 1);
  v += __shfl_xor_sync(gmask, v, 2);
  if (HPC == 1) v += __shfl_xor_sync(gmask, v, 4);      // HPC == 2: lanes 0-3 / 4-7 of the row hold different heads
  return v;
}
// The slice of one row owned by a thread (16 elements as 8 packed pairs; 8 threads per row): 0.1 x -> LayerNorm(128)
// with the affine pairs w2 / b2 -> optional L2 normalisation of the whole row.  FFMA2 / FMUL2 / FADD2 throughout.
// v rows at HPC == 2: the thread's pairs j < 4 belong to the first head, j >= 4 to the second (columns 2 sub + 16 j), and all
// eight lanes of the row contribute to both: two sums per reduction.
__device__ __forceinline__ void norm_slice_v2(float2 (&x)[8], const float2 (&w2)[8], const float2 (&b2)[8], unsigned gmask) {
  const float2 tenth = make_float2(0.1f, 0.1f);
  float2 sa = make_float2(0.f, 0.f), sb = sa;
#pragma unroll
  for (int i = 0; i < 4; ++i) { x[i] = mul2(x[i], tenth); sa = add2(sa, x[i]); x[i + 4] = mul2(x[i + 4], tenth); sb = add2(sb, x[i + 4]); }
  const float ma = group8_sum<1>(sa.x + sa.y, gmask) / 64.f, mb = group8_sum<1>(sb.x + sb.y, gmask) / 64.f;
  const float2 na = make_float2(-ma, -ma), nb2 = make_float2(-mb, -mb);
  float2 qa = make_float2(0.f, 0.f), qb = qa;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    x[i] = add2(x[i], na); qa = fma2(x[i], x[i], qa);
    x[i + 4] = add2(x[i + 4], nb2); qb = fma2(x[i + 4], x[i + 4], qb);
  }
  const float ra = rsqrtf(group8_sum<1>(qa.x + qa.y, gmask) / 64.f + 1e-5f), rb = rsqrtf(group8_sum<1>(qb.x + qb.y, gmask) / 64.f + 1e-5f);
  const float2 ra2 = make_float2(ra, ra), rb2 = make_float2(rb, rb);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    x[i] = fma2(x[i], mul2(w2[i], ra2), b2[i]);
    x[i + 4] = fma2(x[i + 4], mul2(w2[i + 4], rb2), b2[i + 4]);
  }
}

template <bool L2>
__device__ __forceinline__ void norm_slice(float2 (&x)[8], const float2 (&w2)[8], const float2 (&b2)[8], unsigned gmask) {
  const float2 tenth = make_float2(0.1f, 0.1f);
  float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = mul2(x[i], tenth); s2 = add2(s2, x[i]); }
  const float mean = group8_sum(s2.x + s2.y, gmask) / (float)HD;
  const float2 nm = make_float2(-mean, -mean);
  float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = add2(x[i], nm); q2 = fma2(x[i], x[i], q2); }
  const float rstd = rsqrtf(group8_sum(q2.x + q2.y, gmask) / (float)HD + 1e-5f);
  const float2 r2 = make_float2(rstd, rstd);
  float2 n2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = fma2(x[i], mul2(w2[i], r2), b2[i]);
    if (L2) n2 = fma2(x[i], x[i], n2);
  }
  if (L2) {
    const float inv = 1.0f / fmaxf(sqrtf(group8_sum(n2.x + n2.y, gmask)), 1e-12f);
    const float2 i2 = make_float2(inv, inv);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = mul2(x[i], i2);
  }
}

// norm_slice<true> on two rows at once (independent dependency chains interleaved by hand)
template <int HPC>
__device__ __forceinline__ void norm_slice2(float2 (&x)[8], float2 (&y)[8], const float2 (&w2)[8], const float2 (&b2)[8],
                                            unsigned gmask) {
  const float2 tenth = make_float2(0.1f, 0.1f);
  float2 sx = make_float2(0.f, 0.f), sy = sx;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = mul2(x[i], tenth); y[i] = mul2(y[i], tenth);
    sx = add2(sx, x[i]); sy = add2(sy, y[i]);
  }
  float ax = sx.x + sx.y, ay = sy.x + sy.y;
#pragma unroll
  for (int o = 1; o < 8 / HPC; o <<= 1) {
    const float tx = __shfl_xor_sync(gmask, ax, o), ty = __shfl_xor_sync(gmask, ay, o);
    ax += tx; ay += ty;
  }
  const float2 nmx = make_float2(-ax / (float)(HD / HPC), -ax / (float)(HD / HPC)), nmy = make_float2(-ay / (float)(HD / HPC), -ay / (float)(HD / HPC));
  float2 qx = make_float2(0.f, 0.f), qy = qx;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = add2(x[i], nmx); y[i] = add2(y[i], nmy);
    qx = fma2(x[i], x[i], qx); qy = fma2(y[i], y[i], qy);
  }
  ax = qx.x + qx.y; ay = qy.x + qy.y;
#pragma unroll
  for (int o = 1; o < 8 / HPC; o <<= 1) {
    const float tx = __shfl_xor_sync(gmask, ax, o), ty = __shfl_xor_sync(gmask, ay, o);
    ax += tx; ay += ty;
  }
  const float rx = rsqrtf(ax / (float)(HD / HPC) + 1e-5f), ry = rsqrtf(ay / (float)(HD / HPC) + 1e-5f);
  const float2 rx2 = make_float2(rx, rx), ry2 = make_float2(ry, ry);
  float2 nx = make_float2(0.f, 0.f), ny = nx;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = fma2(x[i], mul2(w2[i], rx2), b2[i]); y[i] = fma2(y[i], mul2(w2[i], ry2), b2[i]);
    nx = fma2(x[i], x[i], nx); ny = fma2(y[i], y[i], ny);
  }
  ax = nx.x + nx.y; ay = ny.x + ny.y;
#pragma unroll
  for (int o = 1; o < 8 / HPC; o <<= 1) {
    const float tx = __shfl_xor_sync(gmask, ax, o), ty = __shfl_xor_sync(gmask, ay, o);
    ax += tx; ay += ty;
  }
  const float ix = 1.0f / fmaxf(sqrtf(ax), 1e-12f), iy = 1.0f / fmaxf(sqrtf(ay), 1e-12f);
  const float2 ix2 = make_float2(ix, ix), iy2 = make_float2(iy, iy);
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = mul2(x[i], ix2); y[i] = mul2(y[i], iy2); }
}

template <int TP>   // padded frames: 128 or 256
struct Smem {
  static constexpr int QS = 0;                        // 2 K-halves x [TP rows x 128 B]
  static constexpr int KS = QS + 2 * TP * 128;        // 2 K-halves x [TP rows x 128 B]; later KT: TP/64 chunks x [128 x 128 B]
  static constexpr int PT = KS + 2 * TP * 128;        // 2 K-halves x [128 rows x 128 B]; later kvT
  // V^T: TP/64 chunks x [128 rows x 128 B].  For TP = 128 it shares the 32 KB of Pt (v waits in registers until both
  // P products are done, then kvT takes the place of V^T once the kv product is done): 97 KB -> two CTAs per SM.
  static constexpr int VT = TP == 128 ? PT : PT + 2 * 128 * 128;
  static constexpr int NW = VT + (TP / 64) * 128 * 128;   // float[128] x 2
  static constexpr int BAR = NW + 2 * HD * 4;         // 4 mbarriers + tmem pointer
  static constexpr int TOTAL = BAR + 64 + 1024;       // + alignment slack
};

template <int TP, int HPC>    // HPC heads per CTA: 1 (hd = 128) or 2 (hd = 64, block-diagonal P^T)
__global__ void __launch_bounds__(NTHR, TP == 128 ? 2 : 1)
fastattn_umma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ Ptg, const float* __restrict__ nw,
                     const float* __restrict__ nb, const int64_t* __restrict__ length, int length_shift, int H, int T,
                     bf16* __restrict__ out, const int* __restrict__ seq_order,
                     const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmPt, int use_tma) {
  using L = Smem<TP>;
  constexpr int MT = TP / 128;        // 128-row tiles of the query side
  constexpr int KC = TP / 64;         // 64-frame chunks (K of the kv product)
  constexpr int QCOL = TP;            // TMEM: K'^T / kv in columns [0, TP), Q'' / out tiles from QCOL on
  constexpr int NCOLS = 2 * TP;       // 256 columns per CTA at TP = 128 (two CTAs per SM), 512 at TP = 256
  constexpr int HW = HD / HPC;        // columns of one head: the width of every row statistic
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* Qs = smem + L::QS;
  uint8_t* Ks = smem + L::KS;
  uint8_t* Pt = smem + L::PT;
  uint8_t* VT = smem + L::VT;
  float* nw_s = reinterpret_cast<float*>(smem + L::NW);
  float* nb_s = nw_s + HD;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR);     // kt, q, kv, out
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
  uint64_t* load_bar = bars + 5;                                   // q, k, P^T tiles by TMA

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  FAU_INIT();
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    mbar_init(load_bar, 1);
    fence_mbar_init();
    if (use_tma) { tma_prefetch_desc(&tmQKV); tma_prefetch_desc(&tmPt); }
  }
  if (warp == 0) tmem_alloc(tmem_ptr, NCOLS);
  if (use_tma) __syncthreads();                      // load_bar is polled before the first block-wide barrier below
  pdl_enter();                                       // first global access below
  const int b = seq_order ? seq_order[blockIdx.x / H] : blockIdx.x / H, h = blockIdx.x % H;
  const int D = H * HD;
  const int len = length ? (int)min((long)T, (long)(length[b] >> length_shift)) : T;
  const int kc_used = min(KC, (len + 63) / 64);      // frame chunks that contain unmasked keys

  // ------------------------------------------------------------------ P0: operands
  // Every global load of the CTA is issued before anything waits: v by register prefetch (it is consumed first and
  // leaves transposed), raw q / k by cp.async straight into their final swizzled slots, where the SAME thread
  // normalises them in place afterwards (so a per-thread cp.async.wait_group is all the ordering needed).
  const int sub = tid & 7, rr = tid >> 3;            // 8 lanes per row, 32 rows per pass
  const unsigned gmask = 0xffu << (lane & 24);       // the 8 lanes of this row (rows of a warp may diverge)
  const bf16* base = qkv + (long)b * T * 3 * D + h * HD;
  constexpr int NP = TP / 32;
  uint32_t vraw[NP][8];                              // v: thread owns the column pairs {2 sub + 16 j, + 1}
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int t = rr + 32 * p;
    if (t < len) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(base + (long)t * 3 * D + 2 * D) + sub;
#pragma unroll
      for (int j = 0; j < 8; ++j) vraw[p][j] = __ldg(src + 8 * j);
    }
  }
  // q, k, P^T: six TMA boxes (one K-half of an operand tile each: [TP frames] x 64 columns, SWIZZLE_128B = sw_off; the
  // frames beyond T arrive as zeros; key rows in [len, T) arrive raw and are masked in P2).  As cp.async these were 40
  // 16-byte copies per thread: the load phase was bound by LSU issue (6 k of the CTA's 34 k cycles, tools/fa_prof.py).
  if (use_tma) {
    if (tid == 0) {
      mbar_expect_tx(load_bar, 4 * TP * 128 + 2 * 128 * 128);
#pragma unroll
      for (int kh = 0; kh < 2; ++kh) {
        tma_load_3d(&tmQKV, load_bar, Ks + kh * (TP * 128), D + h * HD + kh * 64, 0, b);
        tma_load_2d(&tmPt, load_bar, Pt + kh * (128 * 128), kh * 64, 0);
        tma_load_3d(&tmQKV, load_bar, Qs + kh * (TP * 128), h * HD + kh * 64, 0, b);
      }
    }
  } else {
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
    uint8_t* dst = (which == 0 ? Qs : Ks) + (sub >> 2) * (TP * 128);
    const int rows_live = which == 0 ? T : len;      // key rows >= len are masked: no need to load them
#pragma unroll 1
    for (int t = rr; t < TP; t += 32) {
      uint8_t* d0 = dst + sw_off(t, (sub & 3) * 16);
      uint8_t* d1 = dst + sw_off(t, (sub & 3) * 16 + 8);
      if (t < rows_live) {
        const bf16* src = base + (long)t * 3 * D + which * D + sub * 16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d0)), "l"(src) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d1)), "l"(src + 8) : "memory");
      } else {
        *reinterpret_cast<uint4*>(d0) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(d1) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
  for (int i = tid; i < HD * 16; i += NTHR) {        // Pt [m][d]: 16-byte chunks, needed by the MMAs only
    const int m = i >> 4, c = i & 15;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                 ::"r"(smem_u32(Pt + (c >> 3) * (128 * 128) + sw_off(m, (c & 7) * 8))), "l"(Ptg + m * HD + c * 8) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int i = tid; i < HD; i += NTHR) { nw_s[i] = nw[i % HW]; nb_s[i] = nb[i % HW]; }   // read again after the barrier below
  FAU_MARK(0);
  // v: LayerNorm, transposed 2-byte stores (the pair ownership costs a 2-way bank conflict, the loads are 4 bytes)
  auto v_pass = [&]() {
    float2 w2[8], b2[8];                               // straight from global: no barrier before this pass
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      w2[j] = __ldg(reinterpret_cast<const float2*>(nw + (2 * sub + 16 * j) % HW));
      b2[j] = __ldg(reinterpret_cast<const float2*>(nb + (2 * sub + 16 * j) % HW));
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int t = rr + 32 * p;
      uint8_t* dcol = VT + (t >> 6) * (128 * 128);
      if (t >= len) {                                  // masked keys: their value rows only need to be finite
#pragma unroll
        for (int j = 0; j < 16; ++j)
          *reinterpret_cast<uint16_t*>(dcol + sw_off(2 * sub + 16 * (j >> 1) + (j & 1), t & 63)) = 0;
        continue;
      }
      float2 x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = bf2_to_f2(vraw[p][j]);
      if constexpr (HPC == 2) norm_slice_v2(x, w2, b2, gmask);
      else norm_slice<false>(x, w2, b2, gmask);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t pk = pack2u(x[j].x, x[j].y);
        *reinterpret_cast<uint16_t*>(dcol + sw_off(2 * sub + 16 * j, t & 63)) = (uint16_t)(pk & 0xffffu);
        *reinterpret_cast<uint16_t*>(dcol + sw_off(2 * sub + 16 * j + 1, t & 63)) = (uint16_t)(pk >> 16);
      }
    }
  };
  if (TP != 128) v_pass();
  FAU_MARK(1);
  // q and k in place: thread owns columns [16 sub, 16 sub + 16) of its rows (the chunks it requested itself);
  // two rows per iteration so that the shuffle / rsqrt latencies of one row hide behind the other's arithmetic
  if (use_tma) mbar_wait(load_bar, 0);
  else asm volatile("cp.async.wait_group 0;" ::: "memory");
  {
    float2 w2[8], b2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      w2[i] = __ldg(reinterpret_cast<const float2*>(nw + (sub * 16 + 2 * i) % HW));
      b2[i] = __ldg(reinterpret_cast<const float2*>(nb + (sub * 16 + 2 * i) % HW));
    }
    auto load8 = [&](uint8_t* dst, int t, float2 (&x)[8]) {
      const uint4 r0 = *reinterpret_cast<const uint4*>(dst + sw_off(t, (sub & 3) * 16));
      const uint4 r1 = *reinterpret_cast<const uint4*>(dst + sw_off(t, (sub & 3) * 16 + 8));
      x[0] = bf2_to_f2(r0.x); x[1] = bf2_to_f2(r0.y); x[2] = bf2_to_f2(r0.z); x[3] = bf2_to_f2(r0.w);
      x[4] = bf2_to_f2(r1.x); x[5] = bf2_to_f2(r1.y); x[6] = bf2_to_f2(r1.z); x[7] = bf2_to_f2(r1.w);
    };
    auto store8 = [&](uint8_t* dst, int t, const float2 (&x)[8]) {
      *reinterpret_cast<uint4*>(dst + sw_off(t, (sub & 3) * 16)) =
          make_uint4(pack2u(x[0].x, x[0].y), pack2u(x[1].x, x[1].y), pack2u(x[2].x, x[2].y), pack2u(x[3].x, x[3].y));
      *reinterpret_cast<uint4*>(dst + sw_off(t, (sub & 3) * 16 + 8)) =
          make_uint4(pack2u(x[4].x, x[4].y), pack2u(x[5].x, x[5].y), pack2u(x[6].x, x[6].y), pack2u(x[7].x, x[7].y));
    };
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      uint8_t* dst = (which == 0 ? Qs : Ks) + (sub >> 2) * (TP * 128);
      const int rows_live = which == 0 ? T : len;
#pragma unroll 1
      for (int t = rr; t < rows_live; t += 64) {
        const bool two = t + 32 < rows_live;           // uniform over the 8 lanes of a row
        float2 xa[8], xb[8];
        load8(dst, t, xa);
        load8(dst, two ? t + 32 : t, xb);
        norm_slice2<HPC>(xa, xb, w2, b2, gmask);
        store8(dst, t, xa);
        if (two) store8(dst, t + 32, xb);
      }
    }
  }
  FAU_MARK(2);
  fence_proxy_async();                 // generic-proxy writes of the operands -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t qs_a = smem_u32(Qs), ks_a = smem_u32(Ks), pt_a = smem_u32(Pt), vt_a = smem_u32(VT);

  // ------------------------------------------------------------------ P1: K'^T and Q'' products
  if (tid == 0) {
    constexpr uint32_t id_kt = make_idesc_bf16(128, TP);
    constexpr uint32_t id_q = make_idesc_bf16(128, 128);
#pragma unroll
    for (int kc = 0; kc < 2; ++kc) {
      const uint64_t ad = make_sw128_kmajor_desc(pt_a + kc * (128 * 128));
      const uint64_t bd = make_sw128_kmajor_desc(ks_a + kc * (TP * 128));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, ad + 2 * k, bd + 2 * k, id_kt, (kc | k) != 0);
    }
    umma_commit(&bars[0]);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
        const uint64_t ad = make_sw128_kmajor_desc(qs_a + kc * (TP * 128) + mt * (128 * 128));
        const uint64_t bd = make_sw128_kmajor_desc(pt_a + kc * (128 * 128));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + QCOL + mt * 128, ad + 2 * k, bd + 2 * k, id_q, (kc | k) != 0);
      }
    }
    umma_commit(&bars[1]);
  }
  const int quad = warp & 3, hi = warp >> 2;          // TMEM lane quadrant, second role index
  const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);

  // ------------------------------------------------------------------ P2: K'^T epilogue -> KT (over Ks)
  {
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const int m = quad * 32 + lane;
    constexpr int CH = TP / 64;                        // 32-column chunks per warp: columns [hi * TP/2, +TP/2)
#pragma unroll 1
    for (int c = 0; c < CH; ++c) {
      const int t0 = hi * (TP / 2) + c * 32;
      uint8_t* dst = Ks + (t0 >> 6) * (128 * 128);
      if (t0 >= len) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(dst + sw_off(m, (t0 & 63) + 8 * j)) = make_uint4(0u, 0u, 0u, 0u);
        continue;
      }
      uint32_t raw[32];
      tmem_ld32(t_lane + t0, raw);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = (t0 + 8 * j + e) < len ? expfeat_u(__uint_as_float(raw[8 * j + e])) : 0.f;
        uint4 pk;
        pk.x = pack2u(f[0], f[1]); pk.y = pack2u(f[2], f[3]); pk.z = pack2u(f[4], f[5]); pk.w = pack2u(f[6], f[7]);
        *reinterpret_cast<uint4*>(dst + sw_off(m, (t0 & 63) + 8 * j)) = pk;
      }
    }
  }
  if (TP == 128) {            // V^T shares Pt's memory: both P products must have retired before it is written
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    v_pass();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  FAU_MARK(3);
  // ------------------------------------------------------------------ P3: kv product; Q' epilogue + denominators
  if (tid == 0) {
    constexpr uint32_t id_kv = make_idesc_bf16(128, 128);
    for (int kc = 0; kc < kc_used; ++kc) {
      const uint64_t ad = make_sw128_kmajor_desc(ks_a + kc * (128 * 128));
      const uint64_t bd = make_sw128_kmajor_desc(vt_a + kc * (128 * 128));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, ad + 2 * k, bd + 2 * k, id_kv, (kc | k) != 0);
    }
    umma_commit(&bars[2]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  float den_r[HPC];                                    // denominator(s) of this thread's frame (same thread in P6)
#pragma unroll
  for (int i = 0; i < HPC; ++i) den_r[i] = 1.f;
  if (hi < MT) {
    const int t = hi * 128 + quad * 32 + lane;
    float den = 0.f, den_b = 0.f;                      // den_b: second head (columns 64..127) at HPC == 2
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t raw[32];
      tmem_ld32(t_lane + QCOL + hi * 128 + c * 32, raw);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) pk[e] = pack2u(expfeat_u(__uint_as_float(raw[2 * e])), expfeat_u(__uint_as_float(raw[2 * e + 1])));
      if (t < len) {                                   // k'[t][:] == 0 for masked frames: den stays 0 -> 1e-6
        const uint8_t* kt = Ks + (t >> 6) * (128 * 128);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const __nv_bfloat162 q2 = *reinterpret_cast<const __nv_bfloat162*>(&pk[e]);
          const int m = c * 32 + 2 * e;
          const float k0 = __bfloat162float(*reinterpret_cast<const bf16*>(kt + sw_off(m, t & 63)));
          const float k1 = __bfloat162float(*reinterpret_cast<const bf16*>(kt + sw_off(m + 1, t & 63)));
          if (HPC == 2 && c >= 2) {
            den_b = fmaf(__low2float(q2), k0, den_b);
            den_b = fmaf(__high2float(q2), k1, den_b);
          } else {
            den = fmaf(__low2float(q2), k0, den);
            den = fmaf(__high2float(q2), k1, den);
          }
        }
      }
      uint8_t* dst = Qs + (c >> 1) * (TP * 128);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(dst + sw_off(t, (c & 1) * 32 + 8 * j)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    }
    den_r[0] = fmaxf(den, 1e-6f);
    if (HPC == 2) den_r[HPC - 1] = fmaxf(den_b, 1e-6f);
  }

  FAU_MARK(4);
  // ------------------------------------------------------------------ P4: kv epilogue -> kvT (over Pt)
  {
    mbar_wait(&bars[2], 0);
    tc_fence_after();
    const int m = quad * 32 + lane;
    uint8_t* dst = Pt + (m >> 6) * (128 * 128);
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int l0 = hi * 64 + c * 32;
      uint32_t raw[32];
      if (kc_used > 0) {
        tmem_ld32(t_lane + l0, raw);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) raw[e] = 0u;     // no unmasked key: no MMA was issued, kv == 0
      }
      if (HPC == 2 && (quad >> 1) != hi) {             // kv[m of one head][l of the other]: not part of the op
#pragma unroll
        for (int e = 0; e < 32; ++e) raw[e] = 0u;
      }
#pragma unroll
      for (int e = 0; e < 32; ++e)
        *reinterpret_cast<bf16*>(dst + sw_off(l0 + e, m & 63)) = __float2bfloat16_rn(__uint_as_float(raw[e]) * 0.1f);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  FAU_MARK(5);
  // ------------------------------------------------------------------ P5: out = Q' . kv
  if (tid == 0) {
    constexpr uint32_t id_o = make_idesc_bf16(128, 128);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
        const uint64_t ad = make_sw128_kmajor_desc(qs_a + kc * (TP * 128) + mt * (128 * 128));
        const uint64_t bd = make_sw128_kmajor_desc(pt_a + kc * (128 * 128));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + QCOL + mt * 128, ad + 2 * k, bd + 2 * k, id_o, (kc | k) != 0);
      }
    }
    umma_commit(&bars[3]);
  }

  // ------------------------------------------------------------------ P6: out epilogue
  // x 0.1 / den -> LayerNorm over the thread's own row -> bf16 row in a staging area (KT and kvT are dead: their
  // regions are contiguous; 272-byte pitch keeps the row-per-lane 16-byte stores conflict-free) -> coalesced stores.
  mbar_wait(&bars[3], 0);
  tc_fence_after();
  constexpr int PITCH = 272;
  static_assert((TP == 128 ? 128 : 256) * PITCH <= L::NW - L::KS - (TP == 128 ? 0 : (TP / 64) * 128 * 128), "staging area");
  uint8_t* stage = Ks;
  if (hi < MT) {
    // three passes over the thread's TMEM row (mean, variance, normalise) instead of 128 live registers; at HPC == 2 once
    // per head (64-column half, its own denominator and statistics)
    const int t = hi * 128 + quad * 32 + lane;
    const uint32_t t_row = t_lane + QCOL + hi * 128;
    constexpr int CPH = 4 / HPC;                       // 32-column chunks per head
#pragma unroll 1
    for (int hf = 0; hf < HPC; ++hf) {
      const float sc = 0.1f / den_r[hf];
      const float2 sc2 = make_float2(sc, sc);
      float2 s2 = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int c = hf * CPH; c < (hf + 1) * CPH; ++c) {
        uint32_t raw[32];
        tmem_ld32(t_row + c * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e)
          s2 = fma2(make_float2(__uint_as_float(raw[2 * e]), __uint_as_float(raw[2 * e + 1])), sc2, s2);
      }
      const float mean = (s2.x + s2.y) / (float)HW;
      const float2 nm = make_float2(-mean, -mean);
      float2 q2 = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int c = hf * CPH; c < (hf + 1) * CPH; ++c) {
        uint32_t raw[32];
        tmem_ld32(t_row + c * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float2 d = fma2(make_float2(__uint_as_float(raw[2 * e]), __uint_as_float(raw[2 * e + 1])), sc2, nm);
          q2 = fma2(d, d, q2);
        }
      }
      const float rstd = rsqrtf((q2.x + q2.y) / (float)HW + 1e-5f);
      const float2 r2 = make_float2(rstd, rstd);
#pragma unroll 1
      for (int c = hf * CPH; c < (hf + 1) * CPH; ++c) {
        uint32_t raw[32];
        tmem_ld32(t_row + c * 32, raw);
        tmem_ld_wait();
        if (t < T) {
          uint4* dst = reinterpret_cast<uint4*>(stage + t * PITCH + c * 64);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int col = c * 32 + 8 * j + 2 * e;
              const float2 d = fma2(make_float2(__uint_as_float(raw[8 * j + 2 * e]), __uint_as_float(raw[8 * j + 2 * e + 1])), sc2, nm);
              const float2 y = fma2(d, mul2(make_float2(nw_s[col], nw_s[col + 1]), r2), make_float2(nb_s[col], nb_s[col + 1]));
              pk[e] = pack2u(y.x, y.y);
            }
            dst[j] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < T * 16; i += NTHR) {
    const int r = i >> 4, c = i & 15;
    *reinterpret_cast<uint4*>(out + ((long)(b * T + r)) * D + h * HD + c * 8) =
        *reinterpret_cast<const uint4*>(stage + r * PITCH + c * 16);
  }
  FAU_MARK(6);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, NCOLS);
  }
}

// H counts the CTAs per sequence: heads of 128, or pairs of heads of 64 (HPC = 2)
template <int TP, int HPC>
int launch(const bf16* qkv, const bf16* Pt, const float* nw, const float* nb, const int64_t* length, int shift, int B,
           int H, int T, bf16* out, const int* seq_order, cudaStream_t st) {
  using L = Smem<TP>;
  static_assert(L::TOTAL <= 227 * 1024, "shared memory budget");
  static unsigned long long attr = 0;   // one bit per device ordinal: the attribute is per (function, device)
  const unsigned long long dev_bit = 1ull << mdm_cur_dev();
  if (!(attr & dev_bit)) {
    if (cudaFuncSetAttribute(fastattn_umma_kernel<TP, HPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) != cudaSuccess)
      return MDM_ERR_CUDA;
    attr |= dev_bit;
  }
  // MDM_FA_TMA=0: q / k / P^T by cp.async instead of TMA (A/B runs; also the path taken if a tensor map cannot be built)
  static const int tma_env = [] { const char* e = getenv("MDM_FA_TMA"); return e ? atoi(e) : 1; }();
  CUtensorMap tq, tp;
  memset(&tq, 0, sizeof(tq)); memset(&tp, 0, sizeof(tp));
  const int use_tma = tma_env && make_seq_map(&tq, qkv, B, T, 3L * H * HD, 3L * H * HD, TP) && make_map(&tp, Pt, HD, HD, HD, 128);
  mdm_launch(fastattn_umma_kernel<TP, HPC>, B * H, NTHR, L::TOTAL, st, qkv, Pt, nw, nb, length, shift, H, T, out, seq_order, tq, tp,
             use_tma);
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// LinearTemporalCrossAttention, motion side (fast_attention.py:248,253) on tcgen05:
//   y[t,h,:] = softmax_hd(q[t,h,:]) @ ctx[b,h]
// One CTA per (sequence, head), 97 KB of shared memory and 256 TMEM columns (two CTAs per SM): raw q rows arrive by
// cp.async in their swizzled A-operand slots and are soft-maxed in place by the thread that requested them; the
// step-invariant state ctx^T [l][d] (bf16, packed once per sampling loop) is the B operand; the result leaves
// through a staging area as coalesced stores.
template <int TP>
struct SmemLC {
  static constexpr int QS = 0;                        // 2 K-halves x [TP rows x 128 B]; later the output staging
  static constexpr int CT = QS + 2 * TP * 128;        // 2 K-halves x [128 rows x 128 B]
  static constexpr int STG = CT + 2 * 128 * 128;      // staging tail (rows beyond what fits over QS): none needed, see PITCH
  static constexpr int BAR = STG;
  static constexpr int STY = BAR + 64;                // STYLE: ln_w | ln_b | 1 + scale | shift of this head (4 x 128 floats),
  static constexpr int RED = STY + 4 * 128 * 4;       //        then float2 [2][TP] partial row statistics (read by the peers)
  static constexpr int TOTAL = RED + 2 * TP * 8 + 1024;
};

// STYLE: the StylizationBlock that consumes this core's output (models/fast_attention.py:248-272 -> stylization.py:27-30:
// LayerNorm over the WHOLE row, FiLM, SiLU) in the epilogue.  A row's D = H * 128 values are spread over the H CTAs of
// one sequence: they form a thread-block cluster, every CTA publishes the partial sum / sum of squares of its 128
// columns per row in its own shared memory, and after one cluster barrier every thread reads the H partials of its row
// through distributed shared memory (the same order in every CTA).  TMEM lane = row, so the statistics themselves are
// per-thread sums.  Replaces the rowop launch (LN, FiLM, SiLU) after the core: the bf16 rounding of y in between is gone.
// HPC = 2 (head size 64): the CTA takes two adjacent heads, ctx^T is the block-diagonal [128 x 128] of the pair, the
// softmax runs over each 64-column half (four lanes of a row instead of eight).
template <int TP, bool STYLE, int HPC>
__global__ void __launch_bounds__(NTHR, 2)
lincross_umma_kernel(const bf16* __restrict__ q, const bf16* __restrict__ ctxT, int H, int T, bf16* __restrict__ y,
                     const float* __restrict__ ln_w, const float* __restrict__ ln_b, const float* __restrict__ film) {
  using L = SmemLC<TP>;
  constexpr int MT = TP / 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* Qs = smem + L::QS;
  uint8_t* Ct = smem + L::CT;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::BAR);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = H * HD;
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(tmem_ptr, TP);
  pdl_enter();
  if (STYLE && tid < HD) {
    float* sty = reinterpret_cast<float*>(smem + L::STY);
    sty[tid] = ln_w[h * HD + tid];
    sty[HD + tid] = ln_b[h * HD + tid];
    sty[2 * HD + tid] = 1.0f + film[(long)b * 2 * D + h * HD + tid];
    sty[3 * HD + tid] = film[(long)b * 2 * D + D + h * HD + tid];
  }
  const int sub = tid & 7, rr = tid >> 3;
  const unsigned gmask = 0xffu << (lane & 24);
  const bf16* base = q + (long)b * T * D + h * HD;
  uint8_t* dstq = Qs + (sub >> 2) * (TP * 128);
#pragma unroll 1
  for (int t = rr; t < TP; t += 32) {
    uint8_t* d0 = dstq + sw_off(t, (sub & 3) * 16);
    uint8_t* d1 = dstq + sw_off(t, (sub & 3) * 16 + 8);
    if (t < T) {
      const bf16* src = base + (long)t * D + sub * 16;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d0)), "l"(src) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d1)), "l"(src + 8) : "memory");
    } else {
      *reinterpret_cast<uint4*>(d0) = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(d1) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  const bf16* cg = ctxT + ((long)(b * H + h)) * HD * HD;
  for (int i = tid; i < HD * 16; i += NTHR) {        // ctx^T [l][d]: 16-byte chunks
    const int l = i >> 4, c = i & 15;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                 ::"r"(smem_u32(Ct + (c >> 3) * (128 * 128) + sw_off(l, (c & 7) * 8))), "l"(cg + l * HD + c * 8) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // softmax over the head dimension, in place; two rows per iteration
  {
    auto load8 = [&](int t, float2 (&x)[8]) {
      const uint4 r0 = *reinterpret_cast<const uint4*>(dstq + sw_off(t, (sub & 3) * 16));
      const uint4 r1 = *reinterpret_cast<const uint4*>(dstq + sw_off(t, (sub & 3) * 16 + 8));
      x[0] = bf2_to_f2(r0.x); x[1] = bf2_to_f2(r0.y); x[2] = bf2_to_f2(r0.z); x[3] = bf2_to_f2(r0.w);
      x[4] = bf2_to_f2(r1.x); x[5] = bf2_to_f2(r1.y); x[6] = bf2_to_f2(r1.z); x[7] = bf2_to_f2(r1.w);
    };
    auto store8 = [&](int t, const float2 (&x)[8]) {
      *reinterpret_cast<uint4*>(dstq + sw_off(t, (sub & 3) * 16)) =
          make_uint4(pack2u(x[0].x, x[0].y), pack2u(x[1].x, x[1].y), pack2u(x[2].x, x[2].y), pack2u(x[3].x, x[3].y));
      *reinterpret_cast<uint4*>(dstq + sw_off(t, (sub & 3) * 16 + 8)) =
          make_uint4(pack2u(x[4].x, x[4].y), pack2u(x[5].x, x[5].y), pack2u(x[6].x, x[6].y), pack2u(x[7].x, x[7].y));
    };
    const float2 l2e = make_float2(1.4426950408889634f, 1.4426950408889634f);
#pragma unroll 1
    for (int t = rr; t < T; t += 64) {
      const bool two = t + 32 < T;
      float2 xa[8], xb[8];
      load8(t, xa);
      load8(two ? t + 32 : t, xb);
      float ma = fmaxf(xa[0].x, xa[0].y), mb = fmaxf(xb[0].x, xb[0].y);
#pragma unroll
      for (int i = 1; i < 8; ++i) { ma = fmaxf(ma, fmaxf(xa[i].x, xa[i].y)); mb = fmaxf(mb, fmaxf(xb[i].x, xb[i].y)); }
#pragma unroll
      for (int o = 1; o < 8 / HPC; o <<= 1) {
        ma = fmaxf(ma, __shfl_xor_sync(gmask, ma, o));
        mb = fmaxf(mb, __shfl_xor_sync(gmask, mb, o));
      }
      const float2 na = make_float2(-ma * 1.4426950408889634f, -ma * 1.4426950408889634f);
      const float2 nb2 = make_float2(-mb * 1.4426950408889634f, -mb * 1.4426950408889634f);
      float sa = 0.f, sb = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 ea = fma2(xa[i], l2e, na), eb = fma2(xb[i], l2e, nb2);
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(xa[i].x) : "f"(ea.x));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(xa[i].y) : "f"(ea.y));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(xb[i].x) : "f"(eb.x));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(xb[i].y) : "f"(eb.y));
        sa += xa[i].x + xa[i].y;
        sb += xb[i].x + xb[i].y;
      }
#pragma unroll
      for (int o = 1; o < 8 / HPC; o <<= 1) {
        sa += __shfl_xor_sync(gmask, sa, o);
        sb += __shfl_xor_sync(gmask, sb, o);
      }
      const float ia = 1.0f / sa, ib = 1.0f / sb;
      const float2 ia2 = make_float2(ia, ia), ib2 = make_float2(ib, ib);
#pragma unroll
      for (int i = 0; i < 8; ++i) { xa[i] = mul2(xa[i], ia2); xb[i] = mul2(xb[i], ib2); }
      store8(t, xa);
      if (two) store8(t + 32, xb);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (tid == 0) {
    constexpr uint32_t id = make_idesc_bf16(128, 128);
    const uint32_t qs_a = smem_u32(Qs), ct_a = smem_u32(Ct);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
        const uint64_t ad = make_sw128_kmajor_desc(qs_a + kc * (TP * 128) + mt * (128 * 128));
        const uint64_t bd = make_sw128_kmajor_desc(ct_a + kc * (128 * 128));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + mt * 128, ad + 2 * k, bd + 2 * k, id, (kc | k) != 0);
      }
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  // epilogue: each warp drains 32 rows x 64 columns; staging over Qs (the product has retired), 272-byte pitch
  constexpr int PITCH = 272;
  static_assert(TP * PITCH <= L::STG - L::QS, "staging area");
  {
    const int quad = warp & 3, hi = warp >> 2;
    constexpr int CH = MT == 2 ? 4 : 2;                // 32-column chunks per warp
    const int mt = MT == 2 ? hi : 0, c0 = MT == 2 ? 0 : hi * 2;
    const int t = mt * 128 + quad * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + mt * 128;
    float mean = 0.f, rstd = 1.f;
    const float* sty = reinterpret_cast<const float*>(smem + L::STY);
    if constexpr (STYLE) {
      // pass 1: partial statistics of this thread's columns -> own shared memory -> cluster barrier -> all H partials
      float2* red = reinterpret_cast<float2*>(smem + L::RED);
      float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int c = c0; c < c0 + CH; ++c) {
        uint32_t raw[32];
        tmem_ld32(t_row + c * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float2 v = make_float2(__uint_as_float(raw[j]), __uint_as_float(raw[j + 1]));
          s2 = add2(s2, v);
          q2 = fma2(v, v, q2);
        }
      }
      red[(MT == 2 ? 0 : hi) * TP + t] = make_float2(s2.x + s2.y, q2.x + q2.y);
      cluster_sync_all();
      float S = 0.f, Q = 0.f;
      for (int rk = 0; rk < H; ++rk) {
        float2 o;
        asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(o.x), "=f"(o.y) : "r"(mapa_u32(smem_u32(&red[t]), (uint32_t)rk)));
        S += o.x; Q += o.y;
        if (MT == 1) {
          asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(o.x), "=f"(o.y) : "r"(mapa_u32(smem_u32(&red[TP + t]), (uint32_t)rk)));
          S += o.x; Q += o.y;
        }
      }
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");   // done with the peers' memory (waited for at exit)
      const float inv_d = 1.0f / (float)D;
      mean = S * inv_d;
      rstd = rsqrtf(fmaxf(fmaf(Q, inv_d, -mean * mean), 0.f) + 1e-5f);
    }
#pragma unroll 1
    for (int c = c0; c < c0 + CH; ++c) {
      uint32_t raw[32];
      tmem_ld32(t_row + c * 32, raw);
      tmem_ld_wait();
      if constexpr (STYLE) {
        const float2 r2 = make_float2(rstd, rstd), m2 = make_float2(-mean * rstd, -mean * rstd), half = make_float2(0.5f, 0.5f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 w4 = *reinterpret_cast<const float4*>(sty + c * 32 + 4 * j);
          const float4 b4 = *reinterpret_cast<const float4*>(sty + HD + c * 32 + 4 * j);
          const float4 g4 = *reinterpret_cast<const float4*>(sty + 2 * HD + c * 32 + 4 * j);
          const float4 h4 = *reinterpret_cast<const float4*>(sty + 3 * HD + c * 32 + 4 * j);
          float2 z0 = fma2(fma2(make_float2(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1])), r2, m2),
                           make_float2(w4.x, w4.y), make_float2(b4.x, b4.y));
          float2 z1 = fma2(fma2(make_float2(__uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3])), r2, m2),
                           make_float2(w4.z, w4.w), make_float2(b4.z, b4.w));
          z0 = fma2(z0, make_float2(g4.x, g4.y), make_float2(h4.x, h4.y));
          z1 = fma2(z1, make_float2(g4.z, g4.w), make_float2(h4.z, h4.w));
          const float2 a0 = mul2(z0, half), a1 = mul2(z1, half);       // SiLU(z) = z/2 + z/2 * tanh(z/2)
          float2 t0, t1;
          asm("tanh.approx.f32 %0, %1;" : "=f"(t0.x) : "f"(a0.x));
          asm("tanh.approx.f32 %0, %1;" : "=f"(t0.y) : "f"(a0.y));
          asm("tanh.approx.f32 %0, %1;" : "=f"(t1.x) : "f"(a1.x));
          asm("tanh.approx.f32 %0, %1;" : "=f"(t1.y) : "f"(a1.y));
          z0 = fma2(a0, t0, a0);
          z1 = fma2(a1, t1, a1);
          raw[4 * j] = __float_as_uint(z0.x); raw[4 * j + 1] = __float_as_uint(z0.y);
          raw[4 * j + 2] = __float_as_uint(z1.x); raw[4 * j + 3] = __float_as_uint(z1.y);
        }
      }
      if (t < T) {
        uint4* dst = reinterpret_cast<uint4*>(Qs + t * PITCH + c * 64);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j] = make_uint4(pack2u(__uint_as_float(raw[8 * j]), __uint_as_float(raw[8 * j + 1])),
                              pack2u(__uint_as_float(raw[8 * j + 2]), __uint_as_float(raw[8 * j + 3])),
                              pack2u(__uint_as_float(raw[8 * j + 4]), __uint_as_float(raw[8 * j + 5])),
                              pack2u(__uint_as_float(raw[8 * j + 6]), __uint_as_float(raw[8 * j + 7])));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  for (int i = tid; i < T * 16; i += NTHR) {
    const int r = i >> 4, c = i & 15;
    *reinterpret_cast<uint4*>(y + ((long)(b * T + r)) * D + h * HD + c * 8) =
        *reinterpret_cast<const uint4*>(Qs + r * PITCH + c * 16);
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TP);
  }
  if (STYLE) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // no peer still reads this CTA's statistics
}

template <int TP, bool STYLE, int HPC = 1>    // H counts the CTAs per sequence (heads of 128 or pairs of heads of 64)
int launch_lc(const bf16* q, const bf16* ctxT, int B, int H, int T, bf16* y, const float* ln_w, const float* ln_b,
              const float* film, cudaStream_t st) {
  using L = SmemLC<TP>;
  static unsigned long long attr = 0;   // one bit per device ordinal: the attribute is per (function, device)
  const unsigned long long dev_bit = 1ull << mdm_cur_dev();
  if (!(attr & dev_bit)) {
    if (cudaFuncSetAttribute(lincross_umma_kernel<TP, STYLE, HPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) != cudaSuccess)
      return MDM_ERR_CUDA;
    attr |= dev_bit;
  }
  if (STYLE)   // the H head-CTAs of a sequence form a cluster
    return mdm_launch_cluster(lincross_umma_kernel<TP, STYLE, HPC>, B * H, NTHR, L::TOTAL, st, H, q, ctxT, H, T, y, ln_w, ln_b, film) ==
                   cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
  mdm_launch(lincross_umma_kernel<TP, STYLE, HPC>, B * H, NTHR, L::TOTAL, st, q, ctxT, H, T, y, ln_w, ln_b, film);
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// MemoryEfficientCrossAttentionBlock core (fast_attention.py:313-325) on tcgen05:
//   o[t,h,:] = softmax_n((q[t,h,:] * hd^-0.5) . k[b,n,h,:]) @ v[b,n,h,:],  n < nt[b] <= 96
// One CTA per (sequence, head), 97 KB of shared memory, TP TMEM columns (two CTAs per SM).  S = Q K^T with raw q / k
// rows brought by cp.async straight into operand layout (K = head dim); the softmax over the keys is lane-local in
// TMEM (lane == frame) and its bf16 result P overwrites Q as the A operand of O = P V (K = keys, only the NK/16
// k-steps that hold keys); v waits in registers and is stored transposed over K once S has retired.
// HDIM = head size: 128, or 64 (one K-half of the Q / K tiles, HDIM / 16 lanes per row in the copies, a 64-wide O product)
template <int TP, int HDIM>
__global__ void __launch_bounds__(NTHR, 2)
softmax_cross_umma_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                          const int* __restrict__ nt, int H, int T, int Nt_max, int NK, float scale,
                          bf16* __restrict__ o) {
  using L = SmemLC<TP>;                  // Qs (later P, later staging) + a 32 KB operand region (K, later V^T)
  constexpr int MT = TP / 128;
  constexpr int KH = HDIM / 64;          // 64-column K-halves of the Q / K operand tiles
  constexpr int LPR = HDIM / 16;         // lanes per row in the q / k / v copies (16 columns each)
  constexpr int RPP = NTHR / LPR;        // rows per pass
  constexpr int NPV = (96 + RPP - 1) / RPP;   // passes over the (at most 96) value rows
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* Qs = smem + L::QS;
  uint8_t* Kv = smem + L::CT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = H * HDIM;
  if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(tmem_ptr, TP);
  pdl_enter();
  const int n_tok = nt ? min(nt[b], Nt_max) : Nt_max;
  const int sub = tid & (LPR - 1), rr = tid / LPR;
  // v rows -> registers (pairs {2 sub + 2 LPR j, +1} of row n = rr + RPP p)
  uint32_t vraw[NPV][8];
  const bf16* vb = v + (long)b * Nt_max * D + h * HDIM;
#pragma unroll
  for (int p = 0; p < NPV; ++p) {
    const int n = rr + RPP * p;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      vraw[p][j] = n < n_tok ? __ldg(reinterpret_cast<const uint32_t*>(vb + (long)n * D) + sub + LPR * j) : 0u;
  }
  // raw q -> A tiles, raw k -> B tiles (rows = keys; keys >= n_tok are zero rows, masked in the softmax)
  {
    const bf16* qb = q + (long)b * T * D + h * HDIM;
    uint8_t* dstq = Qs + (sub >> 2) * (TP * 128);
#pragma unroll 1
    for (int t = rr; t < TP; t += RPP) {
      uint8_t* d0 = dstq + sw_off(t, (sub & 3) * 16);
      uint8_t* d1 = dstq + sw_off(t, (sub & 3) * 16 + 8);
      if (t < T) {
        const bf16* src = qb + (long)t * D + sub * 16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d0)), "l"(src) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d1)), "l"(src + 8) : "memory");
      } else {
        *reinterpret_cast<uint4*>(d0) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(d1) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    const bf16* kb = k + (long)b * Nt_max * D + h * HDIM;
    uint8_t* dstk = Kv + (sub >> 2) * (128 * 128);
#pragma unroll 1
    for (int n = rr; n < NK; n += RPP) {
      uint8_t* d0 = dstk + sw_off(n, (sub & 3) * 16);
      uint8_t* d1 = dstk + sw_off(n, (sub & 3) * 16 + 8);
      if (n < n_tok) {
        const bf16* src = kb + (long)n * D + sub * 16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d0)), "l"(src) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d1)), "l"(src + 8) : "memory");
      } else {
        *reinterpret_cast<uint4*>(d0) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(d1) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t qs_a = smem_u32(Qs), kv_a = smem_u32(Kv);
  if (tid == 0) {                        // S[mt] = Q[mt] . K^T   (N = NK keys)
    const uint32_t id = make_idesc_bf16(128, NK);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
      for (int kc = 0; kc < KH; ++kc) {
        const uint64_t ad = make_sw128_kmajor_desc(qs_a + kc * (TP * 128) + mt * (128 * 128));
        const uint64_t bd = make_sw128_kmajor_desc(kv_a + kc * (128 * 128));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem_base + mt * 128, ad + 2 * kk, bd + 2 * kk, id, (kc | kk) != 0);
      }
    }
    umma_commit(&bars[0]);
  }
  mbar_wait(&bars[0], 0);
  tc_fence_after();
  // v^T over the K region (S has retired): V^T[l][n], zero columns for the masked keys
#pragma unroll
  for (int p = 0; p < NPV; ++p) {
    const int n = rr + RPP * p;
    if (n < NK) {
      uint8_t* dcol = Kv + (n >> 6) * (128 * 128);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        *reinterpret_cast<uint16_t*>(dcol + sw_off(2 * sub + 2 * LPR * j, n & 63)) = (uint16_t)(vraw[p][j] & 0xffffu);
        *reinterpret_cast<uint16_t*>(dcol + sw_off(2 * sub + 2 * LPR * j + 1, n & 63)) = (uint16_t)(vraw[p][j] >> 16);
      }
    }
  }
  // masked softmax over the keys of the thread's own frame -> P (bf16) over Q
  const int quad = warp & 3, hi = warp >> 2;
  const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
  if (hi < MT) {
    const int t = hi * 128 + quad * 32 + lane;
    const uint32_t t_row = t_lane + hi * 128;
    const float sl2 = scale * 1.4426950408889634f;
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < NK / 32; ++c) {
      uint32_t raw[32];
      tmem_ld32(t_row + c * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e)
        if (c * 32 + e < n_tok) mx = fmaxf(mx, __uint_as_float(raw[e]));
    }
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < NK / 32; ++c) {
      uint32_t raw[32];
      tmem_ld32(t_row + c * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        float ev;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ev) : "f"((__uint_as_float(raw[e]) - mx) * sl2));
        sum += (c * 32 + e < n_tok) ? ev : 0.f;
      }
    }
    const float inv = 1.0f / sum;
#pragma unroll 1
    for (int c = 0; c < NK / 32; ++c) {
      uint32_t raw[32];
      tmem_ld32(t_row + c * 32, raw);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        float e0, e1;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"((__uint_as_float(raw[2 * e]) - mx) * sl2));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"((__uint_as_float(raw[2 * e + 1]) - mx) * sl2));
        e0 = (c * 32 + 2 * e < n_tok) ? e0 * inv : 0.f;
        e1 = (c * 32 + 2 * e + 1 < n_tok) ? e1 * inv : 0.f;
        pk[e] = pack2u(e0, e1);
      }
      uint8_t* dst = Qs + (c >> 1) * (TP * 128);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(dst + sw_off(t, (c & 1) * 32 + 8 * j)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {                        // O[mt] = P[mt] . V   (K = NK keys)
    constexpr uint32_t id = make_idesc_bf16(128, HDIM);
    const int ksteps = NK / 16;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      for (int ks = 0; ks < ksteps; ++ks) {
        const int kc = ks >> 2, kk = ks & 3;
        const uint64_t ad = make_sw128_kmajor_desc(qs_a + kc * (TP * 128) + mt * (128 * 128));
        const uint64_t bd = make_sw128_kmajor_desc(kv_a + kc * (128 * 128));
        umma_bf16(tmem_base + mt * 128, ad + 2 * kk, bd + 2 * kk, id, ks != 0);
      }
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  constexpr int PITCH = 272;
  {
    constexpr int NCH = HDIM / 32;                     // 32-column chunks of an output row
    constexpr int CH = MT == 2 ? NCH : NCH / 2;
    const int mt = MT == 2 ? hi : 0, c0 = MT == 2 ? 0 : hi * CH;
    const int t = mt * 128 + quad * 32 + lane;
    const uint32_t t_row = t_lane + mt * 128;
#pragma unroll 1
    for (int c = c0; c < c0 + CH; ++c) {
      uint32_t raw[32];
      tmem_ld32(t_row + c * 32, raw);
      tmem_ld_wait();
      if (t < T) {
        uint4* dst = reinterpret_cast<uint4*>(Qs + t * PITCH + c * 64);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j] = make_uint4(pack2u(__uint_as_float(raw[8 * j]), __uint_as_float(raw[8 * j + 1])),
                              pack2u(__uint_as_float(raw[8 * j + 2]), __uint_as_float(raw[8 * j + 3])),
                              pack2u(__uint_as_float(raw[8 * j + 4]), __uint_as_float(raw[8 * j + 5])),
                              pack2u(__uint_as_float(raw[8 * j + 6]), __uint_as_float(raw[8 * j + 7])));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  constexpr int CPR = HDIM / 8;                        // 16-byte pieces of an output row
  for (int i = tid; i < T * CPR; i += NTHR) {
    const int r = i / CPR, c = i % CPR;
    *reinterpret_cast<uint4*>(o + ((long)(b * T + r)) * D + h * HDIM + c * 8) =
        *reinterpret_cast<const uint4*>(Qs + r * PITCH + c * 16);
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TP);
  }
}

template <int TP, int HDIM = 128>
int launch_sc(const bf16* q, const bf16* k, const bf16* v, const int* nt, int B, int H, int T, int Nt_max, float scale,
              bf16* o, cudaStream_t st) {
  using L = SmemLC<TP>;
  static unsigned long long attr = 0;   // one bit per device ordinal: the attribute is per (function, device)
  const unsigned long long dev_bit = 1ull << mdm_cur_dev();
  if (!(attr & dev_bit)) {
    if (cudaFuncSetAttribute(softmax_cross_umma_kernel<TP, HDIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) !=
        cudaSuccess)
      return MDM_ERR_CUDA;
    attr |= dev_bit;
  }
  const int NK = (Nt_max + 31) / 32 * 32;
  mdm_launch(softmax_cross_umma_kernel<TP, HDIM>, B * H, NTHR, L::TOTAL, st, q, k, v, nt, H, T, Nt_max, NK, scale, o);
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

}  // namespace

// hd == M == 128 (P^T [128 x 128] in bf16) or hd == M == 64 with an even head count (diag(P^T, P^T) [128 x 128] in bf16),
// bf16, T <= 256; MDM_ERR_UNSUPPORTED otherwise
// (the caller falls back to the mma.sync kernels of attention_tc.cu).
int mdm_fastattn_umma(const void* qkv, const void* Pt_bf16, const float* norm_w, const float* norm_b,
                      const int64_t* length, int length_shift, int B, int H, int T, int hd, int M, void* out,
                      const int* seq_order, cudaStream_t st) {
  const bool two = hd == 64 && M == 64 && (H & 1) == 0;      // two heads of 64 per CTA: Pt_bf16 is diag(P^T, P^T) [128 x 128]
  if (!((hd == HD && M == HD) || two) || T > 256 || !Pt_bf16) return MDM_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (reinterpret_cast<uintptr_t>(Pt_bf16) & 15))
    return MDM_ERR_UNSUPPORTED;
  const bf16* q = reinterpret_cast<const bf16*>(qkv);
  const bf16* p = reinterpret_cast<const bf16*>(Pt_bf16);
  bf16* o = reinterpret_cast<bf16*>(out);
  if (two) {
    if (T <= 128) return launch<128, 2>(q, p, norm_w, norm_b, length, length_shift, B, H / 2, T, o, seq_order, st);
    return launch<256, 2>(q, p, norm_w, norm_b, length, length_shift, B, H / 2, T, o, seq_order, st);
  }
  if (T <= 128) return launch<128, 1>(q, p, norm_w, norm_b, length, length_shift, B, H, T, o, seq_order, st);
  return launch<256, 1>(q, p, norm_w, norm_b, length, length_shift, B, H, T, o, seq_order, st);
}

// hd == 128, bf16, T <= 256, with ctx^T in bf16 ([B, H, l, d]), or hd == 64 with an even head count and the block-diagonal
// ctx^T of each pair of heads ([B, H / 2, 128, 128]); MDM_ERR_UNSUPPORTED otherwise.
int mdm_lincross_apply_umma(const void* q, const void* ctxT_bf16, int B, int T, int H, int hd, void* y, cudaStream_t st) {
  if (!(hd == HD || (hd == 64 && (H & 1) == 0)) || T > 256 || !ctxT_bf16) return MDM_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(y) & 15) ||
      (reinterpret_cast<uintptr_t>(ctxT_bf16) & 15))
    return MDM_ERR_UNSUPPORTED;
  const bf16* qq = reinterpret_cast<const bf16*>(q);
  const bf16* cc = reinterpret_cast<const bf16*>(ctxT_bf16);
  bf16* yy = reinterpret_cast<bf16*>(y);
  if (hd == 64) {     // ctxT_bf16: [B, H / 2, 128, 128], the block-diagonal ctx^T of each pair of heads
    if (T <= 128) return launch_lc<128, false, 2>(qq, cc, B, H / 2, T, yy, nullptr, nullptr, nullptr, st);
    return launch_lc<256, false, 2>(qq, cc, B, H / 2, T, yy, nullptr, nullptr, nullptr, st);
  }
  if (T <= 128) return launch_lc<128, false>(qq, cc, B, H, T, yy, nullptr, nullptr, nullptr, st);
  return launch_lc<256, false>(qq, cc, B, H, T, yy, nullptr, nullptr, nullptr, st);
}

// C-ABI: see include/mdm_b200.h
extern "C" MDM_API int mdm_lincross_apply_style(const void* q, const void* ctxT_bf16, int B, int T, int H, int hd,
                                                const float* ln_w, const float* ln_b, const float* film, void* y, void* stream) {
  if (!q || !ctxT_bf16 || !ln_w || !ln_b || !film || !y || B <= 0 || T <= 0) return MDM_ERR_ARG;
  const bool two = hd == 64 && (H & 1) == 0;      // pairs of heads of 64: ctxT_bf16 is [B, H / 2, 128, 128] block-diagonal
  if (!(hd == HD || two) || T > 256 || H < 1 || H / (two ? 2 : 1) > 8) return MDM_ERR_UNSUPPORTED;     // portable cluster size
  if ((reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(y) & 15) ||
      (reinterpret_cast<uintptr_t>(ctxT_bf16) & 15))
    return MDM_ERR_UNSUPPORTED;
  const bf16* qq = reinterpret_cast<const bf16*>(q);
  const bf16* cc = reinterpret_cast<const bf16*>(ctxT_bf16);
  bf16* yy = reinterpret_cast<bf16*>(y);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (hd == 64) {
    if (T <= 128) return launch_lc<128, true, 2>(qq, cc, B, H / 2, T, yy, ln_w, ln_b, film, st);
    return launch_lc<256, true, 2>(qq, cc, B, H / 2, T, yy, ln_w, ln_b, film, st);
  }
  if (T <= 128) return launch_lc<128, true>(qq, cc, B, H, T, yy, ln_w, ln_b, film, st);
  return launch_lc<256, true>(qq, cc, B, H, T, yy, ln_w, ln_b, film, st);
}

// hd == 128 or 64, bf16, T <= 256, at most 96 keys; MDM_ERR_UNSUPPORTED otherwise.
int mdm_softmax_cross_umma(const void* q, const void* k, const void* v, const int* nt, int B, int T, int Nt_max, int H,
                           int hd, float scale, void* o, cudaStream_t st) {
  if ((hd != HD && hd != 64) || T > 256 || Nt_max > 96 || Nt_max < 1) return MDM_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(k) & 15) ||
      (reinterpret_cast<uintptr_t>(v) & 15) || (reinterpret_cast<uintptr_t>(o) & 15))
    return MDM_ERR_UNSUPPORTED;
  const bf16 *qq = reinterpret_cast<const bf16*>(q), *kk = reinterpret_cast<const bf16*>(k),
             *vv = reinterpret_cast<const bf16*>(v);
  bf16* oo = reinterpret_cast<bf16*>(o);
  if (hd == 64) {
    if (T <= 128) return launch_sc<128, 64>(qq, kk, vv, nt, B, H, T, Nt_max, scale, oo, st);
    return launch_sc<256, 64>(qq, kk, vv, nt, B, H, T, Nt_max, scale, oo, st);
  }
  if (T <= 128) return launch_sc<128>(qq, kk, vv, nt, B, H, T, Nt_max, scale, oo, st);
  return launch_sc<256>(qq, kk, vv, nt, B, H, T, Nt_max, scale, oo, st);
}
