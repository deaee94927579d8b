// Backward / optimizer kernels of the DDPM training step (BASELINE.json configs[4]):
// reference trainers/ddpm_trainer.py:201-244 (backward_G, clip_grad_norm_, Adam) and the autograd graph of
// models/transformer.py:291-361.  Everything GEMM-shaped over tokens runs on the forward GEMMs (gemm_tc.cu / gemm_simt.cu);
// this file holds what those cannot express:
//   * mdm_bgemm: strided batched GEMM over (sequence, head) pairs - the small per-head products of the three
//     attention cores' backward passes (operands are slices of token-major [N, H*hd] tensors or head-major scratch);
//   * the row kernels of those backward passes (LayerNorm(hd) / L2 norm / exp feature map / softmax and their
//     derivatives, fast_attention.py:29-92, 242-258, 305-325);
//   * the MoE routing backward (gate softmax / top-2 weights, token un-permute, switch_moe.py:53-109);
//   * activation derivatives, the masked-MSE gradient, global-norm clipping and the fused Adam update.
// All kernels take fp32 or bf16 activations (MDM_F32 / MDM_BF16) and accumulate in fp32; parameter gradients are
// deterministic (fixed-order partial sums, no float atomics).
#include "common.cuh"
#include <type_traits>

namespace {

template <typename T> __device__ __forceinline__ float ldf(const T* p) { return to_f<T>(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v) { *p = from_f<T>(v); }

// warp-per-row mapping for narrow rows (W = 32 * VPT, VPT >= 1): lane owns columns lane + 32 * i
template <int VPT, typename T>
__device__ __forceinline__ void ldrow(const T* __restrict__ p, int lane, float (&v)[VPT]) {
#pragma unroll
  for (int i = 0; i < VPT; ++i) v[i] = to_f<T>(p[lane + 32 * i]);
}
template <int VPT, typename T>
__device__ __forceinline__ void strow(T* __restrict__ p, int lane, const float (&v)[VPT]) {
#pragma unroll
  for (int i = 0; i < VPT; ++i) p[lane + 32 * i] = from_f<T>(v[i]);
}
template <int VPT> __device__ __forceinline__ float rowsum(const float (&v)[VPT]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) s += v[i];
  return warp_sum(s);
}
template <int VPT> __device__ __forceinline__ float rowdot(const float (&a)[VPT], const float (&b)[VPT]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) s = fmaf(a[i], b[i], s);
  return warp_sum(s);
}

// LayerNorm forward of one row in registers: xh = (x - mean) * rstd (returned in x), rstd returned.
template <int VPT> __device__ __forceinline__ float ln_normalize(float (&x)[VPT], int W) {
  const float mean = rowsum<VPT>(x) / (float)W;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) { x[i] -= mean; q = fmaf(x[i], x[i], q); }
  const float rstd = rsqrtf(warp_sum(q) / (float)W + 1e-5f);
#pragma unroll
  for (int i = 0; i < VPT; ++i) x[i] *= rstd;
  return rstd;
}
// LayerNorm backward of one row: g = dL/dy (y = xh * w + b) -> g = dL/dx; accumulates dw, db.
template <int VPT>
__device__ __forceinline__ void ln_backward(float (&g)[VPT], const float (&xh)[VPT], float rstd, const float (&w)[VPT],
                                            float (&dw)[VPT], float (&db)[VPT], int W) {
  float m1 = 0.f, m2 = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    dw[i] = fmaf(g[i], xh[i], dw[i]); db[i] += g[i];
    g[i] *= w[i];
    m1 += g[i]; m2 = fmaf(g[i], xh[i], m2);
  }
  m1 = warp_sum(m1) / (float)W; m2 = warp_sum(m2) / (float)W;
#pragma unroll
  for (int i = 0; i < VPT; ++i) g[i] = rstd * (g[i] - m1 - xh[i] * m2);
}

// per-block partial sums of NV vectors of width W held per warp in registers -> part[blockIdx.x][v][W] (fixed order)
template <int VPT, int NV, int WARPS>
__device__ __forceinline__ void block_reduce_vectors(float (&vec)[NV][VPT], float* __restrict__ part, int W, float* smem) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int v = 0; v < NV; ++v) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < VPT; ++i) smem[warp * W + lane + 32 * i] = vec[v][i];
    __syncthreads();
    for (int c = threadIdx.x; c < W; c += WARPS * 32) {
      float s = 0.f;
      for (int w = 0; w < WARPS; ++w) s += smem[w * W + c];
      part[((long)blockIdx.x * NV + v) * W + c] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// strided batched GEMM:  C[z][m][n] (+)= alpha * sum_k A[z][m][k] * B[z][k][n],  z = (z1, z2)
// ---------------------------------------------------------------------------------------------------------------
constexpr int BT = 64, BK = 16;
template <typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(256)
bgemm_kernel(const MdmBgemm g) {
  __shared__ float As[BK][BT + 4];
  __shared__ float Bs[BK][BT + 4];
  const int z = blockIdx.z, z1 = z / g.Z2, z2 = z - z1 * g.Z2;
  const TA* A = reinterpret_cast<const TA*>(g.A) + z1 * g.a_z1 + z2 * g.a_z2;
  const TB* B = reinterpret_cast<const TB*>(g.B) + z1 * g.b_z1 + z2 * g.b_z2;
  TC* C = reinterpret_cast<TC*>(g.C) + z1 * g.c_z1 + z2 * g.c_z2;
  int M = g.M, K = g.K;
  if (g.m_limit) M = min(M, max(0, (int)g.m_limit[z1] >> g.limit_shift));
  if (g.k_limit) K = min(K, max(0, (int)g.k_limit[z1] >> g.limit_shift));
  const int m0 = blockIdx.y * BT, n0 = blockIdx.x * BT;
  if (m0 >= g.M) return;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4] = {};
  const bool a_kfast = g.a_cs == 1, b_nfast = g.b_cs == 1;
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m, k;
      if (a_kfast) { k = tid & 15; m = (tid >> 4) + 16 * i; } else { m = tid & 63; k = (tid >> 6) + 4 * i; }
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < M && gk < K) ? ldf<TA>(A + (long)gm * g.a_rs + (long)gk * g.a_cs) : 0.f;
      int n, kb;
      if (b_nfast) { n = tid & 63; kb = (tid >> 6) + 4 * i; } else { kb = tid & 15; n = (tid >> 4) + 16 * i; }
      const int gn = n0 + n, gkb = k0 + kb;
      Bs[kb][n] = (gn < g.N && gkb < K) ? ldf<TB>(B + (long)gkb * g.b_rs + (long)gn * g.b_cs) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      TC* c = C + (long)m * g.c_rs + (long)n * g.c_cs;
      float v = (m < M) ? g.alpha * acc[i][j] : 0.f;
      if (g.accumulate) v += ldf<TC>(c);
      stf<TC>(c, v);
    }
  }
}

// Tensor-core flavour for the bf16 training path: the same contract, operands rounded to bf16 while they are staged in
// shared memory (whatever their storage type), mma.sync.m16n8k16 with fp32 accumulators.  These products are 1-2 % of the
// step's FLOPs and have no fixed operand layout (slices of token-major tensors, transposed head-major scratch), so they
// use register-fragment MMAs on generic strided tiles rather than TMA-fed tcgen05 tiles.  64 x 64 x 32 tiles, 8 warps
// (2 x 4), the next k-tile prefetched into registers while the current one is multiplied.
constexpr int TBK = 32, TPITCH = TBK + 8;      // bf16 row pitch 80 B: conflict-free 32-bit fragment loads
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(256, 3)      // <= 85 registers: three CTAs per SM (ncu: 118 registers gave 23 % warps active)
bgemm_tc_kernel(const MdmBgemm g) {
  __shared__ __align__(16) bf16 As[BT][TPITCH];     // [m][k]
  __shared__ __align__(16) bf16 Bs[BT][TPITCH];     // [n][k]
  const int z = blockIdx.z, z1 = z / g.Z2, z2 = z - z1 * g.Z2;
  const TA* A = reinterpret_cast<const TA*>(g.A) + z1 * g.a_z1 + z2 * g.a_z2;
  const TB* B = reinterpret_cast<const TB*>(g.B) + z1 * g.b_z1 + z2 * g.b_z2;
  TC* C = reinterpret_cast<TC*>(g.C) + z1 * g.c_z1 + z2 * g.c_z2;
  int M = g.M, K = g.K;
  if (g.m_limit) M = min(M, max(0, (int)g.m_limit[z1] >> g.limit_shift));
  if (g.k_limit) K = min(K, max(0, (int)g.k_limit[z1] >> g.limit_shift));
  const int m0 = blockIdx.y * BT, n0 = blockIdx.x * BT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = (warp >> 2) * 32, wn = (warp & 3) * 16;      // warp tile 32 x 16
  const int gq = lane >> 2, tq = lane & 3;
  const bool a_kfast = g.a_cs == 1, b_nfast = g.b_cs == 1;
  float acc[2][2][4] = {};
  float ra[8], rb[8];                                           // 64 x 32 elements / 256 threads = 8 each
  // Loads: 4 consecutive elements of the operand's contiguous dimension per access when strides / alignment allow
  // (every product of the attention backward does: row pitches and batch strides are multiples of 4 elements);
  // element by element otherwise.  A tile = 64 (m) x 32 (k), B tile = 32 (k) x 64 (n).
  constexpr int VA = 16 / sizeof(TA) >= 4 ? 4 : 1, VB = 16 / sizeof(TB) >= 4 ? 4 : 1;
  auto al = [](const void* p, long a, long b, long c, int esz) {
    return ((reinterpret_cast<uintptr_t>(p) | (uintptr_t)(a * esz) | (uintptr_t)(b * esz) | (uintptr_t)(c * esz)) & (4 * esz - 1)) == 0;
  };
  // (only where the shared-memory store is contiguous too, i.e. the operand is k-fast: the transposing cases scatter four
  //  2-byte stores 4 rows apart - 16-way bank conflicts under ncu - and stay element-wise, lanes along the fast dimension)
  const bool a_vec = VA == 4 && a_kfast && al(g.A, g.a_z1, g.a_z2, g.a_rs, sizeof(TA));
  const bool b_vec = VB == 4 && !b_nfast && g.b_rs == 1 && al(g.B, g.b_z1, g.b_z2, g.b_cs, sizeof(TB));
  auto ld4 = [](const auto* p, float (&v)[4]) {
    typedef typename std::remove_cv<typename std::remove_pointer<decltype(p)>::type>::type E;
    if constexpr (sizeof(E) == 4) {
      const float4 t = *reinterpret_cast<const float4*>(p);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
      const uint2 t = *reinterpret_cast<const uint2*>(p);
      const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&t.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
      v[0] = __low2float(lo); v[1] = __high2float(lo); v[2] = __low2float(hi); v[3] = __high2float(hi);
    }
  };
  // index maps: (i-th access of this thread) -> (slow index, first fast index); vector form covers 4 fast indices
  auto fetch = [&](int k0) {
    if (a_vec) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int q = tid + 256 * i;                            // 512 vectors
        float v[4];
        if (a_kfast) {                                          // 8 vectors per row of 32 k
          const int m = q >> 3, k = (q & 7) * 4, gm = m0 + m, gk = k0 + k;
          const TA* p = A + (long)gm * g.a_rs + gk;
          if (gm < M && gk + 3 < K) ld4(p, v);
          else
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (gm < M && gk + e < K) ? ldf<TA>(p + e) : 0.f;
        } else {                                                // m fast: 16 vectors per k row of 64 m
          const int k = q >> 4, m = (q & 15) * 4, gm = m0 + m, gk = k0 + k;
          const TA* p = A + (long)gk * g.a_cs + gm;
          if (gk < K && gm + 3 < M) ld4(p, v);
          else
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (gk < K && gm + e < M) ? ldf<TA>(p + e) : 0.f;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) ra[4 * i + e] = v[e];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int m, k;
        if (a_kfast) { k = tid & 31; m = (tid >> 5) + 8 * i; } else { m = tid & 63; k = (tid >> 6) + 4 * i; }
        const int gm = m0 + m, gk = k0 + k;
        ra[i] = (gm < M && gk < K) ? ldf<TA>(A + (long)gm * g.a_rs + (long)gk * g.a_cs) : 0.f;
      }
    }
    if (b_vec) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int q = tid + 256 * i;
        float v[4];
        if (b_nfast) {                                          // 16 vectors per k row of 64 n
          const int k = q >> 4, n = (q & 15) * 4, gn = n0 + n, gk = k0 + k;
          const TB* p = B + (long)gk * g.b_rs + gn;
          if (gk < K && gn + 3 < g.N) ld4(p, v);
          else
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (gk < K && gn + e < g.N) ? ldf<TB>(p + e) : 0.f;
        } else {                                                // k fast: 8 vectors per n row of 32 k
          const int n = q >> 3, k = (q & 7) * 4, gn = n0 + n, gk = k0 + k;
          const TB* p = B + (long)gn * g.b_cs + gk;
          if (gn < g.N && gk + 3 < K) ld4(p, v);
          else
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (gn < g.N && gk + e < K) ? ldf<TB>(p + e) : 0.f;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) rb[4 * i + e] = v[e];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int n, kb;
        if (b_nfast) { n = tid & 63; kb = (tid >> 6) + 4 * i; } else { kb = tid & 31; n = (tid >> 5) + 8 * i; }
        const int gn = n0 + n, gkb = k0 + kb;
        rb[i] = (gn < g.N && gkb < K) ? ldf<TB>(B + (long)gkb * g.b_rs + (long)gn * g.b_cs) : 0.f;
      }
    }
  };
  auto stage = [&]() {
    if (a_vec) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int q = tid + 256 * i;
        if (a_kfast) {
          const int m = q >> 3, k = (q & 7) * 4;
          *reinterpret_cast<__nv_bfloat162*>(&As[m][k]) = __floats2bfloat162_rn(ra[4 * i], ra[4 * i + 1]);
          *reinterpret_cast<__nv_bfloat162*>(&As[m][k + 2]) = __floats2bfloat162_rn(ra[4 * i + 2], ra[4 * i + 3]);
        } else {
          const int k = q >> 4, m = (q & 15) * 4;
#pragma unroll
          for (int e = 0; e < 4; ++e) As[m + e][k] = __float2bfloat16_rn(ra[4 * i + e]);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int m, k;
        if (a_kfast) { k = tid & 31; m = (tid >> 5) + 8 * i; } else { m = tid & 63; k = (tid >> 6) + 4 * i; }
        As[m][k] = __float2bfloat16_rn(ra[i]);
      }
    }
    if (b_vec) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int q = tid + 256 * i;
        if (b_nfast) {
          const int k = q >> 4, n = (q & 15) * 4;
#pragma unroll
          for (int e = 0; e < 4; ++e) Bs[n + e][k] = __float2bfloat16_rn(rb[4 * i + e]);
        } else {
          const int n = q >> 3, k = (q & 7) * 4;
          *reinterpret_cast<__nv_bfloat162*>(&Bs[n][k]) = __floats2bfloat162_rn(rb[4 * i], rb[4 * i + 1]);
          *reinterpret_cast<__nv_bfloat162*>(&Bs[n][k + 2]) = __floats2bfloat162_rn(rb[4 * i + 2], rb[4 * i + 3]);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int n, kb;
        if (b_nfast) { n = tid & 63; kb = (tid >> 6) + 4 * i; } else { kb = tid & 31; n = (tid >> 5) + 8 * i; }
        Bs[n][kb] = __float2bfloat16_rn(rb[i]);
      }
    }
  };
  if (m0 < g.M) {
    fetch(0);
    for (int k0 = 0; k0 < K; k0 += TBK) {
      stage();
      __syncthreads();
      if (k0 + TBK < K) fetch(k0 + TBK);
#pragma unroll
      for (int kk = 0; kk < TBK; kk += 16) {
        uint32_t af[2][4], bfr[2][2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const bf16* ap = &As[wm + 16 * i + gq][kk + 2 * tq];
          af[i][0] = *reinterpret_cast<const uint32_t*>(ap);
          af[i][1] = *reinterpret_cast<const uint32_t*>(ap + 8 * TPITCH);
          af[i][2] = *reinterpret_cast<const uint32_t*>(ap + 8);
          af[i][3] = *reinterpret_cast<const uint32_t*>(ap + 8 * TPITCH + 8);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const bf16* bp = &Bs[wn + 8 * j + gq][kk + 2 * tq];
          bfr[j][0] = *reinterpret_cast<const uint32_t*>(bp);
          bfr[j][1] = *reinterpret_cast<const uint32_t*>(bp + 8);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) mma_bf16_16816(acc[i][j], af[i], bfr[j]);
      }
      __syncthreads();
    }
  }
  if (m0 >= g.M) return;
  // the two columns a thread holds per accumulator row are adjacent: one 8-byte (fp32) / 4-byte (bf16) store when C is
  // row-major and aligned (a warp store then fills whole 32-byte sectors instead of half of each)
  const bool pair_ok = g.c_cs == 1 && !g.accumulate && !(g.c_rs & 1) && !(g.c_z1 & 1) && !(g.c_z2 & 1) &&
                       !(reinterpret_cast<uintptr_t>(g.C) & (2 * sizeof(TC) - 1));
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int m = m0 + wm + 16 * i + gq + rr * 8, n = n0 + wn + 8 * j + 2 * tq;
        if (m >= g.M || n >= g.N) continue;
        const float v0 = (m < M) ? g.alpha * acc[i][j][2 * rr] : 0.f, v1 = (m < M) ? g.alpha * acc[i][j][2 * rr + 1] : 0.f;
        TC* c = C + (long)m * g.c_rs + (long)n * g.c_cs;
        if (pair_ok && n + 1 < g.N) {
          if constexpr (sizeof(TC) == 4) *reinterpret_cast<float2*>(c) = make_float2(v0, v1);
          else *reinterpret_cast<__nv_bfloat162*>(c) = __floats2bfloat162_rn(v0, v1);
        } else {
          stf<TC>(c, g.accumulate ? v0 + ldf<TC>(c) : v0);
          if (n + 1 < g.N) stf<TC>(c + g.c_cs, g.accumulate ? v1 + ldf<TC>(c + g.c_cs) : v1);
        }
      }
}

// ---------------------------------------------------------------------------------------------------------------
// FastAttention backward pieces (fast_attention.py:29-92 with the pre-scale of :155-157)
// ---------------------------------------------------------------------------------------------------------------
// prep (forward recompute): per (token, head): x0 = 0.1 * raw; xn = LN(x0; w, b); q, k additionally L2-normalised.
// raw qkv [N, 3D] token-major -> qh, kh, vn fp32 head-major [B, H, T, hd].
template <int VPT, typename T>
__global__ void __launch_bounds__(256)
fa_prep_kernel(const T* __restrict__ qkv, const float* __restrict__ nw, const float* __restrict__ nb, int B, int H, int Tn,
               float* __restrict__ qh, float* __restrict__ kh, float* __restrict__ vn) {
  constexpr int W = 32 * VPT;
  const long r = (long)blockIdx.x * 8 + (threadIdx.x >> 5);       // (token, head)
  if (r >= (long)B * Tn * H) return;
  const int lane = threadIdx.x & 31;
  const long tok = r / H;
  const int h = (int)(r - tok * H);
  const int b = (int)(tok / Tn), t = (int)(tok - (long)b * Tn);
  const int D = H * W;
  float w[VPT], bb[VPT];
  ldrow<VPT, float>(nw, lane, w); ldrow<VPT, float>(nb, lane, bb);
  float* outs[3] = {qh, kh, vn};
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    float x[VPT];
    ldrow<VPT, T>(qkv + tok * 3 * D + s * D + h * W, lane, x);
#pragma unroll
    for (int i = 0; i < VPT; ++i) x[i] *= 0.1f;
    ln_normalize<VPT>(x, W);
#pragma unroll
    for (int i = 0; i < VPT; ++i) x[i] = x[i] * w[i] + bb[i];
    if (s < 2) {
      const float nrm = fmaxf(sqrtf(rowdot<VPT>(x, x)), 1e-12f);
#pragma unroll
      for (int i = 0; i < VPT; ++i) x[i] = x[i] / nrm;
    }
    strow<VPT, float>(outs[s] + (((long)b * H + h) * Tn + t) * W, lane, x);
  }
}

// feature maps: qp = 0.1 exp(clamp(uq, +-15)); kp = 0.1 exp(clamp(uk, +-15)) * [t < length[b] >> shift]
// (qp / kp may be the same buffers as uq / uk: no __restrict__ on them)
__global__ void fa_feat_kernel(const float* uq, const float* uk, const int64_t* __restrict__ length,
                               int shift, int H, int Tn, int M, long total, float* qp, float* kp) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long row = i / M;                       // (b, h, t)
  const int t = (int)(row % Tn);
  const int b = (int)(row / ((long)Tn * H));
  const long len = length ? (length[b] >> shift) : (long)Tn;
  qp[i] = expf(fminf(fmaxf(uq[i], -15.f), 15.f)) * 0.1f;
  kp[i] = t < len ? expf(fminf(fmaxf(uk[i], -15.f), 15.f)) * 0.1f : 0.f;
}

// output stage backward: o = 0.1 q' kv (given), den = max(sum_m q'k', 1e-6), r = o / den, out = LN(r).
// In: dout token-major [N, D] (T).  Out: d_o (over o, head-major), dden [BH, T]; partial (dw, db) per block.
template <int VPT, typename T>
__global__ void __launch_bounds__(256)
fa_out_bwd_kernel(float* __restrict__ o, const float* __restrict__ qp, const float* __restrict__ kp,
                  const T* __restrict__ dout, const float* __restrict__ nw, int B, int H, int Tn,
                  float* __restrict__ dden, float* __restrict__ part) {
  constexpr int W = 32 * VPT;
  __shared__ float red[8 * W];
  const int lane = threadIdx.x & 31;
  float acc[2][VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) acc[0][i] = acc[1][i] = 0.f;
  // capped grid, fixed row -> CTA assignment: one partial per CTA (deterministic), a few thousand instead of rows / 8
  for (long r = (long)blockIdx.x * 8 + (threadIdx.x >> 5); r < (long)B * H * Tn; r += (long)gridDim.x * 8) {
    const int t = (int)(r % Tn);
    const long bh = r / Tn;
    const int h = (int)(bh % H), b = (int)(bh / H);
    float ov[VPT], q[VPT], k[VPT], g[VPT], w[VPT];
    ldrow<VPT, float>(o + r * W, lane, ov);
    ldrow<VPT, float>(qp + r * W, lane, q);
    ldrow<VPT, float>(kp + r * W, lane, k);
    ldrow<VPT, T>(dout + ((long)b * Tn + t) * H * W + h * W, lane, g);
    ldrow<VPT, float>(nw, lane, w);
    const float dsum = rowdot<VPT>(q, k);
    const float den = fmaxf(dsum, 1e-6f);
    float xh[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) xh[i] = ov[i] / den;
    const float rstd = ln_normalize<VPT>(xh, W);
    ln_backward<VPT>(g, xh, rstd, w, acc[0], acc[1], W);         // g = d r
    const float dot = rowdot<VPT>(g, ov);
#pragma unroll
    for (int i = 0; i < VPT; ++i) g[i] = g[i] / den;
    strow<VPT, float>(o + r * W, lane, g);
    if (lane == 0) dden[r] = dsum > 1e-6f ? -dot / (den * den) : 0.f;
  }
  block_reduce_vectors<VPT, 2, 8>(acc, part, W, red);
}

// output stage FORWARD (generic head sizes, e.g. hd = 256 of model_size="big"): den = max(sum_m q'k', 1e-6),
// out = LN(o / den; w, b) written token-major [N, H*hd] (T)
template <int VPT, typename T>
__global__ void __launch_bounds__(256)
fa_out_fwd_kernel(const float* __restrict__ o, const float* __restrict__ qp, const float* __restrict__ kp,
                  const float* __restrict__ nw, const float* __restrict__ nb, int B, int H, int Tn, T* __restrict__ out) {
  constexpr int W = 32 * VPT;
  const long r = (long)blockIdx.x * 8 + (threadIdx.x >> 5);       // head-major row (b, h, t)
  if (r >= (long)B * H * Tn) return;
  const int lane = threadIdx.x & 31;
  const int t = (int)(r % Tn);
  const long bh = r / Tn;
  const int h = (int)(bh % H), b = (int)(bh / H);
  float ov[VPT], q[VPT], k[VPT], w[VPT], bb[VPT];
  ldrow<VPT, float>(o + r * W, lane, ov);
  ldrow<VPT, float>(qp + r * W, lane, q);
  ldrow<VPT, float>(kp + r * W, lane, k);
  ldrow<VPT, float>(nw, lane, w); ldrow<VPT, float>(nb, lane, bb);
  const float den = fmaxf(rowdot<VPT>(q, k), 1e-6f);
#pragma unroll
  for (int i = 0; i < VPT; ++i) ov[i] = ov[i] / den;
  ln_normalize<VPT>(ov, W);
#pragma unroll
  for (int i = 0; i < VPT; ++i) ov[i] = ov[i] * w[i] + bb[i];
  strow<VPT, T>(out + ((long)b * Tn + t) * H * W + h * W, lane, ov);
}

// feature-map backward (in place): duq = (dqp + dden * kp) * qp * [|uq| <= 15]; duk = (dkp + dden * qp) * kp * [|uk| <= 15]
__global__ void fa_feat_bwd_kernel(const float* __restrict__ uq, const float* __restrict__ uk, const float* __restrict__ qp,
                                   const float* __restrict__ kp, const float* __restrict__ dden, int M, long total,
                                   float* __restrict__ dqp, float* __restrict__ dkp) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float dd = dden[i / M];
  const float q = qp[i], k = kp[i];
  dqp[i] = fabsf(uq[i]) <= 15.f ? (dqp[i] + dd * k) * q : 0.f;
  dkp[i] = fabsf(uk[i]) <= 15.f ? (dkp[i] + dd * q) * k : 0.f;
}

// prep backward: recompute LN / L2 from the raw rows; dqh, dkh, dvn head-major fp32 -> dqkv token-major (T), with the
// 0.1 pre-scale and the reference's gradient clamp to [-1, 1] on q, k, v (fast_attention.py:150-152).
template <int VPT, typename T>
__global__ void __launch_bounds__(256)
fa_prep_bwd_kernel(const T* __restrict__ qkv, const float* __restrict__ nw, const float* __restrict__ nb, int B, int H, int Tn,
                   const float* __restrict__ dqh, const float* __restrict__ dkh, const float* __restrict__ dvn,
                   T* __restrict__ dqkv, float* __restrict__ part) {
  constexpr int W = 32 * VPT;
  __shared__ float red[8 * W];
  const int lane = threadIdx.x & 31;
  float acc[2][VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) acc[0][i] = acc[1][i] = 0.f;
  for (long r = (long)blockIdx.x * 8 + (threadIdx.x >> 5); r < (long)B * Tn * H; r += (long)gridDim.x * 8) {   // (token, head)
    const long tok = r / H;
    const int h = (int)(r - tok * H);
    const int b = (int)(tok / Tn), t = (int)(tok - (long)b * Tn);
    const int D = H * W;
    float w[VPT], bb[VPT];
    ldrow<VPT, float>(nw, lane, w); ldrow<VPT, float>(nb, lane, bb);
    const float* gin[3] = {dqh, dkh, dvn};
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      float xh[VPT], g[VPT];
      ldrow<VPT, T>(qkv + tok * 3 * D + s * D + h * W, lane, xh);
#pragma unroll
      for (int i = 0; i < VPT; ++i) xh[i] *= 0.1f;
      const float rstd = ln_normalize<VPT>(xh, W);
      ldrow<VPT, float>(gin[s] + (((long)b * H + h) * Tn + t) * W, lane, g);
      if (s < 2) {                                  // L2 normalisation backward
        float xn[VPT];
#pragma unroll
        for (int i = 0; i < VPT; ++i) xn[i] = xh[i] * w[i] + bb[i];
        const float nrm = fmaxf(sqrtf(rowdot<VPT>(xn, xn)), 1e-12f);
        const float dot = rowdot<VPT>(xn, g) / (nrm * nrm);
#pragma unroll
        for (int i = 0; i < VPT; ++i) g[i] = (g[i] - xn[i] * dot) / nrm;
      }
      ln_backward<VPT>(g, xh, rstd, w, acc[0], acc[1], W);
#pragma unroll
      for (int i = 0; i < VPT; ++i) g[i] = fminf(fmaxf(0.1f * g[i], -1.f), 1.f);
      strow<VPT, T>(dqkv + tok * 3 * D + s * D + h * W, lane, g);
    }
  }
  block_reduce_vectors<VPT, 2, 8>(acc, part, W, red);
}

// ---------------------------------------------------------------------------------------------------------------
// row softmax helpers (LinearTemporalCrossAttention q side, MemoryEfficientCrossAttention scores)
// ---------------------------------------------------------------------------------------------------------------
// P[(b,h),t,:] = softmax(q[t, h*W : (h+1)*W]) : token-major (T) -> head-major fp32
template <int VPT, typename T>
__global__ void __launch_bounds__(256)
head_softmax_kernel(const T* __restrict__ q, int B, int H, int Tn, float* __restrict__ P) {
  constexpr int W = 32 * VPT;
  const long r = (long)blockIdx.x * 8 + (threadIdx.x >> 5);       // (token, head)
  if (r >= (long)B * Tn * H) return;
  const int lane = threadIdx.x & 31;
  const long tok = r / H;
  const int h = (int)(r - tok * H);
  const int b = (int)(tok / Tn), t = (int)(tok - (long)b * Tn);
  float x[VPT];
  ldrow<VPT, T>(q + tok * H * W + h * W, lane, x);
  float mx = x[0];
#pragma unroll
  for (int i = 1; i < VPT; ++i) mx = fmaxf(mx, x[i]);
  mx = warp_max(mx);
#pragma unroll
  for (int i = 0; i < VPT; ++i) x[i] = expf(x[i] - mx);
  const float s = rowsum<VPT>(x);
#pragma unroll
  for (int i = 0; i < VPT; ++i) x[i] = x[i] / s;
  strow<VPT, float>(P + (((long)b * H + h) * Tn + t) * W, lane, x);
}
// dq[t, h, :] = P * (dP - sum(P * dP)) : head-major fp32 -> token-major (T)
template <int VPT, typename T>
__global__ void __launch_bounds__(256)
head_softmax_bwd_kernel(const float* __restrict__ P, const float* __restrict__ dP, int B, int H, int Tn, T* __restrict__ dq) {
  constexpr int W = 32 * VPT;
  const long r = (long)blockIdx.x * 8 + (threadIdx.x >> 5);       // head-major row (b, h, t)
  if (r >= (long)B * Tn * H) return;
  const int lane = threadIdx.x & 31;
  const int t = (int)(r % Tn);
  const long bh = r / Tn;
  const int h = (int)(bh % H), b = (int)(bh / H);
  float p[VPT], g[VPT];
  ldrow<VPT, float>(P + r * W, lane, p);
  ldrow<VPT, float>(dP + r * W, lane, g);
  const float dot = rowdot<VPT>(p, g);
#pragma unroll
  for (int i = 0; i < VPT; ++i) g[i] = p[i] * (g[i] - dot);
  strow<VPT, T>(dq + ((long)b * Tn + t) * H * W + h * W, lane, g);
}
// masked softmax over the last dimension (len <= 96 keys, n < nt[b] valid), in place on fp32 rows [BH, T, NK]
__global__ void __launch_bounds__(256)
key_softmax_kernel(float* __restrict__ S, const int* __restrict__ nt, int H, int Tn, int NK, long rows) {
  const long r = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const int b = (int)(r / ((long)Tn * H));
  const int n_valid = nt ? min(nt[b], NK) : NK;
  float x[3], mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int n = lane + 32 * i;
    x[i] = n < n_valid ? S[r * NK + n] : -INFINITY;
    mx = fmaxf(mx, x[i]);
  }
  mx = warp_max(mx);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) { x[i] = (lane + 32 * i) < n_valid ? expf(x[i] - mx) : 0.f; s += x[i]; }
  s = warp_sum(s);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int n = lane + 32 * i;
    if (n < NK) S[r * NK + n] = x[i] / s;
  }
}
// dS = P * (dP - sum(P dP)) in place over dP
__global__ void __launch_bounds__(256)
key_softmax_bwd_kernel(const float* __restrict__ P, float* __restrict__ dP, int NK, long rows) {
  const long r = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  float p[3], g[3], dot = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int n = lane + 32 * i;
    p[i] = n < NK ? P[r * NK + n] : 0.f;
    g[i] = n < NK ? dP[r * NK + n] : 0.f;
    dot = fmaf(p[i], g[i], dot);
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int n = lane + 32 * i;
    if (n < NK) dP[r * NK + n] = p[i] * (g[i] - dot);
  }
}
// column softmax over the text tokens (n < nt[b]) of k [B, Nt, C] (T) -> Ks fp32 (0 for padded rows)  [fast_attention.py:251]
template <typename T>
__global__ void col_softmax_kernel(const T* __restrict__ k, const int* __restrict__ nt, int Nt, int C, long total,
                                   float* __restrict__ Ks) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;       // (b, c)
  if (i >= total) return;
  const int b = (int)(i / C), c = (int)(i - (long)b * C);
  const int nv = nt ? min(nt[b], Nt) : Nt;
  const T* kp = k + (long)b * Nt * C + c;
  float mx = -INFINITY;
  for (int n = 0; n < nv; ++n) mx = fmaxf(mx, ldf<T>(kp + (long)n * C));
  float s = 0.f;
  for (int n = 0; n < nv; ++n) s += expf(ldf<T>(kp + (long)n * C) - mx);
  float* o = Ks + (long)b * Nt * C + c;
  for (int n = 0; n < Nt; ++n) o[(long)n * C] = n < nv ? expf(ldf<T>(kp + (long)n * C) - mx) / s : 0.f;
}
// dk[b, n, c] = Ks * (dKs - sum_n Ks dKs)
template <typename T>
__global__ void col_softmax_bwd_kernel(const float* __restrict__ Ks, const float* __restrict__ dKs, int Nt, int C, long total,
                                       T* __restrict__ dk) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = (int)(i / C), c = (int)(i - (long)b * C);
  const long base = (long)b * Nt * C + c;
  float dot = 0.f;
  for (int n = 0; n < Nt; ++n) dot = fmaf(Ks[base + (long)n * C], dKs[base + (long)n * C], dot);
  for (int n = 0; n < Nt; ++n) stf<T>(dk + base + (long)n * C, Ks[base + (long)n * C] * (dKs[base + (long)n * C] - dot));
}

// ---------------------------------------------------------------------------------------------------------------
// MoE backward pieces (switch_moe.py:53-109, multi_branch.py:52-61)
// ---------------------------------------------------------------------------------------------------------------
// training combine: m[token] = sum over the NBK routed rows of rowscale[pos] * z[pos]
template <int VPT, typename T>
__global__ void __launch_bounds__(256)
moe_combine_sum_kernel(const T* __restrict__ z, const float* __restrict__ rowscale, const int* __restrict__ perm, long N, int NBK,
                       T* __restrict__ m) {
  constexpr int W = 32 * VPT;
  const long tok = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (tok >= N) return;
  const int lane = threadIdx.x & 31;
  float acc[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) acc[i] = 0.f;
  for (int j = 0; j < NBK; ++j) {
    const int pos = perm[tok * NBK + j];
    const float rs = rowscale[pos];
    float v[VPT];
    ldrow<VPT, T>(z + (long)pos * W, lane, v);
#pragma unroll
    for (int i = 0; i < VPT; ++i) acc[i] = fmaf(rs, v[i], acc[i]);
  }
  strow<VPT, T>(m + tok * W, lane, acc);
}
// combine backward: dz[pos] = rowscale[pos] * dm[token]; drs[pos] = <dm[token], z[pos]>
template <int VPT, typename T>
__global__ void __launch_bounds__(256)
moe_combine_bwd_kernel(const T* __restrict__ z, const float* __restrict__ rowscale, const int* __restrict__ perm, long N, int NBK,
                       const T* __restrict__ dm, T* __restrict__ dz, float* __restrict__ drs) {
  constexpr int W = 32 * VPT;
  const long tok = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (tok >= N) return;
  const int lane = threadIdx.x & 31;
  float g[VPT];
  ldrow<VPT, T>(dm + tok * W, lane, g);
  for (int j = 0; j < NBK; ++j) {
    const int pos = perm[tok * NBK + j];
    const float rs = rowscale[pos];
    float v[VPT], o[VPT];
    ldrow<VPT, T>(z + (long)pos * W, lane, v);
    const float dot = rowdot<VPT>(g, v);
#pragma unroll
    for (int i = 0; i < VPT; ++i) o[i] = rs * g[i];
    strow<VPT, T>(dz + (long)pos * W, lane, o);
    if (lane == 0) drs[pos] = dot;
  }
}
// gate backward, logits part: vals[token, b, k] = probs[idx] (weights NOT renormalised), rowscale = vals / NB.
// recompute logits = LN_b(x) Wg_b^T + bg_b and probs; dprobs[e] = sum_k [idx_k == e] drs[pos_k] / NB;
// dlogits = probs * (dprobs - <probs, dprobs>)  ->  dlogits [N, NB*E] fp32
template <int VPT>
__global__ void __launch_bounds__(256)
moe_gate_bwd_logits_kernel(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ ln_w,
                           const float* __restrict__ ln_b, const float* __restrict__ gate_w, const float* __restrict__ gate_b,
                           const int* __restrict__ idx, const int* __restrict__ perm, const float* __restrict__ drs, long N, int NB,
                           int E, float* __restrict__ dlogits) {
  constexpr int W = 32 * VPT;
  const long tok = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (tok >= N) return;
  const int lane = threadIdx.x & 31;
  float xv[VPT];
  ldrow<VPT, float>(x + tok * W, lane, xv);
  const float mean = stats[tok * 2], rstd = stats[tok * 2 + 1];
  for (int br = 0; br < NB; ++br) {
    float h[VPT], w[VPT], b[VPT];
    ldrow<VPT, float>(ln_w + br * W, lane, w); ldrow<VPT, float>(ln_b + br * W, lane, b);
#pragma unroll
    for (int i = 0; i < VPT; ++i) h[i] = (xv[i] - mean) * rstd * w[i] + b[i];
    float logit[16], mx = -INFINITY;
    for (int e = 0; e < E; ++e) {
      float gw[VPT];
      ldrow<VPT, float>(gate_w + (long)(br * E + e) * W, lane, gw);
      logit[e] = rowdot<VPT>(h, gw) + gate_b[br * E + e];
      mx = fmaxf(mx, logit[e]);
    }
    float s = 0.f;
    for (int e = 0; e < E; ++e) { logit[e] = expf(logit[e] - mx); s += logit[e]; }
    const int i0 = idx[(tok * NB + br) * 2], i1 = idx[(tok * NB + br) * 2 + 1];
    const float d0 = drs[perm[(tok * NB + br) * 2]] / (float)NB, d1 = drs[perm[(tok * NB + br) * 2 + 1]] / (float)NB;
    float dot = 0.f;
    for (int e = 0; e < E; ++e) {
      logit[e] /= s;                                   // probs
      dot += logit[e] * ((e == i0 ? d0 : 0.f) + (e == i1 ? d1 : 0.f));
    }
    if (lane < E) {
      const int e = lane;
      float pe = 0.f;
      for (int q = 0; q < E; ++q) if (q == e) pe = logit[q];
      dlogits[(tok * NB + br) * E + e] = pe * (((e == i0 ? d0 : 0.f) + (e == i1 ? d1 : 0.f)) - dot);
    }
  }
}
// token un-permute backward for branch br: dh[token] = d_xp[pos0] + d_xp[pos1] + dlogits[token, br, :] . Wg_br
template <int VPT, typename T>
__global__ void __launch_bounds__(256)
moe_unpermute_bwd_kernel(const T* __restrict__ dxp, const int* __restrict__ perm, const float* __restrict__ dlogits,
                         const float* __restrict__ gate_w, long N, int NB, int E, int br, T* __restrict__ dh) {
  constexpr int W = 32 * VPT;
  const long tok = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (tok >= N) return;
  const int lane = threadIdx.x & 31;
  float a[VPT], b[VPT];
  ldrow<VPT, T>(dxp + (long)perm[(tok * NB + br) * 2] * W, lane, a);
  ldrow<VPT, T>(dxp + (long)perm[(tok * NB + br) * 2 + 1] * W, lane, b);
#pragma unroll
  for (int i = 0; i < VPT; ++i) a[i] += b[i];
  for (int e = 0; e < E; ++e) {
    const float dl = dlogits[(tok * NB + br) * E + e];
    float gw[VPT];
    ldrow<VPT, float>(gate_w + (long)(br * E + e) * W, lane, gw);
#pragma unroll
    for (int i = 0; i < VPT; ++i) a[i] = fmaf(dl, gw[i], a[i]);
  }
  strow<VPT, T>(dh + tok * W, lane, a);
}
// device-side tables for the expert weight gradients (dW_g = dY_g^T X_g contracts over the rows of segment g):
// tile_k[(g * mt + i)] = {seg_off[g], max(cnt[g], 1)} - for an empty expert k0 points at `zero_row0`, a region of
// zero rows at the end of the activation buffers, so that its gradient is exactly zero.  cnt[g] = seg_off[g+1]-padding
// is taken from the scan's per-group totals (`seg_cnt`).
__global__ void moe_wgrad_tables_kernel(const int* __restrict__ seg_off, const int* __restrict__ seg_cnt, int G, int mt,
                                        int zero_row0, int* __restrict__ tile_k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G * mt) return;
  const int g = i / mt;
  const int cnt = seg_cnt[g];
  tile_k[2 * i] = cnt > 0 ? seg_off[g] : zero_row0;
  tile_k[2 * i + 1] = cnt > 0 ? cnt : 64;
}
// per-group row counts from the 128-aligned segment offsets and the permutation (count rows whose rowscale was written):
// simpler: cnt[g] = number of (token, slot) pairs routed to g = histogram of idx
__global__ void moe_group_counts_kernel(const int* __restrict__ idx, long N, int NB, int E, int* __restrict__ cnt) {
  __shared__ int h[64];
  if (threadIdx.x < 64) h[threadIdx.x] = 0;
  __syncthreads();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < N * NB * 2; i += (long)gridDim.x * blockDim.x) {
    const int br = (int)((i >> 1) % NB);
    atomicAdd(&h[br * E + idx[i]], 1);                 // integer atomics: exact, order-independent
  }
  __syncthreads();
  if (threadIdx.x < NB * E && h[threadIdx.x]) atomicAdd(&cnt[threadIdx.x], h[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------------------------
// elementwise
// ---------------------------------------------------------------------------------------------------------------
// 16-byte vectors: 8 bf16 / 4 fp32 per thread (the scalar versions moved 2 bytes per thread: 37 ms of a 400 ms step)
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Vec16<bf16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[8]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[8]) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};
__device__ __forceinline__ float act_f(float p, int act) { return act == MDM_ACT_GELU ? gelu_erf(p) : silu_f(p); }
__device__ __forceinline__ float act_df(float p, int act) {
  if (act == MDM_ACT_GELU) {
    const float cdf = 0.5f * (1.0f + erff(p * 0.70710678118654752440f));
    return cdf + p * 0.3989422804014327f * expf(-0.5f * p * p);
  }
  const float sg = 1.f / (1.f + expf(-p));
  return sg * (1.f + p * (1.f - sg));
}
template <typename T>
__global__ void act_fwd_kernel(const T* __restrict__ pre, long n, int act, T* __restrict__ out) {
  constexpr int V = Vec16<T>::N;
  const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (i + V <= n) {
    float v[V];
    Vec16<T>::load(pre + i, v);
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = act_f(v[j], act);
    Vec16<T>::store(out + i, v);
  } else {
    for (long j = i; j < n; ++j) stf<T>(out + j, act_f(ldf<T>(pre + j), act));
  }
}
template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ pre, const T* __restrict__ dy, long n, int act, T* __restrict__ dx) {
  constexpr int V = Vec16<T>::N;
  const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (i + V <= n) {
    float p[V], g[V];
    Vec16<T>::load(pre + i, p);
    Vec16<T>::load(dy + i, g);
#pragma unroll
    for (int j = 0; j < V; ++j) g[j] *= act_df(p[j], act);
    Vec16<T>::store(dx + i, g);
  } else {
    for (long j = i; j < n; ++j) stf<T>(dx + j, ldf<T>(dy + j) * act_df(ldf<T>(pre + j), act));
  }
}
template <typename TX, typename TY, typename TO>
__global__ void axpby_kernel(const TX* __restrict__ x, float a, const TY* __restrict__ y, float b, long n, TO* __restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stf<TO>(out + i, a * ldf<TX>(x + i) + (y ? b * ldf<TY>(y + i) : 0.f));
}
// GatedFusion mix backward (models/gate.py:18-19): out = g t + (1 - g) x, g = sigmoid(t + x)
__global__ void gated_mix_bwd_kernel(const float* __restrict__ t, const float* __restrict__ x, const float* __restrict__ dout,
                                     long n, float* __restrict__ dt, float* __restrict__ dx) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float tv = t[i], xv = x[i], g = 1.f / (1.f + expf(-(tv + xv))), d = dout[i];
  const float dg = d * (tv - xv) * g * (1.f - g);
  dt[i] = d * g + dg;
  dx[i] = d * (1.f - g) + dg;
}
// gradient of the masked reconstruction loss (ddpm_trainer.py:207-214): loss = sum_{b, t < len_b} mean_f (pred - target)^2
// / sum_b min(T, len_b)  ->  dpred = 2 (pred - target) [t < len_b] / (F * count)
__global__ void masked_mse_grad_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                       const int64_t* __restrict__ length, int B, int Tn, int F, float scale,
                                       float* __restrict__ dpred) {
  __shared__ float cnt_s;
  if (threadIdx.x == 0) {
    long c = 0;
    for (int b = 0; b < B; ++b) c += min((long)Tn, max(0L, (long)length[b]));
    cnt_s = (float)c;
  }
  __syncthreads();
  const long n = (long)B * Tn * F;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long row = i / F;
    const int b = (int)(row / Tn), t = (int)(row - (long)b * Tn);
    dpred[i] = t < length[b] ? scale * 2.f * (pred[i] - target[i]) / ((float)F * cnt_s) : 0.f;
  }
}
// column sums of a [M, C] matrix (bias gradients): partials [slabs, C]
template <typename T>
__global__ void colsum_any_kernel(const T* __restrict__ src, long M, int Cc, long ld, int rows_per_blk, float* __restrict__ part) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
  const long r0 = (long)blockIdx.y * rows_per_blk, r1 = min(M, r0 + rows_per_blk);
  float a = 0.f;
  if (c < Cc)
    for (long r = r0 + ty; r < r1; r += 8) a += ldf<T>(src + r * ld + c);
  red[ty][threadIdx.x & 31] = a;
  __syncthreads();
  if (ty == 0 && c < Cc) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x & 31];
    part[(long)blockIdx.y * Cc + c] = s;
  }
}
// column sums of the elementwise product of two [M, C] matrices (e.g. d cs = sum_t dx2 * (x2 - x1))
__global__ void colsum_prod_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c2,
                                   long M, int Cc, int rows_per_blk, float* __restrict__ part) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
  const long r0 = (long)blockIdx.y * rows_per_blk, r1 = min(M, r0 + rows_per_blk);
  float s = 0.f;
  if (c < Cc)
    for (long r = r0 + ty; r < r1; r += 8) s = fmaf(a[r * Cc + c], b[r * Cc + c] - (c2 ? c2[r * Cc + c] : 0.f), s);
  red[ty][threadIdx.x & 31] = s;
  __syncthreads();
  if (ty == 0 && c < Cc) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
    part[(long)blockIdx.y * Cc + c] = t;
  }
}
// 64 x 64 tiles, two elements per access on both sides (C, ld and Ks even): the 32 x 32 scalar version moved 2 bytes per
// thread per access (38 ms of a 400 ms training step)
template <typename T> struct Pair;
template <> struct Pair<float> { typedef float2 V; static __device__ __forceinline__ float2 mk(float a, float b) { return make_float2(a, b); }
  static __device__ __forceinline__ float lo(float2 v) { return v.x; } static __device__ __forceinline__ float hi(float2 v) { return v.y; } };
template <> struct Pair<bf16> { typedef __nv_bfloat162 V; static __device__ __forceinline__ __nv_bfloat162 mk(float a, float b) { return __floats2bfloat162_rn(a, b); }
  static __device__ __forceinline__ float lo(__nv_bfloat162 v) { return __low2float(v); } static __device__ __forceinline__ float hi(__nv_bfloat162 v) { return __high2float(v); } };
template <typename T>
__global__ void __launch_bounds__(256)
transpose_split_pair_kernel(const T* __restrict__ src, long M, int Cc, long ld, int Ks, T* __restrict__ dst) {
  typedef typename Pair<T>::V V2;
  __shared__ float tile[64][65];
  const int s = blockIdx.z;
  const int ml0 = blockIdx.x * 64;                               // row inside the slab
  const long m0 = (long)s * Ks + ml0;
  const int c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = ty + 8 * i, c = c0 + 2 * tx;
    const long m = m0 + r;
    float a = 0.f, b = 0.f;
    if (m < M && ml0 + r < Ks && c < Cc) {                       // Cc even: c + 1 < Cc too
      const V2 v = *reinterpret_cast<const V2*>(src + m * ld + c);
      a = Pair<T>::lo(v); b = Pair<T>::hi(v);
    }
    tile[r][2 * tx] = a; tile[r][2 * tx + 1] = b;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = ty + 8 * i, ml = ml0 + 2 * tx;
    if (c0 + c < Cc && ml < Ks)
      *reinterpret_cast<V2*>(dst + ((long)s * Cc + c0 + c) * Ks + ml) = Pair<T>::mk(tile[2 * tx][c], tile[2 * tx + 1][c]);
  }
}
// 32 x 32 tiled transpose with optional split in S token slabs (see misc.cu: transpose_split_kernel), any dtype
template <typename T>
__global__ void transpose_split_any_kernel(const T* __restrict__ src, long M, int Cc, long ld, int Ks, T* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int s = blockIdx.z;
  const long m0 = (long)s * Ks + (long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty + 8 * i;
    const long m = m0 + r;
    const bool ok = m < M && (blockIdx.x * 32 + r) < Ks && c0 + tx < Cc;
    tile[r][tx] = ok ? ldf<T>(src + m * ld + c0 + tx) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = ty + 8 * i;
    const int ml = blockIdx.x * 32 + tx;
    if (c0 + c < Cc && ml < Ks) stf<T>(dst + ((long)s * Cc + c0 + c) * Ks + ml, tile[tx][c]);
  }
}
// out[g, c] = sum of src[r, c] over rows [seg_off[g], seg_off[g] + seg_cnt[g]); blockIdx.z = row slab of the segment
// (gridDim.z slabs -> out is [slabs, G, C] partials; one slab: the sums themselves)
template <typename T>
__global__ void seg_colsum_any_kernel(const T* __restrict__ src, int Cc, const int* __restrict__ seg_off,
                                      const int* __restrict__ seg_cnt, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int g = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
  const int cnt = seg_cnt[g], per = (cnt + gridDim.z - 1) / gridDim.z;
  const long r0 = (long)seg_off[g] + (long)blockIdx.z * per, r1 = min((long)seg_off[g] + cnt, r0 + per);
  out += (long)blockIdx.z * gridDim.y * Cc;
  float a = 0.f;
  if (c < Cc)
    for (long r = r0 + ty; r < r1; r += 8) a += ldf<T>(src + r * Cc + c);
  red[ty][threadIdx.x & 31] = a;
  __syncthreads();
  if (ty == 0 && c < Cc) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x & 31];
    out[(long)g * Cc + c] = s;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// optimizer: clip_grad_norm_(max_norm) + Adam (ddpm_trainer.py:228-244, torch.optim.Adam defaults)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, long n, float* __restrict__ part) {
  __shared__ float red[8];
  float s = 0.f;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) s = fmaf(g[i], g[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    part[blockIdx.x] = t;
  }
}
// norm = sqrt(sum of partials); coef = min(1, max_norm / (norm + 1e-6))   (torch.nn.utils.clip_grad_norm_)
__global__ void clip_coef_kernel(const float* __restrict__ part, int n, float max_norm, float* __restrict__ out) {
  if (threadIdx.x || blockIdx.x) return;
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += (double)part[i];
  const float norm = (float)sqrt(s);
  out[0] = norm;
  out[1] = max_norm > 0.f ? fminf(1.f, max_norm / (norm + 1e-6f)) : 1.f;
}
__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long n,
                            float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, const float* __restrict__ coef,
                            bf16* __restrict__ mirror) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float c = coef ? coef[1] : 1.f;
  const float gi = g[i] * c;
  g[i] = gi;                                            // clip_grad_norm_ scales .grad in place
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  const float pn = p[i] - (lr / bc1) * (mi / denom);
  p[i] = pn;
  if (mirror) mirror[i] = __float2bfloat16_rn(pn);      // the bf16 operand copy the GEMMs read: refreshed in the same pass
}

inline unsigned blocks_for(long n, int per) { return (unsigned)((n + per - 1) / per); }

}  // namespace

#define TRY(x) do { if (!(x)) return MDM_ERR_ARG; } while (0)
#define DONE() return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA
#define ST(stream) reinterpret_cast<cudaStream_t>(stream)
#define HD_SWITCH(hd, ...)                                   \
  switch (hd) {                                              \
    case 32: { constexpr int V = 1; __VA_ARGS__; break; }    \
    case 64: { constexpr int V = 2; __VA_ARGS__; break; }    \
    case 128: { constexpr int V = 4; __VA_ARGS__; break; }   \
    case 256: { constexpr int V = 8; __VA_ARGS__; break; }   \
    case 512: { constexpr int V = 16; __VA_ARGS__; break; }  \
    case 1024: { constexpr int V = 32; __VA_ARGS__; break; } \
    default: return MDM_ERR_UNSUPPORTED;                     \
  }

extern "C" MDM_API int mdm_bgemm(const MdmBgemm* g, void* stream) {
  TRY(g && g->A && g->B && g->C && g->Z1 > 0 && g->Z2 > 0 && g->M > 0 && g->N > 0 && g->K >= 0);
  dim3 grid((g->N + BT - 1) / BT, (g->M + BT - 1) / BT, g->Z1 * g->Z2);
  cudaStream_t st = ST(stream);
#define BG(TA_, TB_, TC_)                                                     \
  do {                                                                        \
    if (g->tensor_cores) bgemm_tc_kernel<TA_, TB_, TC_><<<grid, 256, 0, st>>>(*g); \
    else bgemm_kernel<TA_, TB_, TC_><<<grid, 256, 0, st>>>(*g);               \
  } while (0)
  const int key = g->a_dt * 4 + g->b_dt * 2 + g->c_dt;
  switch (key) {
    case 0: BG(float, float, float); break;
    case 1: BG(float, float, bf16); break;
    case 2: BG(float, bf16, float); break;
    case 3: BG(float, bf16, bf16); break;
    case 4: BG(bf16, float, float); break;
    case 5: BG(bf16, float, bf16); break;
    case 6: BG(bf16, bf16, float); break;
    case 7: BG(bf16, bf16, bf16); break;
    default: return MDM_ERR_ARG;
  }
#undef BG
  DONE();
}

extern "C" MDM_API int mdm_fa_prep(const void* qkv, int dt, const float* nw, const float* nb, int B, int H, int T, int hd,
                                   float* qh, float* kh, float* vn, void* stream) {
  TRY(qkv && nw && nb && qh && kh && vn);
  const unsigned grid = blocks_for((long)B * T * H, 8);
  HD_SWITCH(hd, {
    if (dt == MDM_F32) fa_prep_kernel<V, float><<<grid, 256, 0, ST(stream)>>>(reinterpret_cast<const float*>(qkv), nw, nb, B, H, T, qh, kh, vn);
    else fa_prep_kernel<V, bf16><<<grid, 256, 0, ST(stream)>>>(reinterpret_cast<const bf16*>(qkv), nw, nb, B, H, T, qh, kh, vn);
  });
  DONE();
}
extern "C" MDM_API int mdm_fa_feat(const float* uq, const float* uk, const int64_t* length, int shift, int B, int H, int T, int M,
                                   float* qp, float* kp, void* stream) {
  TRY(uq && uk && qp && kp);
  const long total = (long)B * H * T * M;
  fa_feat_kernel<<<blocks_for(total, 256), 256, 0, ST(stream)>>>(uq, uk, length, shift, H, T, M, total, qp, kp);
  DONE();
}
extern "C" MDM_API int mdm_fa_out_bwd(float* o, const float* qp, const float* kp, const void* dout, int dt, const float* nw,
                                      int B, int H, int T, int hd, float* dden, float* part, int* n_parts, void* stream) {
  TRY(n_parts);
  const unsigned grid = min(blocks_for((long)B * H * T, 8), 2048u);
  *n_parts = (int)grid;
  if (!o) return MDM_OK;                                 // size query
  TRY(qp && kp && dout && nw && dden && part);
  HD_SWITCH(hd, {
    if (dt == MDM_F32) fa_out_bwd_kernel<V, float><<<grid, 256, 0, ST(stream)>>>(o, qp, kp, reinterpret_cast<const float*>(dout), nw, B, H, T, dden, part);
    else fa_out_bwd_kernel<V, bf16><<<grid, 256, 0, ST(stream)>>>(o, qp, kp, reinterpret_cast<const bf16*>(dout), nw, B, H, T, dden, part);
  });
  DONE();
}
extern "C" MDM_API int mdm_fa_out_fwd(const float* o, const float* qp, const float* kp, const float* nw, const float* nb, int B,
                                      int H, int T, int hd, void* out, int dt, void* stream) {
  TRY(o && qp && kp && nw && nb && out);
  const unsigned grid = blocks_for((long)B * H * T, 8);
  HD_SWITCH(hd, {
    if (dt == MDM_F32) fa_out_fwd_kernel<V, float><<<grid, 256, 0, ST(stream)>>>(o, qp, kp, nw, nb, B, H, T, reinterpret_cast<float*>(out));
    else fa_out_fwd_kernel<V, bf16><<<grid, 256, 0, ST(stream)>>>(o, qp, kp, nw, nb, B, H, T, reinterpret_cast<bf16*>(out));
  });
  DONE();
}
extern "C" MDM_API int mdm_fa_feat_bwd(const float* uq, const float* uk, const float* qp, const float* kp, const float* dden,
                                       int B, int H, int T, int M, float* dqp, float* dkp, void* stream) {
  TRY(uq && uk && qp && kp && dden && dqp && dkp);
  const long total = (long)B * H * T * M;
  fa_feat_bwd_kernel<<<blocks_for(total, 256), 256, 0, ST(stream)>>>(uq, uk, qp, kp, dden, M, total, dqp, dkp);
  DONE();
}
extern "C" MDM_API int mdm_fa_prep_bwd(const void* qkv, int dt, const float* nw, const float* nb, int B, int H, int T, int hd,
                                       const float* dqh, const float* dkh, const float* dvn, void* dqkv, float* part,
                                       int* n_parts, void* stream) {
  TRY(n_parts);
  const unsigned grid = min(blocks_for((long)B * T * H, 8), 2048u);
  *n_parts = (int)grid;
  if (!qkv) return MDM_OK;
  TRY(nw && nb && dqh && dkh && dvn && dqkv && part);
  HD_SWITCH(hd, {
    if (dt == MDM_F32) fa_prep_bwd_kernel<V, float><<<grid, 256, 0, ST(stream)>>>(reinterpret_cast<const float*>(qkv), nw, nb, B, H, T, dqh, dkh, dvn, reinterpret_cast<float*>(dqkv), part);
    else fa_prep_bwd_kernel<V, bf16><<<grid, 256, 0, ST(stream)>>>(reinterpret_cast<const bf16*>(qkv), nw, nb, B, H, T, dqh, dkh, dvn, reinterpret_cast<bf16*>(dqkv), part);
  });
  DONE();
}

extern "C" MDM_API int mdm_head_softmax(const void* q, int dt, int B, int H, int T, int hd, float* P, void* stream) {
  TRY(q && P);
  const unsigned grid = blocks_for((long)B * T * H, 8);
  HD_SWITCH(hd, {
    if (dt == MDM_F32) head_softmax_kernel<V, float><<<grid, 256, 0, ST(stream)>>>(reinterpret_cast<const float*>(q), B, H, T, P);
    else head_softmax_kernel<V, bf16><<<grid, 256, 0, ST(stream)>>>(reinterpret_cast<const bf16*>(q), B, H, T, P);
  });
  DONE();
}
extern "C" MDM_API int mdm_head_softmax_bwd(const float* P, const float* dP, int B, int H, int T, int hd, void* dq, int dt,
                                            void* stream) {
  TRY(P && dP && dq);
  const unsigned grid = blocks_for((long)B * T * H, 8);
  HD_SWITCH(hd, {
    if (dt == MDM_F32) head_softmax_bwd_kernel<V, float><<<grid, 256, 0, ST(stream)>>>(P, dP, B, H, T, reinterpret_cast<float*>(dq));
    else head_softmax_bwd_kernel<V, bf16><<<grid, 256, 0, ST(stream)>>>(P, dP, B, H, T, reinterpret_cast<bf16*>(dq));
  });
  DONE();
}
extern "C" MDM_API int mdm_key_softmax(float* S, const int* nt, int B, int H, int T, int NK, void* stream) {
  TRY(S && NK > 0 && NK <= 96);
  const long rows = (long)B * H * T;
  key_softmax_kernel<<<blocks_for(rows, 8), 256, 0, ST(stream)>>>(S, nt, H, T, NK, rows);
  DONE();
}
extern "C" MDM_API int mdm_key_softmax_bwd(const float* P, float* dP, int B, int H, int T, int NK, void* stream) {
  TRY(P && dP && NK > 0 && NK <= 96);
  const long rows = (long)B * H * T;
  key_softmax_bwd_kernel<<<blocks_for(rows, 8), 256, 0, ST(stream)>>>(P, dP, NK, rows);
  DONE();
}
extern "C" MDM_API int mdm_col_softmax(const void* k, int dt, const int* nt, int B, int Nt, int C, float* Ks, void* stream) {
  TRY(k && Ks);
  const long total = (long)B * C;
  if (dt == MDM_F32) col_softmax_kernel<float><<<blocks_for(total, 128), 128, 0, ST(stream)>>>(reinterpret_cast<const float*>(k), nt, Nt, C, total, Ks);
  else col_softmax_kernel<bf16><<<blocks_for(total, 128), 128, 0, ST(stream)>>>(reinterpret_cast<const bf16*>(k), nt, Nt, C, total, Ks);
  DONE();
}
extern "C" MDM_API int mdm_col_softmax_bwd(const float* Ks, const float* dKs, int B, int Nt, int C, void* dk, int dt, void* stream) {
  TRY(Ks && dKs && dk);
  const long total = (long)B * C;
  if (dt == MDM_F32) col_softmax_bwd_kernel<float><<<blocks_for(total, 128), 128, 0, ST(stream)>>>(Ks, dKs, Nt, C, total, reinterpret_cast<float*>(dk));
  else col_softmax_bwd_kernel<bf16><<<blocks_for(total, 128), 128, 0, ST(stream)>>>(Ks, dKs, Nt, C, total, reinterpret_cast<bf16*>(dk));
  DONE();
}

extern "C" MDM_API int mdm_moe_combine_sum(const void* z, int dt, const float* rowscale, const int* perm, long N, int D, int NBK,
                                           void* m, void* stream) {
  TRY(z && rowscale && perm && m);
  const unsigned grid = blocks_for(N, 8);
  HD_SWITCH(D, {
    if (dt == MDM_F32) moe_combine_sum_kernel<V, float><<<grid, 256, 0, ST(stream)>>>(reinterpret_cast<const float*>(z), rowscale, perm, N, NBK, reinterpret_cast<float*>(m));
    else moe_combine_sum_kernel<V, bf16><<<grid, 256, 0, ST(stream)>>>(reinterpret_cast<const bf16*>(z), rowscale, perm, N, NBK, reinterpret_cast<bf16*>(m));
  });
  DONE();
}
extern "C" MDM_API int mdm_moe_combine_bwd(const void* z, int dt, const float* rowscale, const int* perm, long N, int D, int NBK,
                                           const void* dm, void* dz, float* drs, void* stream) {
  TRY(z && rowscale && perm && dm && dz && drs);
  const unsigned grid = blocks_for(N, 8);
  HD_SWITCH(D, {
    if (dt == MDM_F32) moe_combine_bwd_kernel<V, float><<<grid, 256, 0, ST(stream)>>>(reinterpret_cast<const float*>(z), rowscale, perm, N, NBK, reinterpret_cast<const float*>(dm), reinterpret_cast<float*>(dz), drs);
    else moe_combine_bwd_kernel<V, bf16><<<grid, 256, 0, ST(stream)>>>(reinterpret_cast<const bf16*>(z), rowscale, perm, N, NBK, reinterpret_cast<const bf16*>(dm), reinterpret_cast<bf16*>(dz), drs);
  });
  DONE();
}
extern "C" MDM_API int mdm_moe_gate_bwd_logits(const float* x, const float* stats, const float* ln_w, const float* ln_b,
                                               const float* gate_w, const float* gate_b, const int* idx, const int* perm,
                                               const float* drs, long N, int D, int NB, int E, float* dlogits, void* stream) {
  TRY(x && stats && ln_w && ln_b && gate_w && gate_b && idx && perm && drs && dlogits && E <= 16);
  const unsigned grid = blocks_for(N, 8);
  HD_SWITCH(D, { moe_gate_bwd_logits_kernel<V><<<grid, 256, 0, ST(stream)>>>(x, stats, ln_w, ln_b, gate_w, gate_b, idx, perm, drs, N, NB, E, dlogits); });
  DONE();
}
extern "C" MDM_API int mdm_moe_unpermute_bwd(const void* dxp, int dt, const int* perm, const float* dlogits, const float* gate_w,
                                             long N, int D, int NB, int E, int br, void* dh, void* stream) {
  TRY(dxp && perm && dlogits && gate_w && dh);
  const unsigned grid = blocks_for(N, 8);
  HD_SWITCH(D, {
    if (dt == MDM_F32) moe_unpermute_bwd_kernel<V, float><<<grid, 256, 0, ST(stream)>>>(reinterpret_cast<const float*>(dxp), perm, dlogits, gate_w, N, NB, E, br, reinterpret_cast<float*>(dh));
    else moe_unpermute_bwd_kernel<V, bf16><<<grid, 256, 0, ST(stream)>>>(reinterpret_cast<const bf16*>(dxp), perm, dlogits, gate_w, N, NB, E, br, reinterpret_cast<bf16*>(dh));
  });
  DONE();
}
extern "C" MDM_API int mdm_moe_wgrad_tables(const int* seg_off, const int* idx, long N, int NB, int E, int mt_up, int mt_down,
                                            int zero_row0, int* seg_cnt, int* tile_k_up, int* tile_k_down, void* stream) {
  TRY(seg_off && idx && seg_cnt && tile_k_up && tile_k_down && NB * E <= 64);
  cudaStream_t st = ST(stream);
  const int G = NB * E;
  cudaMemsetAsync(seg_cnt, 0, sizeof(int) * G, st);
  moe_group_counts_kernel<<<64, 256, 0, st>>>(idx, N, NB, E, seg_cnt);
  moe_wgrad_tables_kernel<<<blocks_for((long)G * mt_up, 128), 128, 0, st>>>(seg_off, seg_cnt, G, mt_up, zero_row0, tile_k_up);
  moe_wgrad_tables_kernel<<<blocks_for((long)G * mt_down, 128), 128, 0, st>>>(seg_off, seg_cnt, G, mt_down, zero_row0, tile_k_down);
  DONE();
}

extern "C" MDM_API int mdm_act_fwd(const void* pre, int dt, long n, int act, void* out, void* stream) {
  TRY(pre && out && (act == MDM_ACT_GELU || act == MDM_ACT_SILU));
  if ((reinterpret_cast<uintptr_t>(pre) | reinterpret_cast<uintptr_t>(out)) & 15) return MDM_ERR_ARG;   // 16-byte vectors
  if (dt == MDM_F32) act_fwd_kernel<float><<<blocks_for(n, 256 * 4), 256, 0, ST(stream)>>>(reinterpret_cast<const float*>(pre), n, act, reinterpret_cast<float*>(out));
  else act_fwd_kernel<bf16><<<blocks_for(n, 256 * 8), 256, 0, ST(stream)>>>(reinterpret_cast<const bf16*>(pre), n, act, reinterpret_cast<bf16*>(out));
  DONE();
}
extern "C" MDM_API int mdm_act_bwd(const void* pre, const void* dy, int dt, long n, int act, void* dx, void* stream) {
  TRY(pre && dy && dx && (act == MDM_ACT_GELU || act == MDM_ACT_SILU));
  if ((reinterpret_cast<uintptr_t>(pre) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) return MDM_ERR_ARG;
  if (dt == MDM_F32) act_bwd_kernel<float><<<blocks_for(n, 256 * 4), 256, 0, ST(stream)>>>(reinterpret_cast<const float*>(pre), reinterpret_cast<const float*>(dy), n, act, reinterpret_cast<float*>(dx));
  else act_bwd_kernel<bf16><<<blocks_for(n, 256 * 8), 256, 0, ST(stream)>>>(reinterpret_cast<const bf16*>(pre), reinterpret_cast<const bf16*>(dy), n, act, reinterpret_cast<bf16*>(dx));
  DONE();
}
extern "C" MDM_API int mdm_axpby(const void* x, int x_dt, float a, const void* y, int y_dt, float b, long n, void* out, int out_dt,
                                 void* stream) {
  TRY(x && out);
  const unsigned grid = blocks_for(n, 256);
  cudaStream_t st = ST(stream);
#define AX(TX_, TY_, TO_) axpby_kernel<TX_, TY_, TO_><<<grid, 256, 0, st>>>(reinterpret_cast<const TX_*>(x), a, reinterpret_cast<const TY_*>(y), b, n, reinterpret_cast<TO_*>(out))
  const int key = x_dt * 4 + (y ? y_dt : 0) * 2 + out_dt;
  switch (key) {
    case 0: AX(float, float, float); break;
    case 1: AX(float, float, bf16); break;
    case 2: AX(float, bf16, float); break;
    case 3: AX(float, bf16, bf16); break;
    case 4: AX(bf16, float, float); break;
    case 5: AX(bf16, float, bf16); break;
    case 6: AX(bf16, bf16, float); break;
    case 7: AX(bf16, bf16, bf16); break;
    default: return MDM_ERR_ARG;
  }
#undef AX
  DONE();
}
extern "C" MDM_API int mdm_gated_mix_bwd(const float* t, const float* x, const float* dout, long n, float* dt_, float* dx,
                                         void* stream) {
  TRY(t && x && dout && dt_ && dx);
  gated_mix_bwd_kernel<<<blocks_for(n, 256), 256, 0, ST(stream)>>>(t, x, dout, n, dt_, dx);
  DONE();
}
extern "C" MDM_API int mdm_masked_mse_grad(const float* pred, const float* target, const int64_t* length, int B, int T, int F,
                                           float scale, float* dpred, void* stream) {
  TRY(pred && target && length && dpred);
  masked_mse_grad_kernel<<<296, 256, 0, ST(stream)>>>(pred, target, length, B, T, F, scale, dpred);
  DONE();
}
extern "C" MDM_API int mdm_colsum(const void* src, int dt, long M, int C, long ld, int slabs, float* part, void* stream) {
  TRY(src && part && slabs > 0);
  const int rpb = (int)((M + slabs - 1) / slabs);
  dim3 grid((C + 31) / 32, slabs);
  if (dt == MDM_F32) colsum_any_kernel<float><<<grid, 256, 0, ST(stream)>>>(reinterpret_cast<const float*>(src), M, C, ld, rpb, part);
  else colsum_any_kernel<bf16><<<grid, 256, 0, ST(stream)>>>(reinterpret_cast<const bf16*>(src), M, C, ld, rpb, part);
  DONE();
}
extern "C" MDM_API int mdm_colsum_prod(const float* a, const float* b, const float* c, long M, int C, int slabs, float* part,
                                       void* stream) {
  TRY(a && b && part && slabs > 0);
  const int rpb = (int)((M + slabs - 1) / slabs);
  dim3 grid((C + 31) / 32, slabs);
  colsum_prod_kernel<<<grid, 256, 0, ST(stream)>>>(a, b, c, M, C, rpb, part);
  DONE();
}
extern "C" MDM_API int mdm_transpose_split(const void* src, int dt, long M, int C, long ld, int S, int Ks, void* dst, void* stream) {
  TRY(src && dst && S > 0 && Ks > 0);
  const int esz = dt == MDM_F32 ? 4 : 2;
  if (!((C | Ks) & 1) && !(ld & 1) && !((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & (2 * esz - 1))) {
    dim3 g2((Ks + 63) / 64, (C + 63) / 64, S);
    if (dt == MDM_F32) transpose_split_pair_kernel<float><<<g2, 256, 0, ST(stream)>>>(reinterpret_cast<const float*>(src), M, C, ld, Ks, reinterpret_cast<float*>(dst));
    else transpose_split_pair_kernel<bf16><<<g2, 256, 0, ST(stream)>>>(reinterpret_cast<const bf16*>(src), M, C, ld, Ks, reinterpret_cast<bf16*>(dst));
    DONE();
  }
  dim3 grid((Ks + 31) / 32, (C + 31) / 32, S);
  if (dt == MDM_F32) transpose_split_any_kernel<float><<<grid, 256, 0, ST(stream)>>>(reinterpret_cast<const float*>(src), M, C, ld, Ks, reinterpret_cast<float*>(dst));
  else transpose_split_any_kernel<bf16><<<grid, 256, 0, ST(stream)>>>(reinterpret_cast<const bf16*>(src), M, C, ld, Ks, reinterpret_cast<bf16*>(dst));
  DONE();
}
extern "C" MDM_API int mdm_seg_colsum(const void* src, int dt, int C, const int* seg_off, const int* seg_cnt, int G, int slabs,
                                      float* out, void* stream) {
  TRY(src && seg_off && seg_cnt && out && slabs >= 1);
  dim3 grid((C + 31) / 32, G, slabs);
  if (dt == MDM_F32) seg_colsum_any_kernel<float><<<grid, 256, 0, ST(stream)>>>(reinterpret_cast<const float*>(src), C, seg_off, seg_cnt, out);
  else seg_colsum_any_kernel<bf16><<<grid, 256, 0, ST(stream)>>>(reinterpret_cast<const bf16*>(src), C, seg_off, seg_cnt, out);
  DONE();
}

extern "C" MDM_API int mdm_grad_clip_coef(const float* g, long n, float max_norm, float* part, int n_part, float* norm_coef,
                                          void* stream) {
  TRY(g && part && norm_coef && n_part > 0);
  sumsq_kernel<<<n_part, 256, 0, ST(stream)>>>(g, n, part);
  clip_coef_kernel<<<1, 32, 0, ST(stream)>>>(part, n_part, max_norm, norm_coef);
  DONE();
}
extern "C" MDM_API int mdm_adam_step(float* p, float* g, float* m, float* v, long n, float lr, float beta1, float beta2, float eps,
                                     int step, const float* norm_coef, void* bf16_mirror, void* stream) {
  TRY(p && g && m && v && step >= 1);
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = sqrtf(1.f - powf(beta2, (float)step));
  adam_kernel<<<blocks_for(n, 256), 256, 0, ST(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, bc2, norm_coef,
                                                          reinterpret_cast<bf16*>(bf16_mirror));
  DONE();
}
