// Warp-per-row register math shared by the row pipeline, the MoE gate/permute/combine kernels.
// A row of D = 32*VPT elements is spread over a warp in 4-element chunks:
//   lane owns columns (j*32 + lane)*4 + c,  j < VPT/4, c < 4   (coalesced 16-byte accesses).
#pragma once
#include "common.cuh"

template <int VPT> __device__ __forceinline__ int row_col(int lane, int i) {
  return ((i >> 2) * 32 + lane) * 4 + (i & 3);
}

template <int VPT, typename T>
__device__ __forceinline__ void load_row(const T* __restrict__ p, int lane, float (&v)[VPT]);

template <int VPT, typename T>
__device__ __forceinline__ void store_row(T* __restrict__ p, int lane, const float (&v)[VPT]);

template <int VPT>
__device__ __forceinline__ void load_row_f32(const float* __restrict__ p, int lane, float (&v)[VPT]) {
#pragma unroll
  for (int j = 0; j < VPT / 4; ++j) {
    const float4 t = *reinterpret_cast<const float4*>(p + (j * 32 + lane) * 4);
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}
template <int VPT>
__device__ __forceinline__ void load_row_bf16(const bf16* __restrict__ p, int lane, float (&v)[VPT]) {
#pragma unroll
  for (int j = 0; j < VPT / 4; ++j) {
    const uint2 t = *reinterpret_cast<const uint2*>(p + (j * 32 + lane) * 4);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
    v[4 * j] = __low2float(a); v[4 * j + 1] = __high2float(a);
    v[4 * j + 2] = __low2float(b); v[4 * j + 3] = __high2float(b);
  }
}
template <int VPT>
__device__ __forceinline__ void store_row_f32(float* __restrict__ p, int lane, const float (&v)[VPT]) {
#pragma unroll
  for (int j = 0; j < VPT / 4; ++j)
    *reinterpret_cast<float4*>(p + (j * 32 + lane) * 4) =
        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
template <int VPT>
__device__ __forceinline__ void store_row_bf16(bf16* __restrict__ p, int lane, const float (&v)[VPT]) {
#pragma unroll
  for (int j = 0; j < VPT / 4; ++j) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[4 * j], v[4 * j + 1]);
    const __nv_bfloat162 b = __floats2bfloat162_rn(v[4 * j + 2], v[4 * j + 3]);
    uint2 t;
    t.x = *reinterpret_cast<const uint32_t*>(&a);
    t.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p + (j * 32 + lane) * 4) = t;
  }
}

template <int VPT, typename T> struct RowIO;
template <int VPT> struct RowIO<VPT, float> {
  static __device__ __forceinline__ void load(const float* p, int lane, float (&v)[VPT]) { load_row_f32<VPT>(p, lane, v); }
  static __device__ __forceinline__ void store(float* p, int lane, const float (&v)[VPT]) { store_row_f32<VPT>(p, lane, v); }
};
template <int VPT> struct RowIO<VPT, bf16> {
  static __device__ __forceinline__ void load(const bf16* p, int lane, float (&v)[VPT]) { load_row_bf16<VPT>(p, lane, v); }
  static __device__ __forceinline__ void store(bf16* p, int lane, const float (&v)[VPT]) { store_row_bf16<VPT>(p, lane, v); }
};
template <int VPT, typename T>
__device__ __forceinline__ void load_row(const T* __restrict__ p, int lane, float (&v)[VPT]) {
  RowIO<VPT, T>::load(p, lane, v);
}
template <int VPT, typename T>
__device__ __forceinline__ void store_row(T* __restrict__ p, int lane, const float (&v)[VPT]) {
  RowIO<VPT, T>::store(p, lane, v);
}

// mean / rstd of a row (biased variance, eps = 1e-5: nn.LayerNorm defaults), two-pass in registers.
template <int VPT>
__device__ __forceinline__ void row_stats(const float (&v)[VPT], int D, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) s += v[i];
  mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const float d = v[i] - mean;
    q = fmaf(d, d, q);
  }
  rstd = rsqrtf(warp_sum(q) / (float)D + 1e-5f);
}

template <int VPT>
__device__ __forceinline__ void affine_row(float (&v)[VPT], float mean, float rstd,
                                           const float* __restrict__ w, const float* __restrict__ b,
                                           int lane) {
#pragma unroll
  for (int j = 0; j < VPT / 4; ++j) {
    const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + (j * 32 + lane) * 4));
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(b + (j * 32 + lane) * 4));
    v[4 * j] = (v[4 * j] - mean) * rstd * w4.x + b4.x;
    v[4 * j + 1] = (v[4 * j + 1] - mean) * rstd * w4.y + b4.y;
    v[4 * j + 2] = (v[4 * j + 2] - mean) * rstd * w4.z + b4.z;
    v[4 * j + 3] = (v[4 * j + 3] - mean) * rstd * w4.w + b4.w;
  }
}

template <int VPT>
__device__ __forceinline__ void layernorm_row(float (&v)[VPT], const float* __restrict__ w,
                                              const float* __restrict__ b, int lane, int D) {
  float mean, rstd;
  row_stats<VPT>(v, D, mean, rstd);
  affine_row<VPT>(v, mean, rstd, w, b, lane);
}

// F.normalize(v, dim=-1) * sqrt(D)   (models/fast_attention.py:172)
template <int VPT>
__device__ __forceinline__ void l2norm_row(float (&v)[VPT], int D) {
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) q = fmaf(v[i], v[i], q);
  const float denom = fmaxf(sqrtf(warp_sum(q)), 1e-12f);
  const float s = sqrtf((float)D) / denom;   // one IEEE divide per row
#pragma unroll
  for (int i = 0; i < VPT; ++i) v[i] = v[i] * s;
}

// v * (1 + scale) + shift, film = [scale(D) | shift(D)]   (models/stylization.py:27-29)
template <int VPT>
__device__ __forceinline__ void film_row(float (&v)[VPT], const float* __restrict__ film, int lane, int D) {
#pragma unroll
  for (int j = 0; j < VPT / 4; ++j) {
    const float4 sc = *reinterpret_cast<const float4*>(film + (j * 32 + lane) * 4);
    const float4 sh = *reinterpret_cast<const float4*>(film + D + (j * 32 + lane) * 4);
    v[4 * j] = v[4 * j] * (1.f + sc.x) + sh.x;
    v[4 * j + 1] = v[4 * j + 1] * (1.f + sc.y) + sh.y;
    v[4 * j + 2] = v[4 * j + 2] * (1.f + sc.z) + sh.z;
    v[4 * j + 3] = v[4 * j + 3] * (1.f + sc.w) + sh.w;
  }
}
