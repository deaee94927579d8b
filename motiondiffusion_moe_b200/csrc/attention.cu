// Attention cores of MoEExtendedDecoderLayer (reference: models/fast_attention.py).
//   mdm_fastattn        FastAttention.forward :29-92  (Performer random-feature linear attention)
//   mdm_lincross_ctx    LinearTemporalCrossAttention :249-252 (text side, step-invariant)
//   mdm_lincross_apply  LinearTemporalCrossAttention :248,253 (motion side)
//   mdm_softmax_cross   MemoryEfficientCrossAttentionBlock :313-325
// fp32-compute version: one CTA per (sequence, head), the [M x hd] state lives in registers
// (one column per thread).  Activations are read/written as fp32 or bf16.
#include <stdlib.h>
#include "common.cuh"

// tensor-core (mma.sync) kernel for the bf16 path, attention_tc.cu
int mdm_fastattn_tc(const void* qkv, const float* P, const float* norm_w, const float* norm_b,
                    const int64_t* length, int length_shift, int B, int H, int T, int hd, int M, void* out,
                    const int* seq_order, const void* Pt_bf16, cudaStream_t st);

// tcgen05 kernel (attention_umma.cu): hd == M == 128, T <= 256, needs the pre-transposed bf16 projection matrix
int mdm_fastattn_umma(const void* qkv, const void* Pt_bf16, const float* norm_w, const float* norm_b,
                      const int64_t* length, int length_shift, int B, int H, int T, int hd, int M, void* out,
                      const int* seq_order, cudaStream_t st);

int mdm_lincross_apply_umma(const void* q, const void* ctxT_bf16, int B, int T, int H, int hd, void* y, cudaStream_t st);
int mdm_lincross_apply_tc(const void* q, const float* ctx, int B, int T, int H, int hd, void* y, cudaStream_t st);
int mdm_softmax_cross_umma(const void* q, const void* k, const void* v, const int* nt, int B, int T, int Nt_max, int H,
                           int hd, float scale, void* o, cudaStream_t st);
int mdm_softmax_cross_tc(const void* q, const void* k, const void* v, const int* nt, int B, int T, int Nt_max, int H,
                         int hd, float scale, void* o, cudaStream_t st);

namespace {

constexpr int AT = 128;  // threads per CTA == max(hd, M)
constexpr int TC = 4;    // tokens per chunk == warps per CTA

// LayerNorm over hd (shared affine, eps 1e-5) by one warp; x in smem, in place.
__device__ __forceinline__ void warp_ln(float* x, int hd, const float* __restrict__ w,
                                        const float* __restrict__ b, int lane) {
  float s = 0.f;
  for (int i = lane; i < hd; i += 32) s += x[i];
  const float mean = warp_sum(s) / (float)hd;
  float q = 0.f;
  for (int i = lane; i < hd; i += 32) { const float d = x[i] - mean; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(warp_sum(q) / (float)hd + 1e-5f);
  for (int i = lane; i < hd; i += 32) x[i] = (x[i] - mean) * rstd * w[i] + b[i];
}
// F.normalize(x, dim=-1): x / max(||x||, 1e-12)
__device__ __forceinline__ void warp_l2(float* x, int hd, int lane) {
  float q = 0.f;
  for (int i = lane; i < hd; i += 32) q = fmaf(x[i], x[i], q);
  const float denom = fmaxf(sqrtf(warp_sum(q)), 1e-12f);
  for (int i = lane; i < hd; i += 32) x[i] = x[i] / denom;
}

template <typename T>
__global__ void __launch_bounds__(AT)
fastattn_kernel(const T* __restrict__ qkv, const float* __restrict__ P, const float* __restrict__ nw,
                const float* __restrict__ nb, const int64_t* __restrict__ length, int length_shift,
                int H, int Tn, int hd, int M, T* __restrict__ out) {
  extern __shared__ float sm[];
  float* Ps = sm;                  // [hd][M]
  float* xa = Ps + hd * M;         // [TC][hd]  q (phase 2) / k (phase 1), normalised in place
  float* xb = xa + TC * hd;        // [TC][hd]  v (phase 1) / k (phase 2)
  float* fa = xb + TC * hd;        // [M][TC]   feature map A (transposed for float4 broadcast)
  float* fb = fa + M * TC;         // [M][TC]   feature map B
  float* ob = fb + M * TC;         // [TC][hd]  un-normalised output rows
  float* red = ob + TC * hd;       // [TC][4]   denominator partials
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = H * hd;
  const long len = length ? (long)(length[b] >> length_shift) : (long)Tn;
  for (int i = tid; i < hd * M; i += AT) Ps[i] = P[i];
  float kv[AT];  // column `tid` of kv[M][hd]; only the first M entries are used
#pragma unroll
  for (int m = 0; m < AT; ++m) kv[m] = 0.f;
  __syncthreads();

  // ---------------- phase 1: kv[m][n] = sum_t k_proj[t][m] * v[t][n]
  for (int t0 = 0; t0 < Tn; t0 += TC) {
    const int t = t0 + warp;  // one warp per token of the chunk
    if (t < Tn) {
      const T* row = qkv + ((long)(b * Tn + t)) * 3 * D + h * hd;
      for (int i = lane; i < hd; i += 32) {
        xa[warp * hd + i] = to_f<T>(row[D + i]) * 0.1f;
        xb[warp * hd + i] = to_f<T>(row[2 * D + i]) * 0.1f;
      }
      __syncwarp();
      warp_ln(xa + warp * hd, hd, nw, nb, lane);
      warp_ln(xb + warp * hd, hd, nw, nb, lane);
      __syncwarp();
      warp_l2(xa + warp * hd, hd, lane);
    } else {
      for (int i = lane; i < hd; i += 32) { xa[warp * hd + i] = 0.f; xb[warp * hd + i] = 0.f; }
    }
    __syncthreads();
    if (tid < M) {
      float a[TC] = {0.f, 0.f, 0.f, 0.f};
      for (int n = 0; n < hd; ++n) {
        const float p = Ps[n * M + tid];
#pragma unroll
        for (int c = 0; c < TC; ++c) a[c] = fmaf(xa[c * hd + n], p, a[c]);
      }
      float4 f;
      float* fp = reinterpret_cast<float*>(&f);
#pragma unroll
      for (int c = 0; c < TC; ++c) {
        const bool live = (t0 + c) < Tn && (t0 + c) < len;
        fp[c] = live ? expf(fminf(fmaxf(a[c], -15.f), 15.f)) * 0.1f : 0.f;
      }
      *reinterpret_cast<float4*>(fa + tid * TC) = f;
    }
    __syncthreads();
    if (tid < hd) {
      float vn[TC];
#pragma unroll
      for (int c = 0; c < TC; ++c) vn[c] = xb[c * hd + tid];
#pragma unroll
      for (int m = 0; m < AT; ++m) {
        if (m < M) {
          const float4 f = *reinterpret_cast<const float4*>(fa + m * TC);
          kv[m] = fmaf(f.x, vn[0], kv[m]);
          kv[m] = fmaf(f.y, vn[1], kv[m]);
          kv[m] = fmaf(f.z, vn[2], kv[m]);
          kv[m] = fmaf(f.w, vn[3], kv[m]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int m = 0; m < AT; ++m) kv[m] *= 0.1f;

  // ---------------- phase 2: out[t] = LN( (q_proj[t] . kv) * 0.1 / clamp(q_proj[t].k_proj[t], 1e-6) )
  for (int t0 = 0; t0 < Tn; t0 += TC) {
    const int t = t0 + warp;
    if (t < Tn) {
      const T* row = qkv + ((long)(b * Tn + t)) * 3 * D + h * hd;
      for (int i = lane; i < hd; i += 32) {
        xa[warp * hd + i] = to_f<T>(row[i]) * 0.1f;
        xb[warp * hd + i] = to_f<T>(row[D + i]) * 0.1f;
      }
      __syncwarp();
      warp_ln(xa + warp * hd, hd, nw, nb, lane);
      warp_ln(xb + warp * hd, hd, nw, nb, lane);
      __syncwarp();
      warp_l2(xa + warp * hd, hd, lane);
      warp_l2(xb + warp * hd, hd, lane);
    } else {
      for (int i = lane; i < hd; i += 32) { xa[warp * hd + i] = 0.f; xb[warp * hd + i] = 0.f; }
    }
    __syncthreads();
    float dpart[TC] = {0.f, 0.f, 0.f, 0.f};
    if (tid < M) {
      float a[TC] = {0.f, 0.f, 0.f, 0.f}, c2[TC] = {0.f, 0.f, 0.f, 0.f};
      for (int n = 0; n < hd; ++n) {
        const float p = Ps[n * M + tid];
#pragma unroll
        for (int c = 0; c < TC; ++c) {
          a[c] = fmaf(xa[c * hd + n], p, a[c]);
          c2[c] = fmaf(xb[c * hd + n], p, c2[c]);
        }
      }
      float4 f;
      float* fp = reinterpret_cast<float*>(&f);
#pragma unroll
      for (int c = 0; c < TC; ++c) {
        const float qp = expf(fminf(fmaxf(a[c], -15.f), 15.f)) * 0.1f;
        const bool live = (t0 + c) < len;
        const float kp = live ? expf(fminf(fmaxf(c2[c], -15.f), 15.f)) * 0.1f : 0.f;
        fp[c] = qp;
        dpart[c] = qp * kp;
      }
      *reinterpret_cast<float4*>(fa + tid * TC) = f;
    }
#pragma unroll
    for (int c = 0; c < TC; ++c) {
      const float s = warp_sum(dpart[c]);
      if (lane == 0) red[c * 4 + warp] = s;
    }
    __syncthreads();
    if (tid < hd) {
      float o[TC] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int m = 0; m < AT; ++m) {
        if (m < M) {
          const float4 f = *reinterpret_cast<const float4*>(fa + m * TC);
          o[0] = fmaf(f.x, kv[m], o[0]);
          o[1] = fmaf(f.y, kv[m], o[1]);
          o[2] = fmaf(f.z, kv[m], o[2]);
          o[3] = fmaf(f.w, kv[m], o[3]);
        }
      }
#pragma unroll
      for (int c = 0; c < TC; ++c) {
        const float den = fmaxf((red[c * 4] + red[c * 4 + 1]) + (red[c * 4 + 2] + red[c * 4 + 3]), 1e-6f);
        ob[c * hd + tid] = (o[c] * 0.1f) / den;
      }
    }
    __syncthreads();
    if (t < Tn) {
      warp_ln(ob + warp * hd, hd, nw, nb, lane);
      __syncwarp();
      T* orow = out + ((long)(b * Tn + t)) * D + h * hd;
      for (int i = lane; i < hd; i += 32) orow[i] = from_f<T>(ob[warp * hd + i]);
    }
    __syncthreads();
  }
}

// ctx[b,h,d,l] = sum_n softmax_n(k[b,n,h,d]) * v[b,n,h,l]
template <typename T>
__global__ void __launch_bounds__(AT)
lincross_ctx_kernel(const T* __restrict__ k, const T* __restrict__ v, const int* __restrict__ nt,
                    int Nt_max, int H, int hd, float* __restrict__ ctx) {
  extern __shared__ float sm[];
  float* ks = sm;                 // [n][hd] softmaxed keys
  float* vs = ks + Nt_max * hd;   // [n][hd]
  const int b = blockIdx.x / H, h = blockIdx.x % H, tid = threadIdx.x;
  const int D = H * hd;
  const int n_tok = nt ? nt[b] : Nt_max;
  if (tid < hd) {
    float mx = -INFINITY;
    for (int n = 0; n < n_tok; ++n) {
      const long off = ((long)(b * Nt_max + n)) * D + h * hd + tid;
      const float kk = to_f<T>(k[off]);
      ks[n * hd + tid] = kk;
      vs[n * hd + tid] = to_f<T>(v[off]);
      mx = fmaxf(mx, kk);
    }
    float s = 0.f;
    for (int n = 0; n < n_tok; ++n) {
      const float e = expf(ks[n * hd + tid] - mx);
      ks[n * hd + tid] = e;
      s += e;
    }
    for (int n = 0; n < n_tok; ++n) ks[n * hd + tid] = ks[n * hd + tid] / s;
  }
  __syncthreads();
  if (tid < hd) {
    float* c = ctx + ((long)(b * H + h)) * hd * hd;
    for (int d = 0; d < hd; ++d) {
      float a = 0.f;
      for (int n = 0; n < n_tok; ++n) a = fmaf(ks[n * hd + d], vs[n * hd + tid], a);
      c[d * hd + tid] = a;
    }
  }
}

// y[t,h,:] = softmax_hd(q[t,h,:]) @ ctx[b,h]
template <typename T>
__global__ void __launch_bounds__(AT)
lincross_apply_kernel(const T* __restrict__ q, const float* __restrict__ ctx, int Tn, int H, int hd,
                      T* __restrict__ y) {
  __shared__ float xs[TC][AT];
  __shared__ __align__(16) float pT[AT][TC];
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = H * hd;
  float cc[AT];  // column `tid` of ctx[b,h] ([d][l])
  {
    const float* c = ctx + ((long)(b * H + h)) * hd * hd;
#pragma unroll
    for (int d = 0; d < AT; ++d) cc[d] = (d < hd && tid < hd) ? c[d * hd + tid] : 0.f;
  }
  for (int t0 = 0; t0 < Tn; t0 += TC) {
    const int t = t0 + warp;
    if (t < Tn) {
      const T* row = q + ((long)(b * Tn + t)) * D + h * hd;
      float mx = -INFINITY;
      for (int i = lane; i < hd; i += 32) { const float x = to_f<T>(row[i]); xs[warp][i] = x; mx = fmaxf(mx, x); }
      mx = warp_max(mx);
      float s = 0.f;
      for (int i = lane; i < hd; i += 32) { const float e = expf(xs[warp][i] - mx); xs[warp][i] = e; s += e; }
      s = warp_sum(s);
      for (int i = lane; i < hd; i += 32) pT[i][warp] = xs[warp][i] / s;
    } else {
      for (int i = lane; i < hd; i += 32) pT[i][warp] = 0.f;
    }
    __syncthreads();
    if (tid < hd) {
      float o[TC] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int d = 0; d < AT; ++d) {
        if (d < hd) {
          const float4 f = *reinterpret_cast<const float4*>(&pT[d][0]);
          o[0] = fmaf(f.x, cc[d], o[0]);
          o[1] = fmaf(f.y, cc[d], o[1]);
          o[2] = fmaf(f.z, cc[d], o[2]);
          o[3] = fmaf(f.w, cc[d], o[3]);
        }
      }
#pragma unroll
      for (int c = 0; c < TC; ++c)
        if (t0 + c < Tn) y[((long)(b * Tn + t0 + c)) * D + h * hd + tid] = from_f<T>(o[c]);
    }
    __syncthreads();
  }
}

// o[t,h,:] = softmax_n((q[t,h,:] * scale) . k[b,n,h,:]) @ v[b,n,h,:]
template <typename T>
__global__ void __launch_bounds__(AT)
softmax_cross_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                     const int* __restrict__ nt, int Tn, int Nt_max, int H, int hd, float scale,
                     T* __restrict__ o) {
  extern __shared__ float sm[];
  const int ldk = hd + 1;             // padded: lanes index different keys
  float* ks = sm;                     // [Nt_max][hd+1]
  float* vs = ks + Nt_max * ldk;      // [Nt_max][hd]
  float* qs = vs + Nt_max * hd;       // [TC][hd]
  float* ps = qs + TC * hd;           // [TC][96]
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = H * hd;
  const int n_tok = nt ? nt[b] : Nt_max;
  for (int i = tid; i < n_tok * hd; i += AT) {
    const int n = i / hd, d = i - n * hd;
    const long off = ((long)(b * Nt_max + n)) * D + h * hd + d;
    ks[n * ldk + d] = to_f<T>(k[off]);
    vs[n * hd + d] = to_f<T>(v[off]);
  }
  __syncthreads();
  for (int t = warp; t < Tn; t += TC) {
    const T* row = q + ((long)(b * Tn + t)) * D + h * hd;
    for (int i = lane; i < hd; i += 32) qs[warp * hd + i] = to_f<T>(row[i]) * scale;
    __syncwarp();
    float sc[3];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int n = lane + 32 * j;
      float a = -INFINITY;
      if (n < n_tok) {
        a = 0.f;
        for (int d = 0; d < hd; ++d) a = fmaf(qs[warp * hd + d], ks[n * ldk + d], a);
      }
      sc[j] = a;
      mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int n = lane + 32 * j;
      sc[j] = (n < n_tok) ? expf(sc[j] - mx) : 0.f;
      s += sc[j];
    }
    s = warp_sum(s);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int n = lane + 32 * j;
      if (n < 96) ps[warp * 96 + n] = sc[j] / s;
    }
    __syncwarp();
    T* orow = o + ((long)(b * Tn + t)) * D + h * hd;
    for (int i = lane; i < hd; i += 32) {
      float a = 0.f;
      for (int n = 0; n < n_tok; ++n) a = fmaf(ps[warp * 96 + n], vs[n * hd + i], a);
      orow[i] = from_f<T>(a);
    }
    __syncwarp();
  }
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
      return MDM_ERR_CUDA;
  }
  return MDM_OK;
}

}  // namespace

extern "C" MDM_API int mdm_fastattn_ordered(const void* qkv, int dt, const float* P, const float* norm_w,
                                            const float* norm_b, const int64_t* length, int length_shift, int B,
                                            int H, int T, int hd, int M, void* out, const int* seq_order,
                                            const void* Pt_bf16, void* stream);
extern "C" MDM_API int mdm_fastattn(const void* qkv, int dt, const float* P, const float* norm_w,
                                    const float* norm_b, const int64_t* length, int length_shift, int B,
                                    int H, int T, int hd, int M, void* out, void* stream) {
  return mdm_fastattn_ordered(qkv, dt, P, norm_w, norm_b, length, length_shift, B, H, T, hd, M, out, nullptr, nullptr,
                              stream);
}

extern "C" MDM_API int mdm_fastattn_ordered(const void* qkv, int dt, const float* P, const float* norm_w,
                                            const float* norm_b, const int64_t* length, int length_shift, int B,
                                            int H, int T, int hd, int M, void* out, const int* seq_order,
                                            const void* Pt_bf16, void* stream) {
  if (!qkv || !P || !norm_w || !norm_b || !out) return MDM_ERR_ARG;
  if (hd > AT || M > AT || (hd & 3) || (M & 3)) return MDM_ERR_UNSUPPORTED;
  if (B * H == 0 || T == 0) return MDM_OK;
  const size_t smem = sizeof(float) * ((size_t)hd * M + 3 * TC * hd + 2 * M * TC + TC * 4);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dt == MDM_BF16) {
    // MDM_FA_UMMA=0 keeps the mma.sync kernels (A/B runs)
    static const int umma_env = [] { const char* e = getenv("MDM_FA_UMMA"); return e ? atoi(e) : 1; }();
    if (umma_env && Pt_bf16) {
      const int r = mdm_fastattn_umma(qkv, Pt_bf16, norm_w, norm_b, length, length_shift, B, H, T, hd, M, out, seq_order, st);
      if (r != MDM_ERR_UNSUPPORTED) return r;
    }
    const int r = mdm_fastattn_tc(qkv, P, norm_w, norm_b, length, length_shift, B, H, T, hd, M, out, seq_order, Pt_bf16, st);
    if (r != MDM_ERR_UNSUPPORTED) return r;  // otherwise: shape outside the tensor-core kernel
  }
  if (dt == MDM_F32) {
    if (set_smem(fastattn_kernel<float>, smem)) return MDM_ERR_CUDA;
    fastattn_kernel<float><<<B * H, AT, smem, st>>>(reinterpret_cast<const float*>(qkv), P, norm_w, norm_b,
                                                     length, length_shift, H, T, hd, M,
                                                     reinterpret_cast<float*>(out));
  } else {
    if (set_smem(fastattn_kernel<bf16>, smem)) return MDM_ERR_CUDA;
    fastattn_kernel<bf16><<<B * H, AT, smem, st>>>(reinterpret_cast<const bf16*>(qkv), P, norm_w, norm_b,
                                                    length, length_shift, H, T, hd, M,
                                                    reinterpret_cast<bf16*>(out));
  }
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_lincross_ctx(const void* k, const void* v, int dt, const int* nt, int B,
                                        int Nt_max, int H, int hd, float* ctx, void* stream) {
  if (!k || !v || !ctx) return MDM_ERR_ARG;
  if (hd > AT) return MDM_ERR_UNSUPPORTED;
  if (B * H == 0) return MDM_OK;
  const size_t smem = sizeof(float) * 2 * (size_t)Nt_max * hd;
  if (smem > 200 * 1024) return MDM_ERR_UNSUPPORTED;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dt == MDM_F32) {
    if (set_smem(lincross_ctx_kernel<float>, smem)) return MDM_ERR_CUDA;
    lincross_ctx_kernel<float><<<B * H, AT, smem, st>>>(reinterpret_cast<const float*>(k),
                                                         reinterpret_cast<const float*>(v), nt, Nt_max, H,
                                                         hd, ctx);
  } else {
    if (set_smem(lincross_ctx_kernel<bf16>, smem)) return MDM_ERR_CUDA;
    lincross_ctx_kernel<bf16><<<B * H, AT, smem, st>>>(reinterpret_cast<const bf16*>(k),
                                                        reinterpret_cast<const bf16*>(v), nt, Nt_max, H,
                                                        hd, ctx);
  }
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_lincross_apply_ex(const void* q, int dt, const float* ctx, const void* ctxT_bf16, int B, int T,
                                             int H, int hd, void* y, void* stream);
extern "C" MDM_API int mdm_lincross_apply(const void* q, int dt, const float* ctx, int B, int T, int H,
                                          int hd, void* y, void* stream) {
  return mdm_lincross_apply_ex(q, dt, ctx, nullptr, B, T, H, hd, y, stream);
}

extern "C" MDM_API int mdm_lincross_apply_ex(const void* q, int dt, const float* ctx, const void* ctxT_bf16, int B, int T,
                                             int H, int hd, void* y, void* stream) {
  if (!q || !ctx || !y) return MDM_ERR_ARG;
  if (hd > AT) return MDM_ERR_UNSUPPORTED;
  if (B * H == 0 || T == 0) return MDM_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dt == MDM_BF16 && ctxT_bf16) {
    static const int umma_env = [] { const char* e = getenv("MDM_LC_UMMA"); return e ? atoi(e) : 1; }();
    if (umma_env) {
      const int r = mdm_lincross_apply_umma(q, ctxT_bf16, B, T, H, hd, y, st);
      if (r != MDM_ERR_UNSUPPORTED) return r;
    }
  }
  if (dt == MDM_BF16) {
    const int r = mdm_lincross_apply_tc(q, ctx, B, T, H, hd, y, st);
    if (r != MDM_ERR_UNSUPPORTED) return r;
  }
  if (dt == MDM_F32)
    lincross_apply_kernel<float><<<B * H, AT, 0, st>>>(reinterpret_cast<const float*>(q), ctx, T, H, hd,
                                                        reinterpret_cast<float*>(y));
  else
    lincross_apply_kernel<bf16><<<B * H, AT, 0, st>>>(reinterpret_cast<const bf16*>(q), ctx, T, H, hd,
                                                       reinterpret_cast<bf16*>(y));
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_softmax_cross(const void* q, const void* k, const void* v, int dt,
                                         const int* nt, int B, int T, int Nt_max, int H, int hd, void* o,
                                         void* stream) {
  if (!q || !k || !v || !o) return MDM_ERR_ARG;
  if (hd > AT || Nt_max > 96) return MDM_ERR_UNSUPPORTED;
  if (B * H == 0 || T == 0) return MDM_OK;
  const size_t smem = sizeof(float) * ((size_t)Nt_max * (hd + 1) + (size_t)Nt_max * hd + TC * hd + TC * 96);
  const float scale = (float)(1.0 / sqrt((double)hd));  // python: head_dim ** -0.5, then fp32
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dt == MDM_BF16) {
    static const int umma_env = [] { const char* e = getenv("MDM_SC_UMMA"); return e ? atoi(e) : 1; }();
    if (umma_env) {
      const int r = mdm_softmax_cross_umma(q, k, v, nt, B, T, Nt_max, H, hd, scale, o, st);
      if (r != MDM_ERR_UNSUPPORTED) return r;
    }
    const int r = mdm_softmax_cross_tc(q, k, v, nt, B, T, Nt_max, H, hd, scale, o, st);
    if (r != MDM_ERR_UNSUPPORTED) return r;
  }
  if (dt == MDM_F32) {
    if (set_smem(softmax_cross_kernel<float>, smem)) return MDM_ERR_CUDA;
    softmax_cross_kernel<float><<<B * H, AT, smem, st>>>(
        reinterpret_cast<const float*>(q), reinterpret_cast<const float*>(k),
        reinterpret_cast<const float*>(v), nt, T, Nt_max, H, hd, scale, reinterpret_cast<float*>(o));
  } else {
    if (set_smem(softmax_cross_kernel<bf16>, smem)) return MDM_ERR_CUDA;
    softmax_cross_kernel<bf16><<<B * H, AT, smem, st>>>(
        reinterpret_cast<const bf16*>(q), reinterpret_cast<const bf16*>(k),
        reinterpret_cast<const bf16*>(v), nt, T, Nt_max, H, hd, scale, reinterpret_cast<bf16*>(o));
  }
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}
