// Small per-sequence / elementwise kernels of MotionTransformer.forward and the CFG DDPM sampler.
// Reference: models/time.py:15-26, models/gate.py:18-19, models/transformer.py:324,
//            models/gaussian_diffusion.py:449-475,538-558,1042-1098.
#include "common.cuh"

namespace {

template <typename T>
__global__ void timestep_embedding_kernel(const int64_t* __restrict__ t, int B, int D, T* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = D / 2;
  if (i >= B * half) return;
  const int b = i / half, j = i - b * half;
  // freqs = exp(-ln(max_period) * arange(half) / half), evaluated in fp32 in the reference's order
  const float f = expf(__fdiv_rn(__fmul_rn(-9.210340371976184f, (float)j), (float)half));
  const float arg = __fmul_rn((float)t[b], f);
  out[(long)b * D + j] = from_f<T>(cosf(arg));
  out[(long)b * D + half + j] = from_f<T>(sinf(arg));
  if ((D & 1) && j == 0) out[(long)b * D + D - 1] = from_f<T>(0.f);
}

template <typename T>
__global__ void gated_mix_kernel(const float* __restrict__ t, const float* __restrict__ x, long n, T* __restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float tv = t[i], xv = x[i];
  const float g = 1.f / (1.f + expf(-(tv + xv)));
  out[i] = from_f<T>(g * tv + (1.f - g) * xv);
}

template <typename T>
__global__ void pad_cast_kernel(const float* __restrict__ x, long rows, int F, T* __restrict__ out, int ld) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ld) return;
  const long r = i / ld;
  const int c = (int)(i - r * ld);
  out[i] = from_f<T>(c < F ? x[r * F + c] : 0.f);
}

// torch-eager arithmetic order, no FMA contraction, so that the update is bit-identical to
// p_sample_with_cfg given the same eps.
__global__ void cfg_update_kernel(const float* __restrict__ x, const float* __restrict__ eps_c,
                                  const float* __restrict__ eps_u, const float* __restrict__ noise,
                                  const int64_t* __restrict__ t, const float* __restrict__ tables, int n_steps,
                                  float s, int clip, long per_sample, long total, float* __restrict__ x_prev,
                                  float* __restrict__ x0_out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = (int)(i / per_sample);
  const int64_t ts = t[b];
  const float c_recip = tables[ts], c_recipm1 = tables[n_steps + ts];
  const float coef1 = tables[2 * n_steps + ts], coef2 = tables[3 * n_steps + ts];
  const float logvar = tables[4 * n_steps + ts];
  const float xv = x[i];
  float x0c = __fsub_rn(__fmul_rn(c_recip, xv), __fmul_rn(c_recipm1, eps_c[i]));
  float x0u = __fsub_rn(__fmul_rn(c_recip, xv), __fmul_rn(c_recipm1, eps_u[i]));
  if (clip) {
    x0c = fminf(fmaxf(x0c, -1.f), 1.f);
    x0u = fminf(fmaxf(x0u, -1.f), 1.f);
  }
  const float guided = __fadd_rn(x0u, __fmul_rn(s, __fsub_rn(x0c, x0u)));
  const float mean = __fadd_rn(__fmul_rn(coef1, guided), __fmul_rn(coef2, xv));
  const float nz = ts != 0 ? 1.f : 0.f;
  const float sd = expf(__fmul_rn(0.5f, logvar));
  x_prev[i] = __fadd_rn(mean, __fmul_rn(__fmul_rn(nz, sd), noise[i]));
  if (x0_out) x0_out[i] = guided;
}

// p_mean_variance after the model call + (optionally) the p_sample update, torch-eager op order
__global__ void p_mean_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                              const float* __restrict__ noise, const int64_t* __restrict__ t,
                              const float* __restrict__ tables, int n_steps, int clip, long per_sample, long total,
                              float* __restrict__ mean_out, float* __restrict__ x0_out, float* __restrict__ sample) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t ts = t[i / per_sample];
  const float xv = x[i];
  float x0 = __fsub_rn(__fmul_rn(tables[ts], xv), __fmul_rn(tables[n_steps + ts], eps[i]));
  if (clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
  const float mean = __fadd_rn(__fmul_rn(tables[2 * n_steps + ts], x0), __fmul_rn(tables[3 * n_steps + ts], xv));
  if (mean_out) mean_out[i] = mean;
  if (x0_out) x0_out[i] = x0;
  if (sample) {
    const float nz = ts != 0 ? 1.f : 0.f;
    const float sd = expf(__fmul_rn(0.5f, tables[4 * n_steps + ts]));
    sample[i] = __fadd_rn(mean, __fmul_rn(__fmul_rn(nz, sd), noise[i]));
  }
}

__global__ void q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                const int64_t* __restrict__ t, const float* __restrict__ tables, int n_steps,
                                long per_sample, long total, float* __restrict__ x_t) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t ts = t[i / per_sample];
  x_t[i] = __fadd_rn(__fmul_rn(tables[ts], x0[i]), __fmul_rn(tables[n_steps + ts], noise[i]));
}

inline unsigned blocks(long n) { return (unsigned)((n + 255) / 256); }

}  // namespace

extern "C" MDM_API int mdm_timestep_embedding(const int64_t* t, int B, int D, void* out, int dt, void* stream) {
  if (!t || !out || D < 2) return MDM_ERR_ARG;
  if (B == 0) return MDM_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long n = (long)B * (D / 2);
  if (dt == MDM_F32) timestep_embedding_kernel<float><<<blocks(n), 256, 0, st>>>(t, B, D, reinterpret_cast<float*>(out));
  else timestep_embedding_kernel<bf16><<<blocks(n), 256, 0, st>>>(t, B, D, reinterpret_cast<bf16*>(out));
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_gated_mix(const float* t, const float* x, long n, void* out, int dt, void* stream) {
  if (!t || !x || !out) return MDM_ERR_ARG;
  if (n == 0) return MDM_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dt == MDM_F32) gated_mix_kernel<float><<<blocks(n), 256, 0, st>>>(t, x, n, reinterpret_cast<float*>(out));
  else gated_mix_kernel<bf16><<<blocks(n), 256, 0, st>>>(t, x, n, reinterpret_cast<bf16*>(out));
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_pad_cast(const float* x, long rows, int F, void* out, int ld_out, int dt, void* stream) {
  if (!x || !out || ld_out < F) return MDM_ERR_ARG;
  if (rows == 0) return MDM_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long n = rows * ld_out;
  if (dt == MDM_F32) pad_cast_kernel<float><<<blocks(n), 256, 0, st>>>(x, rows, F, reinterpret_cast<float*>(out), ld_out);
  else pad_cast_kernel<bf16><<<blocks(n), 256, 0, st>>>(x, rows, F, reinterpret_cast<bf16*>(out), ld_out);
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_cfg_update(const float* x, const float* eps_c, const float* eps_u, const float* noise,
                                      const int64_t* t, const float* tables, int n_steps, float cfg_scale, int clip,
                                      int B, long per_sample, float* x_prev, float* x0, void* stream) {
  if (!x || !eps_c || !eps_u || !noise || !t || !tables || !x_prev) return MDM_ERR_ARG;
  const long total = (long)B * per_sample;
  if (total == 0) return MDM_OK;
  cfg_update_kernel<<<blocks(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, eps_c, eps_u, noise, t, tables, n_steps, cfg_scale, clip, per_sample, total, x_prev, x0);
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_q_sample(const float* x0, const float* noise, const int64_t* t, const float* tables2,
                                    int n_steps, int B, long per_sample, float* x_t, void* stream) {
  if (!x0 || !noise || !t || !tables2 || !x_t) return MDM_ERR_ARG;
  const long total = (long)B * per_sample;
  if (total == 0) return MDM_OK;
  q_sample_kernel<<<blocks(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x0, noise, t, tables2, n_steps,
                                                                                     per_sample, total, x_t);
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_p_mean_variance(const float* x, const float* eps, const float* noise, const int64_t* t,
                                           const float* tables, int n_steps, int clip, int B, long per_sample,
                                           float* mean, float* x0, float* sample, void* stream) {
  if (!x || !eps || !t || !tables || (sample && !noise)) return MDM_ERR_ARG;
  const long total = (long)B * per_sample;
  if (total == 0) return MDM_OK;
  p_mean_kernel<<<blocks(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, eps, noise, t, tables, n_steps, clip,
                                                                                   per_sample, total, mean, x0, sample);
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}
