// Small per-sequence / elementwise kernels of MotionTransformer.forward and the CFG DDPM sampler.
// Reference: models/time.py:15-26, models/gate.py:18-19, models/transformer.py:324,
//            models/gaussian_diffusion.py:449-475,538-558,1042-1098.
#include "common.cuh"

namespace {

template <typename T>
__global__ void timestep_embedding_kernel(const int64_t* __restrict__ t, int B, int D, T* __restrict__ out) {
  pdl_enter();
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
  pdl_enter();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float tv = t[i], xv = x[i];
  const float g = 1.f / (1.f + expf(-(tv + xv)));
  out[i] = from_f<T>(g * tv + (1.f - g) * xv);
}

template <typename T>
__global__ void pad_cast_kernel(const float* __restrict__ x, long rows, int F, T* __restrict__ out, int ld) {
  pdl_enter();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ld) return;
  const long r = i / ld;
  const int c = (int)(i - r * ld);
  out[i] = from_f<T>(c < F ? x[r * F + c] : 0.f);
}

// torch-eager arithmetic order, no FMA contraction, so that the update is bit-identical to
// p_sample_with_cfg given the same eps.
// x and x_prev may be the same buffer (CFGStepper updates its state in place: every thread reads element i before it
// writes element i), so neither carries __restrict__.
__global__ void cfg_update_kernel(const float* x, const float* __restrict__ eps_c,
                                  const float* __restrict__ eps_u, const float* __restrict__ noise,
                                  const int64_t* __restrict__ t, const float* __restrict__ tables, int n_steps,
                                  float s, int clip, long per_sample, long total, float* x_prev,
                                  float* __restrict__ x0_out) {
  pdl_enter();
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

// DDIM update (ddim_sample, gaussian_diffusion.py:721-742) after the model call(s), torch-eager op order.
// (sqrt.rn / div.rn like torch's CUDA kernels: bit-identical to the reference run on a GPU; torch's CPU sqrt is off by
// one ulp for 0.6 % of inputs, so a CPU run of the reference agrees to 1e-6 only.)
// eps_u != NULL: classifier-free guidance on pred_xstart exactly as p_sample_with_cfg combines it (:1074-1079).
// t_prev == NULL: alpha_bar_prev = alphas_cumprod_prev[t] (the reference's 1000-step loop); otherwise the
// previous timestep of a strided schedule (alpha_bar_prev = alphas_cumprod[t_prev], 1 for t_prev < 0).
// tables: [4][n_steps] = sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod, alphas_cumprod, alphas_cumprod_prev.
__global__ void ddim_update_kernel(const float* x /* may alias x_prev */, const float* __restrict__ eps_c,
                                   const float* __restrict__ eps_u, const float* __restrict__ noise,
                                   const int64_t* __restrict__ t, const int64_t* __restrict__ t_prev,
                                   const float* __restrict__ tables, int n_steps, float s, float eta, int clip,
                                   long per_sample, long total, float* x_prev, float* __restrict__ x0_out) {
  pdl_enter();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = (int)(i / per_sample);
  const int64_t ts = t[b];
  const float c_recip = tables[ts], c_recipm1 = tables[n_steps + ts];
  const float ab = tables[2 * n_steps + ts];
  float ab_prev = tables[3 * n_steps + ts];
  if (t_prev) { const int64_t tp = t_prev[b]; ab_prev = tp < 0 ? 1.0f : tables[2 * n_steps + tp]; }
  const float xv = x[i];
  float x0 = __fsub_rn(__fmul_rn(c_recip, xv), __fmul_rn(c_recipm1, eps_c[i]));       // :554-558
  if (clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
  if (eps_u) {
    float x0u = __fsub_rn(__fmul_rn(c_recip, xv), __fmul_rn(c_recipm1, eps_u[i]));
    if (clip) x0u = fminf(fmaxf(x0u, -1.f), 1.f);
    x0 = __fadd_rn(x0u, __fmul_rn(s, __fsub_rn(x0, x0u)));                             // :1074-1079
  }
  const float eps = __fdiv_rn(__fsub_rn(__fmul_rn(c_recip, xv), x0), c_recipm1);        // :567-571
  const float sigma = __fmul_rn(__fmul_rn(eta, __fsqrt_rn(__fdiv_rn(__fsub_rn(1.f, ab_prev), __fsub_rn(1.f, ab)))),
                                __fsqrt_rn(__fsub_rn(1.f, __fdiv_rn(ab, ab_prev))));    // :729-733
  const float mean = __fadd_rn(__fmul_rn(x0, __fsqrt_rn(ab_prev)),
                               __fmul_rn(__fsqrt_rn(__fsub_rn(__fsub_rn(1.f, ab_prev), __fmul_rn(sigma, sigma))), eps));
  const float nz = ts != 0 ? 1.f : 0.f;
  const float nv = noise ? noise[i] : 0.f;
  x_prev[i] = __fadd_rn(mean, __fmul_rn(__fmul_rn(nz, sigma), nv));                     // :740-741
  if (x0_out) x0_out[i] = x0;
}

// Inclusive scan of v over the block's threads (one value per thread), result of thread i = sum_{j <= i} v_j.
// `carry` is added to every result; the block total (incl. carry) is returned to every thread.  warp_tot: [32].
__device__ __forceinline__ float block_scan_incl(float v, float carry, float* warp_tot, float& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  __syncthreads();                      // warp_tot may still be read by the previous call
  if (lane == 31) warp_tot[warp] = v;
  __syncthreads();
  float base = carry;
  for (int w = 0; w < warp; ++w) base += warp_tot[w];
  float tot = carry;
  for (int w = 0; w < nw; ++w) tot += warp_tot[w];
  total = tot;
  return v + base;
}

// Generated features -> joint positions, the step after the sampler in the reference's pipeline
// (tools/visualization.py:91 `motion * std + mean`, utils/motion_process.py:362-417 recover_root_rot_pos +
// recover_from_ric, utils/quaternion.py:16-20,54-73 qinv / qrot).  One block per sequence:
//   ang[t]  = sum_{s < t} rot_vel[s]                          (root yaw, cumulative)
//   rpos[t] = sum_{s <= t} qrot(qinv(q[s]), (vx[s-1], 0, vz[s-1]))   (root trajectory; y = data[..., 3])
//   joint j >= 1: qrot(qinv(q[t]), ric[t, j-1]) + (rpos.x, 0, rpos.z);  joint 0 = rpos
// with q[t] = (cos ang, 0, sin ang, 0).  x: [B, T, F] normalised features; mean / std: [F] or NULL.
__global__ void recover_ric_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                   const float* __restrict__ stdv, int T, int F, int J, float* __restrict__ out) {
  extern __shared__ float sh[];
  float* ang = sh;            // [T]
  float* px = ang + T;        // [T]
  float* pz = px + T;         // [T]
  __shared__ float warp_tot[32];
  const int b = blockIdx.x;
  const float* xb = x + (long)b * T * F;
  auto feat = [&](int t, int f) {
    const float v = xb[(long)t * F + f];
    return mean ? __fadd_rn(__fmul_rn(v, stdv[f]), mean[f]) : v;
  };
  // pass 1: root yaw (exclusive prefix sum of the rotation velocity)
  float carry = 0.f;
  for (int t0 = 0; t0 < T; t0 += blockDim.x) {
    const int t = t0 + threadIdx.x;
    const float v = (t >= 1 && t < T) ? feat(t - 1, 0) : 0.f;
    float tot;
    const float inc = block_scan_incl(v, carry, warp_tot, tot);
    if (t < T) ang[t] = inc;
    carry = tot;
  }
  __syncthreads();
  // pass 2: root trajectory (inclusive prefix sums of the rotated planar velocity)
  float cx = 0.f, cz = 0.f;
  for (int t0 = 0; t0 < T; t0 += blockDim.x) {
    const int t = t0 + threadIdx.x;
    float rx = 0.f, rz = 0.f;
    if (t >= 1 && t < T) {
      // qrot(qinv(q), v), v = (vx, 0, vz), q = (c, 0, s, 0): qvec = (0, -s, 0)
      const float c = cosf(ang[t]), s = sinf(ang[t]);
      const float vx = feat(t - 1, 1), vz = feat(t - 1, 2);
      const float qy = -s;
      const float uvx = qy * vz, uvz = -(qy * vx);           // cross(qvec, v)   (y component is 0)
      const float uuvx = qy * uvz, uuvz = -(qy * uvx);       // cross(qvec, uv)
      rx = vx + 2.f * (c * uvx + uuvx);
      rz = vz + 2.f * (c * uvz + uuvz);
    }
    float tx, tz;
    const float ix = block_scan_incl(rx, cx, warp_tot, tx);
    const float iz = block_scan_incl(rz, cz, warp_tot, tz);
    if (t < T) { px[t] = ix; pz[t] = iz; }
    cx = tx; cz = tz;
  }
  __syncthreads();
  // pass 3: joints
  float* ob = out + (long)b * T * J * 3;
  for (int i = threadIdx.x; i < T * J; i += blockDim.x) {
    const int t = i / J, j = i - t * J;
    float ox, oy, oz;
    if (j == 0) {
      ox = px[t]; oy = feat(t, 3); oz = pz[t];
    } else {
      const float c = cosf(ang[t]), s = sinf(ang[t]);
      const float vx = feat(t, 4 + (j - 1) * 3), vy = feat(t, 5 + (j - 1) * 3), vz = feat(t, 6 + (j - 1) * 3);
      const float qy = -s;
      const float uvx = qy * vz, uvz = -(qy * vx);
      const float uuvx = qy * uvz, uuvz = -(qy * uvx);
      ox = vx + 2.f * (c * uvx + uuvx) + px[t];
      oy = vy;
      oz = vz + 2.f * (c * uvz + uuvz) + pz[t];
    }
    ob[(long)i * 3] = ox; ob[(long)i * 3 + 1] = oy; ob[(long)i * 3 + 2] = oz;
  }
}

// dst[n][c][r] (bf16) = src[n][r][c] (fp32): packs step-invariant fp32 state as a K-major bf16 tensor-core operand
__global__ void transpose_cast_kernel(const float* __restrict__ src, int R, int Cc, long total, bf16* __restrict__ dst) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long n = i / ((long)R * Cc);
  const int rem = (int)(i - n * R * Cc), c = rem / R, r = rem - c * R;
  dst[i] = __float2bfloat16_rn(src[(n * R + r) * Cc + c]);
}

// The reconstruction term of DDPMTrainer.backward_G (trainers/ddpm_trainer.py:207-214):
//   per_frame[b, t] = mean_f (pred - target)^2 ;  loss = sum_{t < min(T, len_b)} per_frame / sum_b min(T, len_b)
// One block per sequence (fixed reduction order inside the block), per-sequence sums to `partial[B]`; the block that
// finishes last adds them in index order, so the scalar is deterministic from launch to launch.
__global__ void masked_mse_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                  const int64_t* __restrict__ length, int T, int F, int B, float* __restrict__ partial,
                                  unsigned* __restrict__ counter, float* __restrict__ loss) {
  __shared__ float red[32];
  __shared__ bool last;
  const int b = blockIdx.x;
  const int len = (int)min((long)T, (long)length[b]);
  const float* p = pred + (long)b * T * F;
  const float* q = target + (long)b * T * F;
  float acc = 0.f;
  const long n = (long)len * F;
  for (long i = threadIdx.x; i < n; i += blockDim.x) { const float d = p[i] - q[i]; acc = fmaf(d, d, acc); }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    partial[b] = s / (float)F;
    __threadfence();
    last = atomicAdd(counter, 1u) == (unsigned)(B - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    float s = 0.f;
    long cnt = 0;
    for (int i = 0; i < B; ++i) { s += partial[i]; cnt += min((long)T, (long)length[i]); }
    *loss = s / (float)cnt;
    *counter = 0u;
  }
}

// Backward building blocks of a Linear (y = x W^T + b) on the forward GEMM kernel:
//   dX = dY . W          : the forward GEMM with W^T ([in, out], mdm_transpose_split_bf16 with S = 1) as its weight
//   dW = dY^T . X        : contraction over the M tokens.  Both operands must be K-major in the token index, i.e.
//                          transposed, and a 512 x 512 result alone would occupy 8 of 148 SMs, so the token range is
//                          split in S slabs: dst[s*C + c][m'] = src[s*Ks + m'][c] lays the slabs out as the row groups
//                          of a grouped GEMM (one group per slab, partial products [S, out, in] in fp32), which
//                          mdm_sum_partials then adds up (+ the bias gradient as column sums of dY).
// 32 x 32 tiles through shared memory: coalesced 64-byte reads and writes.
__global__ void transpose_split_kernel(const bf16* __restrict__ src, long M, int Cc, int Ks, bf16* __restrict__ dst) {
  __shared__ bf16 tile[32][33];
  const int s = blockIdx.z;
  const long m0 = (long)s * Ks + (long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8 threads
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty + 8 * i;
    const long m = m0 + r;
    const bool ok = m < M && (blockIdx.x * 32 + r) < Ks && c0 + tx < Cc;
    tile[r][tx] = ok ? src[m * Cc + c0 + tx] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = ty + 8 * i;
    const int ml = blockIdx.x * 32 + tx;
    if (c0 + c < Cc && ml < Ks) dst[((long)s * Cc + c0 + c) * Ks + ml] = tile[tx][c];
  }
}

// The same for MANY partials of a SHORT vector (e.g. the per-CTA LayerNorm-affine partials of the row pipelines' backward:
// S ~ 1800 rows of 2048 floats): 32 columns x 8 row groups per block, every thread sums the partials s = ty, ty + 8, ...
// of its column, then the 8 groups are added in a fixed order.  (The 1-D kernel below gives such a sum 8 blocks of
// threads that each walk S rows serially: 80 us per call, 24 % of the training step when it was measured.)
__global__ void __launch_bounds__(256)
sum_partials_tall_kernel(const float* __restrict__ part, int S, long n, int accumulate, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long c = (long)blockIdx.x * 32 + tx;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (c < n) {
    int s = ty;
    for (; s + 24 < S; s += 32) {
      a0 += part[(long)s * n + c]; a1 += part[(long)(s + 8) * n + c];
      a2 += part[(long)(s + 16) * n + c]; a3 += part[(long)(s + 24) * n + c];
    }
    for (; s < S; s += 8) a0 += part[(long)s * n + c];
  }
  red[ty][tx] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (ty == 0 && c < n) {
    float t = accumulate ? out[c] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    out[c] = t;
  }
}

// out[i] (+)= sum_s part[s][i]; fixed order
__global__ void sum_partials_kernel(const float* __restrict__ part, int S, long n, int accumulate, float* __restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = accumulate ? out[i] : 0.f;
  for (int s = 0; s < S; ++s) a += part[(long)s * n + i];
  out[i] = a;
}

// column sums of a [M, C] bf16 matrix (bias gradient): block per 32 columns x slab of rows, partials [slabs, C]
__global__ void colsum_kernel(const bf16* __restrict__ src, long M, int Cc, int rows_per_blk, float* __restrict__ part) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
  const long r0 = (long)blockIdx.y * rows_per_blk, r1 = min(M, r0 + rows_per_blk);
  float a = 0.f;
  if (c < Cc)
    for (long r = r0 + ty; r < r1; r += 8) a += __bfloat162float(src[r * Cc + c]);
  red[ty][threadIdx.x & 31] = a;
  __syncthreads();
  if (ty == 0 && c < Cc) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x & 31];
    part[(long)blockIdx.y * Cc + c] = s;
  }
}

// Exact-erf GELU forward / backward on bf16 rows (training keeps the pre-activation of the expert up-projection):
//   h = gelu(p)           dp = dh * (Phi(p) + p * phi(p))
__global__ void gelu_fwd_kernel(const bf16* __restrict__ pre, long n, bf16* __restrict__ h) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) h[i] = __float2bfloat16_rn(gelu_erf(__bfloat162float(pre[i])));
}
__global__ void gelu_bwd_kernel(const bf16* __restrict__ pre, const bf16* __restrict__ dh, long n, bf16* __restrict__ dp) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float p = __bfloat162float(pre[i]);
  const float cdf = 0.5f * (1.0f + erff(p * 0.70710678118654752440f));
  const float pdf = 0.3989422804014327f * expf(-0.5f * p * p);
  dp[i] = __float2bfloat16_rn(__bfloat162float(dh[i]) * (cdf + p * pdf));
}
// dst[r, :] = src[r, :] * scale[r]   (the gate weight / NB of a routed row: gradient of the down-projection output)
__global__ void rowscale_kernel(const bf16* __restrict__ src, const float* __restrict__ scale, long rows, int Cc,
                                bf16* __restrict__ dst) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * Cc) return;
  dst[i] = __float2bfloat16_rn(__bfloat162float(src[i]) * scale[i / Cc]);
}
// out[g, c] = sum of src[r, c] over the rows r of segment g = [seg_off[g], seg_off[g] + seg_cnt[g])  (bias gradients
// of the experts); one block per (32 columns, segment), fixed order
__global__ void seg_colsum_kernel(const bf16* __restrict__ src, int Cc, const int* __restrict__ seg_off,
                                  const int* __restrict__ seg_cnt, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int g = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
  const long r0 = seg_off[g], r1 = r0 + seg_cnt[g];
  float a = 0.f;
  if (c < Cc)
    for (long r = r0 + ty; r < r1; r += 8) a += __bfloat162float(src[r * Cc + c]);
  red[ty][threadIdx.x & 31] = a;
  __syncthreads();
  if (ty == 0 && c < Cc) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x & 31];
    out[(long)g * Cc + c] = s;
  }
}

inline unsigned blocks(long n) { return (unsigned)((n + 255) / 256); }

}  // namespace

extern "C" MDM_API int mdm_timestep_embedding(const int64_t* t, int B, int D, void* out, int dt, void* stream) {
  if (!t || !out || D < 2) return MDM_ERR_ARG;
  if (B == 0) return MDM_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long n = (long)B * (D / 2);
  if (dt == MDM_F32) mdm_launch(timestep_embedding_kernel<float>, blocks(n), 256, 0, st, t, B, D, reinterpret_cast<float*>(out));
  else mdm_launch(timestep_embedding_kernel<bf16>, blocks(n), 256, 0, st, t, B, D, reinterpret_cast<bf16*>(out));
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_gated_mix(const float* t, const float* x, long n, void* out, int dt, void* stream) {
  if (!t || !x || !out) return MDM_ERR_ARG;
  if (n == 0) return MDM_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dt == MDM_F32) mdm_launch(gated_mix_kernel<float>, blocks(n), 256, 0, st, t, x, n, reinterpret_cast<float*>(out));
  else mdm_launch(gated_mix_kernel<bf16>, blocks(n), 256, 0, st, t, x, n, reinterpret_cast<bf16*>(out));
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_pad_cast(const float* x, long rows, int F, void* out, int ld_out, int dt, void* stream) {
  if (!x || !out || ld_out < F) return MDM_ERR_ARG;
  if (rows == 0) return MDM_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long n = rows * ld_out;
  if (dt == MDM_F32) mdm_launch(pad_cast_kernel<float>, blocks(n), 256, 0, st, x, rows, F, reinterpret_cast<float*>(out), ld_out);
  else mdm_launch(pad_cast_kernel<bf16>, blocks(n), 256, 0, st, x, rows, F, reinterpret_cast<bf16*>(out), ld_out);
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_cfg_update(const float* x, const float* eps_c, const float* eps_u, const float* noise,
                                      const int64_t* t, const float* tables, int n_steps, float cfg_scale, int clip,
                                      int B, long per_sample, float* x_prev, float* x0, void* stream) {
  if (!x || !eps_c || !eps_u || !noise || !t || !tables || !x_prev) return MDM_ERR_ARG;
  const long total = (long)B * per_sample;
  if (total == 0) return MDM_OK;
  mdm_launch(cfg_update_kernel, blocks(total), 256, 0, reinterpret_cast<cudaStream_t>(stream), x, eps_c, eps_u, noise, t, tables, n_steps, cfg_scale, clip, per_sample, total, x_prev, x0);
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

extern "C" MDM_API int mdm_ddim_update(const float* x, const float* eps_c, const float* eps_u, const float* noise,
                                       const int64_t* t, const int64_t* t_prev, const float* tables4, int n_steps,
                                       float cfg_scale, float eta, int clip, int B, long per_sample, float* x_prev,
                                       float* x0, void* stream) {
  if (!x || !eps_c || !t || !tables4 || !x_prev || (eta != 0.f && !noise)) return MDM_ERR_ARG;
  const long total = (long)B * per_sample;
  if (total == 0) return MDM_OK;
  mdm_launch(ddim_update_kernel, blocks(total), 256, 0, reinterpret_cast<cudaStream_t>(stream), x, eps_c, eps_u, noise, t, t_prev, tables4, n_steps, cfg_scale, eta, clip, per_sample, total, x_prev, x0);
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_recover_from_ric(const float* x, const float* mean, const float* stdv, int B, int T, int F,
                                            int joints, float* out, void* stream) {
  if (!x || !out || (mean == nullptr) != (stdv == nullptr) || joints < 1 || F < 4 + (joints - 1) * 3 || T < 1)
    return MDM_ERR_ARG;
  if (B == 0) return MDM_OK;
  const size_t smem = sizeof(float) * 3 * (size_t)T;
  if (smem > 48 * 1024) return MDM_ERR_UNSUPPORTED;
  recover_ric_kernel<<<B, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(x, mean, stdv, T, F, joints, out);
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_transpose_cast_bf16(const float* src, long n, int R, int Cc, void* dst, void* stream) {
  if (!src || !dst || R <= 0 || Cc <= 0) return MDM_ERR_ARG;
  const long total = n * R * Cc;
  if (total == 0) return MDM_OK;
  transpose_cast_kernel<<<blocks(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, R, Cc, total,
                                                                                         reinterpret_cast<bf16*>(dst));
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_masked_mse(const float* pred, const float* target, const int64_t* length, int B, int T, int F,
                                      float* partial, unsigned* counter, float* loss, void* stream) {
  if (!pred || !target || !length || !partial || !counter || !loss || B <= 0 || T <= 0 || F <= 0) return MDM_ERR_ARG;
  masked_mse_kernel<<<B, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pred, target, length, T, F, B, partial, counter,
                                                                         loss);
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_transpose_split_bf16(const void* src, long M, int Cc, int S, int Ks, void* dst, void* stream) {
  if (!src || !dst || M <= 0 || Cc <= 0 || S <= 0 || Ks <= 0 || (long)S * Ks < M) return MDM_ERR_ARG;
  dim3 grid((Ks + 31) / 32, (Cc + 31) / 32, S);
  transpose_split_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const bf16*>(src), M, Cc, Ks,
                                                                                  reinterpret_cast<bf16*>(dst));
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_sum_partials(const float* part, int S, long n, int accumulate, float* out, void* stream) {
  if (!part || !out || S <= 0) return MDM_ERR_ARG;
  if (n == 0) return MDM_OK;
  if (S >= 32 && n <= (1L << 17))
    sum_partials_tall_kernel<<<(unsigned)((n + 31) / 32), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(part, S, n, accumulate, out);
  else
    sum_partials_kernel<<<blocks(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(part, S, n, accumulate, out);
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_colsum_bf16(const void* src, long M, int Cc, int slabs, float* part, void* stream) {
  if (!src || !part || M <= 0 || Cc <= 0 || slabs <= 0) return MDM_ERR_ARG;
  const int rows_per_blk = (int)((M + slabs - 1) / slabs);
  dim3 grid((Cc + 31) / 32, slabs);
  colsum_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const bf16*>(src), M, Cc, rows_per_blk,
                                                                         part);
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_gelu_fwd(const void* pre, long n, void* h, void* stream) {
  if (!pre || !h) return MDM_ERR_ARG;
  if (n == 0) return MDM_OK;
  gelu_fwd_kernel<<<blocks(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const bf16*>(pre), n,
                                                                               reinterpret_cast<bf16*>(h));
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}
extern "C" MDM_API int mdm_gelu_bwd(const void* pre, const void* dh, long n, void* dp, void* stream) {
  if (!pre || !dh || !dp) return MDM_ERR_ARG;
  if (n == 0) return MDM_OK;
  gelu_bwd_kernel<<<blocks(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(pre), reinterpret_cast<const bf16*>(dh), n, reinterpret_cast<bf16*>(dp));
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}
extern "C" MDM_API int mdm_rowscale_bf16(const void* src, const float* scale, long rows, int Cc, void* dst, void* stream) {
  if (!src || !scale || !dst || Cc <= 0) return MDM_ERR_ARG;
  if (rows == 0) return MDM_OK;
  rowscale_kernel<<<blocks(rows * Cc), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(src), scale, rows, Cc, reinterpret_cast<bf16*>(dst));
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}
extern "C" MDM_API int mdm_seg_colsum_bf16(const void* src, int Cc, const int* seg_off, const int* seg_cnt, int G, float* out,
                                           void* stream) {
  if (!src || !seg_off || !seg_cnt || !out || Cc <= 0 || G <= 0) return MDM_ERR_ARG;
  dim3 grid((Cc + 31) / 32, G);
  seg_colsum_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const bf16*>(src), Cc, seg_off,
                                                                            seg_cnt, out);
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}
