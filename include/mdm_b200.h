/* mdm_b200.h — C-ABI of libmdm_b200.so: the sm_100a kernels behind the MoE motion-denoiser hot path.
 *
 * Drop-in boundary for ltdoanh2004/MotionDiffusion-MoE (reference paths are relative to
 * text2motion/ in that repository).  The reference has no FFI of its own — its hot path is torch
 * eager — so each entry point below names the reference Python code it replaces.  Conventions:
 *   - every function returns an int status (MDM_OK == 0); no allocation, no host synchronisation;
 *   - all pointers are device pointers, tensors are row-major and contiguous unless an `ld` is given;
 *   - `dt` selects the activation/operand element type of `void*` buffers: MDM_F32 or MDM_BF16;
 *     the residual stream, LayerNorm/softmax statistics and all accumulators are always fp32;
 *   - the last argument is the CUDA stream (cudaStream_t passed as void*).
 * The Python host (motiondiffusion_moe_b200/) binds these with ctypes; INTEGRATION.md shows the stub.
 */
#ifndef MDM_B200_H
#define MDM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MDM_API __attribute__((visibility("default")))
#else
#define MDM_API
#endif

#define MDM_OK 0
#define MDM_ERR_ARG 1
#define MDM_ERR_CUDA 2
#define MDM_ERR_UNSUPPORTED 3

#define MDM_F32 0
#define MDM_BF16 1

#define MDM_ACT_NONE 0
#define MDM_ACT_GELU 1    /* exact erf GELU, nn.GELU() default */
#define MDM_ACT_SILU 2
#define MDM_ACT_EXPFEAT 3 /* exp(clamp(v,-15,15))*0.1, models/fast_attention.py:58-66 */

/* GEMM epilogue:  out = alpha * act(acc + bias[n]) * rowscale[m] * rowmask[m] + beta * resid[m'][n],
 * m' = m % resid_mod when resid_mod > 0 (positional embedding), else m. */
typedef struct MdmGemmEpi {
  const float* bias;
  const float* rowscale;
  const float* rowmask;
  const float* resid;
  int ld_resid;
  int resid_mod;
  float alpha, beta;
  int act;
  float* out_f32;
  int ld_f32;
  void* out_bf16; /* secondary output in the operand type: bf16 (mdm_gemm_bf16) / fp32 (mdm_gemm_f32) */
  int ld_bf16;
  int bf16_pre_resid; /* 1: the secondary output omits the residual term */
  int pair_tiles;     /* grouped GEMM only: 1 = tiles 2i and 2i+1 of the table share w_row0 (segments padded
                         to 256 rows), which lets the CTA-pair (cta_group::2) kernel take them together */
  const int* tile_k;  /* grouped GEMM only, optional (device): {k0, klen} per tile - the tile contracts over the columns
                         [k0, k0 + klen) of A and W instead of [0, K) (k0 % 64 == 0; klen > 0, rounded up to 64 with
                         whatever the buffers hold there: pad with zeros).  This is what the weight gradient of an
                         expert needs: dW_e = dY_e^T X_e contracts over the rows of the expert's segment, whose
                         offset and length only exist on the device.  Selects the single-CTA kernel. */
  int mn_major;       /* mdm_gemm_bf16 only: 1 = "TN" contraction C[M, N] = A^T W over the ROWS of A [K rows, >= M columns] and
                         W [K rows, >= N columns] (both row-major, read MN-major by TMA / tcgen05: no transposed copies) - the
                         weight gradient dW = dY^T X of a Linear.  M / N count features, K rows (tokens); a_row0 / w_row0 of a tile
                         are feature offsets, tile_k a row (token) range; fp32 output only. */
} MdmGemmEpi;

/* One 128-row tile of a grouped GEMM: rows [a_row0, a_row0+128) of A times the weight rows starting
 * at w_row0, written to rows [c_row0, c_row0+rows_valid) of C.  bias is indexed bias[w_row0 + n]. */
typedef struct MdmMTile { int a_row0, c_row0, w_row0, rows_valid; } MdmMTile;

/* Fused row pipeline (one warp per row of D elements):
 *   v = in;                         out0_a  = cast(v)
 *   v = LN(v; ln1_w, ln1_b);        if l2norm: v = v / max(||v||, 1e-12) * sqrt(D)
 *                                   out1_f32 / out1_a = v
 *   v = LN(v; ln2_w, ln2_b);        if film: v = v * (1 + scale[b]) + shift[b],  b = row / rows_per_seq
 *   if silu: v = silu(v);           out2_f32 / out2_a = v
 * Stages with null weights are skipped.  Replaces nn.LayerNorm / F.normalize / StylizationBlock's
 * FiLM+SiLU (models/stylization.py:29-30, models/fast_attention.py:142,169-172,210,225,248). */
typedef struct MdmRowOp {
  const void* in;
  int in_dt;
  const float *ln1_w, *ln1_b;
  int l2norm;
  float* out1_f32;
  void* out1_a;
  const float *ln2_w, *ln2_b;
  const float* film; /* [n_seq, 2*D]: scale | shift */
  int rows_per_seq;
  int silu;
  float* out2_f32;
  void* out2_a;
  void* out0_a;
} MdmRowOp;

/* ---- GEMMs: every nn.Linear, expert FFN and k=2,s=2 (transposed) conv of the path ------------- */
/* C = epi(A[rows,K] * W[N,K]^T), bf16 operands, tcgen05/TMEM/TMA.  mtiles == NULL: plain GEMM over
 * M rows.  Otherwise grouped: num_m_tiles entries (or *num_m_tiles_dev if non-NULL, device int).
 * a_rows / w_rows are the total row counts of the A and W allocations (TMA bounds).
 * Replaces F.linear at e.g. models/switch_moe.py:19-25,53,104; models/fast_attention.py:145-147,165. */
MDM_API int mdm_gemm_bf16(const void* A, int lda, long a_rows, const void* W, int ldw, long w_rows,
                          int M, int N, int K, const void* mtiles, int num_m_tiles,
                          const int* num_m_tiles_dev, const MdmGemmEpi* epi, int max_ctas, void* stream);
/* Same contract, fp32 operands, CUDA-core FMA (the reference's fp32 precision mode). */
MDM_API int mdm_gemm_f32(const float* A, int lda, long a_rows, const float* W, int ldw, long w_rows,
                         int M, int N, int K, const void* mtiles, int num_m_tiles,
                         const int* num_m_tiles_dev, const MdmGemmEpi* epi, void* stream);

/* ---- row-wise normalisation / FiLM ----------------------------------------------------------- */
MDM_API int mdm_rowop(const MdmRowOp* op, long rows, int D, int out_dt, void* stream);
/* The row pipeline fused into the GEMM that consumes it (no intermediate in global memory):
 *   out_f32[m, :] = beta * resid[m, :] + alpha * ( pipeline(in[m, :]) . W[N, D]^T + bias )
 * with the stage set of `op` (its outputs are ignored).  The eight epilogue warps build the bf16 A operand of each
 * 128-row tile in shared memory in the tensor core's swizzled K-major layout, W streams by TMA, accumulators in TMEM,
 * fp32 + residual epilogue by TMA.  D == 512, N in {256, 512}, bf16 input rows and weights, fp32 residual / output;
 * stage sets: LN1+L2+LN2+FiLM+SiLU (stylization.py:29-30 after fast_attention.py:169-172) and LN2+FiLM+SiLU.
 * Anything else returns MDM_ERR_UNSUPPORTED (status 3) and the caller uses mdm_rowop + mdm_gemm_bf16. */
MDM_API int mdm_gemm_rowop(const MdmRowOp* op, long rows, int D, const void* W, int ldw, long w_rows, int N,
                           const MdmGemmEpi* epi, void* stream);

/* The row pipeline fused into the EPILOGUE of the GEMM that produces its input (north_star (3): "StylizationBlock
 * timestep-FiLM and LayerNorm fused into the adjacent GEMM epilogues"; csrc/gemm_ln.cu):
 *   y_pre = act(A . W[512, K]^T + bias) * alpha ;  y = y_pre + beta * resid
 *   epi->out_f32 <- y ;  s = epi->bf16_pre_resid ? y_pre : y ;  epi->out_bf16 <- bf16(s)
 *   u = L2norm?(LN1(s)) -> op->out1_f32 | op->out1_a ;  z = SiLU?(FiLM?(LN2(u))) -> op->out2_a
 * (op->in / in_dt / out0_a / out2_f32 are not used: the pipeline's input is the GEMM result, which never leaves the
 * chip unless out_f32 / out_bf16 ask for it).  N == 512 == the whole row in one 128 x 512 TMEM tile, CTA pairs
 * (cta_group::2), K % 64 == 0, bf16 operands.  Stage sets built: the five Linear -> LayerNorm chains of
 * MoEExtendedDecoderLayer (models/fast_attention.py:142,166-176,210-226,248,322-326; models/stylization.py:27-30).
 * Anything else returns MDM_ERR_UNSUPPORTED (status 3) and the caller uses mdm_gemm_bf16 + mdm_rowop. */
MDM_API int mdm_gemm_ln(const void* A, int lda, long a_rows, const void* W, int ldw, long w_rows, int M, int N, int K,
                        const MdmGemmEpi* epi, const MdmRowOp* op, void* stream);

/* mdm_gemm_ln with the MoE GATE as its second pass: the cross-attention output Linear produces the row the gate reads
 * (models/fast_attention.py:270-272 -> models/multi_branch.py:52-55 -> models/switch_moe.py:53-57), so
 *   y = (A . W[512, K]^T + bias) * alpha + beta * resid -> epi->out_f32
 *   per branch br < 2: softmax(LayerNorm_br(y) . gate_w[br]^T + gate_b[br]) -> top-2
 * in one kernel: LayerNorm statistics and the 16 dot products per row are per-thread sums (TMEM lane = row), CTA r of
 * the pair finalises branch r.  Outputs exactly as mdm_moe_gate (idx / vals [M][2][2], stats [M][2] = mean, rstd of the
 * row, per-128-row-block histograms and importance sums: rows of a block counted in token order).  NB == 2, NB * E == 16,
 * N == 512; anything else returns MDM_ERR_UNSUPPORTED and the caller runs mdm_gemm_bf16 + mdm_moe_gate. */
MDM_API int mdm_gemm_gate(const void* A, int lda, long a_rows, const void* W, int ldw, long w_rows, int M, int N, int K,
                          const MdmGemmEpi* epi, int NB, int E, const float* ln_w, const float* ln_b, const float* gate_w,
                          const float* gate_b, int* idx, float* vals, float* stats, int* blk_hist, float* blk_imp,
                          void* stream);



/* ---- FastAttention core, models/fast_attention.py:29-92 (PerformerSelfAttention :155-160) ----
 * qkv: [B*T, 3*H*hd] (q | k | v, already multiplied by nothing: the 0.1 pre-scale of :155-157 is
 * applied inside).  P: [hd, M] fp32 projection_matrix.  norm_w/b: the shared LayerNorm(hd).
 * length[b]: frames t >= length[b] have their key features zeroed (:69-74).  out: [B*T, H*hd]. */
MDM_API int mdm_fastattn(const void* qkv, int dt, const float* P, const float* norm_w,
                         const float* norm_b, const int64_t* length, int length_shift, int B, int H,
                         int T, int hd, int M, void* out, void* stream);
/* Same, with two optional hints that do not change the result:
 *   seq_order [B] int32: a permutation of the sequences by DESCENDING length.  CTAs of long sequences (more
 *     unmasked key windows) are started first, which shortens the tail of the 2-wave grid.  `length` is fixed over
 *     a sampling loop, so the host sorts once per loop (CFGStepper), not per step.
 *   Pt_bf16 [M, hd] bf16: the projection matrix already transposed and rounded (== what the bf16 kernel builds from
 *     P in every CTA); packed once per model.  It selects the tcgen05 kernel (hd == M == 128, T <= 256).  For
 *     hd == M == 64 with an even head count pass the block-diagonal diag(P^T, P^T) [128, 128] instead: one CTA then takes
 *     two adjacent heads through the same 128-wide tcgen05 products (row statistics per 64-column half, the cross-head
 *     blocks of the key-value state zeroed); the [64, 64] matrix selects the mma.sync kernel as before. */
MDM_API int mdm_fastattn_ordered(const void* qkv, int dt, const float* P, const float* norm_w,
                                 const float* norm_b, const int64_t* length, int length_shift, int B, int H,
                                 int T, int hd, int M, void* out, const int* seq_order, const void* Pt_bf16,
                                 void* stream);

/* ---- LinearTemporalCrossAttention, models/fast_attention.py:242-253 --------------------------- */
/* Text side (step-invariant): ctx[b,h,d,l] = sum_n softmax_n(k[b,n,h,d]) * v[b,n,h,l], n < nt[b].
 * k, v: [B, Nt_max, H*hd] of type dt; ctx fp32. */
MDM_API int mdm_lincross_ctx(const void* k, const void* v, int dt, const int* nt, int B, int Nt_max,
                             int H, int hd, float* ctx, void* stream);
/* Motion side: y[t,h,:] = softmax_hd(q[t,h,:]) @ ctx[b,h]. */
MDM_API int mdm_lincross_apply(const void* q, int dt, const float* ctx, int B, int T, int H, int hd,
                               void* y, void* stream);
/* Same, with ctxT_bf16 [B, H, hd(l), hd(d)] = ctx transposed and rounded to bf16 (mdm_transpose_cast_bf16, packed once
 * per sampling loop: ctx is step-invariant): selects the tcgen05 kernel (hd = 128, T <= 256, bf16). */
MDM_API int mdm_lincross_apply_ex(const void* q, int dt, const float* ctx, const void* ctxT_bf16, int B, int T, int H,
                                  int hd, void* y, void* stream);

/* mdm_lincross_apply_ex (tcgen05 kernel) with the StylizationBlock that consumes its output in the epilogue
 * (models/fast_attention.py:248-272 -> models/stylization.py:27-30): y = SiLU(LayerNorm_D(apply(q)) * (1 + scale[b]) + shift[b]),
 * film = [B, 2 * H * hd] (scale | shift).  The LayerNorm spans the H heads of a row: the H CTAs of a sequence run as a
 * thread-block cluster and exchange per-row partial sums through distributed shared memory.  hd == 128, T <= 256,
 * H <= 8, bf16; or hd == 64 with an even H <= 16: ctxT_bf16 is then [B, H / 2, 128, 128], the block-diagonal ctx^T of each
 * pair of adjacent heads (one CTA takes two heads; the same layout selects the tcgen05 kernel in mdm_lincross_apply_ex).
 * MDM_ERR_UNSUPPORTED otherwise (the caller then runs mdm_lincross_apply_ex + mdm_rowop). */
MDM_API int mdm_lincross_apply_style(const void* q, const void* ctxT_bf16, int B, int T, int H, int hd, const float* ln_w,
                                     const float* ln_b, const float* film, void* y, void* stream);

/* dst[n][c][r] (bf16) = src[n][r][c] (fp32). */
MDM_API int mdm_transpose_cast_bf16(const float* src, long n, int R, int C, void* dst, void* stream);

/* ---- MemoryEfficientCrossAttentionBlock core, models/fast_attention.py:305-325 ----------------
 * o[t,h,:] = softmax_n(q[t,h,:]·k[b,n,h,:] * hd^-0.5) @ v[b,n,h,:], n < nt[b] (Nt_max <= 96). */
MDM_API int mdm_softmax_cross(const void* q, const void* k, const void* v, int dt, const int* nt,
                              int B, int T, int Nt_max, int H, int hd, void* o, void* stream);

/* ---- SwitchMoELayer x NB branches, models/switch_moe.py:44-111, models/multi_branch.py:52-61 ---
 * Gate: per token and branch b: h = LN_b(x); probs = softmax(h Wg_b^T + bg_b) (fp32, ATen order);
 * (vals, idx) = top-k(probs) with torch.topk's CUDA tie order.  Writes idx [N,NB,K] int32, vals
 * [N,NB,K] fp32, LN statistics stats [N,2] (mean, rstd) and per-block histograms / importance
 * partial sums: blk_hist [nblk, 2, NB*E] (0: all K slots, 1: top-1 only), blk_imp [nblk, NB*E],
 * nblk = ceil(N/128).  K must be 2 (the reference hard-codes top-2). */
MDM_API int mdm_moe_gate(const float* x, long N, int D, int NB, int E, int K, const float* ln_w,
                         const float* ln_b, const float* gate_w, const float* gate_b, int* idx,
                         float* vals, float* stats, int* blk_hist, float* blk_imp, void* stream);
/* Parity hook (tests): same kernel, but the expert INDICES are taken from forced_idx [N,NB,K] int32 (e.g. the routing
 * the fp32 reference chose for the same tokens, models/switch_moe.py:57) instead of the top-k search; vals are still this
 * kernel's own softmax probabilities of those experts.  Lets a bf16 run be compared with the fp32 reference "with identical
 * routing"  without a flipped near-tie changing a token discontinuously. */
MDM_API int mdm_moe_gate_forced(const float* x, long N, int D, int NB, int E, int K, const float* ln_w,
                                const float* ln_b, const float* gate_w, const float* gate_b, const int* forced_idx,
                                int* idx, float* vals, float* stats, int* blk_hist, float* blk_imp, void* stream);
/* Scan: expert segment offsets (each padded to 128 rows), per-block bases, the two grouped-GEMM tile
 * tables (up: w_row0 = g*F, down: w_row0 = g*D), the tile count, and the usage / importance
 * counters (expert_usage, expert_importance buffers of switch_moe.py:32-34,72-92), updated in place. */
MDM_API int mdm_moe_scan(const int* blk_hist, const float* blk_imp, const int* idx, long N, int NB,
                         int E, int K, int F, int D, int* blk_base, int* seg_offsets, void* tiles_up,
                         void* tiles_down, int* num_tiles, float* usage, float* importance,
                         void* stream);
/* Permute: writes LN_b(x[token]) to its expert-sorted row of xp (type dt), the row position perm
 * [N,NB,K] and rowscale[pos] = vals / NB. */
MDM_API int mdm_moe_permute(const float* x, long N, int D, int NB, int E, int K, const float* ln_w,
                            const float* ln_b, const int* idx, const float* vals, const float* stats,
                            const int* blk_base, const int* seg_offsets, void* xp, int dt, int* perm,
                            float* rowscale, void* stream);
/* Combine + FiLM: m = sum over the NB*K expert rows of a token (already scaled by vals/NB), then
 * out = silu(LN(m; ln_w, ln_b) * (1 + scale[b]) + shift[b])  (multi_branch.py:57-60 +
 * stylization.py:29-30, up to its out Linear). */
MDM_API int mdm_moe_combine_film(const void* yp, int dt, const int* perm, long N, int D, int NBK,
                                 const float* ln_w, const float* ln_b, const float* film,
                                 int rows_per_seq, void* out, void* stream);
/* softmax + top-k alone on given logits [N,E] (parity probe for routing; same device code as the
 * fused gate).  idx64 int64 [N,K], vals/probs fp32. */
MDM_API int mdm_softmax_topk(const float* logits, long N, int E, int K, float* probs, int64_t* idx64,
                             float* vals, void* stream);

/* ---- small per-sequence ops of MotionTransformer.forward, models/transformer.py:318-329 ------- */
/* Sinusoidal timestep embedding, models/time.py:15-26: [cos | sin](t * exp(-ln(1e4) i / (D/2))). */
MDM_API int mdm_timestep_embedding(const int64_t* t, int B, int D, void* out, int dt, void* stream);
/* GatedFusion mix, models/gate.py:18-19: out = sigmoid(t + x) * t + (1 - sigmoid(t + x)) * x. */
MDM_API int mdm_gated_mix(const float* t, const float* x, long n, void* out, int dt, void* stream);
/* x[rows, F] fp32 -> [rows, ld_out] of type dt, zero padded (operand of joint_embed, :324). */
MDM_API int mdm_pad_cast(const float* x, long rows, int F, void* out, int ld_out, int dt, void* stream);

/* ---- GaussianDiffusion, models/gaussian_diffusion.py ------------------------------------------ */
/* One classifier-free-guidance DDPM update (p_sample_with_cfg, :1042-1098) after the two model
 * evaluations: per element, with the fp32 table entries of timestep t[b]:
 *   x0_c = c_recip*x - c_recipm1*eps_c ; x0_u likewise ; x0 = x0_u + s*(x0_c - x0_u)
 *   mean = coef1*x0 + coef2*x ; x_prev = mean + (t != 0) * exp(0.5*logvar) * noise
 * tables: [5, n_steps] fp32 rows = sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod,
 * posterior_mean_coef1, posterior_mean_coef2, posterior_log_variance_clipped.
 * clip != 0 clamps each branch's x0 to [-1, 1] (clip_denoised).  x_prev may be the same buffer as x (in-place step). */
MDM_API int mdm_cfg_update(const float* x, const float* eps_c, const float* eps_u, const float* noise,
                           const int64_t* t, const float* tables, int n_steps, float cfg_scale,
                           int clip, int B, long per_sample, float* x_prev, float* x0, void* stream);
/* p_mean_variance after the model call (EPSILON mean, :538-552 with :554-558 and :462-475) and, when
 * `sample` is non-NULL, the p_sample update (:606-613): x0 = c_recip*x - c_recipm1*eps (clamped to [-1,1]
 * if clip); mean = coef1*x0 + coef2*x; sample = mean + (t != 0)*exp(0.5*logvar)*noise.  Same tables as
 * mdm_cfg_update; mean / x0 / sample may each be NULL. */
MDM_API int mdm_p_mean_variance(const float* x, const float* eps, const float* noise, const int64_t* t,
                                const float* tables, int n_steps, int clip, int B, long per_sample,
                                float* mean, float* x0, float* sample, void* stream);
/* q_sample, :449-460: x_t = sqrt_ac[t]*x0 + sqrt_1mac[t]*noise.  tables2: [2, n_steps]. */
MDM_API int mdm_q_sample(const float* x0, const float* noise, const int64_t* t, const float* tables2,
                         int n_steps, int B, long per_sample, float* x_t, void* stream);

/* DDIM update (ddim_sample, :699-742) after the model call(s): x0 = c_recip*x - c_recipm1*eps_c (clamped if
 * clip); with eps_u != NULL the two x0 are combined as p_sample_with_cfg does (:1074-1079);
 *   eps = (c_recip*x - x0) / c_recipm1                                                  (:567-571)
 *   sigma = eta * sqrt((1 - ab_prev) / (1 - ab)) * sqrt(1 - ab / ab_prev)               (:729-733)
 *   x_prev = x0*sqrt(ab_prev) + sqrt(1 - ab_prev - sigma^2)*eps + (t != 0)*sigma*noise  (:736-741)
 * tables4: [4, n_steps] fp32 rows = sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod, alphas_cumprod,
 * alphas_cumprod_prev.  t_prev == NULL: ab_prev = alphas_cumprod_prev[t] (the reference's full-length loop);
 * otherwise ab_prev = alphas_cumprod[t_prev[b]] (1 when t_prev[b] < 0): strided schedules.  noise may be NULL
 * when eta == 0; eps_u and x0 may be NULL.  Arithmetic in torch-eager order (bit-identical to the reference). */
MDM_API int mdm_ddim_update(const float* x, const float* eps_c, const float* eps_u, const float* noise,
                            const int64_t* t, const int64_t* t_prev, const float* tables4, int n_steps,
                            float cfg_scale, float eta, int clip, int B, long per_sample, float* x_prev,
                            float* x0, void* stream);

/* The step after the sampler in the reference's pipeline: de-normalise the generated features
 * (tools/visualization.py:91: motion * std + mean; mean / std [F] fp32, both NULL = already de-normalised) and
 * recover the joint positions (utils/motion_process.py:362-417 recover_root_rot_pos + recover_from_ric with
 * utils/quaternion.py qinv / qrot).  x [B, T, F] fp32 (F >= 4 + 3*(joints-1): 263 for HumanML3D / 22 joints,
 * 251 for KIT / 21 joints) -> out [B, T, joints, 3] fp32.  One block per sequence, the two cumulative sums over
 * the frames as block scans. */
MDM_API int mdm_recover_from_ric(const float* x, const float* mean, const float* stdv, int B, int T, int F,
                                 int joints, float* out, void* stream);

/* Masked reconstruction loss of DDPMTrainer.backward_G (trainers/ddpm_trainer.py:207-214): mean over features of
 * (pred - target)^2 per frame, summed over the frames t < min(T, length[b]) and divided by the number of such frames
 * (src_mask of models/transformer.py:284-289).  pred / target [B, T, F] fp32; partial [B] fp32 and counter [1] u32
 * (zero before the first call; reset by the kernel) are scratch; loss [1] fp32.  Deterministic. */
MDM_API int mdm_masked_mse(const float* pred, const float* target, const int64_t* length, int B, int T, int F,
                           float* partial, unsigned* counter, float* loss, void* stream);

/* ---- first building blocks of the training step's backward (the full backward is
 * not built): the two gradient GEMMs of a Linear y = x W^T + b (models/*: every nn.Linear) run on mdm_gemm_bf16.
 *   dX = dY . W     -> mdm_gemm_bf16 with W^T [in, out] as the weight (mdm_transpose_split_bf16, S = 1)
 *   dW = dY^T . X   -> contraction over the tokens: operands transposed AND split into S token slabs laid out as the row
 *                      groups of a grouped GEMM (dst[s*C + c][m'] = src[s*Ks + m'][c], zero padded to Ks), fp32 partial
 *                      products [S, out, in] summed by mdm_sum_partials; db = column sums (mdm_colsum_bf16 + sum). */
/* Backward twin of mdm_rowop: `op` describes the FORWARD pipeline (in, in_dt, ln1, l2norm, ln2, film, rows_per_seq, silu;
 * its outputs are ignored); dout / din [rows, D] (grad_dt: MDM_F32 or MDM_BF16) are the gradients of the pipeline's final
 * output and of its input.  The forward intermediates are recomputed in registers.  Parameter gradients come as per-CTA
 * partial sums to be added up by mdm_sum_partials (fixed order, deterministic):
 *   dparam_part [n_param_parts, 4, D]   (ln1_w, ln1_b, ln2_w, ln2_b)
 *   dfilm_part  [n_film_chunks, n_seq, 2*D]   (scale | shift of every sequence)
 * Called with dout == NULL it only reports n_param_parts / n_film_chunks. */
MDM_API int mdm_rowop_bwd(const MdmRowOp* op, long rows, int D, int grad_dt, const void* dout, void* din,
                          float* dparam_part, float* dfilm_part, int* n_param_parts, int* n_film_chunks, void* stream);
/* Generalised form (D in {128, 256, 512, 1024}): dmid (optional, fp32 [rows, D]) is a second gradient arriving at the out1
 * point of the pipeline (after LN1 / L2, before LN2: e.g. the residual-stream gradient of x1 = LN(pre), whose other consumer
 * is LN2); din_flags bit 0 = type of din (MDM_F32 / MDM_BF16), bit 1 = accumulate into din instead of overwriting it. */
MDM_API int mdm_rowop_bwd2(const MdmRowOp* op, long rows, int D, int grad_dt, const void* dout, const void* dmid, int dmid_dt,
                           void* din, float* dparam_part, float* dfilm_part, int din_flags, int* n_param_parts,
                           int* n_film_chunks, void* stream);
/* Elementwise / segment helpers of the expert FFN's backward (models/switch_moe.py:19-25: Linear -> GELU -> Linear per
 * expert): exact-erf GELU forward on the saved pre-activation and its derivative; row scaling by the gate weights; bias
 * gradients as column sums over each expert's row segment [seg_off[g], seg_off[g] + seg_cnt[g]) -> out [G, C] fp32. */
MDM_API int mdm_gelu_fwd(const void* pre, long n, void* h, void* stream);
MDM_API int mdm_gelu_bwd(const void* pre, const void* dh, long n, void* dp, void* stream);
MDM_API int mdm_rowscale_bf16(const void* src, const float* scale, long rows, int C, void* dst, void* stream);
MDM_API int mdm_seg_colsum_bf16(const void* src, int C, const int* seg_off, const int* seg_cnt, int G, float* out,
                                void* stream);
MDM_API int mdm_transpose_split_bf16(const void* src, long M, int C, int S, int Ks, void* dst, void* stream);
MDM_API int mdm_sum_partials(const float* part, int S, long n, int accumulate, float* out, void* stream);
MDM_API int mdm_colsum_bf16(const void* src, long M, int C, int slabs, float* part, void* stream);

/* ---- expert-parallel MoE over NVLink peer memory (BASELINE.json configs[3]) -------------------------
 * The reference has no expert parallelism (experts are a local nn.ModuleList: models/switch_moe.py:
 * 19-25, looped at :97-109); these entry points replace that loop when the E experts of every branch are
 * spread over R ranks of one NVSwitch node (expert e lives on rank e / (E/R)), tokens staying sharded by
 * sequence.  All pointers in MdmEpPeers are valid on the calling GPU: entry [p] is rank p's buffer,
 * mapped with CUDA IPC (mdm_ipc_*); entry [me] is the local buffer.
 *   xp        [cap, D]  expert-sorted LN'ed rows received by rank p     (dt)
 *   rowscale  [cap]     gate weight / NB of each received row           (fp32)
 *   yp        [cap, D]  expert outputs of rank p                        (dt)
 *   cnt       [R, NB*E] rows every rank routes to every group           (int32)
 *   flags     [R]       barrier epochs                                  (uint32)
 * Sequence per MoE call:  mdm_moe_gate -> mdm_ep_counts -> barrier -> mdm_ep_scan -> mdm_ep_dispatch ->
 * barrier -> mdm_gemm_bf16 x2 (local tile tables) -> barrier -> mdm_ep_combine_film. */
#define MDM_EP_MAX_RANKS 8
typedef struct MdmEpPeers {
  void* xp[MDM_EP_MAX_RANKS];
  float* rowscale[MDM_EP_MAX_RANKS];
  void* yp[MDM_EP_MAX_RANKS];
  int* cnt[MDM_EP_MAX_RANKS];
  unsigned* flags[MDM_EP_MAX_RANKS];
} MdmEpPeers;
/* Per-group totals of this rank (from the gate's block histograms) -> row [me] of every peer's cnt table;
 * per-block bases blk_base [nblk, NB*E]; usage / importance counters (switch_moe.py:72-92), per rank as
 * in the reference. */
MDM_API int mdm_ep_counts(const int* blk_hist, const float* blk_imp, long N, int NB, int E, int K, int R,
                          int me, const MdmEpPeers* peers, int* blk_base, float* usage, float* importance,
                          void* stream);
/* From the complete cnt table: dest_base[g] = first row of this rank's rows inside the segment of group g
 * on its owner (segments ordered by local group, rows by source rank, padded to 128); the tile tables and
 * tile count of the groups owned by this rank; *overflow = 1 if a segment exceeds cap. */
MDM_API int mdm_ep_scan(const int* cnt, int NB, int E, int K, int R, int me, int F, int D, int cap,
                        int* dest_base, void* tiles_up, void* tiles_down, int* num_tiles, int* overflow,
                        void* stream);
/* Token dispatch: LN_b(x[token]) -> row of the owning rank's xp (NVLink store), its gate weight / NB ->
 * rowscale; perm[token, b, k] = owner * cap + row. */
MDM_API int mdm_ep_dispatch(const float* x, long N, int D, int NB, int E, int K, int R, int me, int cap,
                            const float* ln_w, const float* ln_b, const int* idx, const float* vals,
                            const float* stats, const int* blk_base, const int* dest_base,
                            const MdmEpPeers* peers, int dt, int* perm, void* stream);
/* Combine: gathers the NB*K expert rows of each token from the owners' yp (NVLink loads), then the same
 * LN + FiLM + SiLU as mdm_moe_combine_film. */
MDM_API int mdm_ep_combine_film(const MdmEpPeers* peers, int dt, const int* perm, long N, int D, int NBK,
                                int cap, const float* ln_w, const float* ln_b, const float* film,
                                int rows_per_seq, void* out, void* stream);
/* Flag barrier of the R ranks in stream order.  *epoch_ctr (device memory, start at 0, private to this
 * rank) is incremented by the kernel, so the call can be captured in a CUDA graph and replayed; every rank
 * must execute the same number of barriers.  If a peer does not arrive within ~2 s, *err is set to 1 and the kernel
 * traps: the stream reports a launch failure at its next synchronisation (a missed barrier is fatal, never silent). */
MDM_API int mdm_ep_barrier(const MdmEpPeers* peers, int R, int me, unsigned* epoch_ctr, int* err, void* stream);
/* CUDA IPC: 64-byte handle of the allocation containing ptr (+ byte offset of ptr inside it); open /
 * close a peer's handle (returns the base of the mapped allocation). */
MDM_API int mdm_ipc_get_handle(const void* ptr, void* handle64, long* offset);
MDM_API int mdm_ipc_open_handle(const void* handle64, void** base);
MDM_API int mdm_ipc_close_handle(void* base);


/* ================================================================================================================
 * Training step (BASELINE.json configs[4]; reference: GaussianDiffusion.training_losses models/gaussian_diffusion.py:923-992,
 * DDPMTrainer.backward_G / update trainers/ddpm_trainer.py:201-244, torch autograd through models/transformer.py:291-361).
 * Every token-level GEMM of the backward pass runs on mdm_gemm_bf16 / mdm_gemm_f32 (dX with the transposed weight, dW as a
 * grouped contraction over token slabs); the entry points below are what those cannot express.  dt arguments select fp32
 * or bf16 activations; parameter gradients are fp32 and deterministic (fixed-order partial sums). */

/* Strided batched GEMM over z = (z1, z2) (e.g. sequence, head):
 *   C[z][m][n] (+)= alpha * sum_k A[z][m][k] * B[z][k][n]
 * element (z, i, j) of X lives at X + z1 * x_z1 + z2 * x_z2 + i * x_rs + j * x_cs.  The small per-head products of the
 * attention cores' backward passes (operands are slices of token-major [N, H*hd] tensors or head-major fp32 scratch).
 * m_limit / k_limit (optional, int64 per z1, shifted right by limit_shift): rows >= limit are written as zero /
 * contraction stops at the limit (sequence lengths). */
typedef struct MdmBgemm {
  const void* A; int a_dt; long a_z1, a_z2, a_rs, a_cs;
  const void* B; int b_dt; long b_z1, b_z2, b_rs, b_cs;
  void* C; int c_dt; long c_z1, c_z2, c_rs, c_cs;
  int Z1, Z2, M, N, K;
  float alpha; int accumulate;
  const int64_t* m_limit; const int64_t* k_limit; int limit_shift;
  int tensor_cores; /* 1: operands rounded to bf16 in shared memory, mma.sync m16n8k16 (bf16 training path); 0: fp32 FMA */
} MdmBgemm;
MDM_API int mdm_bgemm(const MdmBgemm* g, void* stream);
MDM_API int mdm_sizeof_bgemm(void);

/* FastAttention backward (models/fast_attention.py:29-92 with the 0.1 pre-scale of :155-157 and the gradient clamp of
 * :150-152), recomputed from the saved raw qkv [N, 3*H*hd].  Scratch is fp32, head-major [B, H, T, *]:
 *   mdm_fa_prep      qh = L2(LN(0.1 q)), kh = L2(LN(0.1 k)), vn = LN(0.1 v)
 *   (mdm_bgemm)      uq = qh P, uk = kh P
 *   mdm_fa_feat      qp = 0.1 exp(clamp(uq)), kp = 0.1 exp(clamp(uk)) * [t < length >> shift]
 *   (mdm_bgemm)      kv = 0.1 kp^T vn ; o = 0.1 qp kv
 *   mdm_fa_out_bwd   den = max(<qp, kp>, 1e-6); out = LN(o / den): from dout -> d_o (over o), dden; LN affine partials
 *   (mdm_bgemm)      dqp = 0.1 d_o kv^T ; dkv = 0.1 qp^T d_o ; dkp = 0.1 vn dkv^T ; dvn = 0.1 kp dkv
 *   mdm_fa_feat_bwd  duq, duk (in place over dqp, dkp; adds the denominator terms)
 *   (mdm_bgemm)      dqh = duq P^T ; dkh = duk P^T
 *   mdm_fa_prep_bwd  L2 / LN backward, * 0.1, clamp to [-1, 1] -> dqkv [N, 3*H*hd]; LN affine partials
 * Partials are [n_parts, 2, hd] (dw | db), to be summed with mdm_sum_partials; a call with the first pointer NULL only
 * reports n_parts. */
MDM_API int mdm_fa_prep(const void* qkv, int dt, const float* nw, const float* nb, int B, int H, int T, int hd, float* qh,
                        float* kh, float* vn, void* stream);
MDM_API int mdm_fa_feat(const float* uq, const float* uk, const int64_t* length, int shift, int B, int H, int T, int M,
                        float* qp, float* kp, void* stream);
MDM_API int mdm_fa_out_bwd(float* o, const float* qp, const float* kp, const void* dout, int dt, const float* nw, int B, int H,
                           int T, int hd, float* dden, float* part, int* n_parts, void* stream);
/* Output stage FORWARD of the same decomposition: out[t, h, :] = LN(o / max(<qp, kp>, 1e-6)) token-major [N, H*hd].  With
 * mdm_fa_prep / mdm_fa_feat / mdm_bgemm it forms the generic FastAttention forward for head sizes the fused kernels do not
 * cover (hd = 256: model_size="big", models/transformer.py:188-192). */
MDM_API int mdm_fa_out_fwd(const float* o, const float* qp, const float* kp, const float* nw, const float* nb, int B, int H, int T,
                           int hd, void* out, int dt, void* stream);
MDM_API int mdm_fa_feat_bwd(const float* uq, const float* uk, const float* qp, const float* kp, const float* dden, int B, int H,
                            int T, int M, float* dqp, float* dkp, void* stream);
MDM_API int mdm_fa_prep_bwd(const void* qkv, int dt, const float* nw, const float* nb, int B, int H, int T, int hd,
                            const float* dqh, const float* dkh, const float* dvn, void* dqkv, float* part, int* n_parts,
                            void* stream);
/* softmax pieces of the two cross-attention backward passes (fast_attention.py:242-258, 305-325):
 *   head softmax over hd (q of LinearTemporalCrossAttention): token-major q -> head-major P fp32, and its backward;
 *   key softmax over <= 96 text tokens with per-sequence count nt[b] (in place on fp32 [B*H, T, NK]) and its backward;
 *   column softmax over the text tokens of k [B, Nt, C] (:251) and its backward. */
MDM_API int mdm_head_softmax(const void* q, int dt, int B, int H, int T, int hd, float* P, void* stream);
MDM_API int mdm_head_softmax_bwd(const float* P, const float* dP, int B, int H, int T, int hd, void* dq, int dt, void* stream);
MDM_API int mdm_key_softmax(float* S, const int* nt, int B, int H, int T, int NK, void* stream);
MDM_API int mdm_key_softmax_bwd(const float* P, float* dP, int B, int H, int T, int NK, void* stream);
MDM_API int mdm_col_softmax(const void* k, int dt, const int* nt, int B, int Nt, int C, float* Ks, void* stream);
MDM_API int mdm_col_softmax_bwd(const float* Ks, const float* dKs, int B, int Nt, int C, void* dk, int dt, void* stream);
/* MoE (models/switch_moe.py:53-109, multi_branch.py:52-61).  The training forward keeps the un-scaled expert outputs z:
 *   mdm_moe_combine_sum      m[token] = sum_j rowscale[pos_j] * z[pos_j]
 *   mdm_moe_combine_bwd      dz[pos_j] = rowscale[pos_j] * dm[token]; drs[pos_j] = <dm[token], z[pos_j]>
 *   mdm_moe_gate_bwd_logits  vals = probs[idx] (not renormalised): dlogits [N, NB*E] from drs (probabilities recomputed)
 *   mdm_moe_unpermute_bwd    dh_br[token] = d_xp[pos0] + d_xp[pos1] + dlogits[token, br, :] Wg_br
 *   mdm_moe_wgrad_tables     per-group row counts and the MdmGemmEpi.tile_k tables of the experts' weight gradients (an empty
 *                            expert contracts over `zero_row0`, a zero region of the buffers: its gradient is exactly 0) */
MDM_API int mdm_moe_combine_sum(const void* z, int dt, const float* rowscale, const int* perm, long N, int D, int NBK, void* m,
                                void* stream);
MDM_API int mdm_moe_combine_bwd(const void* z, int dt, const float* rowscale, const int* perm, long N, int D, int NBK,
                                const void* dm, void* dz, float* drs, void* stream);
MDM_API int mdm_moe_gate_bwd_logits(const float* x, const float* stats, const float* ln_w, const float* ln_b,
                                    const float* gate_w, const float* gate_b, const int* idx, const int* perm, const float* drs,
                                    long N, int D, int NB, int E, float* dlogits, void* stream);
MDM_API int mdm_moe_unpermute_bwd(const void* dxp, int dt, const int* perm, const float* dlogits, const float* gate_w, long N,
                                  int D, int NB, int E, int br, void* dh, void* stream);
MDM_API int mdm_moe_wgrad_tables(const int* seg_off, const int* idx, long N, int NB, int E, int mt_up, int mt_down,
                                 int zero_row0, int* seg_cnt, int* tile_k_up, int* tile_k_down, void* stream);
/* elementwise / reductions (any dt): activation forward / derivative on a saved pre-activation (GELU exact-erf, SiLU);
 * out = a x + b y; GatedFusion mix backward (models/gate.py:18-19); gradient of the masked reconstruction loss
 * (ddpm_trainer.py:207-214) times `scale`; column sums (bias gradients) as [slabs, C] partials; column sums of
 * a * (b - c); tiled transpose with token-slab split (see mdm_transpose_split_bf16); per-segment column sums. */
MDM_API int mdm_act_fwd(const void* pre, int dt, long n, int act, void* out, void* stream);
MDM_API int mdm_act_bwd(const void* pre, const void* dy, int dt, long n, int act, void* dx, void* stream);
MDM_API int mdm_axpby(const void* x, int x_dt, float a, const void* y, int y_dt, float b, long n, void* out, int out_dt,
                      void* stream);
MDM_API int mdm_gated_mix_bwd(const float* t, const float* x, const float* dout, long n, float* dt_, float* dx, void* stream);
MDM_API int mdm_masked_mse_grad(const float* pred, const float* target, const int64_t* length, int B, int T, int F, float scale,
                                float* dpred, void* stream);
MDM_API int mdm_colsum(const void* src, int dt, long M, int C, long ld, int slabs, float* part, void* stream);
MDM_API int mdm_colsum_prod(const float* a, const float* b, const float* c, long M, int C, int slabs, float* part, void* stream);
MDM_API int mdm_transpose_split(const void* src, int dt, long M, int C, long ld, int S, int Ks, void* dst, void* stream);
MDM_API int mdm_seg_colsum(const void* src, int dt, int C, const int* seg_off, const int* seg_cnt, int G, int slabs, float* out,
                           void* stream);      /* out: [slabs, G, C] partials over `slabs` row slabs of every segment */
/* clip_grad_norm_(max_norm) + Adam (ddpm_trainer.py:228-244; torch.optim.Adam defaults, no weight decay) on flat fp32
 * buffers, no host synchronisation: mdm_grad_clip_coef writes norm_coef = {global L2 norm, min(1, max_norm / (norm + 1e-6))}
 * (part: n_part floats of scratch); mdm_adam_step scales the gradient by norm_coef[1] in place (as clip_grad_norm_ does)
 * and applies bias-corrected Adam step number `step` (1-based). */
MDM_API int mdm_grad_clip_coef(const float* g, long n, float max_norm, float* part, int n_part, float* norm_coef, void* stream);
MDM_API int mdm_adam_step(float* p, float* g, float* m, float* v, long n, float lr, float beta1, float beta2, float eps, int step,
                          const float* norm_coef, void* bf16_mirror /* optional: bf16 copy of p, written in the same pass */,
                          void* stream);

MDM_API int mdm_num_sms(void);
/* Programmatic dependent launch (csrc/common.cuh): the kernels of the hot path are launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization so that the prologue of kernel N+1 overlaps the tail of kernel N
 * (every kernel executes griddepcontrol.wait before its first global-memory access).  on = 1 / 0 switches the launch
 * attribute, on < 0 only queries; returns the previous setting.  Default: on (environment MDM_B200_PDL=0 turns it off).
 * No counterpart in the reference (torch launches every op in plain stream order). */
MDM_API int mdm_set_pdl(int on);
/* sizeof() of the structs above as this library was compiled: a binding checks its own struct definitions against them
 * (a short MdmGemmEpi would make the kernel read garbage as the tile_k device pointer). */
MDM_API int mdm_sizeof_gemm_epi(void);
MDM_API int mdm_sizeof_rowop(void);
MDM_API int mdm_sizeof_ep_peers(void);
MDM_API const char* mdm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MDM_B200_H */
