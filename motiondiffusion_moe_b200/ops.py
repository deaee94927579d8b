"""Thin Python wrappers over the C-ABI (include/mdm_b200.h).  torch is used only for device memory
and the current stream; every computation below runs in libmdm_b200.so.  No fallbacks."""
import ctypes as C

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_GELU, ACT_SILU, ACT_EXPFEAT, MDM_F32, MDM_BF16  # noqa: F401


def _dt(t):
    if t.dtype == torch.float32:
        return MDM_F32
    if t.dtype == torch.bfloat16:
        return MDM_BF16
    raise _lib.MdmError("unsupported dtype %s" % t.dtype)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _dev_check(t):
    """The C library launches on the CURRENT device and stream: a tensor that lives on another GPU would be reached
    through peer access (or fault).  The model / stepper entry points select the tensors' device
    (torch.cuda.device(...)); a direct ops call from the wrong device is refused."""
    if t.device.index != torch.cuda.current_device():
        raise _lib.MdmError("tensor on %s but the current CUDA device is cuda:%d: wrap the call in "
                            "torch.cuda.device(tensor.device)" % (t.device, torch.cuda.current_device()))


def _req_cuda(*ts):
    first = True
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.MdmError("mdm_b200 kernels need CUDA tensors; there is no CPU path")
        if t is not None and first:
            _dev_check(t)
            first = False


def _c(*ts):
    """The C-ABI takes raw pointers: every tensor must be CUDA (on the current device) and dense row-major."""
    first = True
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.MdmError("mdm_b200 kernels need CUDA tensors; there is no CPU path")
        if first:
            _dev_check(t)
            first = False
        if not t.is_contiguous():
            raise _lib.MdmError("non-contiguous tensor (shape %s, strides %s) passed to a raw-pointer kernel"
                                % (tuple(t.shape), t.stride()))


def gemm(A, W, bias=None, *, act=ACT_NONE, alpha=1.0, beta=0.0, resid=None, resid_mod=0, rowscale=None,
         rowmask=None, out_f32=None, out_a=None, a_pre_resid=False, N=None, M=None, tiles=None,
         num_tiles=0, num_tiles_dev=None, a_rows=None, w_rows=None, pair_tiles=False, tile_k=None):
    """C = epi(A @ W^T).  A [rows,K] (row stride may exceed K), W [w_rows,K]; both bf16 (tcgen05 path)
    or both fp32.  out_f32: fp32 output; out_a: secondary output in the operand dtype."""
    _req_cuda(A, W, out_f32, out_a, resid)
    _c(bias, rowscale, rowmask, tiles, num_tiles_dev, tile_k)
    for t_ in (A, W, out_f32, out_a, resid):
        if t_ is not None and (t_.dim() != 2 or t_.stride(1) != 1):
            raise _lib.MdmError("GEMM operands/outputs must be 2-D with unit inner stride")
    lib = _lib.load()
    K = A.shape[1]
    M = A.shape[0] if M is None else M
    N = W.shape[0] if N is None else N
    e = _lib.GemmEpi()
    e.bias, e.rowscale, e.rowmask, e.resid = _ptr(bias), _ptr(rowscale), _ptr(rowmask), _ptr(resid)
    e.ld_resid = resid.stride(0) if resid is not None else 0
    e.resid_mod = resid_mod
    e.alpha, e.beta, e.act = alpha, beta, act
    if out_f32 is not None:
        e.out_f32, e.ld_f32 = out_f32.data_ptr(), out_f32.stride(0)
    if out_a is not None:
        e.out_bf16, e.ld_bf16 = out_a.data_ptr(), out_a.stride(0)
    e.bf16_pre_resid = 1 if a_pre_resid else 0
    e.pair_tiles = 1 if pair_tiles else 0
    e.tile_k = _ptr(tile_k)
    a_rows = A.shape[0] if a_rows is None else a_rows
    w_rows = W.shape[0] if w_rows is None else w_rows
    if A.dtype == torch.bfloat16:
        st = lib.mdm_gemm_bf16(A.data_ptr(), A.stride(0), a_rows, W.data_ptr(), W.stride(0), w_rows, M, N, K,
                               _ptr(tiles), num_tiles, _ptr(num_tiles_dev), C.byref(e), 0, _stream())
    else:
        st = lib.mdm_gemm_f32(A.data_ptr(), A.stride(0), a_rows, W.data_ptr(), W.stride(0), w_rows, M, N, K,
                              _ptr(tiles), num_tiles, _ptr(num_tiles_dev), C.byref(e), _stream())
    _lib.check(st, "mdm_gemm")


def rowop(x, rows, D, out_dt, *, ln1=None, l2norm=False, out1_f32=None, out1_a=None, ln2=None, film=None,
          rows_per_seq=0, silu=False, out2_f32=None, out2_a=None, out0_a=None):
    _c(x, out1_f32, out1_a, out2_f32, out2_a, out0_a, film)
    for pair in (ln1, ln2):
        if pair is not None:
            _c(*pair)
    op = _lib.RowOp()
    op.inp, op.in_dt = x.data_ptr(), _dt(x)
    if ln1 is not None:
        op.ln1_w, op.ln1_b = ln1[0].data_ptr(), ln1[1].data_ptr()
    op.l2norm = 1 if l2norm else 0
    op.out1_f32, op.out1_a = _ptr(out1_f32), _ptr(out1_a)
    if ln2 is not None:
        op.ln2_w, op.ln2_b = ln2[0].data_ptr(), ln2[1].data_ptr()
    op.film, op.rows_per_seq, op.silu = _ptr(film), rows_per_seq, 1 if silu else 0
    op.out2_f32, op.out2_a, op.out0_a = _ptr(out2_f32), _ptr(out2_a), _ptr(out0_a)
    _lib.check(_lib.load().mdm_rowop(C.byref(op), rows, D, out_dt, _stream()), "mdm_rowop")


def gemm_rowop(x, rows, D, W, bias, *, ln1=None, l2norm=False, ln2=None, film=None, rows_per_seq=0, silu=False,
               out_f32=None, resid=None, alpha=1.0, beta=1.0):
    """out_f32 = beta * resid + alpha * (rowop(x) @ W^T + bias) with the row pipeline of `rowop` fused into the GEMM's
    A-operand construction (mdm_gemm_rowop: D == 512, N in {256, 512}, bf16 weights).  Raises MdmError with status
    Returns False (nothing launched) for shapes / stage sets outside the fused kernel: the caller then runs rowop + gemm."""
    _c(x, film, W, bias, out_f32, resid)
    for pair in (ln1, ln2):
        if pair is not None:
            _c(*pair)
    op = _lib.RowOp()
    op.inp, op.in_dt = x.data_ptr(), _dt(x)
    if ln1 is not None:
        op.ln1_w, op.ln1_b = ln1[0].data_ptr(), ln1[1].data_ptr()
    op.l2norm = 1 if l2norm else 0
    if ln2 is not None:
        op.ln2_w, op.ln2_b = ln2[0].data_ptr(), ln2[1].data_ptr()
    op.film, op.rows_per_seq, op.silu = _ptr(film), rows_per_seq, 1 if silu else 0
    e = _lib.GemmEpi()
    e.bias, e.resid = _ptr(bias), _ptr(resid)
    e.ld_resid = resid.stride(0) if resid is not None else 0
    e.alpha, e.beta, e.act = alpha, beta, ACT_NONE
    e.out_f32, e.ld_f32 = out_f32.data_ptr(), out_f32.stride(0)
    st = _lib.load().mdm_gemm_rowop(C.byref(op), rows, D, W.data_ptr(), W.stride(0), W.shape[0], W.shape[0], C.byref(e),
                                    _stream())
    if st == 3:             # MDM_ERR_UNSUPPORTED: shape / stage set outside the fused kernel
        return False
    _lib.check(st, "mdm_gemm_rowop")
    return True


def gemm_ln(A, W, bias, *, ln1, act=ACT_NONE, alpha=1.0, beta=0.0, resid=None, out_f32=None, out_a=None, ln_pre_resid=False,
            l2norm=False, out1_f32=None, out1_a=None, ln2=None, film=None, rows_per_seq=0, silu=False, out2_a=None):
    """Linear + the row pipeline that follows it in ONE kernel (mdm_gemm_ln, csrc/gemm_ln.cu): N == 512 rows stay in
    TMEM between the GEMM and the LayerNorm passes.  y = act(A @ W^T + bias) * alpha + beta * resid -> out_f32;
    s = y (or the pre-residual value with ln_pre_resid) -> out_a (bf16 copy); u = L2norm?(LN1(s)) -> out1_f32 / out1_a;
    z = SiLU?(FiLM?(LN2(u))) -> out2_a.  Returns False (nothing launched) for shapes / stage sets outside the fused
    kernel: the caller then runs gemm + rowop."""
    _req_cuda(A, W, out_f32, out_a, resid)
    _c(bias, film, out1_f32, out1_a, out2_a)
    for pair in (ln1, ln2):
        if pair is not None:
            _c(*pair)
    if A.dtype != torch.bfloat16 or W.shape[0] != 512 or A.shape[1] % 64 or A.dim() != 2 or A.stride(1) != 1:
        return False
    op = _lib.RowOp()
    op.ln1_w, op.ln1_b = ln1[0].data_ptr(), ln1[1].data_ptr()
    op.l2norm = 1 if l2norm else 0
    op.out1_f32, op.out1_a = _ptr(out1_f32), _ptr(out1_a)
    if ln2 is not None:
        op.ln2_w, op.ln2_b = ln2[0].data_ptr(), ln2[1].data_ptr()
    op.film, op.rows_per_seq, op.silu = _ptr(film), rows_per_seq, 1 if silu else 0
    op.out2_a = _ptr(out2_a)
    e = _lib.GemmEpi()
    e.bias, e.resid = _ptr(bias), _ptr(resid)
    e.ld_resid = resid.stride(0) if resid is not None else 0
    e.alpha, e.beta, e.act = alpha, beta, act
    if out_f32 is not None:
        e.out_f32, e.ld_f32 = out_f32.data_ptr(), out_f32.stride(0)
    if out_a is not None:
        e.out_bf16, e.ld_bf16 = out_a.data_ptr(), out_a.stride(0)
    e.bf16_pre_resid = 1 if ln_pre_resid else 0
    st = _lib.load().mdm_gemm_ln(A.data_ptr(), A.stride(0), A.shape[0], W.data_ptr(), W.stride(0), W.shape[0], A.shape[0],
                                 W.shape[0], A.shape[1], C.byref(e), C.byref(op), _stream())
    if st == 3:             # MDM_ERR_UNSUPPORTED
        return False
    _lib.check(st, "mdm_gemm_ln")
    return True


def gemm_gate(A, W, bias, *, resid, out_f32, alpha=1.0, beta=1.0, NB, E, ln_w, ln_b, gate_w, gate_b, idx, vals, stats,
              blk_hist, blk_imp):
    """Linear + residual AND the MoE gate of the row it produces in one kernel (mdm_gemm_gate): out_f32 = alpha * (A @ W^T +
    bias) + beta * resid; idx / vals / stats / blk_hist / blk_imp as moe_gate(out_f32, ...) would write them.  Returns
    False (nothing launched) outside the fused kernel's shapes: the caller then runs gemm + moe_gate."""
    _req_cuda(A, W, out_f32, resid)
    _c(bias, ln_w, ln_b, gate_w, gate_b, idx, vals, stats, blk_hist, blk_imp)
    if (A.dtype != torch.bfloat16 or W.shape[0] != 512 or A.shape[1] % 64 or A.dim() != 2 or A.stride(1) != 1 or NB != 2
            or NB * E != 16):
        return False
    e = _lib.GemmEpi()
    e.bias, e.resid, e.ld_resid = _ptr(bias), resid.data_ptr(), resid.stride(0)
    e.alpha, e.beta, e.act = alpha, beta, ACT_NONE
    e.out_f32, e.ld_f32 = out_f32.data_ptr(), out_f32.stride(0)
    st = _lib.load().mdm_gemm_gate(A.data_ptr(), A.stride(0), A.shape[0], W.data_ptr(), W.stride(0), W.shape[0], A.shape[0],
                                   W.shape[0], A.shape[1], C.byref(e), NB, E, ln_w.data_ptr(), ln_b.data_ptr(),
                                   gate_w.data_ptr(), gate_b.data_ptr(), idx.data_ptr(), vals.data_ptr(), stats.data_ptr(),
                                   blk_hist.data_ptr(), blk_imp.data_ptr(), _stream())
    if st == 3:
        return False
    _lib.check(st, "mdm_gemm_gate")
    return True


def pack_fastattn_pt(P):
    """The bf16 operand that selects the tcgen05 FastAttention kernel: P^T for head size 128; for head size 64 the
    block-diagonal diag(P^T, P^T) [128, 128] (one CTA takes two heads); None where only the other kernels apply."""
    hd, M = P.shape
    Pt = P.float().t().contiguous().to(torch.bfloat16)
    if hd == 128 and M == 128:
        return Pt
    if hd == 64 and M == 64:
        return torch.block_diag(Pt, Pt).contiguous()
    return None


def fastattn(qkv, P, norm_w, norm_b, length, length_shift, B, H, T, hd, out, seq_order=None, Pt=None):
    _c(qkv, P, norm_w, norm_b, length, out, seq_order, Pt)
    if Pt is not None:
        two = tuple(P.shape) == (64, 64) and tuple(Pt.shape) == (128, 128)      # diag(P^T, P^T): pack_fastattn_pt
        if Pt.dtype != torch.bfloat16 or not (two or tuple(Pt.shape) == (P.shape[1], P.shape[0])):
            raise _lib.MdmError("Pt must be P^T in bf16 (pack_fastattn_pt)")
    if hd > 128:       # head sizes outside the fused kernels (model_size="big": 4 heads of 256): the composed generic path
        from . import train_ops
        return train_ops.fastattn_generic(qkv, P, norm_w, norm_b, length, length_shift, B, H, T, hd, out)
    _lib.check(_lib.load().mdm_fastattn_ordered(qkv.data_ptr(), _dt(qkv), P.data_ptr(), norm_w.data_ptr(),
                                                norm_b.data_ptr(), _ptr(length), length_shift, B, H, T, hd,
                                                P.shape[1], out.data_ptr(), _ptr(seq_order), _ptr(Pt), _stream()),
               "mdm_fastattn")


def lincross_ctx(k, v, nt, B, Nt_max, H, hd, ctx):
    _c(k, v, nt, ctx)
    if hd > 128:
        from . import train_ops
        return train_ops.lincross_ctx_generic(k, v, nt, B, Nt_max, H, hd, ctx)
    _lib.check(_lib.load().mdm_lincross_ctx(k.data_ptr(), v.data_ptr(), _dt(k), _ptr(nt), B, Nt_max, H, hd,
                                            ctx.data_ptr(), _stream()), "mdm_lincross_ctx")


def pack_lincross_ctxT(ctx):
    """The bf16 B operand of the tcgen05 linear cross-attention kernel from the fp32 state ctx [B, H, d, l]: ctx^T
    [B, H, l, d] for head size 128; for head size 64 (even head count) the block-diagonal [B, H / 2, 128, 128] of each pair
    of heads (one CTA takes two heads); None where only the other kernels apply."""
    B, H, hd, _ = ctx.shape
    if hd == 128:
        cT = torch.empty(B, H, hd, hd, dtype=torch.bfloat16, device=ctx.device)
        transpose_cast_bf16(ctx, cT)
        return cT
    if hd == 64 and H % 2 == 0:
        ct = ctx.transpose(-1, -2).to(torch.bfloat16)                    # [B, H, l, d]
        out = torch.zeros(B, H // 2, 128, 128, dtype=torch.bfloat16, device=ctx.device)
        out[:, :, :64, :64] = ct[:, 0::2]
        out[:, :, 64:, 64:] = ct[:, 1::2]
        return out
    return None


def _lincross_ctxT_ok(ctxT, B, H, hd):
    if ctxT is None:
        return False
    want = (B, H, 128, 128) if hd == 128 else ((B, H // 2, 128, 128) if (hd == 64 and H % 2 == 0) else None)
    return want is not None and tuple(ctxT.shape) == want and ctxT.dtype == torch.bfloat16


def lincross_apply_style(q, ctxT, B, T, H, hd, ln, film, y):
    """lincross_apply + the StylizationBlock's LayerNorm, FiLM and SiLU in its epilogue (mdm_lincross_apply_style: the H
    head-CTAs of a sequence as a cluster).  Returns False (nothing launched) outside the kernel's shapes."""
    _c(q, ctxT, y, film, *ln)
    if q.dtype != torch.bfloat16 or not _lincross_ctxT_ok(ctxT, B, H, hd) or T > 256 or H // (2 if hd == 64 else 1) > 8:
        return False
    st = _lib.load().mdm_lincross_apply_style(q.data_ptr(), ctxT.data_ptr(), B, T, H, hd, ln[0].data_ptr(), ln[1].data_ptr(),
                                              film.data_ptr(), y.data_ptr(), _stream())
    if st == 3:
        return False
    _lib.check(st, "mdm_lincross_apply_style")
    return True


def lincross_apply(q, ctx, B, T, H, hd, y, ctxT=None):
    _c(q, ctx, y, ctxT)
    if hd > 128:
        from . import train_ops
        return train_ops.lincross_apply_generic(q, ctx, B, T, H, hd, y)
    if not _lincross_ctxT_ok(ctxT, B, H, hd):        # (a [B, H, 64, 64] transpose is not what the two-head kernel reads)
        ctxT = None
    _lib.check(_lib.load().mdm_lincross_apply_ex(q.data_ptr(), _dt(q), ctx.data_ptr(), _ptr(ctxT), B, T, H, hd,
                                                 y.data_ptr(), _stream()), "mdm_lincross_apply")


def transpose_cast_bf16(src, dst):
    """dst[n][c][r] (bf16) = src[n][r][c] (fp32); src [..., R, C] contiguous."""
    _c(src, dst)
    R, Cc = src.shape[-2], src.shape[-1]
    _lib.check(_lib.load().mdm_transpose_cast_bf16(src.data_ptr(), src.numel() // (R * Cc), R, Cc, dst.data_ptr(),
                                                   _stream()), "mdm_transpose_cast_bf16")


def softmax_cross(q, k, v, nt, B, T, Nt_max, H, hd, o):
    _c(q, k, v, nt, o)
    if hd > 128:
        from . import train_ops
        return train_ops.softmax_cross_generic(q, k, v, nt, B, T, Nt_max, H, hd, o)
    _lib.check(_lib.load().mdm_softmax_cross(q.data_ptr(), k.data_ptr(), v.data_ptr(), _dt(q), _ptr(nt), B, T,
                                             Nt_max, H, hd, o.data_ptr(), _stream()), "mdm_softmax_cross")


def moe_gate(x, N, D, NB, E, ln_w, ln_b, gate_w, gate_b, idx, vals, stats, blk_hist, blk_imp, forced_idx=None):
    _c(x, ln_w, ln_b, gate_w, gate_b, idx, vals, stats, blk_hist, blk_imp, forced_idx)
    if forced_idx is not None:
        if forced_idx.dtype != torch.int32 or tuple(forced_idx.shape) != (N, NB, 2):
            raise _lib.MdmError("forced_idx must be int32 [N, NB, 2]")
        _lib.check(_lib.load().mdm_moe_gate_forced(x.data_ptr(), N, D, NB, E, 2, ln_w.data_ptr(), ln_b.data_ptr(),
                                                   gate_w.data_ptr(), gate_b.data_ptr(), forced_idx.data_ptr(),
                                                   idx.data_ptr(), vals.data_ptr(), stats.data_ptr(),
                                                   blk_hist.data_ptr(), blk_imp.data_ptr(), _stream()),
                   "mdm_moe_gate_forced")
        return
    _lib.check(_lib.load().mdm_moe_gate(x.data_ptr(), N, D, NB, E, 2, ln_w.data_ptr(), ln_b.data_ptr(),
                                        gate_w.data_ptr(), gate_b.data_ptr(), idx.data_ptr(), vals.data_ptr(),
                                        stats.data_ptr(), blk_hist.data_ptr(), blk_imp.data_ptr(), _stream()),
               "mdm_moe_gate")


def moe_scan(blk_hist, blk_imp, idx, N, NB, E, F, D, blk_base, seg_offsets, tiles_up, tiles_down, num_tiles,
             usage, importance):
    _c(blk_hist, blk_imp, idx, blk_base, seg_offsets, tiles_up, tiles_down, num_tiles, usage, importance)
    _lib.check(_lib.load().mdm_moe_scan(blk_hist.data_ptr(), blk_imp.data_ptr(), idx.data_ptr(), N, NB, E, 2, F,
                                        D, blk_base.data_ptr(), seg_offsets.data_ptr(), tiles_up.data_ptr(),
                                        tiles_down.data_ptr(), num_tiles.data_ptr(), _ptr(usage),
                                        _ptr(importance), _stream()), "mdm_moe_scan")


def moe_permute(x, N, D, NB, E, ln_w, ln_b, idx, vals, stats, blk_base, seg_offsets, xp, perm, rowscale):
    _c(x, ln_w, ln_b, idx, vals, stats, blk_base, seg_offsets, xp, perm, rowscale)
    _lib.check(_lib.load().mdm_moe_permute(x.data_ptr(), N, D, NB, E, 2, ln_w.data_ptr(), ln_b.data_ptr(),
                                           idx.data_ptr(), vals.data_ptr(), stats.data_ptr(), blk_base.data_ptr(),
                                           seg_offsets.data_ptr(), xp.data_ptr(), _dt(xp), perm.data_ptr(),
                                           rowscale.data_ptr(), _stream()), "mdm_moe_permute")


def moe_combine_film(yp, perm, N, D, NBK, ln_w, ln_b, film, rows_per_seq, out):
    _c(yp, perm, ln_w, ln_b, film, out)
    _lib.check(_lib.load().mdm_moe_combine_film(yp.data_ptr(), _dt(yp), perm.data_ptr(), N, D, NBK,
                                                ln_w.data_ptr(), ln_b.data_ptr(), film.data_ptr(), rows_per_seq,
                                                out.data_ptr(), _stream()), "mdm_moe_combine_film")


def softmax_topk(logits):
    """Routing probe: (probs, idx int64 [N,2], vals) with the fused gate's device code."""
    _c(logits)
    N, E = logits.shape
    probs = torch.empty_like(logits)
    idx = torch.empty(N, 2, dtype=torch.int64, device=logits.device)
    vals = torch.empty(N, 2, dtype=torch.float32, device=logits.device)
    _lib.check(_lib.load().mdm_softmax_topk(logits.data_ptr(), N, E, 2, probs.data_ptr(), idx.data_ptr(),
                                            vals.data_ptr(), _stream()), "mdm_softmax_topk")
    return probs, idx, vals


def timestep_embedding(t, B, D, out):
    _c(t, out)
    _lib.check(_lib.load().mdm_timestep_embedding(t.data_ptr(), B, D, out.data_ptr(), _dt(out), _stream()),
               "mdm_timestep_embedding")


def gated_mix(t, x, out):
    _c(t, x, out)
    _lib.check(_lib.load().mdm_gated_mix(t.data_ptr(), x.data_ptr(), t.numel(), out.data_ptr(), _dt(out),
                                         _stream()), "mdm_gated_mix")


def pad_cast(x, rows, F, out):
    _c(x)
    _req_cuda(out)
    _lib.check(_lib.load().mdm_pad_cast(x.data_ptr(), rows, F, out.data_ptr(), out.stride(0), _dt(out), _stream()),
               "mdm_pad_cast")


def cfg_update(x, eps_c, eps_u, noise, t, tables, n_steps, cfg_scale, clip, x_prev, x0=None):
    _c(x, eps_c, eps_u, noise, t, tables, x_prev, x0)
    B = x.shape[0]
    _lib.check(_lib.load().mdm_cfg_update(x.data_ptr(), eps_c.data_ptr(), eps_u.data_ptr(), noise.data_ptr(),
                                          t.data_ptr(), tables.data_ptr(), n_steps, float(cfg_scale),
                                          1 if clip else 0, B, x.numel() // B, x_prev.data_ptr(), _ptr(x0),
                                          _stream()), "mdm_cfg_update")


def p_mean_variance(x, eps, t, tables, n_steps, clip, mean=None, x0=None, noise=None, sample=None):
    _c(x, eps, t, tables, mean, x0, noise, sample)
    B = x.shape[0]
    _lib.check(_lib.load().mdm_p_mean_variance(x.data_ptr(), eps.data_ptr(), _ptr(noise), t.data_ptr(), tables.data_ptr(),
                                               n_steps, 1 if clip else 0, B, x.numel() // B, _ptr(mean), _ptr(x0),
                                               _ptr(sample), _stream()), "mdm_p_mean_variance")


def ddim_update(x, eps_c, t, tables4, n_steps, eta, clip, x_prev, *, eps_u=None, cfg_scale=1.0, noise=None,
                t_prev=None, x0=None):
    _c(x, eps_c, eps_u, noise, t, t_prev, tables4, x_prev, x0)
    B = x.shape[0]
    _lib.check(_lib.load().mdm_ddim_update(x.data_ptr(), eps_c.data_ptr(), _ptr(eps_u), _ptr(noise), t.data_ptr(),
                                           _ptr(t_prev), tables4.data_ptr(), n_steps, cfg_scale, eta, 1 if clip else 0,
                                           B, x.numel() // B, x_prev.data_ptr(), _ptr(x0), _stream()),
               "mdm_ddim_update")


def recover_from_ric(x, joints, out, mean=None, std=None):
    _c(x, out, mean, std)
    B, T, F = x.shape
    _lib.check(_lib.load().mdm_recover_from_ric(x.data_ptr(), _ptr(mean), _ptr(std), B, T, F, joints, out.data_ptr(),
                                                _stream()), "mdm_recover_from_ric")


def masked_mse(pred, target, length, partial, counter, loss):
    _c(pred, target, length, partial, counter, loss)
    B, T, F = pred.shape
    _lib.check(_lib.load().mdm_masked_mse(pred.data_ptr(), target.data_ptr(), length.data_ptr(), B, T, F,
                                          partial.data_ptr(), counter.data_ptr(), loss.data_ptr(), _stream()),
               "mdm_masked_mse")


def q_sample(x0, noise, t, tables2, n_steps, x_t):
    _c(x0, noise, t, tables2, x_t)
    B = x0.shape[0]
    _lib.check(_lib.load().mdm_q_sample(x0.data_ptr(), noise.data_ptr(), t.data_ptr(), tables2.data_ptr(), n_steps,
                                        B, x0.numel() // B, x_t.data_ptr(), _stream()), "mdm_q_sample")


# ------------------------------------------------------------------ backward building blocks (training step, round 2)
_SLAB_TILES = {}

def linear_backward(x, W, dy, *, dx=None, dW=None, db=None, accumulate=False, splits=None):
    """Gradients of y = x @ W^T + b on the tcgen05 GEMM: x [M, in], W [out, in], dy [M, out] (bf16) ->
    dx [M, in] (bf16, optional), dW [out, in] / db [out] (fp32, optional; `accumulate` adds to them).
    dW contracts over the M tokens: both operands are transposed into S token slabs that form the row groups of one
    grouped GEMM (fp32 partial products), so that the small [out, in] result still fills the machine."""
    _c(x, W, dy, dx, dW, db)
    lib = _lib.load()
    M, K_in = x.shape
    N_out = W.shape[0]
    dev = x.device
    if dx is not None:
        Wt = torch.empty(K_in, N_out, dtype=torch.bfloat16, device=dev)        # W^T: the weight of dX = dY . W
        _lib.check(lib.mdm_transpose_split_bf16(W.data_ptr(), N_out, K_in, 1, N_out, Wt.data_ptr(), _stream()),
                   "mdm_transpose_split_bf16")
        gemm(dy, Wt, None, out_a=dx)
    if dW is not None:
        tiles_out = (N_out + 127) // 128
        S = splits or max(1, min(64, (2 * _lib.load().mdm_num_sms()) // max(1, tiles_out * ((K_in + 255) // 256))))
        Ks = ((M + S - 1) // S + 63) // 64 * 64
        rows_a = tiles_out * 128                                                # slab rows padded to the GEMM tile
        dyT = torch.zeros(S * rows_a, Ks, dtype=torch.bfloat16, device=dev) if rows_a != N_out else \
            torch.empty(S * N_out, Ks, dtype=torch.bfloat16, device=dev)
        xT = torch.empty(S * K_in, Ks, dtype=torch.bfloat16, device=dev)
        if rows_a == N_out:
            _lib.check(lib.mdm_transpose_split_bf16(dy.data_ptr(), M, N_out, S, Ks, dyT.data_ptr(), _stream()), "transpose dY")
        else:
            tmp = torch.empty(S * N_out, Ks, dtype=torch.bfloat16, device=dev)
            _lib.check(lib.mdm_transpose_split_bf16(dy.data_ptr(), M, N_out, S, Ks, tmp.data_ptr(), _stream()), "transpose dY")
            dyT.view(S, rows_a, Ks)[:, :N_out].copy_(tmp.view(S, N_out, Ks))
        _lib.check(lib.mdm_transpose_split_bf16(x.data_ptr(), M, K_in, S, Ks, xT.data_ptr(), _stream()), "transpose X")
        key = (S, tiles_out, rows_a, K_in, N_out, str(dev))
        tt = _SLAB_TILES.get(key)
        if tt is None:                                   # one tile table per shape, built once
            rows = [[s_ * rows_a + i * 128, s_ * rows_a + i * 128, s_ * K_in, min(128, N_out - i * 128)]
                    for s_ in range(S) for i in range(tiles_out)]
            tt = _SLAB_TILES[key] = torch.tensor(rows, dtype=torch.int32).to(dev)
        part = torch.empty(S * rows_a, K_in, dtype=torch.float32, device=dev)
        gemm(dyT, xT, None, out_f32=part, N=K_in, M=S * rows_a, tiles=tt, num_tiles=S * tiles_out, a_rows=S * rows_a,
             w_rows=S * K_in)
        if rows_a == N_out:
            _lib.check(lib.mdm_sum_partials(part.data_ptr(), S, N_out * K_in, 1 if accumulate else 0, dW.data_ptr(), _stream()),
                       "mdm_sum_partials")
        else:
            red = torch.empty(rows_a, K_in, dtype=torch.float32, device=dev)
            _lib.check(lib.mdm_sum_partials(part.data_ptr(), S, rows_a * K_in, 0, red.data_ptr(), _stream()), "mdm_sum_partials")
            if accumulate:
                dW.add_(red[:N_out])
            else:
                dW.copy_(red[:N_out])
    if db is not None:
        slabs = 64
        part = torch.empty(slabs, N_out, dtype=torch.float32, device=dev)
        _lib.check(lib.mdm_colsum_bf16(dy.data_ptr(), M, N_out, slabs, part.data_ptr(), _stream()), "mdm_colsum_bf16")
        _lib.check(lib.mdm_sum_partials(part.data_ptr(), slabs, N_out, 1 if accumulate else 0, db.data_ptr(), _stream()),
                   "mdm_sum_partials")


def rowop_bwd(x, rows, D, dout, *, ln1=None, l2norm=False, ln2=None, film=None, rows_per_seq=0, silu=False):
    """Backward of the row pipeline `rowop(x, ...)` (final output only): returns (din, grads) with grads = {"ln1_w", "ln1_b",
    "ln2_w", "ln2_b"} [D] fp32 and "film" [n_seq, 2D] fp32 (those of the stages that are present)."""
    _c(x, dout, film)
    for pair in (ln1, ln2):
        if pair is not None:
            _c(*pair)
    lib = _lib.load()
    op = _lib.RowOp()
    op.inp, op.in_dt = x.data_ptr(), _dt(x)
    if ln1 is not None:
        op.ln1_w, op.ln1_b = ln1[0].data_ptr(), ln1[1].data_ptr()
    op.l2norm = 1 if l2norm else 0
    if ln2 is not None:
        op.ln2_w, op.ln2_b = ln2[0].data_ptr(), ln2[1].data_ptr()
    op.film, op.rows_per_seq, op.silu = _ptr(film), rows_per_seq, 1 if silu else 0
    npp, nfc = C.c_int(0), C.c_int(0)
    st = lib.mdm_rowop_bwd(C.byref(op), rows, D, _dt(dout), None, None, None, None, C.byref(npp), C.byref(nfc), _stream())
    if st != 0:
        raise _lib.MdmError("mdm_rowop_bwd: status %d" % st)
    dev = x.device
    din = torch.empty_like(dout)
    ppart = torch.empty(npp.value, 4, D, dtype=torch.float32, device=dev)
    n_seq = film.shape[0] if film is not None else 1
    fpart = torch.empty(nfc.value, n_seq, 2 * D, dtype=torch.float32, device=dev) if film is not None else None
    _lib.check(lib.mdm_rowop_bwd(C.byref(op), rows, D, _dt(dout), dout.data_ptr(), din.data_ptr(), ppart.data_ptr(),
                                 _ptr(fpart), C.byref(npp), C.byref(nfc), _stream()), "mdm_rowop_bwd")
    psum = torch.empty(4, D, dtype=torch.float32, device=dev)
    _lib.check(lib.mdm_sum_partials(ppart.data_ptr(), npp.value, 4 * D, 0, psum.data_ptr(), _stream()), "mdm_sum_partials")
    grads = {}
    if ln1 is not None:
        grads["ln1_w"], grads["ln1_b"] = psum[0], psum[1]
    if ln2 is not None:
        grads["ln2_w"], grads["ln2_b"] = psum[2], psum[3]
    if film is not None:
        fsum = torch.empty(n_seq, 2 * D, dtype=torch.float32, device=dev)
        _lib.check(lib.mdm_sum_partials(fpart.data_ptr(), nfc.value, n_seq * 2 * D, 0, fsum.data_ptr(), _stream()),
                   "mdm_sum_partials")
        grads["film"] = fsum
    return din, grads


def expert_ffn_backward(xp, pre, W1, W2, rowscale, d_yp, seg_off, seg_cnt, tiles_rows, *, F, D):
    """Backward of the grouped expert FFN  yp = rowscale * (gelu(xp W1_g^T + b1_g) W2_g^T + b2_g)  over expert-sorted rows
    (models/switch_moe.py:97-109 with the gate weight folded into the down-projection).  Inputs, all on the device:
      xp [cap, D] bf16 (padding rows of every segment ZERO), pre [cap, F] bf16 = saved pre-activation,
      W1 [G*F, D], W2 [G*D, F] bf16 stacked expert weights, rowscale [cap] fp32, d_yp [cap, D] bf16,
      seg_off / seg_cnt [G] int32 (128-aligned offsets, valid rows), tiles_rows: list of (row0, group) 128-row tiles.
    Returns d_xp [cap, D] bf16, dW1 [G*F, D], dW2 [G*D, F], db1 [G, F], db2 [G, D] fp32 (the gradient of the gate weights,
    <d_yp[r], z[r]>, belongs to the gate's backward and is not computed here).
    Every product runs on the tcgen05 grouped GEMM: the data gradients with the transposed stacked weights, the weight
    gradients as contractions over each expert's row segment (MdmGemmEpi.tile_k) of the transposed activations."""
    _c(xp, pre, W1, W2, rowscale, d_yp, seg_off, seg_cnt)
    lib = _lib.load()
    dev = xp.device
    cap = xp.shape[0]
    G = seg_off.shape[0]
    bf = torch.bfloat16
    st = _stream

    def tr(src, rows, cols, groups=1):        # [groups*rows, cols] -> [groups*cols, rows] (per group transposed)
        dst = torch.empty(groups * cols, rows, dtype=bf, device=dev)
        for g_ in range(groups):
            _lib.check(lib.mdm_transpose_split_bf16(src[g_ * rows:].data_ptr(), rows, cols, 1, rows, dst[g_ * cols:].data_ptr(),
                                                    st()), "mdm_transpose_split_bf16")
        return dst

    # tile tables (built once per routing on the device by moe_scan in the fused path; here from the host description)
    def table(w_rows_per_group):
        return torch.tensor([[r0, r0, g_ * w_rows_per_group, 128] for r0, g_ in tiles_rows], dtype=torch.int32, device=dev)

    nt = len(tiles_rows)
    # dz = rowscale * d_yp
    dz = torch.empty_like(d_yp)
    _lib.check(lib.mdm_rowscale_bf16(d_yp.data_ptr(), rowscale.data_ptr(), cap, D, dz.data_ptr(), st()), "mdm_rowscale_bf16")
    hp = torch.empty_like(pre)
    _lib.check(lib.mdm_gelu_fwd(pre.data_ptr(), pre.numel(), hp.data_ptr(), st()), "mdm_gelu_fwd")
    # d_hp = dz . W2_g   (weight operand: W2_g^T [F, D] stacked)
    W2t = tr(W2, D, F, G)                                                   # [G*F, D]
    d_hp = torch.empty(cap, F, dtype=bf, device=dev)
    gemm(dz, W2t, None, out_a=d_hp, N=F, M=cap, tiles=table(F), num_tiles=nt, a_rows=cap, w_rows=G * F)
    d_pre = torch.empty_like(pre)
    _lib.check(lib.mdm_gelu_bwd(pre.data_ptr(), d_hp.data_ptr(), pre.numel(), d_pre.data_ptr(), st()), "mdm_gelu_bwd")
    # d_xp = d_pre . W1_g   (weight operand: W1_g^T [D, F] stacked)
    W1t = tr(W1, F, D, G)                                                   # [G*D, F]
    d_xp = torch.empty(cap, D, dtype=bf, device=dev)
    gemm(d_pre, W1t, None, out_a=d_xp, N=D, M=cap, tiles=table(D), num_tiles=nt, a_rows=cap, w_rows=G * D)
    # weight gradients: contraction over each expert's segment
    def wgrad(dy_rows, x_rows, out_dim, in_dim):
        dyT, xT = tr(dy_rows, cap, out_dim), tr(x_rows, cap, in_dim)        # [out, cap], [in, cap]
        mt = (out_dim + 127) // 128
        rows = [[i * 128, g_ * out_dim + i * 128, 0, min(128, out_dim - i * 128)] for g_ in range(G) for i in range(mt)]
        tt = torch.tensor(rows, dtype=torch.int32, device=dev)
        off, cnt = seg_off.tolist(), seg_cnt.tolist()
        tk = torch.tensor([[off[g_], max(cnt[g_], 1)] for g_ in range(G) for _ in range(mt)], dtype=torch.int32, device=dev)
        out = torch.zeros(G * out_dim, in_dim, dtype=torch.float32, device=dev)
        gemm(dyT, xT, None, out_f32=out, N=in_dim, M=G * out_dim, tiles=tt, num_tiles=len(rows), a_rows=out_dim, w_rows=in_dim,
             tile_k=tk)
        empty = [g_ for g_ in range(G) if cnt[g_] == 0]
        for g_ in empty:
            out[g_ * out_dim:(g_ + 1) * out_dim].zero_()
        return out
    dW2 = wgrad(dz, hp, D, F)
    dW1 = wgrad(d_pre, xp, F, D)
    db2 = torch.empty(G, D, dtype=torch.float32, device=dev)
    db1 = torch.empty(G, F, dtype=torch.float32, device=dev)
    _lib.check(lib.mdm_seg_colsum_bf16(dz.data_ptr(), D, seg_off.data_ptr(), seg_cnt.data_ptr(), G, db2.data_ptr(), st()),
               "mdm_seg_colsum_bf16")
    _lib.check(lib.mdm_seg_colsum_bf16(d_pre.data_ptr(), F, seg_off.data_ptr(), seg_cnt.data_ptr(), G, db1.data_ptr(), st()),
               "mdm_seg_colsum_bf16")
    return d_xp, dW1, dW2, db1, db2
