"""Python wrappers of the training-step kernels (csrc/train.cu; C-ABI in include/mdm_b200.h, "Training step").
torch only provides device memory and the stream; every computation runs in libmdm_b200.so.  The composite functions
below (`linear_bwd`, `fastattn_bwd`, `lincross_bwd`, `softmax_cross_bwd`, ...) are the backward twins of the forward
ops in ops.py; each cites the reference code whose autograd graph it reproduces."""
import ctypes as C
import functools

import torch

from . import _lib, ops
from ._lib import MDM_BF16, MDM_F32, ACT_GELU, ACT_SILU, ACT_NONE  # noqa: F401
from .ops import _c, _dt, _ptr, _stream

f32, bf16 = torch.float32, torch.bfloat16


def _chk(st, what):
    _lib.check(st, what)


def bgemm(A, a_str, B, b_str, Cm, c_str, Z1, Z2, M, N, K, alpha=1.0, accumulate=False, tc=False):
    """C[z][m][n] (+)= alpha * sum_k A[z][m][k] B[z][k][n]; *_str = (z1, z2, row, col) element strides.
    tc: True = bf16 tensor-core flavour (operands rounded to bf16 while staged, csrc/train.cu bgemm_tc_kernel: what the bf16
    path's composite functions below pass), False = fp32 FMA (exact; the fp32 parity path)."""
    g = _lib.Bgemm()
    g.tensor_cores = 1 if tc else 0
    g.A, g.a_dt = A.data_ptr(), _dt(A)
    g.a_z1, g.a_z2, g.a_rs, g.a_cs = a_str
    g.B, g.b_dt = B.data_ptr(), _dt(B)
    g.b_z1, g.b_z2, g.b_rs, g.b_cs = b_str
    g.C, g.c_dt = Cm.data_ptr(), _dt(Cm)
    g.c_z1, g.c_z2, g.c_rs, g.c_cs = c_str
    g.Z1, g.Z2, g.M, g.N, g.K = Z1, Z2, M, N, K
    g.alpha, g.accumulate = alpha, 1 if accumulate else 0
    _chk(_lib.load().mdm_bgemm(C.byref(g), _stream()), "mdm_bgemm")


_bgemm = bgemm


def sum_partials(part, S, n, out, accumulate=True):
    _chk(_lib.load().mdm_sum_partials(part.data_ptr(), S, n, 1 if accumulate else 0, out.data_ptr(), _stream()), "mdm_sum_partials")


def act_fwd(pre, act, out=None):
    out = torch.empty_like(pre) if out is None else out
    _chk(_lib.load().mdm_act_fwd(pre.data_ptr(), _dt(pre), pre.numel(), act, out.data_ptr(), _stream()), "mdm_act_fwd")
    return out


def act_bwd(pre, dy, act, dx=None):
    dx = torch.empty_like(pre) if dx is None else dx
    _chk(_lib.load().mdm_act_bwd(pre.data_ptr(), dy.data_ptr(), _dt(pre), pre.numel(), act, dx.data_ptr(), _stream()), "mdm_act_bwd")
    return dx


def axpby(x, a, y, b, out):
    """out = a * x + b * y (y may be None); any mix of fp32 / bf16."""
    _chk(_lib.load().mdm_axpby(x.data_ptr(), _dt(x), float(a), _ptr(y), _dt(y) if y is not None else 0, float(b), out.numel(),
                               out.data_ptr(), _dt(out), _stream()), "mdm_axpby")
    return out


def colsum_into(src, M, Cc, out, ld=None, slabs=64):
    """out[c] += sum_m src[m, c]   (bias gradient)."""
    part = torch.empty(slabs, Cc, dtype=f32, device=src.device)
    _chk(_lib.load().mdm_colsum(src.data_ptr(), _dt(src), M, Cc, ld if ld is not None else src.stride(0), slabs, part.data_ptr(),
                                _stream()), "mdm_colsum")
    sum_partials(part, slabs, Cc, out)


def transpose_groups(src, rows_per_group, groups=1):
    """[groups * R, Cc] -> [groups * Cc, R]: every group's matrix transposed (weights for the dX GEMMs)."""
    R, Cc = rows_per_group, src.shape[1]
    dst = torch.empty(groups * Cc, R, dtype=src.dtype, device=src.device)
    _chk(_lib.load().mdm_transpose_split(src.data_ptr(), _dt(src), groups * R, Cc, src.stride(0), groups, R, dst.data_ptr(),
                                         _stream()), "mdm_transpose_split")
    return dst


_SLAB_TABLES = {}
USE_TN = [True]      # A/B knob (tests, tools): False = transposed copies + K-major GEMM for the weight gradients


def gemm_tn(A, W, out_f32, M, N, tiles, num_tiles, tile_k):
    """out[c_row0 + m, n] = sum_{k in tile_k range} A[k, a_row0 + m] * W[k, w_row0 + n] per tile: the "TN" contraction of
    mdm_gemm_bf16 (MdmGemmEpi.mn_major): A [K rows, >= M cols], W [K rows, >= N cols] bf16 row-major, read MN-major by
    TMA / tcgen05, i.e. the token contraction of a weight gradient without transposed copies of its operands."""
    _c(tiles, tile_k)
    lib = _lib.load()
    e = _lib.GemmEpi()
    e.alpha, e.beta, e.act = 1.0, 0.0, ACT_NONE
    e.out_f32, e.ld_f32 = out_f32.data_ptr(), out_f32.stride(0)
    e.tile_k = tile_k.data_ptr()
    e.mn_major = 1
    _chk(lib.mdm_gemm_bf16(A.data_ptr(), A.stride(0), A.shape[0], W.data_ptr(), W.stride(0), W.shape[0], M, N, A.shape[0],
                           tiles.data_ptr(), num_tiles, None, C.byref(e), 0, _stream()), "mdm_gemm_bf16 (mn_major)")


def linear_bwd(x, W_t, dy, *, dx_a=None, dx_f32=None, dx_resid=None, dW=None, db=None, M=None):
    """Backward of y = x W^T + b (every nn.Linear of the path).  x [M, in], dy [M, out] in the operand type, W_t = W^T
    [in, out] (pre-transposed).  dx_a: operand-typed dX; dx_f32 (+ dx_resid, may alias): fp32 dX (+ residual gradient).
    dW [out, in] / db [out]: fp32, ACCUMULATED into.  dW contracts over the tokens: both operands are transposed into S
    token slabs which form the row groups of one grouped GEMM (fp32 partial products, summed in a fixed order)."""
    lib = _lib.load()
    M = x.shape[0] if M is None else M
    K_in, N_out = x.shape[1], dy.shape[1]
    dev = x.device
    if dx_a is not None or dx_f32 is not None:
        if dx_f32 is not None:
            ops.gemm(dy, W_t, None, out_f32=dx_f32, resid=dx_resid, alpha=1.0, beta=1.0 if dx_resid is not None else 0.0, M=M)
        else:
            ops.gemm(dy, W_t, None, out_a=dx_a, M=M) if dy.dtype == bf16 else ops.gemm(dy, W_t, None, out_f32=dx_a, M=M)
    tn_ok = (x.dtype == bf16 and USE_TN[0] and not (K_in & 3) and not (x.stride(0) & 7) and not (dy.stride(0) & 7) and
             not ((x.data_ptr() | dy.data_ptr()) & 15))                  # TMA: 16-byte row pitch and base
    if dW is not None and tn_ok:
        # bf16 path: the token contraction reads dY and X as they are (MN-major operands), split in S token slabs whose fp32
        # partial products [S, out, in] are summed in a fixed order
        tiles_out = (N_out + 127) // 128
        sms = lib.mdm_num_sms() or 148
        S = max(1, min(64, (2 * sms) // max(1, tiles_out * ((K_in + 255) // 256)), (M + 255) // 256))
        Ks = ((M + S - 1) // S + 63) // 64 * 64
        S = (M + Ks - 1) // Ks                                       # every slab non-empty
        rows_a = tiles_out * 128
        key = ("tn", S, Ks, M, tiles_out, rows_a, N_out, str(dev))
        tt = _SLAB_TABLES.get(key)
        if tt is None:
            rows = [[i * 128, s_ * rows_a + i * 128, 0, min(128, N_out - i * 128)] for s_ in range(S) for i in range(tiles_out)]
            tk = [[s_ * Ks, min(Ks, M - s_ * Ks)] for s_ in range(S) for _ in range(tiles_out)]
            tt = _SLAB_TABLES[key] = (torch.tensor(rows, dtype=torch.int32).to(dev), torch.tensor(tk, dtype=torch.int32).to(dev))
        part = torch.empty(S * rows_a, K_in, dtype=f32, device=dev)
        if rows_a != N_out:
            part.zero_()
        gemm_tn(dy, x, part, N_out, K_in, tt[0], S * tiles_out, tt[1])
        if rows_a == N_out:
            sum_partials(part, S, N_out * K_in, dW)
        else:
            red = torch.empty(rows_a, K_in, dtype=f32, device=dev)
            sum_partials(part, S, rows_a * K_in, red, accumulate=False)
            axpby(red[:N_out], 1.0, dW, 1.0, dW)
    elif dW is not None:
        tiles_out = (N_out + 127) // 128
        sms = lib.mdm_num_sms() or 148
        S = max(1, min(64, (2 * sms) // max(1, tiles_out * ((K_in + 255) // 256)), (M + 255) // 256))
        Ks = ((M + S - 1) // S + 63) // 64 * 64
        rows_a = tiles_out * 128
        dyT = torch.empty(S * rows_a, Ks, dtype=dy.dtype, device=dev)
        if rows_a != N_out:
            dyT.zero_()
            tmp = torch.empty(S * N_out, Ks, dtype=dy.dtype, device=dev)
            _chk(lib.mdm_transpose_split(dy.data_ptr(), _dt(dy), M, N_out, dy.stride(0), S, Ks, tmp.data_ptr(), _stream()), "transpose dY")
            dyT.view(S, rows_a, Ks)[:, :N_out].copy_(tmp.view(S, N_out, Ks))
        else:
            _chk(lib.mdm_transpose_split(dy.data_ptr(), _dt(dy), M, N_out, dy.stride(0), S, Ks, dyT.data_ptr(), _stream()), "transpose dY")
        xT = torch.empty(S * K_in, Ks, dtype=x.dtype, device=dev)
        _chk(lib.mdm_transpose_split(x.data_ptr(), _dt(x), M, K_in, x.stride(0), S, Ks, xT.data_ptr(), _stream()), "transpose X")
        key = (S, tiles_out, rows_a, K_in, N_out, str(dev))
        tt = _SLAB_TABLES.get(key)
        if tt is None:
            rows = [[s_ * rows_a + i * 128, s_ * rows_a + i * 128, s_ * K_in, min(128, N_out - i * 128)]
                    for s_ in range(S) for i in range(tiles_out)]
            tt = _SLAB_TABLES[key] = torch.tensor(rows, dtype=torch.int32).to(dev)
        part = torch.empty(S * rows_a, K_in, dtype=f32, device=dev)
        ops.gemm(dyT, xT, None, out_f32=part, N=K_in, M=S * rows_a, tiles=tt, num_tiles=S * tiles_out, a_rows=S * rows_a,
                 w_rows=S * K_in)
        if rows_a == N_out:
            sum_partials(part, S, N_out * K_in, dW)
        else:
            red = torch.empty(rows_a, K_in, dtype=f32, device=dev)
            sum_partials(part, S, rows_a * K_in, red, accumulate=False)
            axpby(red[:N_out], 1.0, dW, 1.0, dW)
    if db is not None:
        colsum_into(dy, M, N_out, db)


def rowop_bwd(x, rows, D, dout, *, ln1=None, l2norm=False, ln2=None, film=None, rows_per_seq=0, silu=False, din=None,
              dmid=None, accumulate=False, g_ln1=None, g_ln2=None, g_film=None):
    """Backward of ops.rowop (mdm_rowop_bwd): gradient `dout` of the final output (+ optional `dmid`, the gradient that
    arrives at the out1 point, i.e. after LN1 / L2 and before LN2) -> `din` (operand- or fp32-typed; accumulate=True adds
    to it).  Parameter gradients are ADDED to g_ln1 / g_ln2 (pairs of [D] fp32) and g_film ([n_seq, 2D] fp32)."""
    _c(x, dout, film, din, dmid)
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
    st = lib.mdm_rowop_bwd2(C.byref(op), rows, D, _dt(dout), None, None, 0, None, None, None, 0, C.byref(npp), C.byref(nfc), _stream())
    if st != 0:
        raise _lib.MdmError("mdm_rowop_bwd2: status %d" % st)
    dev = x.device
    if din is None:
        din = torch.empty(rows, D, dtype=dout.dtype, device=dev)
    ppart = torch.empty(npp.value, 4, D, dtype=f32, device=dev)
    n_seq = (rows + rows_per_seq - 1) // rows_per_seq if film is not None else 1
    fpart = torch.empty(nfc.value, n_seq, 2 * D, dtype=f32, device=dev) if film is not None else None
    _chk(lib.mdm_rowop_bwd2(C.byref(op), rows, D, _dt(dout), dout.data_ptr(), _ptr(dmid), _dt(dmid) if dmid is not None else 0,
                            din.data_ptr(), ppart.data_ptr(), _ptr(fpart), (2 if accumulate else 0) | _dt(din),
                            C.byref(npp), C.byref(nfc), _stream()), "mdm_rowop_bwd2")
    if ln1 is not None or ln2 is not None:
        psum = torch.empty(4, D, dtype=f32, device=dev)
        sum_partials(ppart, npp.value, 4 * D, psum, accumulate=False)
        if g_ln1 is not None:
            axpby(psum[0], 1.0, g_ln1[0], 1.0, g_ln1[0]); axpby(psum[1], 1.0, g_ln1[1], 1.0, g_ln1[1])
        if g_ln2 is not None:
            axpby(psum[2], 1.0, g_ln2[0], 1.0, g_ln2[0]); axpby(psum[3], 1.0, g_ln2[1], 1.0, g_ln2[1])
    if film is not None and g_film is not None:
        sum_partials(fpart, nfc.value, n_seq * 2 * D, g_film)
    return din


# ---------------------------------------------------------------------------------------------------------------
# attention cores
# ---------------------------------------------------------------------------------------------------------------
def fastattn_bwd(qkv, P, norm_w, norm_b, length, shift, B, H, T, hd, dout, g_norm):
    """Backward of ops.fastattn (FastAttention.forward, models/fast_attention.py:29-92 + :150-157): returns dqkv [N, 3D] in
    qkv's type (with the reference's [-1, 1] gradient clamp); adds the shared LayerNorm(hd) gradients to g_norm = (dw, db)."""
    lib = _lib.load()
    dev = qkv.device
    bgemm = functools.partial(_bgemm, tc=qkv.dtype == bf16)      # bf16 operands -> tensor-core flavour
    M = P.shape[1]
    BH, st = B * H, _stream
    z = lambda *s: torch.empty(*s, dtype=f32, device=dev)
    qh, kh, vn = z(BH, T, hd), z(BH, T, hd), z(BH, T, hd)
    _chk(lib.mdm_fa_prep(qkv.data_ptr(), _dt(qkv), norm_w.data_ptr(), norm_b.data_ptr(), B, H, T, hd, qh.data_ptr(), kh.data_ptr(),
                         vn.data_ptr(), st()), "mdm_fa_prep")
    hm = lambda w: (H * T * w, T * w, w, 1)                        # head-major [B, H, T, w] strides (z1 = b, z2 = h)
    uq, uk = z(BH, T, M), z(BH, T, M)
    bgemm(qh, hm(hd), P, (0, 0, M, 1), uq, hm(M), B, H, T, M, hd)
    bgemm(kh, hm(hd), P, (0, 0, M, 1), uk, hm(M), B, H, T, M, hd)
    qp, kp = z(BH, T, M), z(BH, T, M)
    _chk(lib.mdm_fa_feat(uq.data_ptr(), uk.data_ptr(), _ptr(length), shift, B, H, T, M, qp.data_ptr(), kp.data_ptr(), st()), "mdm_fa_feat")
    kv = z(BH, M, hd)
    kvs = (H * M * hd, M * hd, hd, 1)
    bgemm(kp, (H * T * M, T * M, 1, M), vn, hm(hd), kv, kvs, B, H, M, hd, T, alpha=0.1)            # kv = 0.1 kp^T vn
    o = z(BH, T, hd)
    bgemm(qp, hm(M), kv, kvs, o, hm(hd), B, H, T, hd, M, alpha=0.1)                                 # o = 0.1 qp kv
    npart = C.c_int(0)
    lib.mdm_fa_out_bwd(None, None, None, None, 0, None, B, H, T, hd, None, None, C.byref(npart), st())
    part = z(npart.value, 2, hd)
    dden = z(BH, T)
    _chk(lib.mdm_fa_out_bwd(o.data_ptr(), qp.data_ptr(), kp.data_ptr(), dout.data_ptr(), _dt(dout), norm_w.data_ptr(), B, H, T, hd,
                            dden.data_ptr(), part.data_ptr(), C.byref(npart), st()), "mdm_fa_out_bwd")
    gsum = z(2, hd)
    sum_partials(part, npart.value, 2 * hd, gsum, accumulate=False)
    do = o                                                                                           # d_o written over o
    dqp, dkp = z(BH, T, M), z(BH, T, M)
    bgemm(do, hm(hd), kv, (H * M * hd, M * hd, 1, hd), dqp, hm(M), B, H, T, M, hd, alpha=0.1)       # dqp = 0.1 do kv^T
    dkv = z(BH, M, hd)
    bgemm(qp, (H * T * M, T * M, 1, M), do, hm(hd), dkv, kvs, B, H, M, hd, T, alpha=0.1)           # dkv = 0.1 qp^T do
    bgemm(vn, hm(hd), dkv, (H * M * hd, M * hd, 1, hd), dkp, hm(M), B, H, T, M, hd, alpha=0.1)     # dkp = 0.1 vn dkv^T
    dvn = z(BH, T, hd)
    bgemm(kp, hm(M), dkv, kvs, dvn, hm(hd), B, H, T, hd, M, alpha=0.1)                              # dvn = 0.1 kp dkv
    _chk(lib.mdm_fa_feat_bwd(uq.data_ptr(), uk.data_ptr(), qp.data_ptr(), kp.data_ptr(), dden.data_ptr(), B, H, T, M,
                             dqp.data_ptr(), dkp.data_ptr(), st()), "mdm_fa_feat_bwd")
    dqh, dkh = z(BH, T, hd), z(BH, T, hd)
    bgemm(dqp, hm(M), P, (0, 0, 1, M), dqh, hm(hd), B, H, T, hd, M)                                 # dqh = duq P^T
    bgemm(dkp, hm(M), P, (0, 0, 1, M), dkh, hm(hd), B, H, T, hd, M)
    lib.mdm_fa_prep_bwd(None, 0, None, None, B, H, T, hd, None, None, None, None, None, C.byref(npart), st())
    part2 = z(npart.value, 2, hd)
    dqkv = torch.empty_like(qkv)
    _chk(lib.mdm_fa_prep_bwd(qkv.data_ptr(), _dt(qkv), norm_w.data_ptr(), norm_b.data_ptr(), B, H, T, hd, dqh.data_ptr(),
                             dkh.data_ptr(), dvn.data_ptr(), dqkv.data_ptr(), part2.data_ptr(), C.byref(npart), st()), "mdm_fa_prep_bwd")
    sum_partials(part2, npart.value, 2 * hd, gsum)
    axpby(gsum[0], 1.0, g_norm[0], 1.0, g_norm[0])
    axpby(gsum[1], 1.0, g_norm[1], 1.0, g_norm[1])
    return dqkv


# ---------------------------------------------------------------------------------------------------------------
# generic FORWARD of the three attention cores from the same pieces: head sizes outside the fused kernels (hd = 256 of
# model_size="big", models/transformer.py:188-192).  ops.fastattn / lincross_* / softmax_cross route here for hd > 128.
# ---------------------------------------------------------------------------------------------------------------
def fastattn_generic(qkv, P, norm_w, norm_b, length, shift, B, H, T, hd, out):
    """FastAttention.forward (models/fast_attention.py:29-92 + the 0.1 pre-scale of :155-157) for any hd in {32 .. 1024}."""
    lib = _lib.load()
    dev = qkv.device
    bgemm = functools.partial(_bgemm, tc=qkv.dtype == bf16)      # bf16 operands -> tensor-core flavour
    M = P.shape[1]
    BH, st = B * H, _stream
    z = lambda *s: torch.empty(*s, dtype=f32, device=dev)
    qh, kh, vn = z(BH, T, hd), z(BH, T, hd), z(BH, T, hd)
    _chk(lib.mdm_fa_prep(qkv.data_ptr(), _dt(qkv), norm_w.data_ptr(), norm_b.data_ptr(), B, H, T, hd, qh.data_ptr(), kh.data_ptr(),
                         vn.data_ptr(), st()), "mdm_fa_prep")
    hm = lambda w: (H * T * w, T * w, w, 1)
    uq, uk = z(BH, T, M), z(BH, T, M)
    bgemm(qh, hm(hd), P, (0, 0, M, 1), uq, hm(M), B, H, T, M, hd)
    bgemm(kh, hm(hd), P, (0, 0, M, 1), uk, hm(M), B, H, T, M, hd)
    _chk(lib.mdm_fa_feat(uq.data_ptr(), uk.data_ptr(), _ptr(length), shift, B, H, T, M, uq.data_ptr(), uk.data_ptr(), st()), "mdm_fa_feat")
    qp, kp = uq, uk                                                        # feature maps written in place
    kv = z(BH, M, hd)
    kvs = (H * M * hd, M * hd, hd, 1)
    bgemm(kp, (H * T * M, T * M, 1, M), vn, hm(hd), kv, kvs, B, H, M, hd, T, alpha=0.1)
    o = z(BH, T, hd)
    bgemm(qp, hm(M), kv, kvs, o, hm(hd), B, H, T, hd, M, alpha=0.1)
    _chk(lib.mdm_fa_out_fwd(o.data_ptr(), qp.data_ptr(), kp.data_ptr(), norm_w.data_ptr(), norm_b.data_ptr(), B, H, T, hd,
                            out.data_ptr(), _dt(out), st()), "mdm_fa_out_fwd")
    return out


def lincross_ctx_generic(k, v, nt, B, Nt, H, hd, ctx):
    """ctx[b, h, d, l] = sum_n softmax_n(k)[n, d] v[n, l], n < nt[b]  (fast_attention.py:249-252)."""
    lib = _lib.load()
    D = H * hd
    bgemm = functools.partial(_bgemm, tc=k.dtype == bf16)      # bf16 operands -> tensor-core flavour
    Ks = torch.empty(B, Nt, D, dtype=f32, device=k.device)
    _chk(lib.mdm_col_softmax(k.data_ptr(), _dt(k), _ptr(nt), B, Nt, D, Ks.data_ptr(), _stream()), "mdm_col_softmax")
    bgemm(Ks, (Nt * D, hd, 1, D), v, (Nt * D, hd, D, 1), ctx, (H * hd * hd, hd * hd, hd, 1), B, H, hd, hd, Nt)
    return ctx


def lincross_apply_generic(q, ctx, B, T, H, hd, y):
    """y[t, h, :] = softmax_hd(q[t, h, :]) @ ctx[b, h]  (fast_attention.py:252-253)."""
    lib = _lib.load()
    D = H * hd
    bgemm = functools.partial(_bgemm, tc=q.dtype == bf16)      # bf16 operands -> tensor-core flavour
    Pm = torch.empty(B * H, T, hd, dtype=f32, device=q.device)
    _chk(lib.mdm_head_softmax(q.data_ptr(), _dt(q), B, H, T, hd, Pm.data_ptr(), _stream()), "mdm_head_softmax")
    bgemm(Pm, (H * T * hd, T * hd, hd, 1), ctx, (H * hd * hd, hd * hd, hd, 1), y, (T * D, hd, D, 1), B, H, T, hd, hd)
    return y


def softmax_cross_generic(q, k, v, nt, B, T, Nt, H, hd, o):
    """o = softmax_n(q k^T hd^-0.5) v over the n < nt[b] text tokens  (fast_attention.py:309-322)."""
    lib = _lib.load()
    D = H * hd
    bgemm = functools.partial(_bgemm, tc=q.dtype == bf16)      # bf16 operands -> tensor-core flavour
    tok, txt, sm = (T * D, hd, D, 1), (Nt * D, hd, D, 1), (H * T * Nt, T * Nt, Nt, 1)
    S = torch.empty(B * H, T, Nt, dtype=f32, device=q.device)
    bgemm(q, tok, k, (Nt * D, hd, 1, D), S, sm, B, H, T, Nt, hd, alpha=hd ** -0.5)
    _chk(lib.mdm_key_softmax(S.data_ptr(), _ptr(nt), B, H, T, Nt, _stream()), "mdm_key_softmax")
    bgemm(S, sm, v, txt, o, tok, B, H, T, hd, Nt)
    return o


def lincross_apply_bwd(q, ctx, B, T, H, hd, dy):
    """Backward of ops.lincross_apply (motion side of LinearTemporalCrossAttention, fast_attention.py:252-253):
    y = softmax_hd(q) @ ctx[b, h].  Returns (dq [N, D] in q's type, dctx [B, H, hd, hd] fp32)."""
    lib = _lib.load()
    dev, D = q.device, H * hd
    bgemm = functools.partial(_bgemm, tc=q.dtype == bf16)      # bf16 operands -> tensor-core flavour
    Pm = torch.empty(B * H, T, hd, dtype=f32, device=dev)
    _chk(lib.mdm_head_softmax(q.data_ptr(), _dt(q), B, H, T, hd, Pm.data_ptr(), _stream()), "mdm_head_softmax")
    hm = (H * T * hd, T * hd, hd, 1)
    tok = (T * D, hd, D, 1)                                     # token-major [B, T, H*hd], head slice
    cs = (H * hd * hd, hd * hd, hd, 1)
    dP = torch.empty_like(Pm)
    bgemm(dy, tok, ctx, (H * hd * hd, hd * hd, 1, hd), dP, hm, B, H, T, hd, hd)                      # dP = dy ctx^T
    dctx = torch.empty(B, H, hd, hd, dtype=f32, device=dev)
    bgemm(Pm, (H * T * hd, T * hd, 1, hd), dy, tok, dctx, cs, B, H, hd, hd, T)                       # dctx = P^T dy
    dq = torch.empty_like(q)
    _chk(lib.mdm_head_softmax_bwd(Pm.data_ptr(), dP.data_ptr(), B, H, T, hd, dq.data_ptr(), _dt(dq), _stream()), "mdm_head_softmax_bwd")
    return dq, dctx


def lincross_ctx_bwd(k, v, nt, B, Nt, H, hd, dctx):
    """Backward of ops.lincross_ctx (text side, fast_attention.py:249-252): ctx[b,h] = softmax_n(k)^T v.  Returns (dk, dv)
    [B*Nt, D] in k's type."""
    lib = _lib.load()
    dev, D = k.device, H * hd
    bgemm = functools.partial(_bgemm, tc=k.dtype == bf16)      # bf16 operands -> tensor-core flavour
    Ks = torch.empty(B, Nt, D, dtype=f32, device=dev)
    _chk(lib.mdm_col_softmax(k.data_ptr(), _dt(k), _ptr(nt), B, Nt, D, Ks.data_ptr(), _stream()), "mdm_col_softmax")
    txt = (Nt * D, hd, D, 1)
    cs = (H * hd * hd, hd * hd, hd, 1)
    dKs = torch.empty_like(Ks)
    bgemm(v, txt, dctx, (H * hd * hd, hd * hd, 1, hd), dKs, txt, B, H, Nt, hd, hd)                   # dKs[n,d] = sum_l v[n,l] dctx[d,l]
    dv = torch.empty_like(v)
    bgemm(Ks, txt, dctx, cs, dv, txt, B, H, Nt, hd, hd)                                              # dv[n,l] = sum_d Ks[n,d] dctx[d,l]
    dk = torch.empty_like(k)
    _chk(lib.mdm_col_softmax_bwd(Ks.data_ptr(), dKs.data_ptr(), B, Nt, D, dk.data_ptr(), _dt(dk), _stream()), "mdm_col_softmax_bwd")
    return dk, dv


def softmax_cross_bwd(q, k, v, nt, B, T, Nt, H, hd, do):
    """Backward of ops.softmax_cross (MemoryEfficientCrossAttentionBlock core, fast_attention.py:305-325):
    o = softmax_n(q k^T hd^-0.5) v.  Returns (dq [N, D], dk, dv [B*Nt, D]) in the operand type."""
    lib = _lib.load()
    dev, D = q.device, H * hd
    bgemm = functools.partial(_bgemm, tc=q.dtype == bf16)      # bf16 operands -> tensor-core flavour
    scale = hd ** -0.5
    tok, txt = (T * D, hd, D, 1), (Nt * D, hd, D, 1)
    sm = (H * T * Nt, T * Nt, Nt, 1)
    S = torch.empty(B * H, T, Nt, dtype=f32, device=dev)
    bgemm(q, tok, k, (Nt * D, hd, 1, D), S, sm, B, H, T, Nt, hd, alpha=scale)                        # S = scale q k^T
    _chk(lib.mdm_key_softmax(S.data_ptr(), _ptr(nt), B, H, T, Nt, _stream()), "mdm_key_softmax")   # P (in place)
    dP = torch.empty_like(S)
    bgemm(do, tok, v, (Nt * D, hd, 1, D), dP, sm, B, H, T, Nt, hd)                                   # dP = do v^T
    dv = torch.empty_like(v)
    bgemm(S, (H * T * Nt, T * Nt, 1, Nt), do, tok, dv, txt, B, H, Nt, hd, T)                         # dv = P^T do
    _chk(lib.mdm_key_softmax_bwd(S.data_ptr(), dP.data_ptr(), B, H, T, Nt, _stream()), "mdm_key_softmax_bwd")   # dS over dP
    dq = torch.empty_like(q)
    bgemm(dP, sm, k, txt, dq, tok, B, H, T, hd, Nt, alpha=scale)                                     # dq = scale dS k
    dk = torch.empty_like(k)
    bgemm(dP, (H * T * Nt, T * Nt, 1, Nt), q, tok, dk, txt, B, H, Nt, hd, T, alpha=scale)            # dk = scale dS^T q
    return dq, dk, dv


# ---------------------------------------------------------------------------------------------------------------
# MoE
# ---------------------------------------------------------------------------------------------------------------
def moe_combine_sum(z, rowscale, perm, N, D, NBK, m):
    _chk(_lib.load().mdm_moe_combine_sum(z.data_ptr(), _dt(z), rowscale.data_ptr(), perm.data_ptr(), N, D, NBK, m.data_ptr(),
                                         _stream()), "mdm_moe_combine_sum")


def moe_combine_bwd(z, rowscale, perm, N, D, NBK, dm, dz, drs):
    _chk(_lib.load().mdm_moe_combine_bwd(z.data_ptr(), _dt(z), rowscale.data_ptr(), perm.data_ptr(), N, D, NBK, dm.data_ptr(),
                                         dz.data_ptr(), drs.data_ptr(), _stream()), "mdm_moe_combine_bwd")


def moe_gate_bwd_logits(x, stats, ln_w, ln_b, gate_w, gate_b, idx, perm, drs, N, D, NB, E, dlogits):
    _chk(_lib.load().mdm_moe_gate_bwd_logits(x.data_ptr(), stats.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), gate_w.data_ptr(),
                                             gate_b.data_ptr(), idx.data_ptr(), perm.data_ptr(), drs.data_ptr(), N, D, NB, E,
                                             dlogits.data_ptr(), _stream()), "mdm_moe_gate_bwd_logits")


def moe_unpermute_bwd(dxp, perm, dlogits, gate_w, N, D, NB, E, br, dh):
    _chk(_lib.load().mdm_moe_unpermute_bwd(dxp.data_ptr(), _dt(dxp), perm.data_ptr(), dlogits.data_ptr(), gate_w.data_ptr(), N, D,
                                           NB, E, br, dh.data_ptr(), _stream()), "mdm_moe_unpermute_bwd")


def expert_ffn_bwd(xp, pre, hp, W1t, W2t, dz, seg_off, idx, N, NB, E, tiles_up, tiles_dn, ntile, max_tiles, cap, F, D,
                   g_w1, g_b1, g_w2, g_b2):
    """Backward of the grouped expert FFN z = gelu(xp W1_g^T + b1_g) W2_g^T + b2_g over expert-sorted rows
    (switch_moe.py:19-25, 97-109).  All buffers have round_up(cap, 128) + 128 rows whose last 128 are zero (the contraction
    range of an empty expert); dz must be zero outside the routed rows.  W1t [G*D, F] / W2t [G*F, D]: per-group
    transposed weights.  Weight / bias gradients are accumulated into g_*.  Returns d_xp [cap + 128, D]."""
    lib = _lib.load()
    dev, adt = xp.device, xp.dtype
    G = NB * E
    rows = xp.shape[0]
    zero_row0 = rows - 128
    kw = dict(num_tiles=max_tiles, num_tiles_dev=ntile, M=cap, a_rows=rows)
    out = lambda t: dict(out_a=t) if adt == bf16 else dict(out_f32=t)
    d_hp = torch.empty(rows, F, dtype=adt, device=dev)
    ops.gemm(dz, W2t, None, N=F, tiles=tiles_up, w_rows=G * F, **out(d_hp), **kw)           # d_hp = dz W2_g  (tiles_up: w_row0 = g*F)
    d_pre = torch.empty(rows, F, dtype=adt, device=dev)      # rows < cap are written below; [cap, rows) is the zero region
    d_pre[cap:].zero_()
    _chk(lib.mdm_act_bwd(pre.data_ptr(), d_hp.data_ptr(), _dt(pre), cap * F, ACT_GELU, d_pre.data_ptr(), _stream()), "mdm_act_bwd")
    d_xp = torch.empty(rows, D, dtype=adt, device=dev)
    ops.gemm(d_pre, W1t, None, N=D, tiles=tiles_dn, w_rows=G * D, **out(d_xp), **kw)        # d_xp = d_pre W1_g (tiles_dn: w_row0 = g*D)
    # weight gradients: contraction over each expert's row segment (tile_k), operands transposed once
    mt_up, mt_dn = (F + 127) // 128, (D + 127) // 128
    seg_cnt = torch.empty(G, dtype=torch.int32, device=dev)
    tk_up = torch.empty(G * mt_up, 2, dtype=torch.int32, device=dev)
    tk_dn = torch.empty(G * mt_dn, 2, dtype=torch.int32, device=dev)
    _chk(lib.mdm_moe_wgrad_tables(seg_off.data_ptr(), idx.data_ptr(), N, NB, E, mt_up, mt_dn, zero_row0, seg_cnt.data_ptr(),
                                  tk_up.data_ptr(), tk_dn.data_ptr(), _stream()), "mdm_moe_wgrad_tables")

    def wgrad(dy_rows, x_rows, out_dim, in_dim, tk, mt, g_w):
        if adt == bf16 and USE_TN[0]:          # token contraction on the operands as they are (MN-major), per expert segment
            key = ("wg_tn", G, out_dim, str(dev))
            tt = _SLAB_TABLES.get(key)
            if tt is None:
                tt = _SLAB_TABLES[key] = torch.tensor([[i * 128, g_ * out_dim + i * 128, 0, min(128, out_dim - i * 128)]
                                                       for g_ in range(G) for i in range(mt)], dtype=torch.int32).to(dev)
            part = torch.empty(G * out_dim, in_dim, dtype=f32, device=dev)
            if out_dim % 128:
                part.zero_()
            gemm_tn(dy_rows, x_rows, part, out_dim, in_dim, tt, G * mt, tk)
            axpby(part, 1.0, g_w, 1.0, g_w)
            return
        dyT, xT = transpose_groups(dy_rows, rows), transpose_groups(x_rows, rows)      # [out, rows], [in, rows]
        key = ("wg", G, out_dim, str(dev))
        tt = _SLAB_TABLES.get(key)
        if tt is None:
            tt = _SLAB_TABLES[key] = torch.tensor([[i * 128, g_ * out_dim + i * 128, 0, min(128, out_dim - i * 128)]
                                                   for g_ in range(G) for i in range(mt)], dtype=torch.int32).to(dev)
        part = torch.empty(G * out_dim, in_dim, dtype=f32, device=dev)
        ops.gemm(dyT, xT, None, out_f32=part, N=in_dim, M=G * out_dim, tiles=tt, num_tiles=G * mt, a_rows=out_dim, w_rows=in_dim,
                 tile_k=tk)
        axpby(part, 1.0, g_w, 1.0, g_w)
    wgrad(dz, hp, D, F, tk_dn, mt_dn, g_w2)
    wgrad(d_pre, xp, F, D, tk_up, mt_up, g_w1)
    SL = 16                                                    # row slabs per segment: [SL, G, C] partials, summed in a fixed order
    pb2 = torch.empty(SL, G, D, dtype=f32, device=dev)
    pb1 = torch.empty(SL, G, F, dtype=f32, device=dev)
    _chk(lib.mdm_seg_colsum(dz.data_ptr(), _dt(dz), D, seg_off.data_ptr(), seg_cnt.data_ptr(), G, SL, pb2.data_ptr(), _stream()), "mdm_seg_colsum")
    _chk(lib.mdm_seg_colsum(d_pre.data_ptr(), _dt(d_pre), F, seg_off.data_ptr(), seg_cnt.data_ptr(), G, SL, pb1.data_ptr(), _stream()), "mdm_seg_colsum")
    sum_partials(pb2, SL, G * D, g_b2)
    sum_partials(pb1, SL, G * F, g_b1)
    return d_xp


def masked_mse_grad(pred, target, length, scale=1.0):
    B, T, F = pred.shape
    out = torch.empty_like(pred)
    _chk(_lib.load().mdm_masked_mse_grad(pred.data_ptr(), target.data_ptr(), length.data_ptr(), B, T, F, float(scale), out.data_ptr(),
                                         _stream()), "mdm_masked_mse_grad")
    return out


def gated_mix_bwd(t, x, dout):
    dt_, dx = torch.empty_like(t), torch.empty_like(x)
    _chk(_lib.load().mdm_gated_mix_bwd(t.data_ptr(), x.data_ptr(), dout.data_ptr(), t.numel(), dt_.data_ptr(), dx.data_ptr(),
                                       _stream()), "mdm_gated_mix_bwd")
    return dt_, dx
