"""GPU parity tests of every C-ABI kernel against the oracle (oracle/motion_oracle.py) or the
corresponding torch fp32 op, on seeded inputs.  Tolerances: fp32 path 1e-5 relative L2 (north_star),
bf16 path 2e-2; routing and the sampler update are bit-exact."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from motiondiffusion_moe_b200 import ops  # noqa: E402
from motiondiffusion_moe_b200._lib import ACT_GELU, ACT_NONE, ACT_SILU  # noqa: E402
from oracle import motion_oracle as mo  # noqa: E402

DEV = "cuda"
TOL = {torch.float32: 1e-5, torch.bfloat16: 2e-2}


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def gen(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


def randn(*shape, seed=0, scale=1.0):
    return (torch.randn(*shape, generator=gen(seed)) * scale).to(DEV)


# ------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 263, 512), (1000, 128, 128), (392, 512, 264),
                                   (2048, 1536, 512), (777, 512, 2048)])
def test_gemm_plain(dtype, M, N, K):
    A = randn(M, K, seed=1).to(dtype)
    W = randn(N, K, seed=2, scale=K ** -0.5).to(dtype)
    b = randn(N, seed=3)
    out = torch.empty(M, N, device=DEV)
    ops.gemm(A, W, b, out_f32=out)
    ref = A.float() @ W.float().t() + b
    assert rel(out, ref) < 1e-5   # operands are identical (already rounded); fp32 accumulate


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("act", [ACT_NONE, ACT_GELU, ACT_SILU])
def test_gemm_epilogue(dtype, act):
    M, N, K = 520, 512, 512
    A = randn(M, K, seed=1).to(dtype)
    W = randn(N, K, seed=2, scale=K ** -0.5).to(dtype)
    b, R, rs = randn(N, seed=3), randn(M, N, seed=4), torch.rand(M, generator=gen(5)).to(DEV)
    o32 = torch.empty(M, N, device=DEV)
    oa = torch.empty(M, N, device=DEV, dtype=dtype)
    ops.gemm(A, W, b, act=act, alpha=0.1, beta=0.7, resid=R, rowscale=rs, out_f32=o32, out_a=oa, a_pre_resid=True)
    v = A.float() @ W.float().t() + b
    v = {ACT_NONE: v, ACT_GELU: F.gelu(v), ACT_SILU: F.silu(v)}[act]
    pre = 0.1 * v * rs[:, None]
    # the bf16 path may use a fast erf (abs err <= 2e-7) in its GELU epilogue
    assert rel(o32, pre + 0.7 * R) < (1e-5 if dtype == torch.float32 else 1e-4)
    assert rel(oa, pre) < (1e-5 if dtype == torch.float32 else 4e-3)


@pytest.mark.parametrize("M,N,K", [(25, 512, 512), (1000, 512, 2048), (333, 128, 512), (777, 264, 576),
                                   (4100, 1024, 1024), (130, 36, 64)])
@pytest.mark.parametrize("act", [ACT_NONE, ACT_GELU])
@pytest.mark.parametrize("secondary", ["none", "pre", "post"])
def test_gemm_f32_residual_tma_epilogue(M, N, K, act, secondary):
    """fp32 output + fp32 residual (the residual-stream GEMMs: s_out / skip / ca_out / ffn_out / sd_o / sd_f3):
    residual in and sum out by TMA, ragged M / N clipped by the tensor maps, optional bf16 copy before or after
    the residual, output aliasing the residual."""
    A = randn(M, K, seed=1).to(torch.bfloat16)
    W = randn(N, K, seed=2, scale=K ** -0.5).to(torch.bfloat16)
    b, R = randn(N, seed=3), randn(M, N, seed=4)
    guard = torch.full((M + 64, N), 3.0, device=DEV)        # rows >= M must stay untouched
    o32 = guard[:M]
    oa = torch.full((M + 64, N), 5.0, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, W, b, act=act, alpha=0.1, beta=0.7, resid=R, out_f32=o32,
             out_a=None if secondary == "none" else oa[:M], a_pre_resid=secondary == "pre")
    v = A.float() @ W.float().t() + b
    pre = 0.1 * (F.gelu(v) if act == ACT_GELU else v)
    assert rel(o32, pre + 0.7 * R) < 1e-4
    assert torch.all(guard[M:] == 3.0)
    if secondary != "none":
        assert rel(oa[:M], pre if secondary == "pre" else pre + 0.7 * R) < 4e-3
        assert torch.all(oa[M:] == 5.0)
    # in place (output == residual buffer) gives the same bits as out of place
    Rc, o2 = R.clone(), torch.empty(M, N, device=DEV)
    ops.gemm(A, W, b, act=act, alpha=0.1, beta=0.7, resid=R, out_f32=o2)
    ops.gemm(A, W, b, act=act, alpha=0.1, beta=0.7, resid=Rc, out_f32=Rc)
    assert torch.equal(Rc, o2)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gemm_positional_residual_and_secondary(dtype):
    T, B, N, K = 12, 5, 256, 72
    A = randn(B * T, K, seed=1).to(dtype)
    W = randn(N, K, seed=2, scale=K ** -0.5).to(dtype)
    b, pos = randn(N, seed=3), randn(20, N, seed=4)
    o32 = torch.empty(B * T, N, device=DEV)
    oa = torch.empty(B * T, N, device=DEV, dtype=dtype)
    ops.gemm(A, W, b, resid=pos, resid_mod=T, beta=1.0, out_f32=o32, out_a=oa)
    ref = (A.float() @ W.float().t() + b).view(B, T, N) + pos[:T]
    assert rel(o32, ref.view(B * T, N)) < 1e-5
    assert rel(oa, ref.view(B * T, N)) < (1e-5 if dtype == torch.float32 else 4e-3)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gemm_grouped_device_tile_count(dtype):
    """Grouped GEMM with a device-resident tile table/count, ragged groups (incl. an empty one)."""
    G, K, N = 5, 256, 384
    counts = [130, 0, 1, 128, 77]
    offs, tiles, off = [], [], 0
    for g, c in enumerate(counts):
        offs.append(off)
        for i in range((c + 127) // 128):
            tiles.append([off + i * 128, off + i * 128, g * N, min(128, c - i * 128)])
        off += (c + 127) // 128 * 128
    cap = off + 256
    A = randn(cap, K, seed=1).to(dtype)
    W = randn(G * N, K, seed=2, scale=K ** -0.5).to(dtype)
    b = randn(G * N, seed=3)
    rs = torch.rand(cap, generator=gen(4)).to(DEV)
    tt = torch.zeros(cap // 128, 4, dtype=torch.int32, device=DEV)
    tt[:len(tiles)] = torch.tensor(tiles, dtype=torch.int32)
    nt = torch.tensor([len(tiles)], dtype=torch.int32, device=DEV)
    out = torch.full((cap, N), 7.0, device=DEV)
    ops.gemm(A, W, b, act=ACT_GELU, rowscale=rs, out_f32=out, N=N, M=cap, tiles=tt, num_tiles=cap // 128,
             num_tiles_dev=nt, a_rows=cap, w_rows=G * N)
    for g, c in enumerate(counts):
        rows = slice(offs[g], offs[g] + c)
        ref = F.gelu(A[rows].float() @ W[g * N:(g + 1) * N].float().t() + b[g * N:(g + 1) * N]) * rs[rows, None]
        if c:
            assert rel(out[rows], ref) < (1e-5 if dtype == torch.float32 else 1e-4)
        pad = slice(offs[g] + c, offs[g] + (c + 127) // 128 * 128)
        assert torch.all(out[pad] == 7.0)      # padding rows are never written


@pytest.mark.parametrize("stages", ["full", "ln2_film_silu"])
@pytest.mark.parametrize("M,T", [(196 * 3, 196), (128, 8), (1000, 100), (25, 5)])
def test_gemm_rowop_fused(stages, M, T):
    """mdm_gemm_rowop (row pipeline fused into the A-operand construction of the residual-stream GEMM) against
    torch and against the unfused rowop + gemm pair."""
    D = N = 512
    x = randn(M, D, seed=1, scale=1.5).bfloat16()
    W = randn(N, D, seed=2, scale=D ** -0.5).bfloat16()
    b, R = randn(N, seed=3), randn(M, N, seed=4)
    ln1 = (torch.rand(D, generator=gen(5)).to(DEV) + 0.5, randn(D, seed=6, scale=0.1))
    ln2 = (torch.rand(D, generator=gen(7)).to(DEV) + 0.5, randn(D, seed=8, scale=0.1))
    film = randn((M + T - 1) // T, 2 * D, seed=9, scale=0.3)
    full = stages == "full"
    kw = dict(ln1=ln1 if full else None, l2norm=full, ln2=ln2, film=film, rows_per_seq=T, silu=True)
    out = torch.full((M + 8, N), 3.0, device=DEV)
    assert ops.gemm_rowop(x, M, D, W, b, out_f32=out[:M], resid=R, alpha=0.1, beta=1.0, **kw)
    assert torch.all(out[M:] == 3.0)
    a2 = torch.empty(M, D, device=DEV, dtype=torch.bfloat16)
    ops.rowop(x, M, D, ops._dt(a2), out2_a=a2, **kw)
    unf = torch.empty(M, N, device=DEV)
    ops.gemm(a2, W, b, out_f32=unf, resid=R, alpha=0.1, beta=1.0)
    v = x.float()
    if full:
        v = F.layer_norm(v, (D,), ln1[0], ln1[1])
        v = F.normalize(v, dim=-1) * math.sqrt(D)
    v = F.layer_norm(v, (D,), ln2[0], ln2[1])
    sc, sh = film[:, :D], film[:, D:]
    seq = torch.arange(M, device=DEV) // T
    v = F.silu(v * (1 + sc[seq]) + sh[seq]).bfloat16().float()
    ref = R + 0.1 * (v @ W.float().t() + b)
    assert rel(out[:M], ref) < 2e-3
    assert rel(out[:M], unf) < 2e-3
    # shapes outside the fused kernel are declined without launching anything
    assert not ops.gemm_rowop(x[:, :256].contiguous(), M, 256, W[:, :256].contiguous(), b, out_f32=out[:M], resid=R, **kw)


@pytest.mark.parametrize("variant", ["proj_style", "sout_ln", "skip_ln_ln", "sdo_ln_pre", "ffn_ln_ln_copy"])
@pytest.mark.parametrize("M,K,T", [(196 * 3, 512, 196), (128, 512, 8), (1000, 2048, 100), (25, 512, 5), (1537, 1024, 196)])
def test_gemm_ln_fused(variant, M, K, T):
    """mdm_gemm_ln (csrc/gemm_ln.cu): the Linear and the row pipeline that follows it in one kernel (the row stays in
    TMEM), for the five Linear -> LayerNorm chains of MoEExtendedDecoderLayer, against torch fp32 on the same bf16
    operands and against the unfused gemm + rowop pair.  Rows past M must stay untouched (TMA clipping), ragged
    M (not a multiple of the 256-row pair tile), T not dividing 128 (several sequences' FiLM rows in one tile)."""
    D = N = 512
    A = randn(M, K, seed=1).bfloat16()
    W = randn(N, K, seed=2, scale=K ** -0.5).bfloat16()
    b, R = randn(N, seed=3, scale=0.5), randn(M, N, seed=4, scale=2.0)
    ln1 = (torch.rand(D, generator=gen(5)).to(DEV) + 0.5, randn(D, seed=6, scale=0.1))
    ln2 = (torch.rand(D, generator=gen(7)).to(DEV) + 0.5, randn(D, seed=8, scale=0.1))
    film = randn((M + T - 1) // T, 2 * D, seed=9, scale=0.3)
    bf = torch.bfloat16

    def buf(dt):
        return torch.full((M + 8, N), 3.0, device=DEV, dtype=dt)

    acc = A.float() @ W.float().t() + b
    got, ref = {}, {}
    if variant == "proj_style":      # Performer p3 -> post LN -> L2 norm -> StylizationBlock LN, FiLM, SiLU
        o2 = buf(bf)
        assert ops.gemm_ln(A, W, b, ln1=ln1, l2norm=True, ln2=ln2, film=film, rows_per_seq=T, silu=True, out2_a=o2[:M])
        v = F.layer_norm(acc, (D,), ln1[0], ln1[1])
        v = F.normalize(v, dim=-1) * math.sqrt(D)
        v = F.layer_norm(v, (D,), ln2[0], ln2[1])
        seq = torch.arange(M, device=DEV) // T
        ref["out2"] = F.silu(v * (1 + film[seq, :D]) + film[seq, D:])
        got["out2"] = o2
    elif variant == "sout_ln":       # s_out + residual -> pre-norm of the next Performer block
        y, o1 = buf(torch.float32), buf(bf)
        assert ops.gemm_ln(A, W, b, ln1=ln1, alpha=0.1, beta=1.0, resid=R, out_f32=y[:M], out1_a=o1[:M])
        ref["y"] = R + 0.1 * acc
        ref["out1"] = F.layer_norm(ref["y"], (D,), ln1[0], ln1[1])
        got["y"], got["out1"] = y, o1
    elif variant == "skip_ln_ln":    # skip Linear + GELU + 0.1 * residual -> post norm (fp32) -> cross-attention norm
        o1, o2 = buf(torch.float32), buf(bf)
        assert ops.gemm_ln(A, W, b, ln1=ln1, act=ACT_GELU, alpha=1.0, beta=0.1, resid=R, out1_f32=o1[:M], ln2=ln2, out2_a=o2[:M])
        pre = F.gelu(acc) + 0.1 * R
        ref["out1"] = F.layer_norm(pre, (D,), ln1[0], ln1[1])
        ref["out2"] = F.layer_norm(ref["out1"], (D,), ln2[0], ln2[1])
        got["out1"], got["out2"] = o1, o2
    elif variant == "sdo_ln_pre":    # cross-attention output projection: residual sum out, LayerNorm of the projection itself
        y, o1 = buf(torch.float32), buf(bf)
        assert ops.gemm_ln(A, W, b, ln1=ln1, alpha=1.0, beta=1.0, resid=R, out_f32=y[:M], ln_pre_resid=True, out1_a=o1[:M])
        ref["y"] = R + acc
        ref["out1"] = F.layer_norm(acc, (D,), ln1[0], ln1[1])
        got["y"], got["out1"] = y, o1
    else:                            # FFN output Linear + residual = layer output -> next layer's pre-norms + its bf16 copy
        y, ya, o1, o2 = buf(torch.float32), buf(bf), buf(torch.float32), buf(bf)
        assert ops.gemm_ln(A, W, b, ln1=ln1, alpha=1.0, beta=1.0, resid=R, out_f32=y[:M], out_a=ya[:M], out1_f32=o1[:M],
                           ln2=ln2, out2_a=o2[:M])
        ref["y"] = R + acc
        ref["ya"] = ref["y"]
        ref["out1"] = F.layer_norm(ref["y"], (D,), ln1[0], ln1[1])
        ref["out2"] = F.layer_norm(ref["out1"], (D,), ln2[0], ln2[1])
        got["y"], got["ya"], got["out1"], got["out2"] = y, ya, o1, o2
    for k, g in got.items():
        assert torch.all(g[M:].float() == 3.0), k                       # nothing written past M
        tol = 1e-4 if g.dtype == torch.float32 else 4e-3                # bf16 outputs: rounding of the output itself
        if variant == "skip_ln_ln":
            tol = max(tol, 1e-3)                                        # bf16-grade GELU (gelu_tanh_fit) under the LayerNorm
        assert rel(g[:M].float(), ref[k]) < tol, (k, rel(g[:M].float(), ref[k]))
    # shapes outside the fused kernel are declined without launching anything
    assert not ops.gemm_ln(A, W[:256].contiguous(), b[:256].contiguous(), ln1=ln1, out1_a=buf(bf)[:M])


@pytest.mark.parametrize("B,H,T,hd", [(3, 4, 196, 128), (5, 4, 98, 128), (2, 4, 8, 128), (4, 2, 130, 128), (64, 4, 196, 128),
                                      (3, 8, 196, 64), (5, 4, 98, 64)])
def test_lincross_apply_style_fused(B, H, T, hd):
    """mdm_lincross_apply_style: the linear cross-attention core with the StylizationBlock (LayerNorm over the whole row =
    all heads, FiLM, SiLU) in its epilogue (the head-CTAs of a sequence as a cluster, partial row statistics through
    distributed shared memory) against the core + rowop pair and against torch on the core's fp32 result."""
    D = H * hd
    N = B * T
    q = randn(N, D, seed=1).bfloat16()
    ctx = randn(B, H, hd, hd, seed=2, scale=hd ** -0.5)
    ctxT = ops.pack_lincross_ctxT(ctx)          # head size 64: block-diagonal pairs of heads, two heads per CTA
    ln = (torch.rand(D, generator=gen(5)).to(DEV) + 0.5, randn(D, seed=6, scale=0.1))
    film = randn(B, 2 * D, seed=9, scale=0.3)
    y = torch.full((N + 4, D), 3.0, device=DEV, dtype=torch.bfloat16)
    assert ops.lincross_apply_style(q, ctxT, B, T, H, hd, ln, film, y[:N])
    assert torch.all(y[N:].float() == 3.0)
    core = torch.empty(N, D, device=DEV, dtype=torch.bfloat16)
    ops.lincross_apply(q, ctx, B, T, H, hd, core, ctxT=ctxT)
    unf = torch.empty(N, D, device=DEV, dtype=torch.bfloat16)
    ops.rowop(core, N, D, ops._dt(unf), ln2=ln, film=film, rows_per_seq=T, silu=True, out2_a=unf)
    # torch: softmax over the head dimension, times ctx (bf16-rounded operands like the kernel), LN over the row, FiLM, SiLU
    p = torch.softmax(q.float().view(B, T, H, hd), -1).bfloat16().float()
    yy = torch.einsum("bthd,bhdl->bthl", p, ctx.bfloat16().float()).reshape(N, D)
    v = F.layer_norm(yy, (D,), ln[0], ln[1])
    seq = torch.arange(N, device=DEV) // T
    ref = F.silu(v * (1 + film[seq, :D]) + film[seq, D:])
    assert rel(y[:N], ref) < 6e-3, rel(y[:N], ref)            # bf16 output + bf16 softmax operand
    assert rel(y[:N], unf) < 1.2e-2, rel(y[:N], unf)          # the pair rounds the core's output to bf16 in between


@pytest.mark.parametrize("M", [196 * 3, 128, 1000, 25, 12544])
def test_gemm_gate_fused(M):
    """mdm_gemm_gate: the cross-attention output Linear + residual with the MoE gate of both branches as its second
    pass, against the unfused pair (mdm_gemm_bf16 + mdm_moe_gate, itself checked against oracle.switch_moe): the same
    x2, the same routing wherever the decision is not a rounding-level near-tie, the same gate values, per-block
    histograms / importance sums consistent with the kernel's own routing."""
    D = N = K = 512
    NB, E = 2, 8
    G = NB * E
    A = randn(M, K, seed=1).bfloat16()
    W = randn(N, K, seed=2, scale=K ** -0.5).bfloat16()
    b, R = randn(N, seed=3, scale=0.5), randn(M, N, seed=4, scale=2.0)
    lnw = torch.rand(NB, D, generator=gen(5)).to(DEV) + 0.5
    lnb = randn(NB, D, seed=6, scale=0.1)
    gw = randn(G, D, seed=7, scale=0.08)
    gb = randn(G, seed=8, scale=0.1)
    nblk = (M + 127) // 128

    def outs():
        return (torch.full((M + 8, NB, 2), -7, dtype=torch.int32, device=DEV), torch.full((M + 8, NB, 2), -7.0, device=DEV),
                torch.full((M + 8, 2), -7.0, device=DEV), torch.zeros(nblk, 2, G, dtype=torch.int32, device=DEV),
                torch.zeros(nblk, G, device=DEV))

    x2f = torch.full((M + 8, N), 3.0, device=DEV)
    idx, vals, stats, hist, imp = outs()
    assert ops.gemm_gate(A, W, b, resid=R, out_f32=x2f[:M], NB=NB, E=E, ln_w=lnw, ln_b=lnb, gate_w=gw, gate_b=gb, idx=idx[:M],
                         vals=vals[:M], stats=stats[:M], blk_hist=hist, blk_imp=imp)
    x2u = torch.empty(M, N, device=DEV)
    ops.gemm(A, W, b, out_f32=x2u, resid=R, alpha=1.0, beta=1.0)
    idx_u, vals_u, stats_u, hist_u, imp_u = outs()
    ops.moe_gate(x2u, M, D, NB, E, lnw, lnb, gw, gb, idx_u[:M], vals_u[:M], stats_u[:M], hist_u, imp_u)
    torch.cuda.synchronize()
    assert torch.all(x2f[M:] == 3.0) and torch.all(idx[M:] == -7) and torch.all(vals[M:] == -7.0) and torch.all(stats[M:] == -7.0)
    assert rel(x2f[:M], x2u) < 1e-6
    # reference probabilities in fp64 from the fused kernel's own x2: decisions with a clear margin must agree
    y = x2f[:M].double()
    mu, var = y.mean(-1, keepdim=True), y.var(-1, unbiased=False, keepdim=True)
    yn = (y - mu) / torch.sqrt(var + 1e-5)
    for br in range(NB):
        h = yn * lnw[br].double() + lnb[br].double()
        p = torch.softmax(h @ gw[br * E:(br + 1) * E].double().t() + gb[br * E:(br + 1) * E].double(), -1)
        top = torch.sort(p, -1, descending=True)
        clear = ((top.values[:, 0] - top.values[:, 1]) > 1e-5) & ((top.values[:, 1] - top.values[:, 2]) > 1e-5)
        assert clear.float().mean() > 0.99
        assert torch.equal(idx[:M, br][clear].long(), top.indices[:, :2][clear])
        assert torch.equal(idx[:M, br][clear], idx_u[:M, br][clear])
        assert (vals[:M, br][clear] - top.values[:, :2][clear].float()).abs().max() < 1e-5
    assert (stats[:M, 0] - mu[:, 0].float()).abs().max() < 1e-5 and rel(stats[:M, 1], (1 / torch.sqrt(var + 1e-5))[:, 0].float()) < 1e-5
    # histograms / importance of the kernel's own routing, per 128-row block
    for blk in range(nblk):
        sl = slice(blk * 128, min(M, blk * 128 + 128))
        for br in range(NB):
            ii, vv = idx[sl, br].long(), vals[sl, br]
            allc = torch.bincount(ii.reshape(-1), minlength=E)
            top1 = torch.bincount(ii[:, 0], minlength=E)
            im = torch.zeros(E, device=DEV).index_add_(0, ii.reshape(-1), vv.reshape(-1))
            assert torch.equal(hist[blk, 0, br * E:(br + 1) * E].long(), allc)
            assert torch.equal(hist[blk, 1, br * E:(br + 1) * E].long(), top1)
            assert (imp[blk, br * E:(br + 1) * E] - im).abs().max() < 1e-4


@pytest.mark.parametrize("M,K_in,N_out", [(1000, 512, 512), (25088, 512, 1536), (300, 264, 128), (4097, 1024, 512)])
def test_linear_backward_building_blocks(M, K_in, N_out):
    """dX = dY W, dW = dY^T X (token slabs as the groups of one grouped GEMM + fp32 partial sums), db = column sums:
    the gradient GEMMs of a Linear on the forward tcgen05 kernel, against torch in fp32 on the same bf16 operands."""
    x = randn(M, K_in, seed=1).bfloat16()
    W = randn(N_out, K_in, seed=2, scale=K_in ** -0.5).bfloat16()
    dy = randn(M, N_out, seed=3).bfloat16()
    dx = torch.empty(M, K_in, device=DEV, dtype=torch.bfloat16)
    dW = torch.empty(N_out, K_in, device=DEV)
    db = torch.empty(N_out, device=DEV)
    ops.linear_backward(x, W, dy, dx=dx, dW=dW, db=db)
    assert rel(dx, dy.float() @ W.float()) < 4e-3                       # bf16 output rounding
    assert rel(dW, dy.float().t() @ x.float()) < 1e-5
    assert rel(db, dy.float().sum(0)) < 1e-5
    dW2, db2 = dW.clone(), db.clone()
    ops.linear_backward(x, W, dy, dW=dW2, db=db2, accumulate=True)
    assert rel(dW2, 2 * dW) < 1e-6 and rel(db2, 2 * db) < 1e-6


def test_gemm_grouped_per_tile_k_range():
    """Grouped GEMM whose tiles contract over different column ranges (MdmGemmEpi.tile_k, device-resident): the shape of
    an expert's weight gradient dW_e = dY_e^T X_e, where the expert's rows are a segment of the permuted buffers."""
    F_, D_, cap = 256, 128, 1024
    offs, lens = [0, 256, 640], [200, 384, 100]
    dyT = torch.zeros(F_, cap, device=DEV, dtype=torch.bfloat16)
    xT = torch.zeros(D_, cap, device=DEV, dtype=torch.bfloat16)
    for e, (o, n) in enumerate(zip(offs, lens)):          # padding columns of a segment stay zero
        dyT[:, o:o + n] = randn(F_, n, seed=10 + e).bfloat16()
        xT[:, o:o + n] = randn(D_, n, seed=20 + e).bfloat16()
    tiles, tk = [], []
    for e, (o, n) in enumerate(zip(offs, lens)):
        for mt in range(F_ // 128):
            tiles.append([mt * 128, e * F_ + mt * 128, 0, 128])
            tk.append([o, n])
    tt = torch.tensor(tiles, dtype=torch.int32, device=DEV)
    tkd = torch.tensor(tk, dtype=torch.int32, device=DEV)
    out = torch.full((3 * F_, D_), 7.0, device=DEV)
    ops.gemm(dyT, xT, None, out_f32=out, N=D_, M=3 * F_, tiles=tt, num_tiles=len(tiles), a_rows=F_, w_rows=D_, tile_k=tkd)
    for e, (o, n) in enumerate(zip(offs, lens)):
        ref = dyT[:, o:o + n].float() @ xT[:, o:o + n].float().t()
        assert rel(out[e * F_:(e + 1) * F_], ref) < 1e-5, e


def test_expert_ffn_backward_building_block():
    """Backward of the grouped expert FFN (Linear -> GELU -> Linear per expert, gate weight folded into the second Linear)
    on the tcgen05 grouped GEMM, including an expert that received no token, against torch autograd."""
    G, D, Fd = 3, 128, 256
    cnt = [200, 0, 130]
    off, tiles_rows, o = [], [], 0
    for g_, c in enumerate(cnt):
        off.append(o)
        for i in range((c + 127) // 128):
            tiles_rows.append((o + i * 128, g_))
        o += (c + 127) // 128 * 128
    cap = o + 128
    bf = torch.bfloat16
    W1 = randn(G * Fd, D, seed=1, scale=D ** -0.5).to(bf)
    b1 = randn(G * Fd, seed=2, scale=0.1)
    W2 = randn(G * D, Fd, seed=3, scale=Fd ** -0.5).to(bf)
    xp = torch.zeros(cap, D, device=DEV, dtype=bf)
    d_yp = torch.zeros(cap, D, device=DEV, dtype=bf)
    rs = torch.zeros(cap, device=DEV)
    for g_, c in enumerate(cnt):
        xp[off[g_]:off[g_] + c] = randn(c, D, seed=10 + g_).to(bf)
        d_yp[off[g_]:off[g_] + c] = randn(c, D, seed=20 + g_).to(bf)
        rs[off[g_]:off[g_] + c] = torch.rand(c, generator=gen(30 + g_)).to(DEV) * 0.5
    # forward pre-activation as the training forward would save it (grouped GEMM without activation)
    tt = torch.tensor([[r0, r0, g_ * Fd, 128] for r0, g_ in tiles_rows], dtype=torch.int32, device=DEV)
    pre = torch.zeros(cap, Fd, device=DEV, dtype=bf)
    ops.gemm(xp, W1, b1, out_a=pre, N=Fd, M=cap, tiles=tt, num_tiles=len(tiles_rows), a_rows=cap, w_rows=G * Fd)
    for g_, c in enumerate(cnt):                          # rows of a tile beyond the segment's count carry the bias: clear
        pre[off[g_] + c:off[g_] + (c + 127) // 128 * 128] = 0
    seg_off = torch.tensor(off, dtype=torch.int32, device=DEV)
    seg_cnt = torch.tensor(cnt, dtype=torch.int32, device=DEV)
    d_xp, dW1, dW2, db1, db2 = ops.expert_ffn_backward(xp, pre, W1, W2, rs, d_yp, seg_off, seg_cnt, tiles_rows, F=Fd, D=D)
    for g_, c in enumerate(cnt):
        sl = slice(off[g_], off[g_] + c)
        w1 = W1[g_ * Fd:(g_ + 1) * Fd].float().requires_grad_(True)
        bb1 = b1[g_ * Fd:(g_ + 1) * Fd].clone().requires_grad_(True)
        w2 = W2[g_ * D:(g_ + 1) * D].float().requires_grad_(True)
        bb2 = torch.zeros(D, device=DEV, requires_grad=True)
        xin = xp[sl].float().requires_grad_(True)
        if c:
            q_ = xin @ w1.t() + bb1
            p_ = q_ + (q_.to(bf).float() - q_).detach()   # the saved pre-activation is bf16 (straight-through rounding)
            y = (F.gelu(p_) @ w2.t() + bb2) * rs[sl, None]
            y.backward(d_yp[sl].float())
            assert rel(d_xp[sl], xin.grad) < 2e-2, g_
            assert rel(dW1[g_ * Fd:(g_ + 1) * Fd], w1.grad) < 2e-2, g_
            assert rel(dW2[g_ * D:(g_ + 1) * D], w2.grad) < 2e-2, g_
            assert rel(db1[g_], bb1.grad) < 2e-2 and rel(db2[g_], bb2.grad) < 2e-2, g_
        else:
            assert dW1[g_ * Fd:(g_ + 1) * Fd].abs().max() == 0 and dW2[g_ * D:(g_ + 1) * D].abs().max() == 0
            assert db1[g_].abs().max() == 0 and db2[g_].abs().max() == 0


# ------------------------------------------------------------------------------------------ row pipeline
@pytest.mark.parametrize("D", [128, 256, 512, 1024])
@pytest.mark.parametrize("in_dtype,out_dtype", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                                (torch.bfloat16, torch.bfloat16)])
def test_rowop_full_pipeline(D, in_dtype, out_dtype):
    rows, T = 77, 11
    x = randn(rows, D, seed=1, scale=3.0).to(in_dtype)
    w1, b1, w2, b2 = (randn(D, seed=s) for s in (2, 3, 4, 5))
    w1, w2 = w1 * 0.1 + 1, w2 * 0.1 + 1
    film = randn((rows + T - 1) // T, 2 * D, seed=6, scale=0.3)
    o0 = torch.empty(rows, D, device=DEV, dtype=out_dtype)
    o1f = torch.empty(rows, D, device=DEV)
    o1a = torch.empty(rows, D, device=DEV, dtype=out_dtype)
    o2f = torch.empty(rows, D, device=DEV)
    o2a = torch.empty(rows, D, device=DEV, dtype=out_dtype)
    ops.rowop(x, rows, D, ops._dt(o0), ln1=(w1, b1), l2norm=True, out1_f32=o1f, out1_a=o1a, ln2=(w2, b2), film=film,
              rows_per_seq=T, silu=True, out2_f32=o2f, out2_a=o2a, out0_a=o0)
    xf = x.float()
    r1 = F.normalize(F.layer_norm(xf, (D,), w1, b1, 1e-5), dim=-1) * (D ** 0.5)
    seq = torch.arange(rows, device=DEV) // T
    r2 = F.silu(F.layer_norm(r1, (D,), w2, b2, 1e-5) * (1 + film[seq, :D]) + film[seq, D:])
    assert rel(o0, xf) < TOL[out_dtype]
    assert rel(o1f, r1) < 1e-5 and rel(o2f, r2) < 1e-5
    assert rel(o1a, r1) < TOL[out_dtype] and rel(o2a, r2) < TOL[out_dtype]


# ------------------------------------------------------------------------------------------ attention cores
def _attn_params(hd, seed):
    g = gen(seed)
    pm = torch.randn(hd, 256, generator=g)
    q, _ = torch.linalg.qr(pm, mode="reduced")
    P = (F.normalize(q, dim=0) * hd ** -0.25).contiguous().to(DEV)   # linalg.qr returns column-major Q
    nw = (1 + 0.1 * torch.randn(hd, generator=g)).to(DEV)
    nb = (0.05 * torch.randn(hd, generator=g)).to(DEV)
    return P, nw, nb


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,T,hd", [(3, 4, 196, 128), (2, 4, 98, 128), (3, 4, 60, 64), (3, 8, 196, 64), (2, 2, 8, 32)])
def test_fastattn(dtype, B, H, T, hd):
    D = H * hd
    P, nw, nb = _attn_params(hd, 1)
    qkv = randn(B * T, 3 * D, seed=2, scale=2.0).to(dtype)
    length = torch.tensor([T, max(1, T // 3), 2][:B], device=DEV)
    out = torch.empty(B * T, D, device=DEV, dtype=dtype)
    ops.fastattn(qkv, P, nw, nb, length, 0, B, H, T, hd, out)
    q, k, v = (t.view(B, T, H, hd).permute(0, 2, 1, 3) * 0.1 for t in qkv.float().view(B, T, 3, D).unbind(2))
    p = {"fa.projection_matrix": P, "fa.norm.weight": nw, "fa.norm.bias": nb}
    ref = mo.fast_attention(p, "fa", q, k, v, mo.src_mask(T, length)).permute(0, 2, 1, 3).reshape(B * T, D)
    assert rel(out, ref) < TOL[dtype]
    if dtype == torch.bfloat16 and hd in (128, 64):
        # with the packed bf16 P^T the tcgen05 kernel (attention_umma.cu) runs instead of the mma.sync one; head size 64:
        # two heads per CTA with the block-diagonal diag(P^T, P^T)
        Pt = ops.pack_fastattn_pt(P)
        assert Pt is not None and tuple(Pt.shape) == (128, 128)
        out2 = torch.full_like(out, float("nan"))
        ops.fastattn(qkv, P, nw, nb, length, 0, B, H, T, hd, out2, Pt=Pt)
        assert rel(out2, ref) < TOL[dtype]
        assert rel(out2, out) < 1e-2


def test_fastattn_fully_masked_windows():
    """Key windows that lie entirely beyond `length` are skipped by the streamed kernel: same result as the
    oracle, which multiplies them by a zero mask (fast_attention.py:60-61), including length 0."""
    B, H, T, hd = 6, 4, 196, 128
    D = H * hd
    P, nw, nb = _attn_params(hd, 1)
    qkv = randn(B * T, 3 * D, seed=5, scale=2.0).bfloat16()
    length = torch.tensor([0, 1, 16, 17, 100, 196], device=DEV)
    out = torch.empty(B * T, D, device=DEV, dtype=torch.bfloat16)
    ops.fastattn(qkv, P, nw, nb, length, 0, B, H, T, hd, out)
    q, k, v = (t.view(B, T, H, hd).permute(0, 2, 1, 3) * 0.1 for t in qkv.float().view(B, T, 3, D).unbind(2))
    p = {"fa.projection_matrix": P, "fa.norm.weight": nw, "fa.norm.bias": nb}
    ref = mo.fast_attention(p, "fa", q, k, v, mo.src_mask(T, length)).permute(0, 2, 1, 3).reshape(B, T, D)
    o = out.view(B, T, D)
    for i in range(B):
        assert rel(o[i], ref[i]) < TOL[torch.bfloat16], (i, rel(o[i], ref[i]))
    # the launch-order hint (longest sequences first) must not change a single bit
    out2 = torch.empty_like(out)
    order = torch.argsort(length, descending=True).to(torch.int32)
    ops.fastattn(qkv, P, nw, nb, length, 0, B, H, T, hd, out2, seq_order=order)
    assert torch.equal(out, out2)
    # tcgen05 kernel (selected by the packed bf16 P^T): same masking semantics, every sequence within tolerance,
    # and independent of the launch order too
    Pt = P.t().contiguous().to(torch.bfloat16)
    out3 = torch.full_like(out, float("nan"))
    ops.fastattn(qkv, P, nw, nb, length, 0, B, H, T, hd, out3, Pt=Pt)
    o3 = out3.view(B, T, D)
    for i in range(B):
        assert rel(o3[i], ref[i]) < TOL[torch.bfloat16], (i, rel(o3[i], ref[i]))
    out4 = torch.empty_like(out)
    ops.fastattn(qkv, P, nw, nb, length, 0, B, H, T, hd, out4, seq_order=order, Pt=Pt)
    assert torch.equal(out3, out4)


def test_fastattn_length_shift():
    B, H, T, hd = 2, 4, 98, 128
    P, nw, nb = _attn_params(hd, 1)
    qkv = randn(B * T, 3 * H * hd, seed=2).bfloat16()
    length = torch.tensor([196, 77], device=DEV)
    a = torch.empty(B * T, H * hd, device=DEV, dtype=torch.bfloat16)
    b = torch.empty_like(a)
    ops.fastattn(qkv, P, nw, nb, length, 1, B, H, T, hd, a)
    ops.fastattn(qkv, P, nw, nb, (length / 2).long(), 0, B, H, T, hd, b)
    assert torch.equal(a, b)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,T,hd,Nt", [(3, 4, 196, 128, 20), (2, 4, 60, 64, 85), (2, 2, 8, 32, 3)])
def test_linear_cross_attention(dtype, B, H, T, hd, Nt):
    D = H * hd
    q = randn(B * T, D, seed=1, scale=2.0).to(dtype)
    k = randn(B, Nt, D, seed=2, scale=2.0).to(dtype)
    v = randn(B, Nt, D, seed=3).to(dtype)
    nt = torch.tensor([Nt, max(1, Nt // 2), 1][:B], dtype=torch.int32, device=DEV)
    ctx = torch.empty(B, H, hd, hd, device=DEV)
    ops.lincross_ctx(k, v, nt, B, Nt, H, hd, ctx)
    y = torch.empty(B * T, D, device=DEV, dtype=dtype)
    ops.lincross_apply(q, ctx, B, T, H, hd, y)
    kf = k.float().view(B, Nt, H, hd)
    pad = torch.arange(Nt, device=DEV)[None, :] >= nt[:, None]
    ks = F.softmax(kf.masked_fill(pad[:, :, None, None], float("-inf")), dim=1)
    att = torch.einsum("bnhd,bnhl->bhdl", ks, v.float().view(B, Nt, H, hd).masked_fill(pad[:, :, None, None], 0))
    assert rel(ctx, att) < 1e-5
    qs = F.softmax(q.float().view(B, T, H, hd), dim=-1)
    ref = torch.einsum("bnhd,bhdl->bnhl", qs, att).reshape(B * T, D)
    assert rel(y, ref) < TOL[dtype]
    if dtype == torch.bfloat16 and hd in (128, 64):
        # with ctx^T packed as bf16 the tcgen05 kernel (attention_umma.cu) runs; also at the half resolution.  Head size
        # 64: two heads per CTA, ctx^T as the block-diagonal [128, 128] of each pair of heads
        ctxT = ops.pack_lincross_ctxT(ctx)
        if hd == 128:
            assert torch.equal(ctxT, ctx.transpose(-1, -2).contiguous().to(torch.bfloat16))
        else:
            assert tuple(ctxT.shape) == (B, H // 2, 128, 128)
            assert torch.equal(ctxT[:, :, 64:, 64:], ctx.transpose(-1, -2)[:, 1::2].to(torch.bfloat16))
            assert float(ctxT[:, :, :64, 64:].abs().max()) == 0.0
        y2 = torch.full_like(y, float("nan"))
        ops.lincross_apply(q, ctx, B, T, H, hd, y2, ctxT=ctxT)
        assert rel(y2, ref) < TOL[dtype]
        Th = T // 2
        qh = q.view(B, T, D)[:, :Th].contiguous().view(B * Th, D)
        yh = torch.full((B * Th, D), float("nan"), device=DEV, dtype=dtype)
        ops.lincross_apply(qh, ctx, B, Th, H, hd, yh, ctxT=ctxT)
        assert rel(yh, ref.view(B, T, D)[:, :Th].reshape(B * Th, D)) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,T,hd,Nt", [(3, 4, 196, 128, 20), (2, 4, 60, 64, 85), (2, 2, 8, 32, 10), (3, 4, 98, 128, 85),
                                         (2, 4, 196, 128, 40), (2, 2, 130, 128, 96), (3, 8, 196, 64, 40), (2, 3, 98, 64, 96)])
def test_softmax_cross_attention(dtype, B, H, T, hd, Nt):
    D = H * hd
    q = randn(B * T, D, seed=1, scale=2.0).to(dtype)
    k = randn(B, Nt, D, seed=2).to(dtype)
    v = randn(B, Nt, D, seed=3).to(dtype)
    nt = torch.tensor([Nt, max(1, Nt // 2), 1][:B], dtype=torch.int32, device=DEV)
    o = torch.empty(B * T, D, device=DEV, dtype=dtype)
    ops.softmax_cross(q, k, v, nt, B, T, Nt, H, hd, o)
    qh = q.float().view(B, T, H, hd).permute(0, 2, 1, 3)
    kh = k.float().view(B, Nt, H, hd).permute(0, 2, 1, 3)
    vh = v.float().view(B, Nt, H, hd).permute(0, 2, 1, 3)
    s = torch.einsum("bhqd,bhkd->bhqk", qh * (hd ** -0.5), kh)
    pad = torch.arange(Nt, device=DEV)[None, :] >= nt[:, None]
    s = s.masked_fill(pad[:, None, None, :], float("-inf"))
    ref = torch.einsum("bhqk,bhkd->bhqd", F.softmax(s, -1), vh).permute(0, 2, 1, 3).reshape(B * T, D)
    assert rel(o, ref) < TOL[dtype]


# ------------------------------------------------------------------------------------------ routing
@pytest.mark.parametrize("stages", ["full", "ln", "ln2_film_silu", "l2"])
@pytest.mark.parametrize("gdt", [torch.float32, torch.bfloat16])
def test_rowop_backward_matches_autograd(stages, gdt):
    """mdm_rowop_bwd (backward twin of the row pipeline, intermediates recomputed) against torch autograd of the same chain:
    input gradient, LayerNorm affine gradients, per-sequence FiLM gradients."""
    D, T, B = 512, 50, 5
    M = B * T - 7                                  # the last sequence is ragged
    x = randn(M, D, seed=1, scale=1.5)
    ln1 = (torch.rand(D, generator=gen(5)).to(DEV) + 0.5, randn(D, seed=6, scale=0.1))
    ln2 = (torch.rand(D, generator=gen(7)).to(DEV) + 0.5, randn(D, seed=8, scale=0.1))
    film = randn(B, 2 * D, seed=9, scale=0.3)
    dout = randn(M, D, seed=10).to(gdt)
    kw = {"full": dict(ln1=ln1, l2norm=True, ln2=ln2, film=film, rows_per_seq=T, silu=True),
          "ln": dict(ln1=ln1), "ln2_film_silu": dict(ln2=ln2, film=film, rows_per_seq=T, silu=True),
          "l2": dict(l2norm=True)}[stages]
    din, grads = ops.rowop_bwd(x, M, D, dout, **kw)
    xr = x.clone().requires_grad_(True)
    leaves = {}
    v = xr
    if "ln1" in kw:
        w, b_ = (t.clone().requires_grad_(True) for t in ln1)
        leaves["ln1_w"], leaves["ln1_b"] = w, b_
        v = F.layer_norm(v, (D,), w, b_)
    if kw.get("l2norm"):
        v = F.normalize(v, dim=-1) * math.sqrt(D)
    if "ln2" in kw:
        w, b_ = (t.clone().requires_grad_(True) for t in ln2)
        leaves["ln2_w"], leaves["ln2_b"] = w, b_
        v = F.layer_norm(v, (D,), w, b_)
    if "film" in kw:
        fl = film.clone().requires_grad_(True)
        leaves["film"] = fl
        seq = torch.arange(M, device=DEV) // T
        v = v * (1 + fl[seq, :D]) + fl[seq, D:]
    if kw.get("silu"):
        v = F.silu(v)
    v.backward(dout.float())
    tol = 1e-4 if gdt == torch.float32 else 6e-3
    assert rel(din, xr.grad) < tol
    for k, leaf in leaves.items():
        assert rel(grads[k], leaf.grad) < 1e-4, k
    assert set(grads) == set(leaves)


@pytest.mark.parametrize("E", [4, 8])
def test_softmax_topk_bit_exact_vs_torch_cuda(E):
    """Expert routing indices bit-exact against torch.softmax + torch.topk ON CUDA (the reference's
    routing on a GPU, models/switch_moe.py:53-57), including heavy ties."""
    g = gen(7)
    n = 1 << 15
    smooth = torch.randn(n, E, generator=g) * 3
    coarse = torch.randint(-2, 3, (n, E), generator=g).float() * 0.5       # many exact ties
    bf = (torch.randn(n, E, generator=g)).bfloat16().float()                # bf16-valued logits
    edge = torch.tensor([[0.0] * E, [1.0] + [0.0] * (E - 1), [0.0] * (E - 1) + [1.0], [-0.0] + [0.0] * (E - 1),
                         [1e-30] * E, [80.0] + [-80.0] * (E - 1)])
    logits = torch.cat([smooth, coarse, bf, edge]).to(DEV)
    probs, idx, vals = ops.softmax_topk(logits)
    rp = torch.softmax(logits, dim=1)
    rv, ri = torch.topk(rp, 2, dim=1)
    assert torch.equal(probs, rp)       # softmax bit-exact (ATen arithmetic order)
    assert torch.equal(idx, ri)         # indices bit-exact, ties included
    assert torch.equal(vals, rv)
    ov, oi = mo.top2_cuda_order(rp)      # and the oracle's explicit tie rule agrees with torch CUDA
    assert torch.equal(oi, ri) and torch.equal(ov, rv)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("N,D,Fd,E", [(1000, 256, 512, 4), (3000, 512, 1024, 8), (37, 128, 256, 4)])
def test_moe_multibranch(dtype, N, D, Fd, E):
    """gate -> scan -> permute -> grouped FFN -> combine+FiLM against the oracle's MoE."""
    T = 10 if N % 10 == 0 else N
    Bn = N // T
    g = gen(3)
    p = {}
    for b in range(2):
        br = "ffn.branches.%d" % b
        p[br + ".layernorm.weight"] = (1 + 0.1 * torch.randn(D, generator=g)).to(DEV)
        p[br + ".layernorm.bias"] = (0.05 * torch.randn(D, generator=g)).to(DEV)
        p[br + ".moe.gate.weight"] = (torch.randn(E, D, generator=g) * 4 / math.sqrt(D)).to(DEV)
        p[br + ".moe.gate.bias"] = (0.05 * torch.randn(E, generator=g)).to(DEV)
        for e in range(E):
            p["%s.moe.experts.%d.0.weight" % (br, e)] = (torch.randn(Fd, D, generator=g) / math.sqrt(D)).to(DEV)
            p["%s.moe.experts.%d.0.bias" % (br, e)] = (0.05 * torch.randn(Fd, generator=g)).to(DEV)
            p["%s.moe.experts.%d.2.weight" % (br, e)] = (torch.randn(D, Fd, generator=g) / math.sqrt(Fd)).to(DEV)
            p["%s.moe.experts.%d.2.bias" % (br, e)] = (0.05 * torch.randn(D, generator=g)).to(DEV)
    sw = (1 + 0.1 * torch.randn(D, generator=g)).to(DEV)
    sb = (0.05 * torch.randn(D, generator=g)).to(DEV)
    film = (0.3 * torch.randn(Bn, 2 * D, generator=g)).to(DEV)
    x = (torch.randn(N, D, generator=g) * 2).to(DEV)
    # oracle
    routing, counters = [], {}
    out = 0
    for b in range(2):
        br = "ffn.branches.%d" % b
        h, idx_o, vals_o = mo.switch_moe(p, br + ".moe", mo._ln(p, br + ".layernorm", x.view(Bn, T, D)), E, counters)
        routing.append((idx_o, vals_o))
        out = out + h
    out = (out / 2).view(N, D)
    seq = torch.arange(N, device=DEV) // T
    ref = F.silu(F.layer_norm(out, (D,), sw, sb, 1e-5) * (1 + film[seq, :D]) + film[seq, D:])
    # product kernels
    NB, NBK, G = 2, 4, 2 * E
    cap, nblk = NBK * N + G * 128, (N + 127) // 128
    i32, f32 = torch.int32, torch.float32
    z = lambda *s, dt=f32: torch.zeros(*s, device=DEV, dtype=dt)
    idx, vals, stats = z(N, NB, 2, dt=i32), z(N, NB, 2), z(N, 2)
    hist, imp, base, seg = z(nblk, 2, G, dt=i32), z(nblk, G), z(nblk, G, dt=i32), z(G + 1, dt=i32)
    t_up, t_dn, ntile = z(cap // 128, 4, dt=i32), z(cap // 128, 4, dt=i32), z(1, dt=i32)
    perm, rscale = z(N, NBK, dt=i32), z(cap)
    xp, hp, yp = z(cap, D, dt=dtype), z(cap, Fd, dt=dtype), z(cap, D, dt=dtype)
    usage, importance = z(G), z(G)
    ln_w = torch.stack([p["ffn.branches.%d.layernorm.weight" % b] for b in range(2)]).contiguous()
    ln_b = torch.stack([p["ffn.branches.%d.layernorm.bias" % b] for b in range(2)]).contiguous()
    gw = torch.cat([p["ffn.branches.%d.moe.gate.weight" % b] for b in range(2)]).contiguous()
    gb = torch.cat([p["ffn.branches.%d.moe.gate.bias" % b] for b in range(2)]).contiguous()
    cat = lambda fmt, dt: torch.cat([p[fmt % (b, e)] for b in range(2) for e in range(E)]).to(dt).contiguous()
    w1, b1 = cat("ffn.branches.%d.moe.experts.%d.0.weight", dtype), cat("ffn.branches.%d.moe.experts.%d.0.bias", f32)
    w2, b2 = cat("ffn.branches.%d.moe.experts.%d.2.weight", dtype), cat("ffn.branches.%d.moe.experts.%d.2.bias", f32)
    ops.moe_gate(x, N, D, NB, E, ln_w, ln_b, gw, gb, idx, vals, stats, hist, imp)
    ops.moe_scan(hist, imp, idx, N, NB, E, Fd, D, base, seg, t_up, t_dn, ntile, usage, importance)
    ops.moe_permute(x, N, D, NB, E, ln_w, ln_b, idx, vals, stats, base, seg, xp, perm, rscale)
    kw = dict(num_tiles=cap // 128, num_tiles_dev=ntile, M=cap, a_rows=cap)
    if dtype == torch.float32:
        ops.gemm(xp, w1, b1, act=ACT_GELU, out_f32=hp, N=Fd, tiles=t_up, w_rows=G * Fd, **kw)
        ops.gemm(hp, w2, b2, out_f32=yp, N=D, rowscale=rscale, tiles=t_dn, w_rows=G * D, **kw)
    else:
        ops.gemm(xp, w1, b1, act=ACT_GELU, out_a=hp, N=Fd, tiles=t_up, w_rows=G * Fd, **kw)
        ops.gemm(hp, w2, b2, out_a=yp, N=D, rowscale=rscale, tiles=t_dn, w_rows=G * D, **kw)
    res = torch.empty(N, D, device=DEV, dtype=dtype)
    ops.moe_combine_film(yp, perm, N, D, NBK, sw, sb, film, T, res)
    # routing: the gate's logits are a different fp32 summation order than cuBLAS, so indices may only
    # differ where the oracle's own top-2/top-3 margin is at rounding level
    for b in range(2):
        io, vo = routing[b]
        mism = (idx[:, b].long() != io).any(dim=1)
        assert mism.float().mean().item() < 2e-3
        if mism.any():
            br = "ffn.branches.%d" % b
            pr = F.softmax(mo._lin(p, br + ".moe.gate", mo._ln(p, br + ".layernorm", x)), dim=1)
            top3 = torch.topk(pr[mism], 3, dim=1).values
            margin = torch.minimum(top3[:, 0] - top3[:, 1], top3[:, 1] - top3[:, 2])
            assert margin.max().item() < 1e-5
    # permutation invariants: perm is a bijection onto the valid rows; segments are 128-aligned
    assert perm.unique().numel() == N * NBK
    assert torch.all(seg % 128 == 0)
    assert int(ntile) == int(seg[-1]) // 128
    u = torch.cat([counters["ffn.branches.%d.moe.expert_usage" % b] for b in range(2)])
    if dtype == torch.float32:
        assert (usage - u).abs().sum().item() <= 2 * N * 2e-3 + 1
    clean = ~((idx[:, 0].long() != routing[0][0]).any(1) | (idx[:, 1].long() != routing[1][0]).any(1))
    assert rel(res[clean], ref[clean]) < TOL[dtype]


# ------------------------------------------------------------------------------------------ small ops / sampler
def test_timestep_embedding_and_gated_mix():
    B, D = 9, 512
    t = torch.tensor([0, 1, 2, 10, 100, 500, 998, 999, 37], device=DEV)
    out = torch.empty(B, D, device=DEV)
    ops.timestep_embedding(t, B, D, out)
    assert rel(out, mo.timestep_embedding(t, D)) < 1e-5
    a, b = randn(B, D, seed=1), randn(B, D, seed=2)
    o = torch.empty(B, D, device=DEV)
    ops.gated_mix(a, b, o)
    gte = torch.sigmoid(a + b)
    assert rel(o, gte * a + (1 - gte) * b) < 1e-6


@pytest.mark.parametrize("clip", [False, True])
def test_cfg_update_bit_exact(clip):
    """The fused CFG/DDPM update equals the torch-eager op sequence of p_sample_with_cfg bit for bit."""
    B, T, Fd = 5, 196, 263
    tab = mo.diffusion_tables(1000)
    x, ec, eu, nz = (randn(B, T, Fd, seed=s) for s in (1, 2, 3, 4))
    t = torch.tensor([999, 500, 1, 0, 37], device=DEV)
    import numpy as np
    from motiondiffusion_moe_b200 import GaussianDiffusion, get_named_beta_schedule
    d = GaussianDiffusion(betas=get_named_beta_schedule("linear", 1000))
    step, _ = d._tables(torch.device(DEV))
    xp, x0 = torch.empty_like(x), torch.empty_like(x)
    ops.cfg_update(x, ec, eu, nz, t, step, 1000, 7.5, clip, xp, x0)
    rs, r0 = mo.cfg_update(tab, x, t, ec, eu, nz, 7.5, clip)
    assert torch.equal(x0, r0)
    assert torch.equal(xp, rs)
    assert np.allclose(d.posterior_log_variance_clipped, tab["posterior_log_variance_clipped"])


@pytest.mark.parametrize("eta", [0.0, 0.5, 1.0])
@pytest.mark.parametrize("clip", [False, True])
def test_ddim_update_bit_exact(eta, clip):
    """mdm_ddim_update against the reference's own ddim_sample outputs (tests/golden/ddim_step.npz) and against
    the oracle for the guided / strided variants: bit for bit."""
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ddim_step.npz"))
    x, eps, t, noise = (torch.from_numpy(g[k]).to(DEV) for k in ("x", "eps", "t", "noise"))
    tab = mo.diffusion_tables(1000)
    tab4 = torch.from_numpy(np.stack([tab["sqrt_recip_alphas_cumprod"], tab["sqrt_recipm1_alphas_cumprod"],
                                      tab["alphas_cumprod"], np.append(1.0, tab["alphas_cumprod"][:-1])])
                            .astype(np.float32)).to(DEV)
    out, x0 = torch.empty_like(x), torch.empty_like(x)
    ops.ddim_update(x, eps, t, tab4, 1000, eta, clip, out, noise=noise, x0=x0)
    assert torch.equal(x0.cpu(), torch.from_numpy(g["x0_eta%g_clip%d" % (eta, int(clip))]))
    # The golden was produced by the reference on the CPU, where torch.sqrt(float32) is NOT correctly rounded
    # (0.6 % of inputs are off by one ulp against IEEE sqrt); the kernel uses sqrt.rn like torch's CUDA sqrt.
    # So: 1e-6 against the CPU golden, bit-identical against the same op sequence evaluated by torch on the GPU.
    assert rel(out.cpu(), torch.from_numpy(g["sample_eta%g_clip%d" % (eta, int(clip))])) < 1e-6
    assert torch.equal(out, mo.ddim_update(tab, x, t, x0, eta, noise))
    # classifier-free guidance on pred_xstart + strided predecessor
    eps_u = randn(*x.shape, seed=8)
    t_prev = torch.tensor([-1, -1, 3, 250, 700, 980], device=DEV)
    ops.ddim_update(x, eps, t, tab4, 1000, eta, clip, out, eps_u=eps_u, cfg_scale=7.5, noise=noise, t_prev=t_prev, x0=x0)
    _, guided = mo.cfg_update(tab, x, t, eps, eps_u, noise, 7.5, clip)
    ref = mo.ddim_update(tab, x, t, guided, eta, noise, t_prev=t_prev)
    assert torch.equal(x0, guided)
    assert torch.equal(out, ref)


@pytest.mark.parametrize("tag,J", [("t2m", 22), ("kit", 21)])
def test_recover_from_ric_matches_reference_golden(tag, J):
    """mdm_recover_from_ric (de-normalise + recover_root_rot_pos + recover_from_ric in one kernel) against the
    reference's own outputs; the cumulative sums run in a different order than torch.cumsum, hence 1e-5."""
    import os
    import numpy as np
    import motiondiffusion_moe_b200 as mdm
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ric.npz"))
    x, mean, std = (torch.from_numpy(g[tag + k]).to(DEV) for k in ("_x", "_mean", "_std"))
    ref = torch.from_numpy(g[tag + "_joints"]).to(DEV)
    got = mdm.recover_from_ric(x, J, mean, std)
    assert got.shape == ref.shape
    assert rel(got, ref) < 1e-5 and (got - ref).abs().max() < 1e-3
    raw = x * std + mean
    assert torch.equal(mdm.recover_from_ric(raw, J), got)              # already de-normalised input
    long = torch.randn(2, 700, 263, generator=gen(3)).to(DEV) * 0.1    # more frames than threads in a block
    assert rel(mdm.recover_from_ric(long, 22), mo.recover_from_ric(long, 22)) < 1e-5


def test_q_sample_bit_exact():
    from motiondiffusion_moe_b200 import GaussianDiffusion, get_named_beta_schedule
    d = GaussianDiffusion(betas=get_named_beta_schedule("linear", 1000))
    x0, nz = randn(4, 60, 251, seed=1), randn(4, 60, 251, seed=2)
    t = torch.tensor([0, 999, 17, 500], device=DEV)
    assert torch.equal(d.q_sample(x0, t, nz), mo.q_sample(mo.diffusion_tables(1000), x0, t, nz))


def test_cpu_and_strided_tensors_fail_loudly():
    from motiondiffusion_moe_b200 import MdmError
    with pytest.raises(MdmError):
        ops.gemm(torch.zeros(4, 8), torch.zeros(4, 8), out_f32=torch.zeros(4, 4))
    P, nw, nb = _attn_params(32, 1)
    qkv = torch.zeros(8, 3 * 64, device=DEV)
    with pytest.raises(MdmError):   # column-major projection matrix
        ops.fastattn(qkv, P.t().contiguous().t(), nw, nb, None, 0, 1, 2, 8, 32, torch.empty(8, 64, device=DEV))
