"""Per-op timing at the config-2 shapes (N = 128 sequences x 196 frames, D 512, F 1024, 16 expert groups),
each op alone, CUDA events, buffers rotated through > 126 MB so nothing is L2-resident between calls.
Prints achieved TFLOP/s (GEMMs) or GB/s of algorithmic bytes (row / routing / attention kernels).
MDM_B200_LIB=<path> selects a library variant (tools/build_variant.sh)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motiondiffusion_moe_b200 import ops
from motiondiffusion_moe_b200._lib import ACT_NONE, ACT_GELU, MDM_BF16
dev = torch.device("cuda")
torch.manual_seed(0)
bf, f32 = torch.bfloat16, torch.float32
ONLY = os.environ.get("ONLY", "")
NSEQ = int(os.environ.get("NSEQ", "128"))
T = int(os.environ.get("T", "196"))
N, D, H = NSEQ * T, 512, 4


def timeit(name, fn, nset, work, unit, iters=24):
    """Device time per call: `iters` calls captured into one CUDA graph (no Python / launch overhead in
    the timed region; several ops here run for less than a ctypes call takes), replayed and timed."""
    if ONLY and ONLY not in name:
        return
    for i in range(2):
        fn(i % nset)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(iters):
                fn(i % nset)
    g.replay()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / iters * 1e3
    print("%-44s %8.1f us  %8.1f %s" % (name, us, work / us / (1e6 if unit == "TFLOP/s" else 1e3), unit), flush=True)


def gemm_case(name, M, Nn, K, act=ACT_NONE, f32out=False, resid=False, bf16out=True, rowscale=False):
    nset = 4
    A = [torch.randn(M, K, device=dev).to(bf) for _ in range(nset)]
    W = (torch.randn(Nn, K, device=dev) / K ** 0.5).to(bf)
    b = torch.randn(Nn, device=dev)
    R = [torch.randn(M, Nn, device=dev) for _ in range(nset)] if resid else None
    Of = [torch.empty(M, Nn, device=dev) for _ in range(nset)] if f32out else None
    Ob = [torch.empty(M, Nn, device=dev, dtype=bf) for _ in range(nset)] if bf16out else None
    rs = torch.rand(M, device=dev) if rowscale else None

    def fn(i):
        ops.gemm(A[i], W, b, act=act, out_f32=Of[i] if Of else None, out_a=Ob[i] if Ob else None,
                 resid=R[i] if R else None, alpha=1.0, beta=1.0 if resid else 0.0, rowscale=rs)
    timeit(name, fn, nset, 2.0 * M * Nn * K, "TFLOP/s")


gemm_case("gemm qkv      N x1536x512  bf16", N, 1536, 512)
gemm_case("gemm p0       N x512 x512  gelu bf16", N, 512, 512, act=ACT_GELU)
gemm_case("gemm p3       N x512 x512  bf16", N, 512, 512)
gemm_case("gemm s_out    N x512 x512  f32+resid", N, 512, 512, f32out=True, resid=True, bf16out=False)
gemm_case("gemm skip     N x512 x512  gelu f32+resid", N, 512, 512, act=ACT_GELU, f32out=True, resid=True, bf16out=False)
gemm_case("gemm ffn_out  N x512 x512  f32+resid+bf16", N, 512, 512, f32out=True, resid=True, bf16out=True)
gemm_case("gemm up       4N x1024x512 gelu bf16", 4 * N, 1024, 512, act=ACT_GELU)
gemm_case("gemm down     4N x512x1024 rowscale bf16", 4 * N, 512, 1024, rowscale=True)
gemm_case("gemm f1       N x2048x512  gelu bf16", N, 2048, 512, act=ACT_GELU)
gemm_case("gemm f3       N x512x2048  f32+resid", N, 512, 2048, f32out=True, resid=True, bf16out=False)

# ---- row pipeline
nset = 4
xb = [torch.randn(N, D, device=dev).to(bf) for _ in range(nset)]
xf = [torch.randn(N, D, device=dev) for _ in range(nset)]
ob = [torch.empty(N, D, device=dev, dtype=bf) for _ in range(nset)]
ob2 = [torch.empty(N, D, device=dev, dtype=bf) for _ in range(nset)]
of = [torch.empty(N, D, device=dev) for _ in range(nset)]
ln = (torch.rand(D, device=dev) + 0.5, torch.randn(D, device=dev))
ln2 = (torch.rand(D, device=dev) + 0.5, torch.randn(D, device=dev))
film = torch.randn(NSEQ, 2 * D, device=dev)
timeit("rowop bf16->bf16 ln+l2+ln+film+silu", lambda i: ops.rowop(xb[i], N, D, MDM_BF16, ln1=ln, l2norm=True, ln2=ln2, film=film,
       rows_per_seq=T, silu=True, out2_a=ob[i]), nset, N * D * 4, "GB/s")
timeit("rowop bf16->bf16 ln", lambda i: ops.rowop(xb[i], N, D, MDM_BF16, ln1=ln, out1_a=ob[i]), nset, N * D * 4, "GB/s")
timeit("rowop f32->f32+bf16+bf16 ln,ln", lambda i: ops.rowop(xf[i], N, D, MDM_BF16, ln1=ln, out1_f32=of[i], ln2=ln2, out2_a=ob[i],
       out0_a=ob2[i]), nset, N * D * 12, "GB/s")
timeit("rowop f32->bf16 ln", lambda i: ops.rowop(xf[i], N, D, MDM_BF16, ln1=ln, out1_a=ob[i]), nset, N * D * 6, "GB/s")

# ---- row pipeline fused into the residual-stream GEMM (a1 -> LN, L2, LN, FiLM, SiLU -> s_out Linear + residual)
Wf = (torch.randn(D, D, device=dev) / D ** 0.5).to(bf)
bfv = torch.randn(D, device=dev)
kwf = dict(ln1=ln, l2norm=True, ln2=ln2, film=film, rows_per_seq=T, silu=True)
timeit("gemm_rowop fused (rowop31 + s_out)", lambda i: ops.gemm_rowop(xb[i], N, D, Wf, bfv, out_f32=of[i], resid=xf[i], alpha=0.1, beta=1.0, **kwf),
       nset, 2.0 * N * D * D, "TFLOP/s")
def _unfused(i):
    ops.rowop(xb[i], N, D, MDM_BF16, out2_a=ob[i], **kwf)
    ops.gemm(ob[i], Wf, bfv, out_f32=of[i], resid=xf[i], alpha=0.1, beta=1.0)
timeit("rowop31 + gemm s_out (unfused pair)", _unfused, nset, 2.0 * N * D * D, "TFLOP/s")

# ---- Linear + row pipeline in one kernel (mdm_gemm_ln: the row stays in TMEM) against the pairs it replaces
W2k = (torch.randn(D, 4 * D, device=dev) / (4 * D) ** 0.5).to(bf)
xb4 = [torch.randn(N, 4 * D, device=dev).to(bf) for _ in range(2)]
of2 = [torch.empty(N, D, device=dev) for _ in range(nset)]
timeit("gemm_ln p3 + ln,l2,ln,film,silu", lambda i: ops.gemm_ln(xb[i], Wf, bfv, out2_a=ob[i], **kwf), nset, 2.0 * N * D * D, "TFLOP/s")
def _pair1(i):
    ops.gemm(xb[i], Wf, bfv, out_a=ob2[i])
    ops.rowop(ob2[i], N, D, MDM_BF16, out2_a=ob[i], **kwf)
timeit("  unfused: gemm p3 + rowop31", _pair1, nset, 2.0 * N * D * D, "TFLOP/s")
timeit("gemm_ln s_out + resid -> y, ln", lambda i: ops.gemm_ln(xb[i], Wf, bfv, ln1=ln, alpha=0.1, beta=1.0, resid=xf[i], out_f32=of[i], out1_a=ob[i]),
       nset, 2.0 * N * D * D, "TFLOP/s")
def _pair2(i):
    ops.gemm(xb[i], Wf, bfv, out_f32=of[i], resid=xf[i], alpha=0.1, beta=1.0)
    ops.rowop(of[i], N, D, MDM_BF16, ln1=ln, out1_a=ob[i])
timeit("  unfused: gemm s_out + rowop ln", _pair2, nset, 2.0 * N * D * D, "TFLOP/s")
timeit("gemm_ln skip gelu + resid -> ln f32, ln", lambda i: ops.gemm_ln(xb[i], Wf, bfv, ln1=ln, act=ACT_GELU, alpha=1.0, beta=0.1, resid=xf[i],
       out1_f32=of[i], ln2=ln2, out2_a=ob[i]), nset, 2.0 * N * D * D, "TFLOP/s")
def _pair3(i):
    ops.gemm(xb[i], Wf, bfv, act=ACT_GELU, out_f32=of2[i], resid=xf[i], alpha=1.0, beta=0.1)
    ops.rowop(of2[i], N, D, MDM_BF16, ln1=ln, out1_f32=of[i], ln2=ln2, out2_a=ob[i])
timeit("  unfused: gemm skip + rowop ln,ln", _pair3, nset, 2.0 * N * D * D, "TFLOP/s")
timeit("gemm_ln sd_o + resid -> y, ln(pre)", lambda i: ops.gemm_ln(xb[i], Wf, bfv, ln1=ln, alpha=1.0, beta=1.0, resid=xf[i], out_f32=of[i],
       ln_pre_resid=True, out1_a=ob[i]), nset, 2.0 * N * D * D, "TFLOP/s")
def _pair4(i):
    ops.gemm(xb[i], Wf, bfv, out_f32=of[i], out_a=ob2[i], a_pre_resid=True, resid=xf[i], alpha=1.0, beta=1.0)
    ops.rowop(ob2[i], N, D, MDM_BF16, ln1=ln, out1_a=ob[i])
timeit("  unfused: gemm sd_o + rowop ln", _pair4, nset, 2.0 * N * D * D, "TFLOP/s")
timeit("gemm_ln f3 (K 2048) + resid -> y, copy, ln f32, ln", lambda i: ops.gemm_ln(xb4[i % 2], W2k, bfv, ln1=ln, alpha=1.0, beta=1.0, resid=xf[i],
       out_f32=of[i], out_a=ob2[i], out1_f32=of2[i], ln2=ln2, out2_a=ob[i]), nset, 2.0 * N * D * 4 * D, "TFLOP/s")
def _pair5(i):
    ops.gemm(xb4[i % 2], W2k, bfv, out_f32=of[i], resid=xf[i], alpha=1.0, beta=1.0)
    ops.rowop(of[i], N, D, MDM_BF16, ln1=ln, out1_f32=of2[i], ln2=ln2, out2_a=ob[i], out0_a=ob2[i])
timeit("  unfused: gemm f3 + rowop ln,ln,copy", _pair5, nset, 2.0 * N * D * 4 * D, "TFLOP/s")

# ---- MoE routing
E, NB = 8, 2
G = NB * E
lnw = torch.rand(NB, D, device=dev) + 0.5
lnb = torch.randn(NB, D, device=dev)
gw = torch.randn(G, D, device=dev) * 0.05
gb = torch.zeros(G, device=dev)
idx = torch.empty(N, NB, 2, dtype=torch.int32, device=dev)
vals = torch.empty(N, NB, 2, device=dev)
stats = torch.empty(N, 2, device=dev)
nblk = (N + 127) // 128
hist = torch.empty(nblk, 2, G, dtype=torch.int32, device=dev)
imp = torch.empty(nblk, G, device=dev)
timeit("moe_gate", lambda i: ops.moe_gate(xf[i], N, D, NB, E, lnw, lnb, gw, gb, idx, vals, stats, hist, imp), nset, N * D * 4, "GB/s")
def _gg(i):
    ops.gemm_gate(xb[i], Wf, bfv, resid=xf[i], out_f32=of[i], NB=NB, E=E, ln_w=lnw, ln_b=lnb, gate_w=gw, gate_b=gb, idx=idx, vals=vals,
                  stats=stats, blk_hist=hist, blk_imp=imp)
timeit("gemm_gate ca_out + resid -> y, gate (both branches)", _gg, nset, 2.0 * N * D * D, "TFLOP/s")
def _gu(i):
    ops.gemm(xb[i], Wf, bfv, out_f32=of[i], resid=xf[i], alpha=1.0, beta=1.0)
    ops.moe_gate(of[i], N, D, NB, E, lnw, lnb, gw, gb, idx, vals, stats, hist, imp)
timeit("  unfused: gemm ca_out + moe_gate", _gu, nset, 2.0 * N * D * D, "TFLOP/s")

# scan / permute / combine on the routing the gate above produced (Fd = 1024 per expert)
if not ONLY or "moe" in ONLY:
    Fd, NBK = 1024, 2 * NB
    ops.moe_gate(xf[0], N, D, NB, E, lnw, lnb, gw, gb, idx, vals, stats, hist, imp)
    cap = NBK * N + G * 128
    base = torch.empty(nblk, G, dtype=torch.int32, device=dev)
    seg = torch.empty(G + 1, dtype=torch.int32, device=dev)
    t_up = torch.empty(cap // 128, 4, dtype=torch.int32, device=dev)
    t_dn = torch.empty(cap // 128, 4, dtype=torch.int32, device=dev)
    ntile = torch.empty(1, dtype=torch.int32, device=dev)
    perm = torch.empty(N, NBK, dtype=torch.int32, device=dev)
    rscale = torch.empty(cap, device=dev)
    usage, importance = torch.zeros(G, device=dev), torch.zeros(G, device=dev)
    xps = [torch.empty(cap, D, device=dev, dtype=bf) for _ in range(2)]
    yps = [torch.randn(cap, D, device=dev).to(bf) for _ in range(2)]
    filmm = torch.randn(NSEQ, 2 * D, device=dev) * 0.1
    timeit("moe_scan", lambda i: ops.moe_scan(hist, imp, idx, N, NB, E, Fd, D, base, seg, t_up, t_dn, ntile, usage, importance),
           nset, N * 4, "GB/s")
    timeit("moe_permute (LN both branches, 4 rows out)", lambda i: ops.moe_permute(xf[i], N, D, NB, E, lnw, lnb, idx, vals, stats, base,
           seg, xps[i % 2], perm, rscale), nset, N * D * (4 + 4 * 2), "GB/s")
    timeit("moe_combine_film (4 rows in, LN, FiLM, SiLU)", lambda i: ops.moe_combine_film(yps[i % 2], perm, N, D, NBK, ln[0], ln[1],
           filmm, T, ob[i]), nset, N * D * (4 * 2 + 2), "GB/s")

# ---- FastAttention core
hd = D // H
qkv = [(torch.randn(N, 3 * D, device=dev)).to(bf) for _ in range(nset)]
P = torch.randn(hd, hd, device=dev) * hd ** -0.5
nw, nb_ = torch.rand(hd, device=dev) + 0.5, torch.randn(hd, device=dev) * 0.1
length = torch.randint(40, T + 1, (NSEQ,), device=dev, dtype=torch.int64)
Pt = P.t().contiguous().to(bf)
order = torch.argsort(length, descending=True).to(torch.int32)
timeit("fastattn tcgen05 (Pt, longest first)", lambda i: ops.fastattn(qkv[i], P, nw, nb_, length, 0, NSEQ, H, T, hd, ob[i], seq_order=order, Pt=Pt), nset, N * D * 8, "GB/s")
timeit("fastattn tcgen05 (Pt)", lambda i: ops.fastattn(qkv[i], P, nw, nb_, length, 0, NSEQ, H, T, hd, ob[i], Pt=Pt), nset, N * D * 8, "GB/s")
timeit("fastattn mma.sync", lambda i: ops.fastattn(qkv[i], P, nw, nb_, length, 0, NSEQ, H, T, hd, ob[i], seq_order=order), nset, N * D * 8, "GB/s")

# head size 64 (the as-shipped tools/train.py shape: D 512, 8 heads): two heads per CTA on the tcgen05 kernel vs mma.sync
if not ONLY or "fastattn" in ONLY:
    H8, hd8 = 8, 64
    P8 = torch.randn(hd8, hd8, device=dev) * hd8 ** -0.5
    Pt8 = ops.pack_fastattn_pt(P8)
    nw8, nb8 = torch.rand(hd8, device=dev) + 0.5, torch.randn(hd8, device=dev) * 0.1
    timeit("fastattn hd 64 tcgen05 (2 heads / CTA)", lambda i: ops.fastattn(qkv[i], P8, nw8, nb8, length, 0, NSEQ, H8, T, hd8, ob[i], seq_order=order, Pt=Pt8), nset, N * D * 8, "GB/s")
    timeit("fastattn hd 64 mma.sync", lambda i: ops.fastattn(qkv[i], P8, nw8, nb8, length, 0, NSEQ, H8, T, hd8, ob[i], seq_order=order), nset, N * D * 8, "GB/s")
ctx = torch.randn(NSEQ, H, hd, hd, device=dev)
ctxT = torch.empty(NSEQ, H, hd, hd, device=dev, dtype=bf)
ops.transpose_cast_bf16(ctx, ctxT)
timeit("lincross_apply tcgen05", lambda i: ops.lincross_apply(xb[i], ctx, NSEQ, T, H, hd, ob[i], ctxT=ctxT), nset, N * D * 4, "GB/s")
timeit("lincross_apply mma.sync", lambda i: ops.lincross_apply(xb[i], ctx, NSEQ, T, H, hd, ob[i]), nset, N * D * 4, "GB/s")
lnD = (torch.rand(D, device=dev) + 0.5, torch.randn(D, device=dev))
timeit("lincross_apply_style (core + LN, FiLM, SiLU)", lambda i: ops.lincross_apply_style(xb[i], ctxT, NSEQ, T, H, hd, lnD, film, ob[i]), nset, N * D * 4, "GB/s")
def _lcu(i):
    ops.lincross_apply(xb[i], ctx, NSEQ, T, H, hd, ob2[i], ctxT=ctxT)
    ops.rowop(ob2[i], N, D, MDM_BF16, ln2=lnD, film=film, rows_per_seq=T, silu=True, out2_a=ob[i])
timeit("  unfused: lincross_apply + rowop ln,film,silu", _lcu, nset, N * D * 4, "GB/s")
Nt = 20
k2 = torch.randn(NSEQ * Nt, D, device=dev).to(bf)
v2 = torch.randn(NSEQ * Nt, D, device=dev).to(bf)
nt = torch.full((NSEQ,), Nt, dtype=torch.int32, device=dev)
timeit("softmax_cross", lambda i: ops.softmax_cross(xb[i], k2, v2, nt, NSEQ, T, Nt, H, hd, ob[i]), nset, N * D * 4, "GB/s")

# head size 64 (8 heads): the two cross-attention cores (MDM_LC_UMMA=0 / MDM_SC_UMMA=0 select the mma.sync kernels)
if not ONLY or "hd 64" in ONLY:
    H8, hd8 = 8, 64
    ctx8 = torch.randn(NSEQ, H8, hd8, hd8, device=dev)
    ctxT8 = ops.pack_lincross_ctxT(ctx8)
    timeit("lincross_apply hd 64 (packed ctx^T)", lambda i: ops.lincross_apply(xb[i], ctx8, NSEQ, T, H8, hd8, ob[i], ctxT=ctxT8), nset, N * D * 4, "GB/s")
    timeit("lincross_apply_style hd 64", lambda i: ops.lincross_apply_style(xb[i], ctxT8, NSEQ, T, H8, hd8, lnD, film, ob[i]), nset, N * D * 4, "GB/s")
    timeit("softmax_cross hd 64", lambda i: ops.softmax_cross(xb[i], k2, v2, nt, NSEQ, T, Nt, H8, hd8, ob[i]), nset, N * D * 4, "GB/s")

# ---- backward building blocks of a Linear (dX, dW with token slabs, db)
xg = torch.randn(N, D, device=dev).to(bf)
Wg = (torch.randn(D, D, device=dev) / D ** 0.5).to(bf)
dyg = torch.randn(N, D, device=dev).to(bf)
dxg = torch.empty(N, D, device=dev, dtype=bf)
dWg, dbg = torch.empty(D, D, device=dev), torch.empty(D, device=dev)
if not ONLY or "linear_backward" in ONLY:
    for _ in range(2):
        ops.linear_backward(xg, Wg, dyg, dx=dxg, dW=dWg, db=dbg)
    torch.cuda.synchronize()
    s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_.record()
    for _ in range(10):
        ops.linear_backward(xg, Wg, dyg, dx=dxg, dW=dWg, db=dbg)
    e_.record()
    torch.cuda.synchronize()
    us = s_.elapsed_time(e_) / 10 * 1e3
    print("%-44s %8.1f us  %8.1f TFLOP/s (dX + dW + db: 4 N D^2 FLOP; 8 eager launches incl. the three transposes)" % ("linear_backward N x512x512", us, 4.0 * N * D * D / us / 1e6))
