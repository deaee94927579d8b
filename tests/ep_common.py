"""Shared set-up of the expert-parallel tests: random MoEMultiBranchFFN weights and the single-GPU
(local) kernel sequence the expert-parallel path must reproduce bit for bit."""
import math

import torch

from motiondiffusion_moe_b200 import ops
from motiondiffusion_moe_b200._lib import ACT_GELU


def make_weights(D, Fd, E, dtype, device, seed=3):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    G = 2 * E
    w = dict(
        ln_w=(1 + 0.1 * r(2, D)), ln_b=0.05 * r(2, D),
        gate_w=r(G, D) * 4 / math.sqrt(D), gate_b=0.05 * r(G),
        w1=r(G * Fd, D) / math.sqrt(D), b1=0.05 * r(G * Fd),
        w2=r(G * D, Fd) / math.sqrt(Fd), b2=0.05 * r(G * D),
        s_w=1 + 0.1 * r(D), s_b=0.05 * r(D))
    out = {k: v.to(device).contiguous() for k, v in w.items()}
    out["w1"], out["w2"] = out["w1"].to(dtype), out["w2"].to(dtype)
    return out


def make_tokens(n_seq, T, D, device, seed):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(n_seq * T, D, generator=g) * 2).to(device)
    film = (0.3 * torch.randn(n_seq, 2 * D, generator=g)).to(device)
    return x, film


def local_moe(w, x, film, T, D, Fd, E, dtype):
    """The single-GPU path of MotionTransformer._layer: gate -> scan -> permute -> grouped FFN -> combine."""
    dev = x.device
    N, NB, NBK, G = x.shape[0], 2, 4, 2 * E
    cap, nblk = NBK * N + G * 128, (N + 127) // 128
    i32, f32 = torch.int32, torch.float32
    z = lambda *s, dt=f32: torch.zeros(*s, device=dev, dtype=dt)
    idx, vals, stats = z(N, NB, 2, dt=i32), z(N, NB, 2), z(N, 2)
    hist, imp, base, seg = z(nblk, 2, G, dt=i32), z(nblk, G), z(nblk, G, dt=i32), z(G + 1, dt=i32)
    t_up, t_dn, ntile = z(cap // 128, 4, dt=i32), z(cap // 128, 4, dt=i32), z(1, dt=i32)
    perm, rscale = z(N, NBK, dt=i32), z(cap)
    xp, hp, yp = z(cap, D, dt=dtype), z(cap, Fd, dt=dtype), z(cap, D, dt=dtype)
    usage, importance = z(G), z(G)
    ops.moe_gate(x, N, D, NB, E, w["ln_w"], w["ln_b"], w["gate_w"], w["gate_b"], idx, vals, stats, hist, imp)
    ops.moe_scan(hist, imp, idx, N, NB, E, Fd, D, base, seg, t_up, t_dn, ntile, usage, importance)
    ops.moe_permute(x, N, D, NB, E, w["ln_w"], w["ln_b"], idx, vals, stats, base, seg, xp, perm, rscale)
    kw = dict(num_tiles=cap // 128, num_tiles_dev=ntile, M=cap, a_rows=cap)
    o1 = dict(out_a=hp) if dtype == torch.bfloat16 else dict(out_f32=hp)
    o2 = dict(out_a=yp) if dtype == torch.bfloat16 else dict(out_f32=yp)
    ops.gemm(xp, w["w1"], w["b1"], act=ACT_GELU, N=Fd, tiles=t_up, w_rows=G * Fd, **o1, **kw)
    ops.gemm(hp, w["w2"], w["b2"], N=D, rowscale=rscale, tiles=t_dn, w_rows=G * D, **o2, **kw)
    res = torch.empty(N, D, device=dev, dtype=dtype)
    ops.moe_combine_film(yp, perm, N, D, NBK, w["s_w"], w["s_b"], film, T, res)
    return res, idx, vals, usage, importance
