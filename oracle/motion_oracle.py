"""ORACLE — test infrastructure only.  Not part of the product; never imported by the package.

A plain torch-fp32 functional restatement of the reference hot path (ltdoanh2004/MotionDiffusion-MoE,
paths relative to text2motion/): MotionTransformer.forward and the classifier-free-guidance DDPM step.
Every function cites the reference lines it follows.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.

Pinning: the reference ships no tests or golden vectors, so this restatement is pinned
against outputs of the *unmodified reference itself*, imported in the build container by
tests/golden/make_golden.py (fixtures + generating script are committed; tests/test_oracle_golden.py
checks them).  Parity status: pinned by reference-generated fixtures.

Parameters are a flat dict keyed like the reference state_dict plus the
tensors the reference keeps outside its state_dict:
  <blk>.dual_self_attn.{local,global}_attn.fast_attention.projection_matrix   (lazy, H3)
  <style>.emb_proj.{weight,bias}  for every StylizationBlock                   (ephemeral, H1)
  text_proj.{weight,bias}                                                      (ephemeral, H2)
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LN_EPS = 1e-5


# ----------------------------------------------------------------------------------------------
# configuration / deterministic parameters
# ----------------------------------------------------------------------------------------------
class Config(dict):
    """input_feats, num_frames, latent_dim, ff_size, num_layers, num_heads, text_latent_dim,
    moe_num_experts — the ctor arguments of MotionTransformer (models/transformer.py:174-186)."""
    __getattr__ = dict.__getitem__


CONFIGS = {
    "tiny": Config(input_feats=12, num_frames=8, latent_dim=128, ff_size=256, num_layers=1, num_heads=4,
                   text_latent_dim=128, moe_num_experts=4),
    # BASELINE.json configs[0]
    "small": Config(input_feats=263, num_frames=196, latent_dim=256, ff_size=512, num_layers=4, num_heads=4,
                    text_latent_dim=256, moe_num_experts=4),
    # BASELINE.json configs[1] ("default": SURVEY.md §0)
    "default": Config(input_feats=263, num_frames=196, latent_dim=512, ff_size=1024, num_layers=8,
                      num_heads=4, text_latent_dim=256, moe_num_experts=8),
}


def block_prefixes(cfg):
    """Execution order of the 2L decoder layers (models/transformer.py:343-357)."""
    return ["decoder_blocks_low.%d.module" % i for i in range(cfg.num_layers)] + \
           ["decoder_blocks_high.%d.module" % i for i in range(cfg.num_layers)]


def style_prefixes(blk):
    """The four StylizationBlocks of a layer in call order (fast_attention.py:175,257; multi_branch.py:60)."""
    return [blk + ".dual_self_attn.local_attn.style_block", blk + ".dual_self_attn.global_attn.style_block",
            blk + ".cross_attn.base_ca.proj_out", blk + ".ffn.proj_out"]


def param_shapes(cfg):
    """Name -> shape of every tensor of the reference state_dict (text encoder excluded)."""
    D, F_, E, Dt, H = cfg.latent_dim, cfg.ff_size, cfg.moe_num_experts, cfg.text_latent_dim, cfg.num_heads
    Te, hd = 4 * D, D // H
    s = {}

    def lin(name, o, i):
        s[name + ".weight"] = (o, i)
        s[name + ".bias"] = (o,)

    def ln(name, d):
        s[name + ".weight"] = (d,)
        s[name + ".bias"] = (d,)

    def style(name):
        lin(name + ".emb_layers.1", 2 * D, Te)
        ln(name + ".norm", D)
        lin(name + ".out_layers.2", D, D)

    s["sequence_embedding"] = (cfg.num_frames, D)
    lin("learnable_time_embed.mlp.0", 2 * D, D)
    lin("learnable_time_embed.mlp.2", D, 2 * D)
    lin("gated_fusion.proj_time", D, D)
    lin("gated_fusion.proj_text", D, D)
    lin("gated_fusion.post_mlp.0", D, D)
    lin("gated_fusion.post_mlp.2", D, D)
    lin("time_embed.0", Te, D)
    lin("time_embed.2", Te, Te)
    lin("time_proj", D, Te)
    lin("joint_embed", D, cfg.input_feats)
    s["downsample.weight"] = (D, D, 2)
    s["downsample.bias"] = (D,)
    s["upsample.weight"] = (D, D, 2)
    s["upsample.bias"] = (D,)
    lin("out", cfg.input_feats, D)
    for blk in block_prefixes(cfg):
        dsa = blk + ".dual_self_attn"
        ln(dsa + ".pre_norm", D)
        ln(dsa + ".post_norm", D)
        for a in ("local_attn", "global_attn"):
            p = dsa + "." + a
            ln(p + ".pre_norm", D)
            ln(p + ".post_norm", D)
            lin(p + ".query", D, D)
            lin(p + ".key", D, D)
            lin(p + ".value", D, D)
            ln(p + ".fast_attention.norm", hd)
            lin(p + ".proj_out.0", D, D)
            lin(p + ".proj_out.3", D, D)
            style(p + ".style_block")
        lin(dsa + ".skip_proj.0", D, D)
        ca = blk + ".cross_attn"
        s[ca + ".gate"] = (D,)
        s[ca + ".base_ca.adaptive_gate"] = (1,)
        ln(ca + ".base_ca.norm", D)
        ln(ca + ".base_ca.text_norm", Dt)
        lin(ca + ".base_ca.query", D, D)
        lin(ca + ".base_ca.key", D, Dt)
        lin(ca + ".base_ca.value", D, Dt)
        style(ca + ".base_ca.proj_out")
        for b in range(2):
            br = blk + ".ffn.branches.%d" % b
            ln(br + ".layernorm", D)
            s[br + ".moe.expert_usage"] = (E,)
            s[br + ".moe.expert_importance"] = (E,)
            lin(br + ".moe.gate", E, D)
            for e in range(E):
                lin(br + ".moe.experts.%d.0" % e, F_, D)
                lin(br + ".moe.experts.%d.2" % e, D, F_)
        style(blk + ".ffn.proj_out")
        sd = blk + ".sd_cross_attn"
        lin(sd + ".query", D, D)
        lin(sd + ".key", D, Dt)
        lin(sd + ".value", D, Dt)
        lin(sd + ".out", D, D)
        ln(sd + ".ffn.0", D)
        lin(sd + ".ffn.1", 4 * D, D)
        lin(sd + ".ffn.3", D, 4 * D)
    return s


def make_params(cfg, seed=0, device="cpu"):
    """Deterministic non-degenerate parameters (H4: nothing left at its zero init).

    Weights ~ N(0, 1/fan_in) (gate weights x4 so routing is decisive), LayerNorm affine near
    (1, 0), counters zero; projection matrices as FastAttention._create_projection
    (fast_attention.py:19-27) from a private generator; ephemeral Linears as nn.Linear's default
    init from a private generator (tests that need RNG-replay equality with the reference use
    draw_ephemerals instead).
    """
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, shape in param_shapes(cfg).items():
        if name.endswith("expert_usage") or name.endswith("expert_importance"):
            t = torch.zeros(shape)
        elif name == "sequence_embedding":
            t = torch.randn(shape, generator=g)
        elif name.endswith(".gate") or name.endswith("adaptive_gate"):
            t = torch.randn(shape, generator=g) * 0.5
        elif len(shape) == 1:
            is_ln_w = name.endswith(".weight") and any(k in name for k in ("norm", "layernorm", ".ffn.0"))
            t = 1.0 + 0.1 * torch.randn(shape, generator=g) if is_ln_w else 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = int(np.prod(shape[1:]))
            scale = 4.0 if name.endswith("moe.gate.weight") else 1.0
            t = torch.randn(shape, generator=g) * (scale / math.sqrt(fan_in))
        p[name] = t
    hd = cfg.latent_dim // cfg.num_heads
    for blk in block_prefixes(cfg):
        for a in ("local_attn", "global_attn"):
            pm = torch.randn(hd, 256, generator=g)
            q, _ = torch.linalg.qr(pm, mode="reduced")
            p["%s.dual_self_attn.%s.fast_attention.projection_matrix" % (blk, a)] = \
                F.normalize(q, dim=0) * (hd ** -0.25)
    D, Te = cfg.latent_dim, 4 * cfg.latent_dim

    def eph(name, o, i):
        bound = 1.0 / math.sqrt(i)
        p[name + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * bound
        p[name + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * bound

    if cfg.text_latent_dim != D:
        eph("text_proj", D, cfg.text_latent_dim)
    for blk in block_prefixes(cfg):
        for sp in style_prefixes(blk):
            eph(sp + ".emb_proj", Te, D)
    return {k: v.to(device) for k, v in p.items()}


def _draw_one_forward(cfg):
    D, Te = cfg.latent_dim, 4 * cfg.latent_dim
    out = {}
    if cfg.text_latent_dim != D:
        l = torch.nn.Linear(cfg.text_latent_dim, D)
        out["text_proj.weight"], out["text_proj.bias"] = l.weight.detach().clone(), l.bias.detach().clone()
    for blk in block_prefixes(cfg):
        for sp in style_prefixes(blk):
            l = torch.nn.Linear(D, Te)
            out[sp + ".emb_proj.weight"] = l.weight.detach().clone()
            out[sp + ".emb_proj.bias"] = l.bias.detach().clone()
    return out


def draw_ephemerals(cfg, seed):
    """Replay of the reference's per-forward RNG consumption (H1/H2): the nn.Linear layers that
    MotionTransformer.forward (transformer.py:313-315) and StylizationBlock.forward
    (stylization.py:22-24) construct on every call, in call order, under torch.manual_seed(seed)."""
    state = torch.get_rng_state()
    torch.manual_seed(seed)
    try:
        return _draw_one_forward(cfg)
    finally:
        torch.set_rng_state(state)


def replay_cfg_rng(cfg, seed, x_shape):
    """RNG stream of one reference p_sample_with_cfg call under torch.manual_seed(seed) on the CPU:
    ephemerals of the conditional forward, of the unconditional forward, then randn_like(x)
    (models/gaussian_diffusion.py:1065-1094)."""
    state = torch.get_rng_state()
    torch.manual_seed(seed)
    try:
        e_c = _draw_one_forward(cfg)
        e_u = _draw_one_forward(cfg)
        return e_c, e_u, torch.randn(x_shape)
    finally:
        torch.set_rng_state(state)


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------
def _lin(p, name, x):
    return F.linear(x, p[name + ".weight"], p[name + ".bias"])


def _ln(p, name, x):
    w = p[name + ".weight"]
    return F.layer_norm(x, (w.shape[0],), w, p[name + ".bias"], LN_EPS)


def top2_cuda_order(probs):
    """(vals, idx) of torch.topk(probs, 2, dim=1) with the tie order torch 2.11 shows on CUDA
    (measured on B200 by tests/test_ops_gpu.py::test_softmax_topk_bit_exact_vs_torch_cuda): the lowest indices among equal values are selected,
    and two equal selected values are emitted higher index first.  models/switch_moe.py:57."""
    n, e = probs.shape
    ar = torch.arange(e, device=probs.device)
    a = torch.argmax((probs == probs.max(dim=1, keepdim=True).values).to(torch.int8), dim=1)  # lowest idx of max
    masked = probs.masked_fill(ar[None, :] == a[:, None], float("-inf"))
    b = torch.argmax((masked == masked.max(dim=1, keepdim=True).values).to(torch.int8), dim=1)
    va, vb = probs.gather(1, a[:, None])[:, 0], probs.gather(1, b[:, None])[:, 0]
    tie = va == vb
    i0 = torch.where(tie, b, a)
    i1 = torch.where(tie, a, b)
    idx = torch.stack([i0, i1], dim=1)
    return probs.gather(1, idx), idx


def stylization(p, name, h, emb):
    """StylizationBlock.forward, models/stylization.py:20-31 (ephemeral emb_proj pinned)."""
    emb = _lin(p, name + ".emb_proj", emb)                        # :22-24
    emb_out = _lin(p, name + ".emb_layers.1", F.silu(emb)).unsqueeze(1)   # :26
    scale, shift = torch.chunk(emb_out, 2, dim=2)                 # :27
    h = _ln(p, name + ".norm", h) * (1 + scale) + shift           # :29
    return _lin(p, name + ".out_layers.2", F.silu(h))             # :30 (dropout inactive)


def fast_attention(p, name, q, k, v, mask):
    """FastAttention.forward, models/fast_attention.py:29-92.  q,k,v: [B,H,T,hd]; mask [B,T,1]."""
    B, H, T, hd = q.shape
    P = p[name + ".projection_matrix"]
    nw, nb = p[name + ".norm.weight"], p[name + ".norm.bias"]
    norm = lambda t: F.layer_norm(t, (hd,), nw, nb, LN_EPS)      # per (b,t,h) row, :44-51
    q, k, v = norm(q), norm(k), norm(v)
    q, k = F.normalize(q, dim=-1), F.normalize(k, dim=-1)         # :54-55
    q_proj = torch.exp(torch.clamp(q @ P, -15, 15)) * 0.1         # :58-61
    k_proj = torch.exp(torch.clamp(k @ P, -15, 15)) * 0.1         # :63-66
    k_proj = k_proj * mask.squeeze(-1)[:, None, :, None].to(q.dtype)   # :69-74
    kv = torch.einsum("bhtm,bhtn->bhmn", k_proj, v) * 0.1         # :77
    qkv = torch.einsum("bhtm,bhmn->bhtn", q_proj, kv) * 0.1       # :78
    den = (q_proj * k_proj).sum(-1, keepdim=True).clamp(min=1e-6)  # :81-82 (same-t product, H11)
    return norm(qkv / den)                                        # :85-90


def performer_self_attention(p, name, x, emb, mask, H):
    """PerformerSelfAttention.forward, models/fast_attention.py:137-179."""
    B, T, D = x.shape
    hd = D // H
    h = _ln(p, name + ".pre_norm", x)
    split = lambda t: t.reshape(B, T, H, hd).permute(0, 2, 1, 3) * 0.1     # :155-157
    q, k, v = _lin(p, name + ".query", h), _lin(p, name + ".key", h), _lin(p, name + ".value", h)
    for tensor in (q, k, v):                                              # :150-152 (training: gradients clamped to [-1, 1])
        if tensor.requires_grad:
            tensor.register_hook(lambda grad: torch.clamp(grad, -1, 1))
    q, k, v = split(q), split(k), split(v)
    a = fast_attention(p, name + ".fast_attention", q, k, v, mask)
    a = a.permute(0, 2, 1, 3).reshape(B, T, D)                    # :162
    a = _lin(p, name + ".proj_out.3", F.gelu(_lin(p, name + ".proj_out.0", a)))   # :165
    a = _ln(p, name + ".post_norm", a)                            # :169
    a = F.normalize(a, dim=-1) * (D ** 0.5)                       # :172
    return x + 0.1 * stylization(p, name + ".style_block", a, emb)   # :175-178


def dual_self_attention(p, name, x, emb, mask, H):
    """DualSelfAttentionBlock.forward, models/fast_attention.py:208-226."""
    h = _ln(p, name + ".pre_norm", x)
    local = performer_self_attention(p, name + ".local_attn", h, emb, mask, H)
    glob = performer_self_attention(p, name + ".global_attn", local, emb, mask, H)
    skip = F.gelu(_lin(p, name + ".skip_proj.0", x))              # Linear -> Dropout -> GELU, :202-206
    return _ln(p, name + ".post_norm", skip + 0.1 * glob)


def gated_cross_attention(p, name, x, xf, emb, H, nt=None):
    """GatedCrossAttention.forward :269-272 over LinearTemporalCrossAttention.forward :242-258.
    nt: optional per-sequence text lengths (rows beyond nt[b] are padding of a batched call and are
    excluded, which equals running each sequence with its own un-padded xf)."""
    B, T, D = x.shape
    N = xf.shape[1]
    base = name + ".base_ca"
    q = F.softmax(_lin(p, base + ".query", _ln(p, base + ".norm", x)).view(B, T, H, -1), dim=-1)
    kl = _lin(p, base + ".key", _ln(p, base + ".text_norm", xf)).view(B, N, H, -1)
    vl = _lin(p, base + ".value", _ln(p, base + ".text_norm", xf)).view(B, N, H, -1)
    if nt is not None:
        pad = torch.arange(N, device=x.device)[None, :] >= nt[:, None]
        kl = kl.masked_fill(pad[:, :, None, None], float("-inf"))
        vl = vl.masked_fill(pad[:, :, None, None], 0.0)
    k = F.softmax(kl, dim=1)
    att = torch.einsum("bnhd,bnhl->bhdl", k, vl)
    y = torch.einsum("bnhd,bhdl->bnhl", q, att).reshape(B, T, D)
    alpha = torch.sigmoid(p[base + ".adaptive_gate"])
    ca = x + alpha * stylization(p, base + ".proj_out", y, emb)
    return x + torch.sigmoid(p[name + ".gate"]).view(1, 1, -1) * (ca - x)


def switch_moe(p, name, x, E, counters=None, tie_order="cuda", forced_idx=None):
    """SwitchMoELayer.forward, models/switch_moe.py:44-111.  Returns (y, idx[N,2], vals[N,2]).
    forced_idx [N,2] (test hook, not in the reference): use these expert indices instead of the top-k search; the
    weights are the softmax probabilities of those experts ("identical routing" comparisons)."""
    B, T, D = x.shape
    xf = x.reshape(-1, D)
    probs = F.softmax(_lin(p, name + ".gate", xf).float(), dim=1)  # :53-54 (softmax runs in fp32 under autocast too)
    if forced_idx is not None:
        idx = forced_idx.to(device=x.device, dtype=torch.long)
        vals = probs.gather(1, idx)
    elif tie_order == "cuda":
        vals, idx = top2_cuda_order(probs)
    else:
        vals, idx = torch.topk(probs, k=2, dim=1)                 # :57 (host tie order)
    if counters is not None:                                      # :72-92
        counters[name + ".expert_usage"] = counters.get(name + ".expert_usage", 0) + \
            torch.bincount(idx[:, 0], minlength=E).float()
        imp = torch.zeros(E, device=x.device).index_add_(0, idx.reshape(-1), vals.reshape(-1))
        counters[name + ".expert_importance"] = counters.get(name + ".expert_importance", 0) + imp
    y = torch.zeros_like(xf)
    for e in range(E):                                            # :97-109
        for j in range(2):
            sel = idx[:, j] == e
            if sel.any():
                h = F.gelu(_lin(p, "%s.experts.%d.0" % (name, e), xf[sel]))
                y[sel] += vals[sel, j:j + 1] * _lin(p, "%s.experts.%d.2" % (name, e), h)
    return y.view(B, T, D), idx, vals


def moe_multibranch_ffn(p, name, x, emb, E, routing=None, counters=None, tie_order="cuda", forced=None):
    """MoEMultiBranchFFN.forward, models/multi_branch.py:52-61.  forced: optional iterator over per-call [N,2]
    expert indices (test hook)."""
    out = 0
    for b in range(2):
        br = "%s.branches.%d" % (name, b)
        h, idx, vals = switch_moe(p, br + ".moe", _ln(p, br + ".layernorm", x), E, counters, tie_order,
                                  None if forced is None else next(forced))
        if routing is not None:
            routing.append((br + ".moe", idx, vals))
        out = out + h
    out = out / 2
    return x + stylization(p, name + ".proj_out", out, emb)


def sd_cross_attention(p, name, x, xf, H, nt=None):
    """MemoryEfficientCrossAttentionBlock.forward, models/fast_attention.py:301-330 (one chunk)."""
    B, T, D = x.shape
    N = xf.shape[1]
    hd = D // H
    q = _lin(p, name + ".query", x).view(B, T, H, hd).permute(0, 2, 1, 3)
    k = _lin(p, name + ".key", xf).view(B, N, H, hd).permute(0, 2, 1, 3)
    v = _lin(p, name + ".value", xf).view(B, N, H, hd).permute(0, 2, 1, 3)
    s = torch.einsum("bhqd,bhkd->bhqk", q * (hd ** -0.5), k)
    if nt is not None:
        pad = torch.arange(N, device=x.device)[None, :] >= nt[:, None]
        s = s.masked_fill(pad[:, None, None, :], float("-inf"))
    o = torch.einsum("bhqk,bhkd->bhqd", F.softmax(s, dim=-1), v)
    o = _lin(p, name + ".out", o.permute(0, 2, 1, 3).reshape(B, T, D))
    f = _lin(p, name + ".ffn.3", F.gelu(_lin(p, name + ".ffn.1", _ln(p, name + ".ffn.0", o))))
    return x + (o + f)


def decoder_layer(p, blk, x, xf, emb, mask, cfg, nt=None, routing=None, counters=None, tie_order="cuda", forced=None):
    """MoEExtendedDecoderLayer.forward, models/transformer.py:55-64."""
    H, E = cfg.num_heads, cfg.moe_num_experts
    x = dual_self_attention(p, blk + ".dual_self_attn", x, emb, mask, H)
    x = gated_cross_attention(p, blk + ".cross_attn", x, xf, emb, H, nt)
    x = moe_multibranch_ffn(p, blk + ".ffn", x, emb, E, routing, counters, tie_order, forced)
    return sd_cross_attention(p, blk + ".sd_cross_attn", x, xf, H, nt)


def timestep_embedding(t, dim, max_period=10000):
    """LearnableTimeEmbedding._sinusoidal_embedding, models/time.py:15-26."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def src_mask(T, length):
    """generate_src_mask, models/transformer.py:284-289 -> [B,T,1]."""
    return (torch.arange(T, device=length.device)[None, :] < length.view(-1, 1)).float().unsqueeze(-1)


def fused_embedding(p, cfg, timesteps, xf_proj):
    """models/transformer.py:313-321 with models/time.py:28-31 and models/gate.py:16-22."""
    if xf_proj.shape[-1] != cfg.latent_dim:
        xf_proj = _lin(p, "text_proj", xf_proj)
    te = timestep_embedding(timesteps, cfg.latent_dim)
    te = _lin(p, "learnable_time_embed.mlp.2", F.silu(_lin(p, "learnable_time_embed.mlp.0", te)))
    te = _lin(p, "time_embed.2", F.silu(_lin(p, "time_embed.0", te)))
    te = _lin(p, "time_proj", te)
    t = _lin(p, "gated_fusion.proj_time", te)
    x = _lin(p, "gated_fusion.proj_text", xf_proj)
    g = torch.sigmoid(t + x)
    fused = g * t + (1 - g) * x
    return _lin(p, "gated_fusion.post_mlp.2", F.silu(_lin(p, "gated_fusion.post_mlp.0", fused)))


def forward(p, cfg, x, timesteps, length, xf_proj, xf_out, nt=None, routing=None, counters=None,
            tie_order="cuda", force_routing=None, skip_layers=()):
    """MotionTransformer.forward, models/transformer.py:291-361 (eval mode, text embeddings given).
    force_routing (test hook): the `routing` list of an earlier call (or its [N,2] index tensors), replayed.
    skip_layers: indices (execution order, 0..2L-1) of the decoder layers that StochasticDepth skips in train mode
    (models/time.py:41-49: the block returns its input unchanged)."""
    forced = None if force_routing is None else iter([r[1] if isinstance(r, (tuple, list)) else r for r in force_routing])
    B, T, _ = x.shape
    if T % 2:
        raise RuntimeError("odd T: the reference's h_up + h fails (H8), T=%d" % T)
    length = length.view(-1)
    emb = fused_embedding(p, cfg, timesteps, xf_proj)
    h = _lin(p, "joint_embed", x) + p["sequence_embedding"].unsqueeze(0)[:, :T, :]
    mask = src_mask(T, length)
    # cuDNN convolutions default to TF32 on CUDA (torch.backends.cudnn.allow_tf32): the oracle is fp32
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        h_low = F.conv1d(h.permute(0, 2, 1), p["downsample.weight"], p["downsample.bias"], stride=2).permute(0, 2, 1)
    mask_low = src_mask(h_low.shape[1], (length / 2).long())
    blks = block_prefixes(cfg)
    for li, blk in enumerate(blks[:cfg.num_layers]):
        if li in skip_layers:
            continue
        h_low = decoder_layer(p, blk, h_low, xf_out, emb, mask_low, cfg, nt, routing, counters, tie_order, forced)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        h_up = F.conv_transpose1d(h_low.permute(0, 2, 1), p["upsample.weight"], p["upsample.bias"], stride=2)
    hc = h_up.permute(0, 2, 1) + h
    for li, blk in enumerate(blks[cfg.num_layers:]):
        if li + cfg.num_layers in skip_layers:
            continue
        hc = decoder_layer(p, blk, hc, xf_out, emb, mask, cfg, nt, routing, counters, tie_order, forced)
    return _lin(p, "out", hc)


def load_balancing_loss(usage, importance, num_experts, epsilon=1e-8):
    """SwitchMoELayer.get_load_balancing_loss, models/switch_moe.py:113-145."""
    fu = usage / usage.sum().clamp_min(epsilon)
    fi = importance / importance.sum().clamp_min(epsilon)
    return num_experts * (1.0 - (fu * fi).sum())


# ----------------------------------------------------------------------------------------------
# GaussianDiffusion (linear beta, eps-prediction, FIXED_SMALL variance: trainers/ddpm_trainer.py:43-50)
# ----------------------------------------------------------------------------------------------
def diffusion_tables(num_steps=1000):
    """GaussianDiffusion.__init__ tables, models/gaussian_diffusion.py:19-33,396-431 (float64)."""
    scale = 1000 / num_steps
    betas = np.linspace(scale * 0.0001, scale * 0.02, num_steps, dtype=np.float64)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    return {
        "betas": betas,
        "alphas_cumprod": ac,
        "sqrt_alphas_cumprod": np.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": np.sqrt(1.0 - ac),
        "sqrt_recip_alphas_cumprod": np.sqrt(1.0 / ac),
        "sqrt_recipm1_alphas_cumprod": np.sqrt(1.0 / ac - 1),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": np.log(np.append(post_var[1], post_var[1:])),
        "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    }


def _extract(arr, t, shape):
    """_extract_into_tensor, models/gaussian_diffusion.py:329-341."""
    res = torch.from_numpy(arr).to(t.device)[t].float()
    while res.dim() < len(shape):
        res = res[..., None]
    return res.expand(shape)


def q_sample(tab, x0, t, noise):
    """GaussianDiffusion.q_sample, :449-460."""
    return _extract(tab["sqrt_alphas_cumprod"], t, x0.shape) * x0 + \
        _extract(tab["sqrt_one_minus_alphas_cumprod"], t, x0.shape) * noise


def cfg_update(tab, x, t, eps_c, eps_u, noise, cfg_scale=7.5, clip_denoised=False):
    """p_sample_with_cfg after the two model calls, :1065-1098 (p_mean_variance :538-542,554-558;
    q_posterior_mean_variance :462-475).  Returns (sample, guided pred_xstart)."""
    def x0_from_eps(eps):
        x0 = _extract(tab["sqrt_recip_alphas_cumprod"], t, x.shape) * x - \
            _extract(tab["sqrt_recipm1_alphas_cumprod"], t, x.shape) * eps
        return x0.clamp(-1, 1) if clip_denoised else x0
    e_c, e_u = x0_from_eps(eps_c), x0_from_eps(eps_u)
    guided = e_u + cfg_scale * (e_c - e_u)
    mean = _extract(tab["posterior_mean_coef1"], t, x.shape) * guided + \
        _extract(tab["posterior_mean_coef2"], t, x.shape) * x
    logvar = _extract(tab["posterior_log_variance_clipped"], t, x.shape)
    nz = (t != 0).float().view(-1, *([1] * (x.dim() - 1)))
    return mean + nz * torch.exp(0.5 * logvar) * noise, guided


def p_mean_variance_update(tab, x, t, eps, clip_denoised=True, noise=None):
    """p_mean_variance after the model call (EPSILON / FIXED_SMALL: gaussian_diffusion.py:510-521,538-552,
    554-558, 462-475) and, with `noise`, the p_sample update (:606-613).  Returns a dict like the reference."""
    x0 = _extract(tab["sqrt_recip_alphas_cumprod"], t, x.shape) * x - \
        _extract(tab["sqrt_recipm1_alphas_cumprod"], t, x.shape) * eps
    if clip_denoised:
        x0 = x0.clamp(-1, 1)
    mean = _extract(tab["posterior_mean_coef1"], t, x.shape) * x0 + _extract(tab["posterior_mean_coef2"], t, x.shape) * x
    out = {"mean": mean, "variance": _extract(tab["posterior_variance"], t, x.shape),
           "log_variance": _extract(tab["posterior_log_variance_clipped"], t, x.shape), "pred_xstart": x0}
    if noise is not None:
        nz = (t != 0).float().view(-1, *([1] * (x.dim() - 1)))
        out["sample"] = mean + nz * torch.exp(0.5 * out["log_variance"]) * noise
    return out


def ddim_update(tab, x, t, x0, eta, noise, t_prev=None):
    """ddim_sample after p_mean_variance, models/gaussian_diffusion.py:724-742, given pred_xstart `x0`.
    t_prev (extension for strided schedules): alpha_bar_prev = alphas_cumprod[t_prev] (1 where t_prev < 0)
    instead of the reference's alphas_cumprod_prev[t]."""
    eps = (_extract(tab["sqrt_recip_alphas_cumprod"], t, x.shape) * x - x0) / \
        _extract(tab["sqrt_recipm1_alphas_cumprod"], t, x.shape)                          # :567-571
    ac = 1.0 / tab["sqrt_recip_alphas_cumprod"] ** 2 if "alphas_cumprod" not in tab else tab["alphas_cumprod"]
    alpha_bar = _extract(ac, t, x.shape)
    if t_prev is None:
        alpha_bar_prev = _extract(np.append(1.0, ac[:-1]), t, x.shape)
    else:
        alpha_bar_prev = _extract(np.append(ac, 1.0), torch.where(t_prev < 0, torch.full_like(t_prev, len(ac)), t_prev),
                                  x.shape)
    sigma = eta * torch.sqrt((1 - alpha_bar_prev) / (1 - alpha_bar)) * torch.sqrt(1 - alpha_bar / alpha_bar_prev)
    mean_pred = x0 * torch.sqrt(alpha_bar_prev) + torch.sqrt(1 - alpha_bar_prev - sigma ** 2) * eps
    nz = (t != 0).float().view(-1, *([1] * (x.dim() - 1)))
    return mean_pred + nz * sigma * noise


def cfg_step(p, cfg, tab, x, t, length, cond, uncond, noise, cfg_scale=7.5, clip_denoised=False,
             p_uncond=None):
    """One p_sample_with_cfg step: cond / uncond are (xf_proj, xf_out) pairs.  p_uncond: parameters of
    the unconditional forward when its ephemerals differ (reference RNG replay); default: same."""
    eps_c = forward(p, cfg, x, t, length, cond[0], cond[1])
    eps_u = forward(p if p_uncond is None else p_uncond, cfg, x, t, length, uncond[0], uncond[1])
    return cfg_update(tab, x, t, eps_c, eps_u, noise, cfg_scale, clip_denoised)


def stub_text(text, dim, device="cpu"):
    """Deterministic stand-in for EnhancedTextEncoder (models/text_encoder.py needs a hub download;
    SURVEY.md Appendix C): (pooled [B,Dt], tokens [B,Nt,Dt]), Nt = 20 for a prompt, 10 for ""."""
    g = torch.Generator().manual_seed(1234 if text[0] else 4321)
    tok = torch.randn(len(text), 8 + (12 if text[0] else 2), dim, generator=g).to(device)
    return tok.mean(1), tok


# ----------------------------------------------------------------------------------------------
# Post-processing after the sampler (tools/visualization.py:72-91, utils/motion_process.py, utils/quaternion.py)
# ----------------------------------------------------------------------------------------------
def _qinv(q):
    """utils/quaternion.py:16-20."""
    mask = torch.ones_like(q)
    mask[..., 1:] = -mask[..., 1:]
    return q * mask


def _qrot(q, v):
    """utils/quaternion.py:54-73."""
    shape = list(v.shape)
    q = q.contiguous().view(-1, 4)
    v = v.contiguous().view(-1, 3)
    qvec = q[:, 1:]
    uv = torch.cross(qvec, v, dim=1)
    uuv = torch.cross(qvec, uv, dim=1)
    return (v + 2 * (q[:, :1] * uv + uuv)).view(shape)


def recover_from_ric(data, joints_num, mean=None, std=None):
    """`motion * std + mean` (tools/visualization.py:91) then recover_root_rot_pos + recover_from_ric
    (utils/motion_process.py:362-380, 401-417).  data [..., T, F] -> [..., T, joints_num, 3]."""
    if mean is not None:
        data = data * std + mean
    rot_vel = data[..., 0]
    r_rot_ang = torch.zeros_like(rot_vel)
    r_rot_ang[..., 1:] = rot_vel[..., :-1]
    r_rot_ang = torch.cumsum(r_rot_ang, dim=-1)
    r_rot_quat = torch.zeros(data.shape[:-1] + (4,), device=data.device)
    r_rot_quat[..., 0] = torch.cos(r_rot_ang)
    r_rot_quat[..., 2] = torch.sin(r_rot_ang)
    r_pos = torch.zeros(data.shape[:-1] + (3,), device=data.device)
    r_pos[..., 1:, [0, 2]] = data[..., :-1, 1:3]
    r_pos = _qrot(_qinv(r_rot_quat), r_pos)
    r_pos = torch.cumsum(r_pos, dim=-2)
    r_pos[..., 1] = data[..., 3]
    positions = data[..., 4:(joints_num - 1) * 3 + 4]
    positions = positions.reshape(positions.shape[:-1] + (-1, 3))
    positions = _qrot(_qinv(r_rot_quat[..., None, :]).expand(positions.shape[:-1] + (4,)), positions)
    positions[..., 0] += r_pos[..., 0:1]
    positions[..., 2] += r_pos[..., 2:3]
    return torch.cat([r_pos.unsqueeze(-2), positions], dim=-2)
