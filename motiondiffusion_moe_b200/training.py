"""The DDPM training step of the reference on the B200 kernels (BASELINE.json configs[4]; SURVEY.md section 8 rows a5 /
a18 / a19 / e3):

    GaussianDiffusion.training_losses   models/gaussian_diffusion.py:923-992   (q_sample, forward, MSE, MoE balance loss)
    DDPMTrainer.backward_G / update     trainers/ddpm_trainer.py:201-244       (masked loss, backward, clip_grad_norm_(1.0), Adam)
    StochasticDepth (train mode)        models/time.py:35-49                   (per-block skip, CPU RNG draw)
    q / k / v gradient clamp            models/fast_attention.py:150-152

`TrainEngine` owns the optimisation state of one MotionTransformer:
  * all parameters live in ONE flat fp32 buffer (the nn.Parameters are views of it), with the tensors that the kernels
    consume stacked (q|k|v, the experts of both MoE branches, the FiLM MLPs of all StylizationBlocks) laid out contiguously,
    so that "packing" the weights for the tcgen05 GEMMs is a single fp32 -> bf16 cast of the flat buffer, and
    clip_grad_norm_ + Adam are two kernels over flat buffers with no host synchronisation;
  * `forward_backward` runs the forward with the activations kept, then the hand-written backward: token-level GEMMs on
    the forward GEMM kernels (dX with transposed weights, dW as grouped contractions over token slabs), the attention cores
    through the strided batched GEMM + row kernels of csrc/train.cu, the MoE routing backward, the row pipelines' backward
    twin.  Gradients land in the flat gradient buffer (each parameter's .grad is a view of it).
Training parity is defined with dropout = 0 (fused kernels cannot replay torch's Philox dropout stream) and
with the ephemeral Linears pinned (H1); the MoE balance loss carries no gradient in the reference (H9) and is a value only.
"""
import math

import torch

from . import ops, train_ops as T
from ._lib import ACT_GELU, ACT_NONE, ACT_SILU, MDM_BF16, MDM_F32, MdmError
from .transformer import _round_up

f32 = torch.float32


class TrainEngine:
    def __init__(self, model, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=1.0):
        self.m = model
        dev = model._t("sequence_embedding").device
        if dev.type != "cuda":
            raise MdmError("TrainEngine needs the model on a CUDA device (no CPU path)")
        if model._ep_on:
            raise MdmError("training with expert parallelism is not built")
        self.dev = dev
        self.lr, self.betas, self.eps, self.max_grad_norm = lr, betas, eps, max_grad_norm
        self.step_count = 0
        self.adt = model._adt()
        with torch.cuda.device(dev):
            self._flatten()
            self._build_pack()
            self.refresh()
        object.__setattr__(model, "_train_engine", self)      # GaussianDiffusion.training_losses keeps activations through it

    # ------------------------------------------------------------------ flat parameter storage
    def _groups(self):
        m = self.m
        E = m.moe_num_experts
        blks = m.block_prefixes()
        styles = [sp for blk in blks for sp in m.style_prefixes(blk)]
        groups = [[sp + ".emb_layers.1.weight" for sp in styles], [sp + ".emb_layers.1.bias" for sp in styles]]
        for blk in blks:
            for a in ("local_attn", "global_attn"):
                p = "%s.dual_self_attn.%s." % (blk, a)
                groups.append([p + "query.weight", p + "key.weight", p + "value.weight"])
                groups.append([p + "query.bias", p + "key.bias", p + "value.bias"])
            br = [blk + ".ffn.branches.%d" % b for b in range(2)]
            groups.append([b + ".layernorm.weight" for b in br])
            groups.append([b + ".layernorm.bias" for b in br])
            groups.append([b + ".moe.gate.weight" for b in br])
            groups.append([b + ".moe.gate.bias" for b in br])
            for leaf in ("0.weight", "0.bias", "2.weight", "2.bias"):
                groups.append(["%s.moe.experts.%d.%s" % (b, e, leaf) for b in br for e in range(E)])
        return groups

    def layout(self):
        """Pure host logic (no CUDA): the flat layout of the parameters.  Returns (order, offset, numel, block_ranges,
        rest_ranges): stacked groups first (contiguous, 64-element aligned starts), then the remaining tensors; per
        decoder layer the (at most two) contiguous ranges holding all of its parameters except the FiLM MLPs, and the ranges
        of everything else.  block_ranges / rest_ranges are the gradient buckets of data-parallel training: a layer's bucket is
        final as soon as that layer's backward is done, the rest at the end of the backward."""
        m = self.m
        groups = self._groups()
        grouped = {n for g in groups for n in g}
        order = [n for g in groups for n in g] + [n for n in m._param_names if n not in grouped]
        starts = {g[0] for g in groups}
        off, offset = 0, {}
        for n in order:
            if n in starts or n not in grouped:
                off = _round_up(off, 64)
            offset[n] = off
            off += m._t(n).numel()
        numel = _round_up(off, 64)
        block_ranges, covered = [], []
        for blk in m.block_prefixes():
            mine = lambda n: n.startswith(blk + ".") and ".emb_layers.1." not in n     # (the FiLM MLPs are a global group)
            rs = []
            for names in ([n for n in order if n in grouped and mine(n)], [n for n in order if n not in grouped and mine(n)]):
                if names:
                    lo = min(offset[n] for n in names)
                    hi = max(offset[n] + m._t(n).numel() for n in names)
                    assert not any(lo <= offset[n] < hi for n in order if not mine(n)), blk
                    rs.append((lo, hi))
            block_ranges.append(rs)
            covered += rs
        covered.sort()
        rest_ranges, pos = [], 0
        for lo, hi in covered:
            if lo > pos:
                rest_ranges.append((pos, lo))
            pos = hi
        if pos < numel:
            rest_ranges.append((pos, numel))
        return order, offset, numel, block_ranges, rest_ranges

    @torch.no_grad()
    def _flatten(self):
        m = self.m
        order, self.offset, self.numel, self.block_ranges, self.rest_ranges = self.layout()
        dev = self.dev
        self.flat = torch.zeros(self.numel, dtype=f32, device=dev)
        self.grad = torch.zeros(self.numel, dtype=f32, device=dev)
        self.exp_avg = torch.zeros(self.numel, dtype=f32, device=dev)
        self.exp_avg_sq = torch.zeros(self.numel, dtype=f32, device=dev)
        self.flat_op = torch.zeros(self.numel, dtype=torch.bfloat16, device=dev) if self.adt == torch.bfloat16 else self.flat
        for n in order:
            p = m._t(n)
            view = self.flat[self.offset[n]:self.offset[n] + p.numel()].view(p.shape)
            view.copy_(p.data.to(dev))
            p.data = view
            p.grad = self.grad[self.offset[n]:self.offset[n] + p.numel()].view(p.shape)
        self._norm_part = torch.empty(1024, dtype=f32, device=dev)
        self.norm_coef = torch.zeros(2, dtype=f32, device=dev)

    def _view(self, buf, names, shape=None):
        if isinstance(names, str):
            names = [names]
        o0 = self.offset[names[0]]
        n = sum(self.m._t(x).numel() for x in names)
        o = o0
        for x in names:                      # contiguity of the group inside the flat layout
            assert self.offset[x] == o, x
            o += self.m._t(x).numel()
        v = buf[o0:o0 + n]
        return v.view(shape) if shape is not None else v.view(self.m._t(names[0]).shape) if len(names) == 1 else v

    def P(self, names, shape=None):          # fp32 master
        return self._view(self.flat, names, shape)

    def W(self, names, shape=None):          # operand-typed mirror (bf16) / master (fp32 mode)
        return self._view(self.flat_op, names, shape)

    def G(self, names, shape=None):          # gradient
        return self._view(self.grad, names, shape)

    # ------------------------------------------------------------------ packed views (model._packed is replaced by them)
    @torch.no_grad()
    def _build_pack(self):
        m, dev = self.m, self.dev
        D, Fd, E, H, Dt = m.latent_dim, m.ff_size, m.moe_num_experts, m.num_heads, m.text_latent_dim
        wdt = self.adt
        ext = lambda n: m._t(n).detach()
        pk = {}
        Kp = _round_up(m.input_feats, 8)
        pk["Kp"] = Kp
        pk["je_w"], pk["je_b"] = torch.zeros(D, Kp, device=dev, dtype=wdt), self.P("joint_embed.bias")
        pk["seq_emb"] = self.P("sequence_embedding")
        pk["down_w"], pk["down_b"] = torch.empty(D, 2 * D, device=dev, dtype=wdt), self.P("downsample.bias")
        pk["up_w"], pk["up_b"] = torch.empty(2 * D, D, device=dev, dtype=wdt), torch.empty(2 * D, device=dev, dtype=f32)
        pk["out_w"], pk["out_b"] = self.W("out.weight"), self.P("out.bias")
        self.small = ("learnable_time_embed.mlp.0", "learnable_time_embed.mlp.2", "time_embed.0", "time_embed.2", "time_proj",
                      "gated_fusion.proj_time", "gated_fusion.proj_text", "gated_fusion.post_mlp.0", "gated_fusion.post_mlp.2")
        for n in self.small:
            pk[n] = (self.W(n + ".weight"), self.P(n + ".bias"))
        if Dt != D:
            pk["text_proj"] = (ext("text_proj.weight").to(wdt).contiguous(), ext("text_proj.bias").float().contiguous())
        blks = m.block_prefixes()
        styles = [sp for blk in blks for sp in m.style_prefixes(blk)]
        self.styles = styles
        S = len(styles)
        pk["eph_w"] = torch.cat([ext(sp + ".emb_proj.weight") for sp in styles]).to(wdt).contiguous()
        pk["eph_b"] = torch.cat([ext(sp + ".emb_proj.bias") for sp in styles]).float().contiguous()
        pk["emb_w"] = self.W([sp + ".emb_layers.1.weight" for sp in styles], (S * 2 * D, 4 * D))
        pk["emb_b"] = self.P([sp + ".emb_layers.1.bias" for sp in styles], (S * 2 * D,))
        pk["n_style"] = S
        LNp = lambda n: (self.P(n + ".weight"), self.P(n + ".bias"))
        Lin = lambda n: (self.W(n + ".weight"), self.P(n + ".bias"))
        layers = []
        for blk in blks:
            L = {"blk": blk}
            dsa = blk + ".dual_self_attn"
            L["dsa_pre"], L["dsa_post"] = LNp(dsa + ".pre_norm"), LNp(dsa + ".post_norm")
            L["skip"] = Lin(dsa + ".skip_proj.0")
            L["perf"] = []
            for a in ("local_attn", "global_attn"):
                p = dsa + "." + a
                Pm = ext(p + ".fast_attention.projection_matrix").float().contiguous()
                L["perf"].append({
                    "name": p, "pre": LNp(p + ".pre_norm"), "post": LNp(p + ".post_norm"),
                    "qkv_w": self.W([p + ".query.weight", p + ".key.weight", p + ".value.weight"], (3 * D, D)),
                    "qkv_b": self.P([p + ".query.bias", p + ".key.bias", p + ".value.bias"], (3 * D,)),
                    "P": Pm, "Pt": Pm.t().contiguous().to(torch.bfloat16) if wdt == torch.bfloat16 else None,
                    "fa_norm": LNp(p + ".fast_attention.norm"), "p0": Lin(p + ".proj_out.0"), "p3": Lin(p + ".proj_out.3"),
                    "s_norm": LNp(p + ".style_block.norm"), "s_out": Lin(p + ".style_block.out_layers.2")})
            ca = blk + ".cross_attn"
            base = ca + ".base_ca"
            L["ca_norm"], L["ca_tnorm"] = LNp(base + ".norm"), LNp(base + ".text_norm")
            L["ca_q"], L["ca_k"], L["ca_v"] = Lin(base + ".query"), Lin(base + ".key"), Lin(base + ".value")
            L["ca_s_norm"] = LNp(base + ".proj_out.norm")
            L["ca_out"] = (torch.empty(D, D, device=dev, dtype=wdt), torch.empty(D, device=dev, dtype=f32))   # folded: refresh()
            L["ca_cs"] = torch.empty(D, device=dev, dtype=f32)
            br = [blk + ".ffn.branches.%d" % b for b in range(2)]
            L["moe_ln_w"] = self.P([b + ".layernorm.weight" for b in br], (2, D))
            L["moe_ln_b"] = self.P([b + ".layernorm.bias" for b in br], (2, D))
            L["gate_w"] = self.P([b + ".moe.gate.weight" for b in br], (2 * E, D))
            L["gate_b"] = self.P([b + ".moe.gate.bias" for b in br], (2 * E,))
            ex = lambda leaf: ["%s.moe.experts.%d.%s" % (b, e, leaf) for b in br for e in range(E)]
            L["w1"], L["b1"] = self.W(ex("0.weight"), (2 * E * Fd, D)), self.P(ex("0.bias"), (2 * E * Fd,))
            L["w2"], L["b2"] = self.W(ex("2.weight"), (2 * E * D, Fd)), self.P(ex("2.bias"), (2 * E * D,))
            L["ffn_s_norm"] = LNp(blk + ".ffn.proj_out.norm")
            L["ffn_out"] = Lin(blk + ".ffn.proj_out.out_layers.2")
            sd = blk + ".sd_cross_attn"
            for k_, n_ in (("sd_q", "query"), ("sd_k", "key"), ("sd_v", "value"), ("sd_o", "out"), ("sd_f1", "ffn.1"),
                           ("sd_f3", "ffn.3")):
                L[k_] = Lin(sd + "." + n_)
            L["sd_ln"] = LNp(sd + ".ffn.0")
            layers.append(L)
        pk["layers"] = layers
        nl = len(blks)
        pk["usage"] = torch.zeros(nl, 2 * E, device=dev)
        pk["importance"] = torch.zeros(nl, 2 * E, device=dev)
        self.pk = pk
        m._packed = pk

    @torch.no_grad()
    def refresh(self, mirror_done=False):
        """Bring the operand-typed weights in line with the fp32 masters (after an optimizer step or a manual edit):
        one cast of the flat buffer (mirror_done: the Adam kernel already wrote it) + the few tensors whose kernel layout is
        a transformation of the parameter."""
        m, pk = self.m, self.pk
        D = m.latent_dim
        if self.flat_op is not self.flat and not mirror_done:
            T.axpby(self.flat, 1.0, None, 0.0, self.flat_op)
        wdt = self.adt
        pk["je_w"][:, :m.input_feats] = self.P("joint_embed.weight").to(wdt)
        pk["down_w"].copy_(self.P("downsample.weight").permute(0, 2, 1).reshape(D, 2 * D))
        pk["up_w"].copy_(self.P("upsample.weight").permute(2, 1, 0).reshape(2 * D, D))
        pk["up_b"].copy_(torch.cat([self.P("upsample.bias"), self.P("upsample.bias")]))
        for L in pk["layers"]:
            ca = L["blk"] + ".cross_attn"
            cs = torch.sigmoid(self.P(ca + ".gate")) * torch.sigmoid(self.P(ca + ".base_ca.adaptive_gate"))
            L["ca_cs"].copy_(cs)
            L["ca_out"][0].copy_(self.P(ca + ".base_ca.proj_out.out_layers.2.weight") * cs[:, None])
            L["ca_out"][1].copy_(self.P(ca + ".base_ca.proj_out.out_layers.2.bias") * cs)
        self._wt = {}
        m._packed = pk        # (a call of model.repack() / load_state_dict would have dropped it)

    def wt(self, key, W, rows_per_group=None, groups=1):
        """W^T (per group) of a packed weight, computed once per step: the weight operand of the dX GEMMs."""
        t = self._wt.get(key)
        if t is None:
            t = self._wt[key] = T.transpose_groups(W, W.shape[0] // groups if rows_per_group is None else rows_per_group, groups)
        return t

    # ------------------------------------------------------------------ helpers
    def _new(self, *shape, dtype=None):
        return torch.empty(*shape, dtype=self.adt if dtype is None else dtype, device=self.dev)

    def _lin(self, A, wb, out, act=ACT_NONE, **kw):
        """out (operand-typed) = act(A W^T + b)."""
        if self.adt == torch.bfloat16:
            ops.gemm(A, wb[0], wb[1], act=act, out_a=out, **kw)
        else:
            ops.gemm(A, wb[0], wb[1], act=act, out_f32=out, **kw)
        return out

    def _lin_bwd(self, key, x, wb_names, W, dy, **kw):
        """Linear backward with parameter gradients accumulated into the flat gradient buffer."""
        wn, bn = wb_names
        T.linear_bwd(x, self.wt(key, W), dy, dW=self.G(wn) if isinstance(wn, str) else wn,
                     db=self.G(bn) if isinstance(bn, str) else bn, **kw)

    # ------------------------------------------------------------------ forward with saved activations
    def _embeddings_train(self, timesteps, xf_proj, Bn):
        """MotionTransformer._embeddings with the pre-activations kept (transformer.py:313-321, stylization.py:22-27)."""
        m, pk = self.m, self.pk
        D, Dt, Te = m.latent_dim, m.text_latent_dim, m.time_embed_dim
        S = {}
        new = self._new
        e0 = new(Bn, D)
        ops.timestep_embedding(timesteps, Bn, D, e0)
        S["e0"] = e0
        S["e1p"] = self._lin(e0, pk["learnable_time_embed.mlp.0"], new(Bn, 2 * D))
        S["e1"] = T.act_fwd(S["e1p"], ACT_SILU)
        S["e2"] = self._lin(S["e1"], pk["learnable_time_embed.mlp.2"], new(Bn, D))
        S["e3p"] = self._lin(S["e2"], pk["time_embed.0"], new(Bn, Te))
        S["e3"] = T.act_fwd(S["e3p"], ACT_SILU)
        S["e4"] = self._lin(S["e3"], pk["time_embed.2"], new(Bn, Te))
        S["e5"] = self._lin(S["e4"], pk["time_proj"], new(Bn, D))
        xpa = new(Bn, Dt)
        ops.pad_cast(xf_proj, Bn, Dt, xpa)
        S["xpj"] = self._lin(xpa, pk["text_proj"], new(Bn, D)) if Dt != D else xpa
        S["tt"], S["xx"] = new(Bn, D, dtype=f32), new(Bn, D, dtype=f32)
        ops.gemm(S["e5"], *pk["gated_fusion.proj_time"], out_f32=S["tt"])
        ops.gemm(S["xpj"], *pk["gated_fusion.proj_text"], out_f32=S["xx"])
        S["fu"] = new(Bn, D)
        ops.gated_mix(S["tt"], S["xx"], S["fu"])
        S["f1p"] = self._lin(S["fu"], pk["gated_fusion.post_mlp.0"], new(Bn, D))
        S["f1"] = T.act_fwd(S["f1p"], ACT_SILU)
        S["emb"] = self._lin(S["f1"], pk["gated_fusion.post_mlp.2"], new(Bn, D))
        t1, t2, Bpad = m._film_tile_tables(Bn, self.dev)
        ns = pk["n_style"]
        S["g1p"] = torch.zeros(ns * Bpad, Te, dtype=self.adt, device=self.dev)
        self._lin(S["emb"], (pk["eph_w"], pk["eph_b"]), S["g1p"], N=Te, M=ns * Bpad, tiles=t1, num_tiles=t1.shape[0], a_rows=Bn,
                  w_rows=ns * Te)
        S["g1"] = T.act_fwd(S["g1p"], ACT_SILU)
        S["film"] = torch.zeros(ns * Bpad, 2 * D, dtype=f32, device=self.dev)
        ops.gemm(S["g1"], pk["emb_w"], pk["emb_b"], out_f32=S["film"], N=2 * D, M=ns * Bpad, tiles=t2, num_tiles=t2.shape[0],
                 a_rows=ns * Bpad, w_rows=ns * 2 * D)
        S["Bpad"] = Bpad
        return S

    def _performer_fwd(self, Pk, resid, hh, film, out, Bn, Tn, length, shift):
        m = self.m
        D, H = m.latent_dim, m.num_heads
        N = Bn * Tn
        new = self._new
        sv = {"hh": hh}
        sv["qkv"] = self._lin(hh, (Pk["qkv_w"], Pk["qkv_b"]), new(N, 3 * D))
        sv["att"] = new(N, D)
        ops.fastattn(sv["qkv"], Pk["P"], Pk["fa_norm"][0], Pk["fa_norm"][1], length, shift, Bn, H, Tn, D // H, sv["att"], Pt=Pk["Pt"])
        sv["gp"] = self._lin(sv["att"], Pk["p0"], new(N, D))
        sv["g"] = T.act_fwd(sv["gp"], ACT_GELU)
        sv["u"] = self._lin(sv["g"], Pk["p3"], new(N, D))
        sv["s"] = new(N, D)
        ops.rowop(sv["u"], N, D, ops._dt(sv["s"]), ln1=Pk["post"], l2norm=True, ln2=Pk["s_norm"], film=film, rows_per_seq=Tn,
                  silu=True, **({"out2_a": sv["s"]} if self.adt == torch.bfloat16 else {"out2_f32": sv["s"]}))
        ops.gemm(sv["s"], Pk["s_out"][0], Pk["s_out"][1], out_f32=out, resid=resid, alpha=0.1, beta=1.0)
        return sv

    def _layer_fwd(self, li, x, ctx, film_all, Bpad, Bn, Tn, length, shift):
        """MoEExtendedDecoderLayer.forward (transformer.py:55-64), activations kept.  x [N, D] fp32 -> new tensor."""
        m, pk = self.m, self.pk
        L = pk["layers"][li]
        D, Fd, E, H = m.latent_dim, m.ff_size, m.moe_num_experts, m.num_heads
        adt = self.adt
        adti = MDM_BF16 if adt == torch.bfloat16 else MDM_F32
        bf = adt == torch.bfloat16
        N = Bn * Tn
        new = self._new
        film = [film_all[(li * 4 + j) * Bpad:] for j in range(4)]
        a_out = lambda t: {"out1_a": t} if bf else {"out1_f32": t}
        a_out2 = lambda t: {"out2_a": t} if bf else {"out2_f32": t}
        sv = {"x": x, "Bn": Bn, "Tn": Tn, "shift": shift, "film": film}
        h, a0, xa = new(N, D, dtype=f32), new(N, D), (new(N, D) if bf else x)
        ops.rowop(x, N, D, adti, ln1=L["dsa_pre"], out1_f32=h, ln2=L["perf"][0]["pre"], **a_out2(a0), **({"out0_a": xa} if bf else {}))
        sv["h"], sv["xa"] = h, xa
        loc = new(N, D, dtype=f32)
        sv["p0"] = self._performer_fwd(L["perf"][0], h, a0, film[0], loc, Bn, Tn, length, shift)
        a0b = new(N, D)
        ops.rowop(loc, N, D, adti, ln1=L["perf"][1]["pre"], **a_out(a0b))
        sv["loc"] = loc
        glb = new(N, D, dtype=f32)
        sv["p1"] = self._performer_fwd(L["perf"][1], loc, a0b, film[1], glb, Bn, Tn, length, shift)
        sv["skp"] = self._lin(xa, L["skip"], new(N, D))
        skg = T.act_fwd(sv["skp"], ACT_GELU)
        pre = new(N, D, dtype=f32)
        T.axpby(skg, 1.0, glb, 0.1, pre)
        sv["pre"] = pre
        x1, a0c = new(N, D, dtype=f32), new(N, D)
        ops.rowop(pre, N, D, adti, ln1=L["dsa_post"], out1_f32=x1, ln2=L["ca_norm"], **a_out2(a0c))
        sv["x1"], sv["a0c"] = x1, a0c
        # ---- GatedCrossAttention
        sv["q"] = self._lin(a0c, L["ca_q"], new(N, D))
        sv["y"] = new(N, D)
        ops.lincross_apply(sv["q"], ctx.lin_ctx[li], Bn, Tn, H, D // H, sv["y"], ctxT=ctx.lin_ctxT[li])
        sv["s3"] = new(N, D)
        ops.rowop(sv["y"], N, D, adti, ln2=L["ca_s_norm"], film=film[2], rows_per_seq=Tn, silu=True, **a_out2(sv["s3"]))
        x2 = new(N, D, dtype=f32)
        ops.gemm(sv["s3"], L["ca_out"][0], L["ca_out"][1], out_f32=x2, resid=x1, alpha=1.0, beta=1.0)
        sv["x2"] = x2
        # ---- MoEMultiBranchFFN (un-scaled expert outputs z are kept; the gate weights are applied by the combine)
        NB, NBK, G = 2, 4, 2 * E
        cap = NBK * N + G * 128
        rows = _round_up(cap, 128) + 128      # + a zero region: the contraction range of an empty expert's weight gradient
        nblk = (N + 127) // 128
        i32 = torch.int32
        z = lambda *s, dtype=f32: torch.zeros(*s, dtype=dtype, device=self.dev)
        idx, vals, stats = new(N, NB, 2, dtype=i32), new(N, NB, 2, dtype=f32), new(N, 2, dtype=f32)
        hist, imp, base = new(nblk, 2, G, dtype=i32), new(nblk, G, dtype=f32), new(nblk, G, dtype=i32)
        seg, ntile = new(G + 1, dtype=i32), new(1, dtype=i32)
        max_tiles = cap // 128
        t_up, t_dn = new(max_tiles, 4, dtype=i32), new(max_tiles, 4, dtype=i32)
        perm, rscale = new(N, NBK, dtype=i32), z(rows)
        xp = z(rows, D, dtype=adt)
        ops.moe_gate(x2, N, D, NB, E, L["moe_ln_w"], L["moe_ln_b"], L["gate_w"], L["gate_b"], idx, vals, stats, hist, imp,
                     forced_idx=None if m.force_routing is None else m.force_routing[li])
        ops.moe_scan(hist, imp, idx, N, NB, E, Fd, D, base, seg, t_up, t_dn, ntile, pk["usage"][li], pk["importance"][li])
        ops.moe_permute(x2, N, D, NB, E, L["moe_ln_w"], L["moe_ln_b"], idx, vals, stats, base, seg, xp, perm, rscale)
        if m.record_routing:
            m.last_routing.append((idx.clone(), vals.clone()))
        kw = dict(num_tiles=max_tiles, num_tiles_dev=ntile, M=cap, a_rows=rows)
        # Only what a later contraction may read has to be defined: every row of a tile is written by the GEMMs (padding rows
        # included: xp is zero there), rows beyond the tiles are never read, and the weight-gradient contraction of an empty
        # expert reads the zero region [cap, rows) of xp / hp / dz / d_pre.  (Zero-filling hpre, hp and zz as well was 1 GB of
        # memset per layer: 4.5 ms of the step.)
        hpre = new(rows, Fd)
        self._lin(xp, (L["w1"], L["b1"]), hpre, N=Fd, w_rows=G * Fd, tiles=t_up, **kw)
        hp = new(rows, Fd)
        hp[cap:].zero_()
        T._chk(T._lib.load().mdm_act_fwd(hpre.data_ptr(), ops._dt(hpre), cap * Fd, ACT_GELU, hp.data_ptr(), ops._stream()), "mdm_act_fwd")
        zz = new(rows, D)
        self._lin(hp, (L["w2"], L["b2"]), zz, N=D, w_rows=G * D, tiles=t_dn, **kw)
        mm = new(N, D)
        T.moe_combine_sum(zz, rscale, perm, N, D, NBK, mm)
        s4 = new(N, D)
        ops.rowop(mm, N, D, adti, ln2=L["ffn_s_norm"], film=film[3], rows_per_seq=Tn, silu=True, **a_out2(s4))
        sv.update(idx=idx, stats=stats, perm=perm, rscale=rscale, xp=xp, hpre=hpre, hp=hp, zz=zz, mm=mm, s4=s4, seg=seg, ntile=ntile,
                  t_up=t_up, t_dn=t_dn, cap=cap, max_tiles=max_tiles)
        x3 = new(N, D, dtype=f32)
        xa3 = new(N, D) if bf else x3
        ops.gemm(s4, L["ffn_out"][0], L["ffn_out"][1], out_f32=x3, resid=x2, alpha=1.0, beta=1.0, **({"out_a": xa3} if bf else {}))
        sv["xa3"] = xa3
        # ---- MemoryEfficientCrossAttentionBlock
        sv["q2"] = self._lin(xa3, L["sd_q"], new(N, D))
        sv["oat"] = new(N, D)
        ops.softmax_cross(sv["q2"], ctx.k2[li], ctx.v2[li], ctx.nt, Bn, Tn, ctx.nt_max, H, D // H, sv["oat"])
        sv["o"] = self._lin(sv["oat"], L["sd_o"], new(N, D))
        rr = new(N, D, dtype=f32)
        T.axpby(x3, 1.0, sv["o"], 1.0, rr)
        sv["lno"] = new(N, D)
        ops.rowop(sv["o"], N, D, adti, ln1=L["sd_ln"], **a_out(sv["lno"]))
        sv["fp"] = self._lin(sv["lno"], L["sd_f1"], new(N, 4 * D))
        sv["f1"] = T.act_fwd(sv["fp"], ACT_GELU)
        out = new(N, D, dtype=f32)
        ops.gemm(sv["f1"], L["sd_f3"][0], L["sd_f3"][1], out_f32=out, resid=rr, alpha=1.0, beta=1.0)
        return out, sv

    def _prepare_text_train(self, xf_proj, xf_out, nt=None):
        """MotionTransformer.prepare_text with the text-side intermediates kept (their Linears / LayerNorm are trained)."""
        m, pk = self.m, self.pk
        ctx = m.prepare_text(xf_proj, xf_out, nt)
        adt, D, Dt = self.adt, m.latent_dim, m.text_latent_dim
        B, Nt, _ = xf_out.shape
        rows = B * Nt
        xf = xf_out.float().contiguous().view(rows, Dt)
        bf = adt == torch.bfloat16
        tx = {"xf": xf, "rows": rows, "Nt": Nt, "B": B}
        xa = self._new(rows, Dt)
        ops.rowop(xf, rows, Dt, ops._dt(xa), **({"out0_a": xa} if bf else {"out1_f32": xa}))      # plain cast (fp32: copy)
        tx["xa"] = xa
        tx["layers"] = []
        for L in pk["layers"]:
            xn = self._new(rows, Dt)
            ops.rowop(xf, rows, Dt, ops._dt(xn), ln1=L["ca_tnorm"], **({"out1_a": xn} if bf else {"out1_f32": xn}))
            k = self._lin(xn, L["ca_k"], self._new(rows, D))
            v = self._lin(xn, L["ca_v"], self._new(rows, D))
            tx["layers"].append({"xn": xn, "k": k, "v": v})
        return ctx, tx

    def forward_train(self, x, timesteps, length, xf_proj, xf_out, nt=None, sd_skip=None):
        """Training-mode forward (activations kept).  sd_skip: optional list of 2L booleans, True = this decoder layer is
        skipped by StochasticDepth (models/time.py:41-49); None = draw torch.rand(1) per layer with survival probability
        linspace(1, 0.8, L)[i] on the CPU generator in call order, as the reference does, when the model is in train() mode."""
        m, pk = self.m, self.pk
        if m.dropout != 0.0 and m.training:
            raise NotImplementedError("the fused training step is defined for dropout = 0.0 (SURVEY.md H12): torch's Philox "
                                      "dropout stream cannot be replayed inside fused kernels")
        Bn, Tn, Fin = x.shape
        if Tn % 2 or Tn > m.num_frames:
            raise RuntimeError("T=%d: must be even and <= num_frames=%d" % (Tn, m.num_frames))
        D = m.latent_dim
        adt = self.adt
        bf = adt == torch.bfloat16
        adti = MDM_BF16 if bf else MDM_F32
        N = Bn * Tn
        nl = m.num_layers
        if sd_skip is None:
            surv = torch.linspace(1.0, 0.8, steps=nl).tolist() if nl > 1 else [1.0]
            sd_skip = []
            for i in range(2 * nl):
                p = surv[i % nl]
                sd_skip.append(bool(m.training and p != 1.0 and not (torch.rand(1).item() < p)))
        timesteps = timesteps.to(torch.int64).contiguous()
        length = length.reshape(-1).to(torch.int64).contiguous()
        m.last_routing = []
        ctx, tx = self._prepare_text_train(xf_proj, xf_out, nt)
        E = self._embeddings_train(timesteps, ctx.xf_proj, Bn)
        film, Bpad = E["film"], E["Bpad"]
        S = {"E": E, "tx": tx, "ctx": ctx, "Bn": Bn, "Tn": Tn, "length": length, "sd_skip": sd_skip, "layers": [None] * (2 * nl)}
        xin = self._new(N, pk["Kp"])
        ops.pad_cast(x.float().contiguous().view(N, Fin), N, Fin, xin)
        S["xin"] = xin
        h = self._new(N, D, dtype=f32)
        ha = self._new(N, D) if bf else h
        ops.gemm(xin, pk["je_w"], pk["je_b"], out_f32=h, out_a=(ha if bf else None), resid=pk["seq_emb"], resid_mod=Tn, alpha=1.0,
                 beta=1.0)
        S["ha"] = ha
        Nl = N // 2
        cur = self._new(Nl, D, dtype=f32)
        ops.gemm(ha.view(Nl, 2 * D), pk["down_w"], pk["down_b"], out_f32=cur)
        for li in range(nl):
            if not sd_skip[li]:
                cur, S["layers"][li] = self._layer_fwd(li, cur, ctx, film, Bpad, Bn, Tn // 2, length, 1)
        if bf:
            hla = self._new(Nl, D)
            ops.rowop(cur, Nl, D, adti, out0_a=hla)
        else:
            hla = cur
        S["hla"] = hla
        hc = self._new(N, D, dtype=f32)
        ops.gemm(hla, pk["up_w"], pk["up_b"], out_f32=hc.view(Nl, 2 * D), resid=h.view(Nl, 2 * D), alpha=1.0, beta=1.0)
        cur = hc
        for li in range(nl, 2 * nl):
            if not sd_skip[li]:
                cur, S["layers"][li] = self._layer_fwd(li, cur, ctx, film, Bpad, Bn, Tn, length, 0)
        if bf:
            hca = self._new(N, D)
            ops.rowop(cur, N, D, adti, out0_a=hca)
        else:
            hca = cur
        S["hca"] = hca
        out = torch.empty(Bn, Tn, Fin, dtype=f32, device=self.dev)
        ops.gemm(hca, pk["out_w"], pk["out_b"], out_f32=out.view(N, Fin))
        return out, S

    # ------------------------------------------------------------------ backward
    def _performer_bwd(self, Pk, sv, film, d_out, d_resid_into, Bn, Tn, length, shift, g_film):
        """Backward of _performer_fwd.  d_out [N, D] fp32 = gradient of the block output; returns d_hh (operand-typed
        gradient of the block's normalised input); the residual path (d_out itself) is the caller's."""
        m = self.m
        D, H = m.latent_dim, m.num_heads
        N = Bn * Tn
        name = Pk["name"]
        ds_ = self._new(N, D)
        T.axpby(d_out, 0.1, None, 0.0, ds_)                                   # out = resid + 0.1 * Lin(s)
        d_s = self._new(N, D)
        self._lin_bwd(name + ".s_out", sv["s"], (name + ".style_block.out_layers.2.weight", name + ".style_block.out_layers.2.bias"),
                      Pk["s_out"][0], ds_, dx_a=d_s)
        d_u = T.rowop_bwd(sv["u"], N, D, d_s, ln1=Pk["post"], l2norm=True, ln2=Pk["s_norm"], film=film, rows_per_seq=Tn, silu=True,
                          g_ln1=(self.G(name + ".post_norm.weight"), self.G(name + ".post_norm.bias")),
                          g_ln2=(self.G(name + ".style_block.norm.weight"), self.G(name + ".style_block.norm.bias")), g_film=g_film)
        d_g = self._new(N, D)
        self._lin_bwd(name + ".p3", sv["g"], (name + ".proj_out.3.weight", name + ".proj_out.3.bias"), Pk["p3"][0], d_u, dx_a=d_g)
        d_gp = T.act_bwd(sv["gp"], d_g, ACT_GELU)
        d_att = self._new(N, D)
        self._lin_bwd(name + ".p0", sv["att"], (name + ".proj_out.0.weight", name + ".proj_out.0.bias"), Pk["p0"][0], d_gp, dx_a=d_att)
        d_qkv = T.fastattn_bwd(sv["qkv"], Pk["P"], Pk["fa_norm"][0], Pk["fa_norm"][1], length, shift, Bn, H, Tn, D // H, d_att,
                               (self.G(name + ".fast_attention.norm.weight"), self.G(name + ".fast_attention.norm.bias")))
        d_hh = self._new(N, D)
        qn = [name + ".query.weight", name + ".key.weight", name + ".value.weight"]
        bn = [name + ".query.bias", name + ".key.bias", name + ".value.bias"]
        T.linear_bwd(sv["hh"], self.wt(name + ".qkv", Pk["qkv_w"]), d_qkv, dx_a=d_hh, dW=self.G(qn, (3 * D, D)), db=self.G(bn, (3 * D,)))
        return d_hh

    def _layer_bwd(self, li, sv, dx, ctx, tx, g_film_all, Bpad):
        """Backward of _layer_fwd.  dx [N, D] fp32 (gradient of the layer output) -> gradient of the layer input (fp32)."""
        m, pk = self.m, self.pk
        L = pk["layers"][li]
        blk = L["blk"]
        D, Fd, E, H = m.latent_dim, m.ff_size, m.moe_num_experts, m.num_heads
        Bn, Tn, shift = sv["Bn"], sv["Tn"], sv["shift"]
        N = Bn * Tn
        hd = D // H
        new = self._new
        G = self.G
        film = sv["film"]
        gfilm = [g_film_all[(li * 4 + j) * Bpad:(li * 4 + j) * Bpad + Bn] for j in range(4)]
        length = self._length
        sd = blk + ".sd_cross_attn"
        adt = self.adt
        # ---- MemoryEfficientCrossAttentionBlock: out = x3 + o + Lin3(gelu(Lin1(LN(o)))),  o = Lin_o(attn)
        d_out = new(N, D)
        T.axpby(dx, 1.0, None, 0.0, d_out)                               # operand-typed copy of the incoming gradient
        d_f1 = new(N, 4 * D)
        self._lin_bwd(sd + ".f3", sv["f1"], (sd + ".ffn.3.weight", sd + ".ffn.3.bias"), L["sd_f3"][0], d_out, dx_a=d_f1)
        d_fp = T.act_bwd(sv["fp"], d_f1, ACT_GELU)
        d_lno = new(N, D)
        self._lin_bwd(sd + ".f1", sv["lno"], (sd + ".ffn.1.weight", sd + ".ffn.1.bias"), L["sd_f1"][0], d_fp, dx_a=d_lno)
        d_o = T.rowop_bwd(sv["o"], N, D, d_lno, ln1=L["sd_ln"], g_ln1=(G(sd + ".ffn.0.weight"), G(sd + ".ffn.0.bias")))
        T.axpby(d_o, 1.0, dx, 1.0, d_o)                                   # + the direct path (rr = x3 + o)
        d_oat = new(N, D)
        self._lin_bwd(sd + ".o", sv["oat"], (sd + ".out.weight", sd + ".out.bias"), L["sd_o"][0], d_o, dx_a=d_oat)
        tl = tx["layers"][li]
        d_q2, d_k2, d_v2 = T.softmax_cross_bwd(sv["q2"], ctx.k2[li], ctx.v2[li], ctx.nt, Bn, Tn, ctx.nt_max, H, hd, d_oat)
        T.linear_bwd(tx["xa"], None, d_k2, dW=G(sd + ".key.weight"), db=G(sd + ".key.bias"))
        T.linear_bwd(tx["xa"], None, d_v2, dW=G(sd + ".value.weight"), db=G(sd + ".value.bias"))
        dx3 = dx                                                            # gradient of x3: dx (direct) + d_xa3 W_q
        self._lin_bwd(sd + ".q", sv["xa3"], (sd + ".query.weight", sd + ".query.bias"), L["sd_q"][0], d_q2, dx_f32=dx3, dx_resid=dx3)
        # ---- MoEMultiBranchFFN: x3 = x2 + Lin_out(silu(film(LN(m)))),  m = sum_j rs_j z_j
        d3a = new(N, D)
        T.axpby(dx3, 1.0, None, 0.0, d3a)
        d_s4 = new(N, D)
        fo = blk + ".ffn.proj_out"
        self._lin_bwd(fo + ".out", sv["s4"], (fo + ".out_layers.2.weight", fo + ".out_layers.2.bias"), L["ffn_out"][0], d3a, dx_a=d_s4)
        d_m = T.rowop_bwd(sv["mm"], N, D, d_s4, ln2=L["ffn_s_norm"], film=film[3], rows_per_seq=Tn, silu=True,
                          g_ln2=(G(fo + ".norm.weight"), G(fo + ".norm.bias")), g_film=gfilm[3])
        cap, rows = sv["cap"], sv["xp"].shape[0]
        dz = torch.zeros(rows, D, dtype=adt, device=self.dev)
        drs = torch.zeros(rows, dtype=f32, device=self.dev)
        T.moe_combine_bwd(sv["zz"], sv["rscale"], sv["perm"], N, D, 4, d_m, dz, drs)
        br = [blk + ".ffn.branches.%d" % b for b in range(2)]
        ex = lambda leaf: ["%s.moe.experts.%d.%s" % (b, e, leaf) for b in br for e in range(E)]
        d_xp = T.expert_ffn_bwd(sv["xp"], sv["hpre"], sv["hp"], self.wt(blk + ".w1t", L["w1"], Fd, 2 * E),
                                self.wt(blk + ".w2t", L["w2"], D, 2 * E), dz, sv["seg"], sv["idx"], N, 2, E, sv["t_up"], sv["t_dn"],
                                sv["ntile"], sv["max_tiles"], cap, Fd, D, G(ex("0.weight"), (2 * E * Fd, D)), G(ex("0.bias"), (2 * E * Fd,)),
                                G(ex("2.weight"), (2 * E * D, Fd)), G(ex("2.bias"), (2 * E * D,)))
        dlog = new(N, 2 * E, dtype=f32)
        T.moe_gate_bwd_logits(sv["x2"], sv["stats"], L["moe_ln_w"], L["moe_ln_b"], L["gate_w"], L["gate_b"], sv["idx"], sv["perm"], drs,
                              N, D, 2, E, dlog)
        dx2 = dx3                                                           # x3 = x2 + ...: the residual path, accumulated below
        gate_gw = G([b + ".moe.gate.weight" for b in br], (2 * E, D))
        gate_gb = G([b + ".moe.gate.bias" for b in br], (2 * E,))
        T.colsum_into(dlog, N, 2 * E, gate_gb)
        for b in range(2):
            dh = new(N, D)
            T.moe_unpermute_bwd(d_xp, sv["perm"], dlog, L["gate_w"], N, D, 2, E, b, dh)
            hb = new(N, D)                                                  # LN_b(x2): the gate's input (for d gate.weight)
            lnb = (L["moe_ln_w"][b], L["moe_ln_b"][b])
            ops.rowop(sv["x2"], N, D, ops._dt(hb), ln1=lnb, **({"out1_a": hb} if adt == torch.bfloat16 else {"out1_f32": hb}))
            dl_b = new(N, E)
            T.axpby(dlog.view(N, 2, E)[:, b].contiguous(), 1.0, None, 0.0, dl_b)
            T.linear_bwd(hb, None, dl_b, dW=gate_gw[b * E:(b + 1) * E])
            T.rowop_bwd(sv["x2"], N, D, dh, ln1=lnb, din=dx2, accumulate=True,
                        g_ln1=(G(br[b] + ".layernorm.weight"), G(br[b] + ".layernorm.bias")))
        # ---- GatedCrossAttention: x2 = x1 + cs * (W s3 + b)  (cs folded into the packed weight)
        ca = blk + ".cross_attn"
        base = ca + ".base_ca"
        d2a = new(N, D)
        T.axpby(dx2, 1.0, None, 0.0, d2a)
        d_s3 = new(N, D)
        gWf, gbf = torch.zeros(D, D, dtype=f32, device=self.dev), torch.zeros(D, dtype=f32, device=self.dev)
        T.linear_bwd(sv["s3"], self.wt(ca + ".out", L["ca_out"][0]), d2a, dx_a=d_s3, dW=gWf, db=gbf)
        cs = L["ca_cs"]
        part = torch.empty(64, D, dtype=f32, device=self.dev)
        T._chk(T._lib.load().mdm_colsum_prod(dx2.data_ptr(), sv["x2"].data_ptr(), sv["x1"].data_ptr(), N, D, 64, part.data_ptr(),
                                             ops._stream()), "mdm_colsum_prod")
        dcs_f = torch.zeros(D, dtype=f32, device=self.dev)
        T.sum_partials(part, 64, D, dcs_f)                                   # sum_t dx2 * (x2 - x1) = cs * d cs
        # parameter-space glue on [D]-sized vectors: un-fold cs = sigmoid(gate) * sigmoid(adaptive_gate)
        dcs = dcs_f / cs
        sg, sa = torch.sigmoid(self.P(ca + ".gate")), torch.sigmoid(self.P(base + ".adaptive_gate"))
        G(ca + ".gate").add_(dcs * sa * sg * (1 - sg))
        G(base + ".adaptive_gate").add_((dcs * sg).sum() * sa * (1 - sa))
        G(base + ".proj_out.out_layers.2.weight").add_(gWf * cs[:, None])
        G(base + ".proj_out.out_layers.2.bias").add_(gbf * cs)
        d_y = T.rowop_bwd(sv["y"], N, D, d_s3, ln2=L["ca_s_norm"], film=film[2], rows_per_seq=Tn, silu=True,
                          g_ln2=(G(base + ".proj_out.norm.weight"), G(base + ".proj_out.norm.bias")), g_film=gfilm[2])
        d_q, d_ctx = T.lincross_apply_bwd(sv["q"], ctx.lin_ctx[li], Bn, Tn, H, hd, d_y)
        d_k, d_v = T.lincross_ctx_bwd(tl["k"], tl["v"], ctx.nt, tx["B"], tx["Nt"], H, hd, d_ctx)
        Dt = m.text_latent_dim
        d_xn = new(tx["rows"], Dt)
        T.linear_bwd(tl["xn"], self.wt(base + ".key", L["ca_k"][0]), d_k, dx_a=d_xn, dW=G(base + ".key.weight"), db=G(base + ".key.bias"))
        d_xn2 = new(tx["rows"], Dt)
        T.linear_bwd(tl["xn"], self.wt(base + ".value", L["ca_v"][0]), d_v, dx_a=d_xn2, dW=G(base + ".value.weight"), db=G(base + ".value.bias"))
        T.axpby(d_xn, 1.0, d_xn2, 1.0, d_xn)
        T.rowop_bwd(tx["xf"], tx["rows"], Dt, d_xn, ln1=L["ca_tnorm"], din=torch.empty(tx["rows"], Dt, dtype=f32, device=self.dev),
                    g_ln1=(G(base + ".text_norm.weight"), G(base + ".text_norm.bias")))
        d_a0c = new(N, D)
        self._lin_bwd(base + ".q", sv["a0c"], (base + ".query.weight", base + ".query.bias"), L["ca_q"][0], d_q, dx_a=d_a0c)
        # x1 = LN_post(pre) (fp32, gradient dx2 via the residual path), a0c = LN_ca(x1)
        dsa = blk + ".dual_self_attn"
        d_pre = T.rowop_bwd(sv["pre"], N, D, d_a0c, ln1=L["dsa_post"], ln2=L["ca_norm"], dmid=dx2,
                            din=new(N, D, dtype=f32), g_ln1=(G(dsa + ".post_norm.weight"), G(dsa + ".post_norm.bias")),
                            g_ln2=(G(base + ".norm.weight"), G(base + ".norm.bias")))
        # ---- DualSelfAttentionBlock: pre = gelu(Lin_skip(xa)) + 0.1 * glb
        d_prea = new(N, D)
        T.axpby(d_pre, 1.0, None, 0.0, d_prea)
        d_skp = T.act_bwd(sv["skp"], d_prea, ACT_GELU)
        d_glb = new(N, D, dtype=f32)
        T.axpby(d_pre, 0.1, None, 0.0, d_glb)
        # global performer: glb = loc + 0.1 * style(...)
        d_hh1 = self._performer_bwd(L["perf"][1], sv["p1"], film[1], d_glb, None, Bn, Tn, length, shift, gfilm[1])
        d_loc = d_glb                                                       # residual path of the global block
        p1 = L["perf"][1]["name"]
        T.rowop_bwd(sv["loc"], N, D, d_hh1, ln1=L["perf"][1]["pre"], din=d_loc, accumulate=True,
                    g_ln1=(G(p1 + ".pre_norm.weight"), G(p1 + ".pre_norm.bias")))
        d_hh0 = self._performer_bwd(L["perf"][0], sv["p0"], film[0], d_loc, None, Bn, Tn, length, shift, gfilm[0])
        # h = LN_pre(x) (fp32; gradient d_loc through the residual path of the local block), a0 = LN_p0pre(h), xa = x
        p0 = L["perf"][0]["name"]
        d_x = T.rowop_bwd(sv["x"], N, D, d_hh0, ln1=L["dsa_pre"], ln2=L["perf"][0]["pre"], dmid=d_loc, din=new(N, D, dtype=f32),
                          g_ln1=(G(dsa + ".pre_norm.weight"), G(dsa + ".pre_norm.bias")),
                          g_ln2=(G(p0 + ".pre_norm.weight"), G(p0 + ".pre_norm.bias")))
        self._lin_bwd(dsa + ".skip", sv["xa"], (dsa + ".skip_proj.0.weight", dsa + ".skip_proj.0.bias"), L["skip"][0], d_skp,
                      dx_f32=d_x, dx_resid=d_x)
        return d_x

    def _embeddings_bwd(self, E, g_film, Bn):
        """Backward of _embeddings_train from the FiLM gradients g_film [S * Bpad, 2D] fp32."""
        m, pk = self.m, self.pk
        D, Te = m.latent_dim, m.time_embed_dim
        Bpad = E["Bpad"]
        S = pk["n_style"]
        adt = self.adt
        G = self.G
        new = self._new
        d_film = new(S * Bpad, 2 * D)
        T.axpby(g_film, 1.0, None, 0.0, d_film)
        d_emb = torch.zeros(Bn, D, dtype=f32, device=self.dev)
        emb_wt = self.wt("emb_w", pk["emb_w"], 2 * D, S)                     # per style [Te, 2D]
        eph_wt = self.wt("eph_w", pk["eph_w"], Te, S)                        # per style [D, Te]
        for s, sp in enumerate(self.styles):
            r0 = s * Bpad
            df = d_film[r0:r0 + Bn]
            d_g1 = new(Bn, Te)
            T.linear_bwd(E["g1"][r0:r0 + Bn], emb_wt[s * Te:(s + 1) * Te], df, dx_a=d_g1, dW=G(sp + ".emb_layers.1.weight"),
                         db=G(sp + ".emb_layers.1.bias"))
            d_g1p = T.act_bwd(E["g1p"][r0:r0 + Bn], d_g1, ACT_SILU)
            T.linear_bwd(E["emb"], eph_wt[s * D:(s + 1) * D], d_g1p, dx_f32=d_emb, dx_resid=d_emb)     # ephemeral: no weight gradient
        d_emba = new(Bn, D)
        T.axpby(d_emb, 1.0, None, 0.0, d_emba)
        lb = lambda key, x, n, dy, **kw: T.linear_bwd(x, self.wt(key, pk[n][0]), dy, dW=G(n + ".weight"), db=G(n + ".bias"), **kw)
        d_f1 = new(Bn, D)
        lb("pm2", E["f1"], "gated_fusion.post_mlp.2", d_emba, dx_a=d_f1)
        d_f1p = T.act_bwd(E["f1p"], d_f1, ACT_SILU)
        d_fu = new(Bn, D, dtype=f32)
        lb("pm0", E["fu"], "gated_fusion.post_mlp.0", d_f1p, dx_f32=d_fu)
        d_tt, d_xx = T.gated_mix_bwd(E["tt"], E["xx"], d_fu)
        d_tta, d_xxa = new(Bn, D), new(Bn, D)
        T.axpby(d_tt, 1.0, None, 0.0, d_tta)
        T.axpby(d_xx, 1.0, None, 0.0, d_xxa)
        T.linear_bwd(E["xpj"], None, d_xxa, dW=G("gated_fusion.proj_text.weight"), db=G("gated_fusion.proj_text.bias"))
        d_e5 = new(Bn, D)
        lb("pt", E["e5"], "gated_fusion.proj_time", d_tta, dx_a=d_e5)
        d_e4 = new(Bn, Te)
        lb("tp", E["e4"], "time_proj", d_e5, dx_a=d_e4)
        d_e3 = new(Bn, Te)
        lb("te2", E["e3"], "time_embed.2", d_e4, dx_a=d_e3)
        d_e3p = T.act_bwd(E["e3p"], d_e3, ACT_SILU)
        d_e2 = new(Bn, D)
        lb("te0", E["e2"], "time_embed.0", d_e3p, dx_a=d_e2)
        d_e1 = new(Bn, 2 * D)
        lb("lt2", E["e1"], "learnable_time_embed.mlp.2", d_e2, dx_a=d_e1)
        d_e1p = T.act_bwd(E["e1p"], d_e1, ACT_SILU)
        T.linear_bwd(E["e0"], None, d_e1p, dW=G("learnable_time_embed.mlp.0.weight"), db=G("learnable_time_embed.mlp.0.bias"))

    def backward(self, S, d_out, grad_ready=None):
        """Backward of forward_train: d_out [B, T, feats] fp32 = dLoss/d prediction.  Gradients are ACCUMULATED into the flat
        gradient buffer (zero_grad() first).  grad_ready (optional): called with a list of (lo, hi) ranges of the flat
        gradient buffer as soon as they are final (one decoder layer at a time, in backward order; the rest at the end):
        the hook data-parallel training uses to all-reduce buckets while the backward is still running."""
        m, pk = self.m, self.pk
        D = m.latent_dim
        Bn, Tn = S["Bn"], S["Tn"]
        N, Nl = Bn * Tn, Bn * Tn // 2
        nl = m.num_layers
        Fin = m.input_feats
        G = self.G
        new = self._new
        self._length = S["length"]
        E = S["E"]
        Bpad = E["Bpad"]
        g_film = torch.zeros(pk["n_style"] * Bpad, 2 * D, dtype=f32, device=self.dev)
        dy = new(N, _round_up(Fin, 8))
        if dy.shape[1] != Fin:
            dy.zero_()
        ops.pad_cast(d_out.contiguous().view(N, Fin), N, Fin, dy)
        dyv = dy[:, :Fin]
        # out = Lin(hca): out.weight [Fin, D]
        d_hc = new(N, D, dtype=f32)
        out_wt = self._wt.get("out")
        if out_wt is None:                                                  # W^T padded to a TMA-friendly row pitch
            Kp = dy.shape[1]
            out_wt = self._wt["out"] = torch.zeros(D, Kp, dtype=self.adt, device=self.dev)
            out_wt[:, :Fin] = pk["out_w"].t()
        T.linear_bwd(S["hca"], out_wt, dy, dx_f32=d_hc)
        gw = torch.zeros(dy.shape[1], D, dtype=f32, device=self.dev)
        T.linear_bwd(S["hca"], None, dy, dW=gw)
        G("out.weight").add_(gw[:Fin])
        T.colsum_into(dyv, N, Fin, G("out.bias"), ld=dy.stride(0))
        cur = d_hc
        for li in reversed(range(nl, 2 * nl)):
            if S["layers"][li] is not None:
                cur = self._layer_bwd(li, S["layers"][li], cur, S["ctx"], S["tx"], g_film, Bpad)
            if grad_ready is not None:
                grad_ready(self.block_ranges[li])
        # hc = up(hla) + h   (ConvTranspose1d as a GEMM over row pairs)
        d_h = cur                                                           # gradient of h through the skip connection
        d_up = new(Nl, 2 * D)
        T.axpby(cur.view(Nl, 2 * D), 1.0, None, 0.0, d_up)
        d_low = new(Nl, D, dtype=f32)
        gup = torch.zeros(2 * D, D, dtype=f32, device=self.dev)
        gub = torch.zeros(2 * D, dtype=f32, device=self.dev)
        T.linear_bwd(S["hla"], self.wt("up", pk["up_w"]), d_up, dx_f32=d_low, dW=gup, db=gub)
        G("upsample.weight").add_(gup.view(2, D, D).permute(2, 1, 0))       # up_w = weight.permute(2, 1, 0).reshape(2D, D)
        G("upsample.bias").add_(gub[:D] + gub[D:])
        cur = d_low
        for li in reversed(range(nl)):
            if S["layers"][li] is not None:
                cur = self._layer_bwd(li, S["layers"][li], cur, S["ctx"], S["tx"], g_film, Bpad)
            if grad_ready is not None:
                grad_ready(self.block_ranges[li])
        # h_low = down(ha pairs): Conv1d(k=2, s=2) as a GEMM over row pairs; h = joint_embed(x) + pos
        d_lowa = new(Nl, D)
        T.axpby(cur, 1.0, None, 0.0, d_lowa)
        gdw = torch.zeros(D, 2 * D, dtype=f32, device=self.dev)
        T.linear_bwd(S["ha"].view(Nl, 2 * D), self.wt("down", pk["down_w"]), d_lowa, dx_f32=d_h.view(Nl, 2 * D),
                     dx_resid=d_h.view(Nl, 2 * D), dW=gdw, db=G("downsample.bias"))
        G("downsample.weight").add_(gdw.view(D, 2, D).permute(0, 2, 1))     # down_w = weight.permute(0, 2, 1).reshape(D, 2D)
        d_ha = new(N, D)
        T.axpby(d_h, 1.0, None, 0.0, d_ha)
        gje = torch.zeros(D, pk["Kp"], dtype=f32, device=self.dev)
        T.linear_bwd(S["xin"], None, d_ha, dW=gje, db=G("joint_embed.bias"))
        G("joint_embed.weight").add_(gje[:, :Fin])
        T.colsum_into(d_h.view(Bn, Tn * D), Bn, Tn * D, G("sequence_embedding").view(-1)[:Tn * D], slabs=1)   # pos. embedding: sum over the batch
        self._embeddings_bwd(E, g_film, Bn)
        if grad_ready is not None:
            grad_ready(self.rest_ranges)

    # ------------------------------------------------------------------ optimizer
    def zero_grad(self):
        self.grad.zero_()

    def optimizer_step(self):
        """clip_grad_norm_(max_norm) + Adam on the flat buffers (ddpm_trainer.py:228-244), then refresh the operand mirror.
        No host synchronisation; the gradient norm is left in self.norm_coef[0]."""
        lib = T._lib.load()
        self.step_count += 1
        with torch.cuda.device(self.dev):
            T._chk(lib.mdm_grad_clip_coef(self.grad.data_ptr(), self.numel, float(self.max_grad_norm or 0.0), self._norm_part.data_ptr(),
                                          1024, self.norm_coef.data_ptr(), ops._stream()), "mdm_grad_clip_coef")
            T._chk(lib.mdm_adam_step(self.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                     self.numel, self.lr, self.betas[0], self.betas[1], self.eps, self.step_count,
                                     self.norm_coef.data_ptr(), self.flat_op.data_ptr() if self.flat_op is not self.flat else None,
                                     ops._stream()), "mdm_adam_step")
            self.refresh(mirror_done=True)

    # ------------------------------------------------------------------ one training-step evaluation
    def loss_and_grads(self, x_start, t, length, xf_proj, xf_out, noise, diffusion, sd_skip=None, nt=None):
        """training_losses (gaussian_diffusion.py:923-992) + the masked loss of backward_G (ddpm_trainer.py:207-217) + backward.
        Returns {"loss_mot_rec", "moe_loss", "pred", "target", "mse"}; gradients are in the flat buffer."""
        m = self.m
        with torch.cuda.device(self.dev):
            x_t = diffusion.q_sample(x_start, t, noise=noise)
            m.reset_all_moe_counters(m)
            pred, S = self.forward_train(x_t, t, length, xf_proj, xf_out, nt=nt, sd_skip=sd_skip)
            B, Tn, F = pred.shape
            cur_len = length.reshape(-1).to(torch.int64).clamp(max=Tn).contiguous()
            ws = (torch.zeros(B, device=self.dev), torch.zeros(1, dtype=torch.int32, device=self.dev), torch.zeros(1, device=self.dev))
            ops.masked_mse(pred, noise.float().contiguous(), cur_len, *ws)
            d_pred = T.masked_mse_grad(pred, noise.float().contiguous(), cur_len)
            self.backward(S, d_pred)
            mse = ((noise - pred) ** 2).mean(dim=(1, 2))
            return {"loss_mot_rec": ws[2][0].clone(), "moe_loss": m.get_moe_loss(m), "pred": pred, "target": noise, "mse": mse}
