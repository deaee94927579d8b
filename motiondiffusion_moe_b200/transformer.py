"""MotionTransformer — B200-native drop-in for the reference's MoE denoiser
(text2motion/models/transformer.py:166-361 in ltdoanh2004/MotionDiffusion-MoE).

Same constructor arguments, same `forward(x, timesteps, length, text, xf_proj, xf_out)`, same
state_dict key names, same helper methods (`encode_text`, `generate_src_mask`,
`reset_all_moe_counters`, `get_moe_loss`, `get_total_moe_loss`).  The computation itself runs in
hand-written sm_100a kernels behind the C-ABI of include/mdm_b200.h; this file only owns parameters,
packs them once into kernel-friendly layouts and sequences the launches on the current CUDA stream.
There is no CPU or eager-PyTorch fallback.

Deliberate, documented differences from the reference:
  * the reference re-creates a randomly initialised nn.Linear inside every StylizationBlock.forward
    and for text_proj.  Here those "ephemeral" layers are pinned tensors owned by
    the module (non-persistent buffers `<style>.emb_proj.*`, `text_proj.*`); `redraw_ephemerals(seed)`
    replays the reference's RNG consumption for seed-for-seed equality.
  * FastAttention.projection_matrix (lazy, unsaved in the reference, H3) is an owned buffer that can
    be copied from a live reference module with `import_reference_extras`.
  * per-sequence text lengths `nt` allow the conditional and unconditional CFG branches (different
    token counts) to run as one batch.
"""
import math
from typing import List, Optional

import os

import torch
import torch.nn as nn

from . import ops
from ._lib import MdmError, MDM_F32, MDM_BF16, ACT_NONE, ACT_GELU, ACT_SILU


class _Node(nn.Module):
    """Anonymous tree node so that parameter paths equal the reference's state_dict keys."""


def _round_up(a, b):
    return (a + b - 1) // b * b


class TextContext:
    """Step-invariant text-side tensors: per layer the [hd x hd] linear
    cross-attention state and the K/V of the softmax cross-attention."""

    def __init__(self):
        self.B = 0
        self.nt = None
        self.nt_max = 0
        self.lin_ctx = []
        self.lin_ctxT = []        # bf16 [B, H, l, d] copies of lin_ctx (B operand of the tcgen05 apply kernel) or None
        self.k2 = []
        self.v2 = []
        self.xf_proj = None
        self.seq_order = None     # optional int32 [B]: sequences by descending length (FastAttention launch order)


class MotionTransformer(nn.Module):
    def __init__(self, input_feats: int, num_frames: int = 60, latent_dim: int = 512, ff_size: int = 1024,
                 num_layers: int = 4, num_heads: int = 4, dropout: float = 0.1, text_latent_dim: int = 256,
                 moe_num_experts: int = 4, model_size: str = "small", chunk_size: int = 256,
                 text_encoder: Optional[nn.Module] = None, precision: str = "bf16", **kwargs):
        super().__init__()
        if model_size == "big":  # transformer.py:188-192
            latent_dim *= 2
            ff_size *= 2
            text_latent_dim *= 2
        if latent_dim % num_heads:
            raise AssertionError("latent_dim must be divisible by num_head")
        self.input_feats = input_feats
        self.num_frames = num_frames
        self.latent_dim = latent_dim
        self.ff_size = ff_size
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.dropout = dropout
        self.text_latent_dim = text_latent_dim
        self.moe_num_experts = moe_num_experts
        self.chunk_size = chunk_size
        self.time_embed_dim = latent_dim * 4
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.precision = precision
        self.record_routing = False
        self.last_routing = []
        # parity hook (tests): per-layer expert indices [N_l, 2 branches, 2] int32 that replace the top-2 search of the
        # gate (set_forced_routing); None = the gate routes, as in production
        self.force_routing = None
        if text_encoder is not None:
            self.text_encoder = text_encoder
        # mdm_gemm_rowop (row pipeline fused into the GEMM's A-operand construction) is correct and tested but slower
        # than rowop + GEMM today (profiles/README.md, step 22): opt-in with MDM_FUSE_ROWOP=1
        self._fuse_rowop = os.environ.get("MDM_FUSE_ROWOP", "0") == "1"
        # Linear + the LayerNorm chain that follows it in one kernel (ops.gemm_ln: the row stays in TMEM; north_star (3)).
        # bf16 mode, latent_dim 512; MDM_FUSE_LN=0 keeps the gemm + rowop pairs (A/B runs, parity tests of both).
        self._fuse_ln = os.environ.get("MDM_FUSE_LN", "1") == "1"
        # ... and the MoE gate as the second pass of the cross-attention output Linear (ops.gemm_gate); MDM_FUSE_GATE=0 = off
        self._fuse_gate = os.environ.get("MDM_FUSE_GATE", "1") == "1"
        # ... and the StylizationBlock after the linear cross-attention core in that core's epilogue; MDM_FUSE_STYLE=0 = off
        self._fuse_style = os.environ.get("MDM_FUSE_STYLE", "1") == "1"
        self._packed = None
        self._ws = {}
        self._film_tiles = {}
        self._ep_group = None      # expert parallelism: off (see enable_expert_parallel)
        self._ep_on = False
        self._ep_inst = {}
        self._build_tree()
        self.reset_parameters()

    # ------------------------------------------------------------------ parameter tree
    def block_prefixes(self):
        L = self.num_layers
        return ["decoder_blocks_low.%d.module" % i for i in range(L)] + \
               ["decoder_blocks_high.%d.module" % i for i in range(L)]

    @staticmethod
    def style_prefixes(blk):
        return [blk + ".dual_self_attn.local_attn.style_block", blk + ".dual_self_attn.global_attn.style_block",
                blk + ".cross_attn.base_ca.proj_out", blk + ".ffn.proj_out"]

    def _shapes(self):
        D, Fd, E, Dt, H = self.latent_dim, self.ff_size, self.moe_num_experts, self.text_latent_dim, self.num_heads
        Te, hd = 4 * D, D // H
        params, buffers, extras = {}, {}, {}

        def lin(n, o, i):
            params[n + ".weight"] = (o, i)
            params[n + ".bias"] = (o,)

        def ln(n, d):
            params[n + ".weight"] = (d,)
            params[n + ".bias"] = (d,)

        def style(n):
            lin(n + ".emb_layers.1", 2 * D, Te)
            ln(n + ".norm", D)
            lin(n + ".out_layers.2", D, D)
            extras[n + ".emb_proj.weight"] = (Te, D)
            extras[n + ".emb_proj.bias"] = (Te,)

        params["sequence_embedding"] = (self.num_frames, D)
        lin("learnable_time_embed.mlp.0", 2 * D, D)
        lin("learnable_time_embed.mlp.2", D, 2 * D)
        for n in ("proj_time", "proj_text", "post_mlp.0", "post_mlp.2"):
            lin("gated_fusion." + n, D, D)
        lin("time_embed.0", Te, D)
        lin("time_embed.2", Te, Te)
        lin("time_proj", D, Te)
        lin("joint_embed", D, self.input_feats)
        params["downsample.weight"] = (D, D, 2)
        params["downsample.bias"] = (D,)
        params["upsample.weight"] = (D, D, 2)
        params["upsample.bias"] = (D,)
        lin("out", self.input_feats, D)
        if Dt != D:
            extras["text_proj.weight"] = (D, Dt)
            extras["text_proj.bias"] = (D,)
        for blk in self.block_prefixes():
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
                extras[p + ".fast_attention.projection_matrix"] = (hd, min(hd, 256))
            lin(dsa + ".skip_proj.0", D, D)
            ca = blk + ".cross_attn"
            params[ca + ".gate"] = (D,)
            params[ca + ".base_ca.adaptive_gate"] = (1,)
            ln(ca + ".base_ca.norm", D)
            ln(ca + ".base_ca.text_norm", Dt)
            lin(ca + ".base_ca.query", D, D)
            lin(ca + ".base_ca.key", D, Dt)
            lin(ca + ".base_ca.value", D, Dt)
            style(ca + ".base_ca.proj_out")
            for b in range(2):
                br = blk + ".ffn.branches.%d" % b
                ln(br + ".layernorm", D)
                buffers[br + ".moe.expert_usage"] = (E,)
                buffers[br + ".moe.expert_importance"] = (E,)
                lin(br + ".moe.gate", E, D)
                for e in range(E):
                    lin(br + ".moe.experts.%d.0" % e, Fd, D)
                    lin(br + ".moe.experts.%d.2" % e, D, Fd)
            style(blk + ".ffn.proj_out")
            sd = blk + ".sd_cross_attn"
            lin(sd + ".query", D, D)
            lin(sd + ".key", D, Dt)
            lin(sd + ".value", D, Dt)
            lin(sd + ".out", D, D)
            ln(sd + ".ffn.0", D)
            lin(sd + ".ffn.1", 4 * D, D)
            lin(sd + ".ffn.3", D, 4 * D)
        return params, buffers, extras

    def _register(self, dotted, tensor, kind):
        node = self
        parts = dotted.split(".")
        for part in parts[:-1]:
            if part not in node._modules:
                node.add_module(part, _Node())
            node = node._modules[part]
        if kind == "param":
            node.register_parameter(parts[-1], nn.Parameter(tensor))
        else:
            node.register_buffer(parts[-1], tensor, persistent=(kind == "buffer"))

    def _build_tree(self):
        params, buffers, extras = self._shapes()
        for n, s in params.items():
            self._register(n, torch.empty(s), "param")
        for n, s in buffers.items():
            self._register(n, torch.zeros(s), "buffer")
        for n, s in extras.items():
            self._register(n, torch.zeros(s), "extra")  # non-persistent: state_dict keys == reference
        self._param_names = list(params)
        self._buffer_names = list(buffers)
        self._extra_names = list(extras)

    def _t(self, dotted):
        node = self
        parts = dotted.split(".")
        for part in parts[:-1]:
            node = node._modules[part]
        last = parts[-1]
        return node._parameters[last] if last in node._parameters else node._buffers[last]

    @torch.no_grad()
    def reset_parameters(self):
        """Initialisation distributions of the reference constructor (nn.Linear / Conv1d defaults,
        LayerNorm (1,0), zero gates: switch_moe.py:28-29; zero_module: stylization.py:17,
        transformer.py:257; xavier_normal_(gain=0.1) inside PerformerSelfAttention:
        fast_attention.py:132-135)."""
        for n in self._param_names:
            t = self._t(n)
            leaf = n.rsplit(".", 1)[-1]
            if n == "sequence_embedding":
                t.normal_()
            elif n.endswith(".gate") or n.endswith("adaptive_gate") or ".moe.gate." in n:
                t.zero_()
            elif ".out_layers.2." in n or n.startswith("out."):
                t.zero_()
            elif t.dim() == 1 and leaf == "weight":
                t.fill_(1.0)   # LayerNorm weight
            elif t.dim() == 1:
                w = self._t(n[:-4] + "weight")
                if w.dim() == 1:
                    t.zero_()  # LayerNorm bias
                else:
                    fan_in = w[0].numel()
                    t.uniform_(-1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in))
            else:
                fan_in = t[0].numel()
                t.uniform_(-1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in))
        for blk in self.block_prefixes():
            for a in ("local_attn", "global_attn"):
                pre = "%s.dual_self_attn.%s." % (blk, a)
                for n in self._param_names:
                    if n.startswith(pre):
                        t = self._t(n)
                        if t.dim() > 1:
                            nn.init.xavier_normal_(t, gain=0.1)
        for n in self._buffer_names:
            self._t(n).zero_()
        hd = self.latent_dim // self.num_heads
        for n in self._extra_names:
            t = self._t(n)
            if n.endswith("projection_matrix"):   # fast_attention.py:19-27
                q, _ = torch.linalg.qr(torch.randn(hd, 256), mode="reduced")
                t.copy_(nn.functional.normalize(q, dim=0) * (hd ** -0.25))
        self.redraw_ephemerals(None)
        self._packed = None

    @torch.no_grad()
    def redraw_ephemerals(self, seed=None):
        """Draw the ephemeral Linears in the reference's per-forward order (transformer.py:313-315,
        stylization.py:22-24).  With a seed this reproduces exactly what the reference draws after
        torch.manual_seed(seed); the global RNG state is restored afterwards."""
        state = torch.get_rng_state() if seed is not None else None
        if seed is not None:
            torch.manual_seed(seed)
        try:
            D, Te = self.latent_dim, self.time_embed_dim
            if self.text_latent_dim != D:
                l = nn.Linear(self.text_latent_dim, D)
                self._t("text_proj.weight").copy_(l.weight)
                self._t("text_proj.bias").copy_(l.bias)
            for blk in self.block_prefixes():
                for sp in self.style_prefixes(blk):
                    l = nn.Linear(D, Te)
                    self._t(sp + ".emb_proj.weight").copy_(l.weight)
                    self._t(sp + ".emb_proj.bias").copy_(l.bias)
        finally:
            if state is not None:
                torch.set_rng_state(state)
        self._packed = None

    @torch.no_grad()
    def load_extras(self, tensors: dict):
        """Set projection matrices / pinned ephemerals from a dict keyed like `_extra_names`."""
        for n in self._extra_names:
            if n in tensors:
                self._t(n).copy_(tensors[n])
        self._packed = None

    @torch.no_grad()
    def import_reference_extras(self, ref_model):
        """Copy every FastAttention.projection_matrix from a live (warmed) reference MotionTransformer."""
        got = {}
        for name, mod in ref_model.named_modules():
            pm = getattr(mod, "projection_matrix", None)
            if name.endswith("fast_attention") and pm is not None:
                got[name + ".projection_matrix"] = pm.detach()
        self.load_extras(got)
        return sorted(got)

    def extras_state(self):
        return {n: self._t(n).detach().clone() for n in self._extra_names}

    # ------------------------------------------------------------------ nn.Module plumbing
    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = {k: v for k, v in state_dict.items() if not k.startswith("text_encoder.")} \
            if "text_encoder" not in self._modules else state_dict
        r = super().load_state_dict(sd, strict=strict, **kw)
        self._packed = None
        return r

    def _apply(self, fn, *a, **kw):
        r = super()._apply(fn, *a, **kw)
        self._packed = None
        self._ws = {}
        self._film_tiles = {}
        return r

    # ------------------------------------------------------------------ expert parallelism
    def enable_expert_parallel(self, group=None):
        """Shard the experts of every MoEMultiBranchFFN over the ranks of `group` (default: the default
        torch.distributed group; one process per GPU on one NVSwitch node): expert e of both branches lives
        on rank e // (E // world).  Tokens stay on the rank that owns their sequence; the dispatch / combine
        kernels move rows over NVLink peer memory (expert_parallel.py, csrc/ep.cu).  Every rank must call
        forward() the same number of times with the same shapes.  The result is bit-identical to the
        single-GPU path.  (The reference has no expert parallelism: models/switch_moe.py:97-109.)"""
        import torch.distributed as dist
        if not dist.is_initialized():
            raise MdmError("enable_expert_parallel needs an initialised torch.distributed process group")
        world = dist.get_world_size(group)
        if self.moe_num_experts % world:
            raise MdmError("moe_num_experts=%d is not divisible by the %d ranks" % (self.moe_num_experts, world))
        self._ep_group, self._ep_on = group, True
        self._ep_inst = {}
        if self._packed is not None:
            for L in self._packed["layers"]:
                L.pop("ep_w", None)

    def _ep_for(self, n_tokens):
        from .expert_parallel import ExpertParallelFFN
        ep = self._ep_inst.get(n_tokens)
        if ep is None:    # collective (IPC handle exchange): first use happens in the same order on every rank
            ep = ExpertParallelFFN.create_distributed(self.latent_dim, self.ff_size, self.moe_num_experts, 2, n_tokens,
                                                      self._adt(), self._t("sequence_embedding").device,
                                                      group=self._ep_group)
            self._ep_inst[n_tokens] = ep
        return ep

    def set_forced_routing(self, routing):
        """Parity hook: make every SwitchMoELayer use the given expert indices instead of its own top-2 search, so that a
        bf16 (or fp32) run can be compared with the fp32 reference "with identical routing" (reference
        routing: models/switch_moe.py:53-57).  `routing`: None (off), or one entry per SwitchMoELayer call in execution
        order (2 per decoder layer: branch 0, branch 1) holding the [N, 2] index tensor (or a tuple whose element [1] is
        it, as oracle.forward(routing=[...]) records), or one [N, 2, 2] tensor per decoder layer.  The gate weights are
        still this model's own softmax probabilities of those experts (mdm_moe_gate_forced)."""
        if routing is None:
            self.force_routing = None
            return
        dev = self._t("sequence_embedding").device
        ent = [r[1] if isinstance(r, (tuple, list)) else r for r in routing]
        nl = 2 * self.num_layers
        if len(ent) == 2 * nl:
            ent = [torch.stack([ent[2 * i], ent[2 * i + 1]], dim=1) for i in range(nl)]
        if len(ent) != nl:
            raise MdmError("forced routing needs %d (per branch) or %d (per layer) entries, got %d" % (2 * nl, nl, len(ent)))
        self.force_routing = [e.to(device=dev, dtype=torch.int32).contiguous() for e in ent]

    def check_health(self):
        """Expert parallelism: raise if a peer missed a barrier or a receive buffer overflowed (reads two device words per
        token count: synchronises).  Called by the sampling loops at their synchronisation points; no-op otherwise."""
        for ep in self._ep_inst.values():
            ep.check_health()

    def repack(self):
        """Call after modifying parameters in place (the packed kernel layouts are cached)."""
        self._packed = None

    # ------------------------------------------------------------------ reference helper API
    def reset_all_moe_counters(self, model=None):      # transformer.py:260-263
        for n in self._buffer_names:
            self._t(n).zero_()
        if self._packed is not None:
            self._packed["usage"].zero_()
            self._packed["importance"].zero_()

    def _pull_counters(self):
        pk = self._packed
        if pk is None:
            return
        E = self.moe_num_experts
        for li, blk in enumerate(self.block_prefixes()):
            for b in range(2):
                br = "%s.ffn.branches.%d.moe." % (blk, b)
                self._t(br + "expert_usage").copy_(pk["usage"][li, b * E:(b + 1) * E])
                self._t(br + "expert_importance").copy_(pk["importance"][li, b * E:(b + 1) * E])

    def state_dict(self, *a, **kw):
        self._pull_counters()
        return super().state_dict(*a, **kw)

    def get_moe_loss(self, model=None):                # transformer.py:272-279, switch_moe.py:113-145
        self._pull_counters()
        E = self.moe_num_experts
        total = 0
        for blk in self.block_prefixes():
            for b in range(2):
                br = "%s.ffn.branches.%d.moe." % (blk, b)
                u, i = self._t(br + "expert_usage"), self._t(br + "expert_importance")
                fu = u / u.sum().clamp_min(1e-8)
                fi = i / i.sum().clamp_min(1e-8)
                total = total + E * (1.0 - (fu * fi).sum())
        return total

    def get_total_moe_loss(self, model=None, moe_coef=0.01):   # transformer.py:265-270
        return moe_coef * self.get_moe_loss(model)

    def encode_text(self, text: List[str], device):
        if "text_encoder" not in self._modules:
            raise MdmError("no text_encoder attached: pass xf_proj/xf_out, or construct with "
                           "text_encoder=<module returning (pooled [B,Dt], tokens [B,Nt,Dt])>. The DeBERTa "
                           "encoder of the reference (models/text_encoder.py) is outside this hot path.")
        return self.text_encoder(text, device)

    def generate_src_mask(self, T: int, length: torch.Tensor) -> torch.Tensor:   # transformer.py:284-289
        length = length.view(-1)
        return (torch.arange(T, device=length.device)[None, :] < length[:, None]).float()

    # ------------------------------------------------------------------ packing
    def _adt(self):
        return torch.bfloat16 if self.precision == "bf16" else torch.float32

    @torch.no_grad()
    def _pack(self):
        dev = self._t("sequence_embedding").device
        if dev.type != "cuda":
            raise MdmError("MotionTransformer must live on a CUDA device (no CPU path): call .cuda() first")
        wdt = self._adt()
        D, Fd, E, H = self.latent_dim, self.ff_size, self.moe_num_experts, self.num_heads
        g = lambda n: self._t(n).detach()
        W = lambda n: g(n + ".weight").to(wdt).contiguous()
        Bf = lambda n: g(n + ".bias").float().contiguous()
        LN = lambda n: (g(n + ".weight").float().contiguous(), g(n + ".bias").float().contiguous())
        pk = {}
        Kp = _round_up(self.input_feats, 8)
        je = torch.zeros(D, Kp, device=dev, dtype=wdt)
        je[:, :self.input_feats] = g("joint_embed.weight").to(wdt)
        pk["Kp"] = Kp
        pk["je_w"], pk["je_b"] = je, Bf("joint_embed")
        pk["seq_emb"] = g("sequence_embedding").float().contiguous()
        pk["down_w"] = g("downsample.weight").permute(0, 2, 1).reshape(D, 2 * D).to(wdt).contiguous()
        pk["down_b"] = Bf("downsample")
        pk["up_w"] = g("upsample.weight").permute(2, 1, 0).reshape(2 * D, D).to(wdt).contiguous()
        pk["up_b"] = torch.cat([g("upsample.bias"), g("upsample.bias")]).float().contiguous()
        pk["out_w"], pk["out_b"] = W("out"), Bf("out")
        for n in ("learnable_time_embed.mlp.0", "learnable_time_embed.mlp.2", "time_embed.0", "time_embed.2",
                  "time_proj", "gated_fusion.proj_time", "gated_fusion.proj_text", "gated_fusion.post_mlp.0",
                  "gated_fusion.post_mlp.2"):
            pk[n] = (W(n), Bf(n))
        if self.text_latent_dim != D:
            pk["text_proj"] = (W("text_proj"), Bf("text_proj"))
        blks = self.block_prefixes()
        styles = [sp for blk in blks for sp in self.style_prefixes(blk)]
        pk["eph_w"] = torch.cat([g(sp + ".emb_proj.weight") for sp in styles]).to(wdt).contiguous()
        pk["eph_b"] = torch.cat([g(sp + ".emb_proj.bias") for sp in styles]).float().contiguous()
        pk["emb_w"] = torch.cat([g(sp + ".emb_layers.1.weight") for sp in styles]).to(wdt).contiguous()
        pk["emb_b"] = torch.cat([g(sp + ".emb_layers.1.bias") for sp in styles]).float().contiguous()
        pk["n_style"] = len(styles)
        layers = []
        for blk in blks:
            L = {}
            dsa = blk + ".dual_self_attn"
            L["dsa_pre"], L["dsa_post"] = LN(dsa + ".pre_norm"), LN(dsa + ".post_norm")
            L["skip"] = (W(dsa + ".skip_proj.0"), Bf(dsa + ".skip_proj.0"))
            L["perf"] = []
            for a in ("local_attn", "global_attn"):
                p = dsa + "." + a
                P_ = {"pre": LN(p + ".pre_norm"), "post": LN(p + ".post_norm"),
                      "qkv_w": torch.cat([g(p + ".query.weight"), g(p + ".key.weight"),
                                          g(p + ".value.weight")]).to(wdt).contiguous(),
                      "qkv_b": torch.cat([g(p + ".query.bias"), g(p + ".key.bias"),
                                          g(p + ".value.bias")]).float().contiguous(),
                      "P": g(p + ".fast_attention.projection_matrix").float().contiguous(),
                      "Pt": (ops.pack_fastattn_pt(g(p + ".fast_attention.projection_matrix"))
                             if wdt == torch.bfloat16 else None),
                      "fa_norm": LN(p + ".fast_attention.norm"),
                      "p0": (W(p + ".proj_out.0"), Bf(p + ".proj_out.0")),
                      "p3": (W(p + ".proj_out.3"), Bf(p + ".proj_out.3")),
                      "s_norm": LN(p + ".style_block.norm"),
                      "s_out": (W(p + ".style_block.out_layers.2"), Bf(p + ".style_block.out_layers.2"))}
                L["perf"].append(P_)
            ca = blk + ".cross_attn"
            base = ca + ".base_ca"
            # x + sigmoid(gate) * ((x + sigmoid(a) * style) - x)  ==  x + cs * style   (fast_attention.py:256-272)
            cs = torch.sigmoid(g(ca + ".gate").float()) * torch.sigmoid(g(base + ".adaptive_gate").float())
            L["ca_norm"], L["ca_tnorm"] = LN(base + ".norm"), LN(base + ".text_norm")
            L["ca_q"] = (W(base + ".query"), Bf(base + ".query"))
            L["ca_k"] = (W(base + ".key"), Bf(base + ".key"))
            L["ca_v"] = (W(base + ".value"), Bf(base + ".value"))
            L["ca_s_norm"] = LN(base + ".proj_out.norm")
            L["ca_out"] = ((g(base + ".proj_out.out_layers.2.weight").float() * cs[:, None]).to(wdt).contiguous(),
                           (g(base + ".proj_out.out_layers.2.bias").float() * cs).contiguous())
            br = [blk + ".ffn.branches.%d" % b for b in range(2)]
            L["moe_ln_w"] = torch.stack([g(b + ".layernorm.weight") for b in br]).float().contiguous()
            L["moe_ln_b"] = torch.stack([g(b + ".layernorm.bias") for b in br]).float().contiguous()
            L["gate_w"] = torch.cat([g(b + ".moe.gate.weight") for b in br]).float().contiguous()
            L["gate_b"] = torch.cat([g(b + ".moe.gate.bias") for b in br]).float().contiguous()
            L["w1"] = torch.cat([g("%s.moe.experts.%d.0.weight" % (b, e)) for b in br for e in range(E)]).to(wdt).contiguous()
            L["b1"] = torch.cat([g("%s.moe.experts.%d.0.bias" % (b, e)) for b in br for e in range(E)]).float().contiguous()
            L["w2"] = torch.cat([g("%s.moe.experts.%d.2.weight" % (b, e)) for b in br for e in range(E)]).to(wdt).contiguous()
            L["b2"] = torch.cat([g("%s.moe.experts.%d.2.bias" % (b, e)) for b in br for e in range(E)]).float().contiguous()
            L["ffn_s_norm"] = LN(blk + ".ffn.proj_out.norm")
            L["ffn_out"] = (W(blk + ".ffn.proj_out.out_layers.2"), Bf(blk + ".ffn.proj_out.out_layers.2"))
            sd = blk + ".sd_cross_attn"
            for k_, n_ in (("sd_q", "query"), ("sd_k", "key"), ("sd_v", "value"), ("sd_o", "out"),
                           ("sd_f1", "ffn.1"), ("sd_f3", "ffn.3")):
                L[k_] = (W(sd + "." + n_), Bf(sd + "." + n_))
            L["sd_ln"] = LN(sd + ".ffn.0")
            layers.append(L)
        pk["layers"] = layers
        nl = len(blks)
        pk["usage"] = torch.zeros(nl, 2 * E, device=dev)
        pk["importance"] = torch.zeros(nl, 2 * E, device=dev)
        for li, blk in enumerate(blks):
            for b in range(2):
                brn = "%s.ffn.branches.%d.moe." % (blk, b)
                pk["usage"][li, b * E:(b + 1) * E] = g(brn + "expert_usage")
                pk["importance"][li, b * E:(b + 1) * E] = g(brn + "expert_importance")
        self._packed = pk
        return pk

    # ------------------------------------------------------------------ workspace
    def _buf(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        t = self._ws.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=self._t("sequence_embedding").device)
            self._ws[key] = t
        return t

    def _film_tile_tables(self, B, dev):
        key = B
        if key in self._film_tiles:
            return self._film_tiles[key]
        S, D = self._packed["n_style"], self.latent_dim
        Bpad = _round_up(B, 128)
        t1, t2 = [], []
        for s in range(S):
            for i in range(Bpad // 128):
                rows = min(128, B - i * 128)
                t1.append([i * 128, s * Bpad + i * 128, s * 4 * D, rows])
                t2.append([s * Bpad + i * 128, s * Bpad + i * 128, s * 2 * D, rows])
        r = (torch.tensor(t1, dtype=torch.int32, device=dev), torch.tensor(t2, dtype=torch.int32, device=dev), Bpad)
        self._film_tiles[key] = r
        return r

    # ------------------------------------------------------------------ text side
    def prepare_text(self, xf_proj, xf_out, nt=None):
        """Project the text tokens for every layer once (they do not depend on x or t):
        fast_attention.py:249-252 (linear cross-attention state) and :306-307 (K/V)."""
        if not xf_out.is_cuda:
            raise MdmError("prepare_text needs CUDA tensors: there is no CPU fallback")
        with torch.cuda.device(xf_out.device):      # the C library launches on the current device
            return self._prepare_text(xf_proj, xf_out, nt)

    @torch.no_grad()
    def _prepare_text(self, xf_proj, xf_out, nt=None):
        pk = self._packed or self._pack()
        adt, D, H, Dt = self._adt(), self.latent_dim, self.num_heads, self.text_latent_dim
        hd = D // H
        B, Nt, _ = xf_out.shape
        if Nt > 96:
            raise MdmError("at most 96 text tokens are supported (reference maximum: 8 + 77)")
        dev = xf_out.device
        xf = xf_out.float().contiguous().view(B * Nt, Dt)
        ctx = TextContext()
        ctx.B, ctx.nt_max = B, Nt
        ctx.nt = (torch.full((B,), Nt, dtype=torch.int32, device=dev) if nt is None
                  else nt.to(device=dev, dtype=torch.int32).contiguous())
        ctx.xf_proj = xf_proj.float().contiguous()
        rows = B * Nt
        xa = torch.empty(rows, Dt, dtype=adt, device=dev)
        ops.rowop(xf, rows, Dt, ops._dt(xa), out0_a=xa)
        xn = torch.empty(rows, Dt, dtype=adt, device=dev)
        k = torch.empty(rows, D, dtype=adt, device=dev)
        v = torch.empty(rows, D, dtype=adt, device=dev)
        for L in pk["layers"]:
            ops.rowop(xf, rows, Dt, ops._dt(xn), ln1=L["ca_tnorm"], out1_a=xn)
            self._lin(xn, L["ca_k"], out_a=k)
            self._lin(xn, L["ca_v"], out_a=v)
            c = torch.empty(B, H, hd, hd, dtype=torch.float32, device=dev)
            ops.lincross_ctx(k, v, ctx.nt, B, Nt, H, hd, c)
            k2 = torch.empty(rows, D, dtype=adt, device=dev)
            v2 = torch.empty(rows, D, dtype=adt, device=dev)
            self._lin(xa, L["sd_k"], out_a=k2)
            self._lin(xa, L["sd_v"], out_a=v2)
            ctx.lin_ctx.append(c)
            ctx.lin_ctxT.append(ops.pack_lincross_ctxT(c) if adt == torch.bfloat16 else None)
            ctx.k2.append(k2)
            ctx.v2.append(v2)
        return ctx

    # ------------------------------------------------------------------ kernels sequencing
    def _lin(self, A, wb, *, out_a=None, out_f32=None, act=ACT_NONE, **kw):
        """Linear in the active precision.  In fp32 mode the 'operand-typed' output is fp32 too."""
        if self.precision == "fp32" and out_a is not None and out_f32 is None and not kw.get("a_pre_resid"):
            out_f32, out_a = out_a, None
        ops.gemm(A, wb[0], wb[1], act=act, out_a=out_a, out_f32=out_f32, **kw)

    def _performer(self, Pk, resid, hh, film, out, Bn, T, length, shift, order=None, next_ln=None, next_out=None):
        """PerformerSelfAttention.forward (fast_attention.py:137-179) after its pre_norm.  next_ln / next_out: the
        LayerNorm that consumes this block's output (the next block's pre-norm) and its bf16 destination; returns True
        when that LayerNorm was computed here (in the epilogue of the output Linear)."""
        adt, D, H = self._adt(), self.latent_dim, self.num_heads
        N = Bn * T
        qkv = self._buf("qkv", (N, 3 * D), adt)
        a1 = self._buf("a1", (N, D), adt)
        a2 = self._buf("a2", (N, D), adt)
        self._lin(hh, (Pk["qkv_w"], Pk["qkv_b"]), out_a=qkv)
        ops.fastattn(qkv, Pk["P"], Pk["fa_norm"][0], Pk["fa_norm"][1], length, shift, Bn, H, T, D // H, a1,
                     seq_order=order, Pt=Pk["Pt"])
        self._lin(a1, Pk["p0"], out_a=a2, act=ACT_GELU)
        fuse = self._fuse_ln and adt == torch.bfloat16
        # projection Linear -> post LN -> L2 norm -> StylizationBlock (LN, FiLM, SiLU) in the Linear's epilogue
        if fuse and ops.gemm_ln(a2, Pk["p3"][0], Pk["p3"][1], ln1=Pk["post"], l2norm=True, ln2=Pk["s_norm"], film=film,
                                rows_per_seq=T, silu=True, out2_a=a1):
            src = a1
        else:
            self._lin(a2, Pk["p3"], out_a=a1)
            # opt-in alternative: the row pipeline builds the A operand of the output Linear in shared memory
            if (self._fuse_rowop and adt == torch.bfloat16 and
                    ops.gemm_rowop(a1, N, D, Pk["s_out"][0], Pk["s_out"][1], ln1=Pk["post"], l2norm=True, ln2=Pk["s_norm"],
                                   film=film, rows_per_seq=T, silu=True, out_f32=out, resid=resid, alpha=0.1, beta=1.0)):
                return False
            ops.rowop(a1, N, D, ops._dt(a2), ln1=Pk["post"], l2norm=True, ln2=Pk["s_norm"], film=film,
                      rows_per_seq=T, silu=True, out2_a=a2)
            src = a2
        # StylizationBlock output Linear + residual (+ the next block's pre-norm in its epilogue)
        if (fuse and next_ln is not None and
                ops.gemm_ln(src, Pk["s_out"][0], Pk["s_out"][1], ln1=next_ln, alpha=0.1, beta=1.0, resid=resid, out_f32=out,
                            out1_a=next_out)):
            return True
        self._lin(src, Pk["s_out"], out_f32=out, resid=resid, alpha=0.1, beta=1.0)
        return False

    def _layer(self, li, x, ctx, film_all, Bpad, Bn, T, length, shift, pre_done=False, nxt=None):
        """MoEExtendedDecoderLayer.forward (transformer.py:55-64); x [Bn*T, D] fp32 is updated in place.
        pre_done: this layer's two pre-norms (h, a0) and the bf16 copy of x (xa) were already produced by the previous
        layer's last Linear; nxt: index of the layer that consumes x next (its pre-norms go into this layer's last
        Linear).  Returns True when the pre-norms of `nxt` were computed here."""
        pk = self._packed
        L = pk["layers"][li]
        adt, D, Fd, E, H = self._adt(), self.latent_dim, self.ff_size, self.moe_num_experts, self.num_heads
        adti = MDM_BF16 if adt == torch.bfloat16 else MDM_F32
        N = Bn * T
        f32 = torch.float32
        film = [film_all[(li * 4 + j) * Bpad:] for j in range(4)]
        h = self._buf("h", (N, D), f32)
        loc = self._buf("loc", (N, D), f32)
        glb = self._buf("glb", (N, D), f32)
        x1 = self._buf("x1", (N, D), f32)
        x2 = self._buf("x2", (N, D), f32)
        x3 = self._buf("x3", (N, D), f32)
        a0 = self._buf("a0", (N, D), adt)
        a1 = self._buf("a1", (N, D), adt)
        a2 = self._buf("a2", (N, D), adt)
        xa = self._buf("xa", (N, D), adt)
        # ---- DualSelfAttentionBlock (fast_attention.py:208-226)
        fuse = self._fuse_ln and adt == torch.bfloat16
        if not pre_done:
            ops.rowop(x, N, D, adti, ln1=L["dsa_pre"], out1_f32=h, ln2=L["perf"][0]["pre"], out2_a=a0, out0_a=xa)
        if not self._performer(L["perf"][0], h, a0, film[0], loc, Bn, T, length, shift, ctx.seq_order,
                               next_ln=L["perf"][1]["pre"], next_out=a0):
            ops.rowop(loc, N, D, adti, ln1=L["perf"][1]["pre"], out1_a=a0)
        self._performer(L["perf"][1], loc, a0, film[1], glb, Bn, T, length, shift, ctx.seq_order)
        # skip Linear + GELU + 0.1 * global branch -> post norm (x1, fp32) -> cross-attention norm (a0)
        if not (fuse and ops.gemm_ln(xa, L["skip"][0], L["skip"][1], ln1=L["dsa_post"], act=ACT_GELU, alpha=1.0, beta=0.1,
                                     resid=glb, out1_f32=x1, ln2=L["ca_norm"], out2_a=a0)):
            pre = h  # h is dead from here on
            self._lin(xa, L["skip"], out_f32=pre, act=ACT_GELU, resid=glb, alpha=1.0, beta=0.1)
            ops.rowop(pre, N, D, adti, ln1=L["dsa_post"], out1_f32=x1, ln2=L["ca_norm"], out2_a=a0)
        # ---- GatedCrossAttention (fast_attention.py:242-272)
        self._lin(a0, L["ca_q"], out_a=a1)
        ca_done = False           # the output Linear of the block may run below, fused with the MoE gate
        # linear cross-attention core with the StylizationBlock (LN over the row, FiLM, SiLU) in its epilogue: the H
        # head-CTAs of a sequence form a cluster (a2 -> a1 directly); otherwise core, then rowop
        if not (fuse and self._fuse_style and ctx.lin_ctxT[li] is not None and
                ops.lincross_apply_style(a1, ctx.lin_ctxT[li], Bn, T, H, D // H, L["ca_s_norm"], film[2], a2)):
            ops.lincross_apply(a1, ctx.lin_ctx[li], Bn, T, H, D // H, a2, ctxT=ctx.lin_ctxT[li])
            if (self._fuse_rowop and adt == torch.bfloat16 and
                    ops.gemm_rowop(a2, N, D, L["ca_out"][0], L["ca_out"][1], ln2=L["ca_s_norm"], film=film[2], rows_per_seq=T,
                                   silu=True, out_f32=x2, resid=x1, alpha=1.0, beta=1.0)):
                ca_done = True
            else:
                ops.rowop(a2, N, D, adti, ln2=L["ca_s_norm"], film=film[2], rows_per_seq=T, silu=True, out2_a=a1)
        else:
            a1, a2 = a2, a1       # the styled rows are in a2: it is the A operand of the output Linear below
        # ---- MoEMultiBranchFFN (multi_branch.py:52-61, switch_moe.py:44-111)
        forced = None
        if self.force_routing is not None:
            forced = self.force_routing[li]
            if self._ep_on:
                raise MdmError("forced routing (parity hook) is not wired into the expert-parallel path")
            if tuple(forced.shape) != (N, 2, 2):
                raise MdmError("forced routing of layer %d has shape %s, expected %s" % (li, tuple(forced.shape), (N, 2, 2)))
        if self._ep_on:
            if not ca_done:
                self._lin(a1, L["ca_out"], out_f32=x2, resid=x1, alpha=1.0, beta=1.0)
            ep = self._ep_for(N)
            if "ep_w" not in L:
                L["ep_w"] = ep.shard_weights(L["moe_ln_w"], L["moe_ln_b"], L["gate_w"], L["gate_b"], L["w1"], L["b1"],
                                             L["w2"], L["b2"])
            ep.use_weights(L["ep_w"])
            ep.forward(x2, L["ffn_s_norm"][0], L["ffn_s_norm"][1], film[3], T, a1, usage=pk["usage"][li],
                       importance=pk["importance"][li])
            if self.record_routing:
                self.last_routing.append((ep.idx.clone(), ep.vals.clone()))
            return self._layer_tail(li, L, x, x2, None, a1, a2, xa, x1, ctx, Bn, T, N, nxt)
        NB, NBK, G = 2, 4, 2 * E
        cap = NBK * N + G * 128
        nblk = (N + 127) // 128
        i32 = torch.int32
        idx = self._buf("moe_idx", (N, NB, 2), i32)
        vals = self._buf("moe_vals", (N, NB, 2), f32)
        stats = self._buf("moe_stats", (N, 2), f32)
        hist = self._buf("moe_hist", (nblk, 2, G), i32)
        imp = self._buf("moe_imp", (nblk, G), f32)
        base = self._buf("moe_base", (nblk, G), i32)
        seg = self._buf("moe_seg", (G + 1,), i32)
        max_tiles = cap // 128
        t_up = self._buf("moe_tup", (max_tiles, 4), i32)
        t_dn = self._buf("moe_tdn", (max_tiles, 4), i32)
        ntile = self._buf("moe_ntile", (1,), i32)
        perm = self._buf("moe_perm", (N, NBK), i32)
        rscale = self._buf("moe_rscale", (cap,), f32)
        xp = self._buf("moe_xp", (cap, D), adt)
        hp = self._buf("moe_hp", (cap, Fd), adt)
        yp = self._buf("moe_yp", (cap, D), adt)
        # cross-attention output Linear + residual (x2) and the gate of both MoE branches on the row it produces: one
        # kernel in bf16 mode (not with injected routing: the parity hook goes through the stand-alone gate)
        if not (not ca_done and fuse and self._fuse_gate and forced is None and
                ops.gemm_gate(a1, L["ca_out"][0], L["ca_out"][1], resid=x1, out_f32=x2, alpha=1.0, beta=1.0, NB=NB, E=E,
                              ln_w=L["moe_ln_w"], ln_b=L["moe_ln_b"], gate_w=L["gate_w"], gate_b=L["gate_b"], idx=idx,
                              vals=vals, stats=stats, blk_hist=hist, blk_imp=imp)):
            if not ca_done:
                self._lin(a1, L["ca_out"], out_f32=x2, resid=x1, alpha=1.0, beta=1.0)
            ops.moe_gate(x2, N, D, NB, E, L["moe_ln_w"], L["moe_ln_b"], L["gate_w"], L["gate_b"], idx, vals, stats,
                         hist, imp, forced_idx=forced)
        ops.moe_scan(hist, imp, idx, N, NB, E, Fd, D, base, seg, t_up, t_dn, ntile, pk["usage"][li],
                     pk["importance"][li])
        ops.moe_permute(x2, N, D, NB, E, L["moe_ln_w"], L["moe_ln_b"], idx, vals, stats, base, seg, xp, perm,
                        rscale)
        if self.record_routing:
            self.last_routing.append((idx.clone(), vals.clone()))
        kw = dict(tiles=None, num_tiles=max_tiles, num_tiles_dev=ntile, M=cap)
        self._lin(xp, (L["w1"], L["b1"]), out_a=hp, act=ACT_GELU, N=Fd, a_rows=cap, w_rows=G * Fd,
                  **dict(kw, tiles=t_up))
        self._lin(hp, (L["w2"], L["b2"]), out_a=yp, N=D, rowscale=rscale, a_rows=cap, w_rows=G * D,
                  **dict(kw, tiles=t_dn))
        ops.moe_combine_film(yp, perm, N, D, NBK, L["ffn_s_norm"][0], L["ffn_s_norm"][1], film[3], T, a1)
        return self._layer_tail(li, L, x, x2, None, a1, a2, xa, x1, ctx, Bn, T, N, nxt)

    def _layer_tail(self, li, L, x, x2, _unused, a1, a2, xa, x1, ctx, Bn, T, N, nxt=None):
        """ffn.proj_out Linear + MemoryEfficientCrossAttentionBlock (fast_attention.py:301-330)."""
        adt, D, H = self._adt(), self.latent_dim, self.num_heads
        adti = MDM_BF16 if adt == torch.bfloat16 else MDM_F32
        x3 = self._buf("x3", (N, D), torch.float32)
        if adt == torch.bfloat16:
            self._lin(a1, L["ffn_out"], out_f32=x3, out_a=xa, resid=x2, alpha=1.0, beta=1.0)
            x3a = xa
        else:
            self._lin(a1, L["ffn_out"], out_f32=x3, resid=x2, alpha=1.0, beta=1.0)
            x3a = x3
        # ---- MemoryEfficientCrossAttentionBlock (fast_attention.py:301-330)
        self._lin(x3a, L["sd_q"], out_a=a1)
        ops.softmax_cross(a1, ctx.k2[li], ctx.v2[li], ctx.nt, Bn, T, ctx.nt_max, H, D // H, a2)
        rr = x1  # x1 is dead from here on
        fuse = self._fuse_ln and adt == torch.bfloat16
        # output projection: rr = proj + x3 (fp32), and the LayerNorm of the projection itself in the same kernel
        if fuse and ops.gemm_ln(a2, L["sd_o"][0], L["sd_o"][1], ln1=L["sd_ln"], alpha=1.0, beta=1.0, resid=x3, out_f32=rr,
                                ln_pre_resid=True, out1_a=a1):
            n_in = a1
        else:
            ops.gemm(a2, L["sd_o"][0], L["sd_o"][1], out_f32=rr, out_a=a1, a_pre_resid=True, resid=x3, alpha=1.0,
                     beta=1.0)
            ops.rowop(a1, N, D, adti, ln1=L["sd_ln"], out1_a=a2)
            n_in = a2
        f1 = self._buf("f1", (N, 4 * D), adt)
        self._lin(n_in, L["sd_f1"], out_a=f1, act=ACT_GELU)
        if fuse and nxt is not None:
            # the layer output + the next layer's two pre-norms (h fp32, a0) and the bf16 copy of x (xa): all three
            # buffers are dead at this point of the layer
            Ln = self._packed["layers"][nxt]
            if ops.gemm_ln(f1, L["sd_f3"][0], L["sd_f3"][1], ln1=Ln["dsa_pre"], alpha=1.0, beta=1.0, resid=rr, out_f32=x,
                           out_a=xa, out1_f32=self._buf("h", (N, D), torch.float32), ln2=Ln["perf"][0]["pre"],
                           out2_a=self._buf("a0", (N, D), adt)):
                return True
        self._lin(f1, L["sd_f3"], out_f32=x, resid=rr, alpha=1.0, beta=1.0)
        return False

    def _embeddings(self, timesteps, xf_proj, Bn):
        """fused_emb (transformer.py:313-321) and the FiLM (scale|shift) of all 8L StylizationBlocks
        (stylization.py:22-27), as two grouped GEMMs over the stacked per-block weights."""
        pk = self._packed
        adt, D, Dt, Te = self._adt(), self.latent_dim, self.text_latent_dim, self.time_embed_dim
        f32 = torch.float32
        dev = timesteps.device
        e0 = self._buf("e_sin", (Bn, D), adt)
        e1 = self._buf("e_1", (Bn, 2 * D), adt)
        e2 = self._buf("e_2", (Bn, D), adt)
        e3 = self._buf("e_3", (Bn, Te), adt)
        e4 = self._buf("e_4", (Bn, Te), adt)
        e5 = self._buf("e_5", (Bn, D), adt)
        ops.timestep_embedding(timesteps, Bn, D, e0)
        self._lin(e0, pk["learnable_time_embed.mlp.0"], out_a=e1, act=ACT_SILU)
        self._lin(e1, pk["learnable_time_embed.mlp.2"], out_a=e2)
        self._lin(e2, pk["time_embed.0"], out_a=e3, act=ACT_SILU)
        self._lin(e3, pk["time_embed.2"], out_a=e4)
        self._lin(e4, pk["time_proj"], out_a=e5)
        xpa = self._buf("e_xp", (Bn, Dt), adt)
        ops.pad_cast(xf_proj, Bn, Dt, xpa)
        if Dt != D:
            xpj = self._buf("e_xpj", (Bn, D), adt)
            self._lin(xpa, pk["text_proj"], out_a=xpj)
        else:
            xpj = xpa
        tt = self._buf("e_tt", (Bn, D), f32)
        xx = self._buf("e_xx", (Bn, D), f32)
        ops.gemm(e5, *pk["gated_fusion.proj_time"], out_f32=tt)
        ops.gemm(xpj, *pk["gated_fusion.proj_text"], out_f32=xx)
        fu = self._buf("e_fu", (Bn, D), adt)
        ops.gated_mix(tt, xx, fu)
        f1 = self._buf("e_f1", (Bn, D), adt)
        emb = self._buf("e_emb", (Bn, D), adt)
        self._lin(fu, pk["gated_fusion.post_mlp.0"], out_a=f1, act=ACT_SILU)
        self._lin(f1, pk["gated_fusion.post_mlp.2"], out_a=emb)
        t1, t2, Bpad = self._film_tile_tables(Bn, dev)
        S = pk["n_style"]
        g1 = self._buf("film_g1", (S * Bpad, Te), adt)
        film = self._buf("film", (S * Bpad, 2 * D), f32)
        self._lin(emb, (pk["eph_w"], pk["eph_b"]), out_a=g1, act=ACT_SILU, N=Te, M=S * Bpad, tiles=t1,
                  num_tiles=t1.shape[0], a_rows=Bn, w_rows=S * Te)
        ops.gemm(g1, pk["emb_w"], pk["emb_b"], out_f32=film, N=2 * D, M=S * Bpad, tiles=t2, num_tiles=t2.shape[0],
                 a_rows=S * Bpad, w_rows=S * 2 * D)
        return film, Bpad

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, x: torch.Tensor, timesteps: torch.Tensor, length: torch.Tensor,
                text: Optional[List[str]] = None, xf_proj=None, xf_out=None, *, nt=None,
                text_ctx: Optional[TextContext] = None) -> torch.Tensor:
        """x [B,T,input_feats] float, timesteps [B] int64, length [B] (or [B,1]) int64 ->
        [B,T,input_feats] fp32.  T must be even and <= num_frames (the reference fails on odd T, H8)."""
        if not x.is_cuda:
            raise MdmError("MotionTransformer.forward needs CUDA tensors: there is no CPU fallback")
        if x.device != self._t("sequence_embedding").device:
            raise MdmError("input on %s but the model lives on %s" % (x.device, self._t("sequence_embedding").device))
        with torch.cuda.device(x.device):           # the C library launches on the current device / stream
            return self._forward(x, timesteps, length, text, xf_proj, xf_out, nt, text_ctx)

    def _forward(self, x, timesteps, length, text, xf_proj, xf_out, nt, text_ctx):
        Bn, T, Fin = x.shape
        if Fin != self.input_feats:
            raise RuntimeError("expected %d input features, got %d" % (self.input_feats, Fin))
        if T % 2 or T > self.num_frames:
            raise RuntimeError("T=%d: must be even (the reference's skip connection h_up + h fails "
                               "otherwise) and <= num_frames=%d" % (T, self.num_frames))
        pk = self._packed or self._pack()
        if text_ctx is None:
            if xf_proj is None or xf_out is None:                       # transformer.py:311-312
                xf_proj, xf_out = self.encode_text(text, x.device)
            text_ctx = self.prepare_text(xf_proj, xf_out, nt)
        if text_ctx.B != Bn:
            raise RuntimeError("text context batch %d != batch %d" % (text_ctx.B, Bn))
        adt, D = self._adt(), self.latent_dim
        adti = MDM_BF16 if adt == torch.bfloat16 else MDM_F32
        f32 = torch.float32
        N = Bn * T
        timesteps = timesteps.to(torch.int64).contiguous()
        length = length.reshape(-1).to(torch.int64).contiguous()
        self.last_routing = []
        film, Bpad = self._embeddings(timesteps, text_ctx.xf_proj, Bn)
        # joint_embed + sequence embedding (transformer.py:324-326)
        xin = self._buf("xin", (N, pk["Kp"]), adt)
        ops.pad_cast(x.float().contiguous().view(N, Fin), N, Fin, xin)
        h = self._buf("h_emb", (N, D), f32)
        ha = self._buf("h_emb_a", (N, D), adt) if adt == torch.bfloat16 else h
        ops.gemm(xin, pk["je_w"], pk["je_b"], out_f32=h, out_a=(ha if ha is not h else None), resid=pk["seq_emb"],
                 resid_mod=T, alpha=1.0, beta=1.0)
        # downsample Conv1d(k=2,s=2) == GEMM over row pairs (transformer.py:332-337)
        Nl = N // 2
        hl = self._buf("x_low", (Nl, D), f32)
        ops.gemm(ha.view(Nl, 2 * D), pk["down_w"], pk["down_b"], out_f32=hl)
        nl = self.num_layers
        done = False
        for li in range(nl):                                             # transformer.py:341-344
            done = self._layer(li, hl, text_ctx, film, Bpad, Bn, T // 2, length, 1, pre_done=done,
                               nxt=li + 1 if li + 1 < nl else None)
        # upsample ConvTranspose1d(k=2,s=2) + skip (transformer.py:347-353)
        if adt == torch.bfloat16:
            hla = self._buf("x_low_a", (Nl, D), adt)
            ops.rowop(hl, Nl, D, adti, out0_a=hla)
        else:
            hla = hl
        hc = self._buf("x_high", (N, D), f32)
        ops.gemm(hla, pk["up_w"], pk["up_b"], out_f32=hc.view(Nl, 2 * D), resid=h.view(Nl, 2 * D), alpha=1.0,
                 beta=1.0)
        done = False
        for li in range(nl, 2 * nl):                                     # transformer.py:356-357
            done = self._layer(li, hc, text_ctx, film, Bpad, Bn, T, length, 0, pre_done=done,
                               nxt=li + 1 if li + 1 < 2 * nl else None)
        if adt == torch.bfloat16:
            hca = self._buf("x_high_a", (N, D), adt)
            ops.rowop(hc, N, D, adti, out0_a=hca)
        else:
            hca = hc
        out = torch.empty(Bn, T, Fin, dtype=f32, device=x.device)
        ops.gemm(hca, pk["out_w"], pk["out_b"], out_f32=out.view(N, Fin))   # transformer.py:360
        return out
