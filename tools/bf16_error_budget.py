"""Where the bf16 path's error against the fp32 oracle comes from: per decoder layer with teacher-forced inputs and
the oracle's routing injected, the relative L2 error of the residual stream after each of the four sub-blocks
(DualSelfAttention -> x1, GatedCrossAttention -> x2, MoEMultiBranchFFN -> x3, MemoryEfficientCrossAttention -> out).
Diagnostic tool (imports the oracle: not part of the product).  python tools/bf16_error_budget.py [case]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import motiondiffusion_moe_b200 as mdm  # noqa: E402
from oracle import cases, motion_oracle as mo  # noqa: E402

DEV = torch.device("cuda")
rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm()).item()


def main(case):
    cfg_name, B, T = cases.CASES[case]
    cfg, p = cases.case_params(case, DEV)
    net = mdm.MotionTransformer(precision="bf16", **cfg)
    net.load_state_dict({k: p[k].cpu() for k in net.state_dict()})
    net.load_extras({k: v.cpu() for k, v in p.items()})
    net.to(DEV)
    x, t, length, xf_proj, xf_out = cases.make_inputs(cfg, B, T, seed=3, device=DEV)
    emb = mo.fused_embedding(p, cfg, t, xf_proj)
    net._packed or net._pack()
    ctx = net.prepare_text(xf_proj, xf_out)
    film, Bpad = net._embeddings(t, ctx.xf_proj, B)
    # the oracle's own per-layer inputs (true trajectory)
    H, E, D = cfg.num_heads, cfg.moe_num_experts, cfg.latent_dim
    with torch.no_grad():
        h = mo._lin(p, "joint_embed", x) + p["sequence_embedding"][None, :T]
        h_low = torch.nn.functional.conv1d(h.permute(0, 2, 1), p["downsample.weight"], p["downsample.bias"], stride=2).permute(0, 2, 1)
        cur = h_low
        print("%-34s %9s %9s %9s %9s" % ("layer", "x1(dsa)", "x2(ca)", "x3(moe)", "out(sd)"))
        for li, blk in enumerate(net.block_prefixes()):
            if li == cfg.num_layers:
                h_up = torch.nn.functional.conv_transpose1d(cur.permute(0, 2, 1), p["upsample.weight"], p["upsample.bias"], stride=2)
                cur = h_up.permute(0, 2, 1) + h
            Tl = cur.shape[1]
            shift = 1 if li < cfg.num_layers else 0
            mask = mo.src_mask(Tl, (length >> shift))
            o1 = mo.dual_self_attention(p, blk + ".dual_self_attn", cur, emb, mask, H)
            o2 = mo.gated_cross_attention(p, blk + ".cross_attn", o1, xf_out, emb, H)
            routing = []
            o3 = mo.moe_multibranch_ffn(p, blk + ".ffn", o2, emb, E, routing)
            o4 = mo.sd_cross_attention(p, blk + ".sd_cross_attn", o3, xf_out, H)
            net.force_routing = [None] * li + [torch.stack([routing[0][1], routing[1][1]], 1).to(torch.int32).contiguous()]
            buf = cur.clone().reshape(B * Tl, D).contiguous()
            net._layer(li, buf, ctx, film, Bpad, B, Tl, length, shift)
            N = B * Tl
            g = lambda n: net._ws[(n, (N, D), torch.float32)]
            # x1 is reused as scratch by the last sub-block: report what is still intact
            e2, e3, e4 = rel(g("x2").view(B, Tl, D), o2), rel(g("x3").view(B, Tl, D), o3), rel(buf.view(B, Tl, D), o4)
            print("%-34s %9s %9.2e %9.2e %9.2e   |o4-o3|/|o4| %.2f |o3-o2|/|o3| %.2f |o2-in|/|o2| %.2f" %
                  (blk, "-", e2, e3, e4, rel(o3, o4), rel(o2, o3), rel(cur, o2)))
            cur = o4
    net.force_routing = None


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "small_b4")
