"""Default-config model + synthetic inputs for the profiling tools (no oracle: weights come from the
model's own initialisers, with the zero-initialised tensors re-randomised like bench.py does)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import motiondiffusion_moe_b200 as mdm

CFG = dict(input_feats=263, num_frames=196, latent_dim=512, ff_size=1024, num_layers=8, num_heads=4,
           text_latent_dim=256, moe_num_experts=8)


def build(dev, B, T=196, seed=0):
    torch.manual_seed(seed)
    net = mdm.MotionTransformer(precision="bf16", **CFG)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for n, p in net.named_parameters():
            if p.abs().sum() == 0 and not n.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    net.repack()
    net.to(dev)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, T, CFG["input_feats"], generator=g).to(dev)
    t = torch.randint(0, 1000, (B,), generator=g).to(dev)
    length = torch.randint(40, T + 1, (B,), generator=g).to(dev)
    xf_out = torch.nn.functional.gelu(torch.randn(B, 20, CFG["text_latent_dim"], generator=g)).to(dev)
    return net, x, t, length, xf_out.mean(1), xf_out
