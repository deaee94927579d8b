"""torchrun worker: data-parallel CFG sampling (motiondiffusion_moe_b200/parallel.py:sample_dp) on the REAL model and
kernels: the batch sharded over the ranks (ragged: 5 sequences over 2 ranks, ...), CUDA-graph steppers, noise drawn on
the device from one seed, final all_gather - against the unsharded loop run by every rank itself (layout=(1, 0)).
north_star (4) / BASELINE configs[2]: the N-GPU sample must equal the 1-GPU sample BIT FOR BIT.  Prints DP_MODEL_OK."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from dist_common import init_dist, all_max  # noqa: E402
import motiondiffusion_moe_b200 as mdm  # noqa: E402
from motiondiffusion_moe_b200 import parallel  # noqa: E402
from oracle import cases, motion_oracle as mo  # noqa: E402


def main():
    rank, world, dev, shared = init_dist()
    cfg, p = cases.case_params("small_b4")
    B, T, steps = world * 2 + 1, 60, 6
    net = mdm.MotionTransformer(precision="bf16", **cfg)
    net.load_state_dict({k: p[k] for k in net.state_dict()})
    net.load_extras(p)
    net.to(dev)

    def encode(text, device):
        # like a real encoder, the embedding of a sequence depends on its own string only (not on its position in the
        # batch): every "" row gets the same embedding, so a shard sees exactly the rows of the global batch
        pooled, tok = mo.stub_text(text[:1], cfg.text_latent_dim, device)
        return pooled.expand(len(text), -1).contiguous(), tok.expand(len(text), -1, -1).contiguous()
    net.encode_text = encode
    _, _, length, xf_proj, xf_out = cases.make_inputs(cfg, B, T, seed=40, device=dev)     # same global batch on every rank
    d = mdm.GaussianDiffusion(betas=mdm.get_named_beta_schedule("linear", 1000))
    kw = {"text": ["a person walks"] * B, "length": length, "xf_proj": xf_proj, "xf_out": xf_out}
    shape = (B, T, cfg.input_feats)
    mk = lambda s, k: d.make_cfg_stepper(net, s, k, cfg_scale=7.5, clip_denoised=False, device=dev)
    ok = True
    for noise in ("device", "host"):
        full = parallel.sample_dp(mk, shape, kw, 1000, seed=9, num_steps=steps, noise=noise, layout=(1, 0))
        sharded = parallel.sample_dp(mk, shape, kw, 1000, seed=9, num_steps=steps, noise=noise)
        same = torch.equal(full, sharded)
        if rank == 0:
            err = ((full - sharded).norm() / full.norm()).item()
            print("dp world=%d noise=%s: sharded vs unsharded sample bit-identical=%s (rel %.2e)" % (world, noise, same, err))
        ok &= same and bool(torch.isfinite(sharded).all())
    # strided DDIM schedule through the same path
    order, prev = d.ddim_timesteps(4)
    mkd = lambda s, k: mdm.CFGStepper(d, net, s, k, 7.5, False, dev, True, sampler="ddim", eta=0.0)
    full = parallel.sample_dp(mkd, shape, kw, 1000, seed=9, schedule=list(zip(order, prev)), noise="device", layout=(1, 0))
    sharded = parallel.sample_dp(mkd, shape, kw, 1000, seed=9, schedule=list(zip(order, prev)), noise="device")
    ok &= torch.equal(full, sharded)
    flag = all_max([0.0 if ok else 1.0], dev, shared)[0]
    if rank == 0:
        print("DP_MODEL_OK" if flag == 0.0 else "DP_MODEL_MISMATCH")
    dist.destroy_process_group()
    sys.exit(0 if flag == 0.0 else 1)


if __name__ == "__main__":
    main()
