"""torchrun worker: MotionTransformer with expert parallelism (enable_expert_parallel) against the same
model without it, on this rank's own batch: forward outputs, routing and MoE counters must be bit-identical;
then CFG sampling steps through the CUDA-graph stepper (the flag barriers replay inside the graph)."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from dist_common import init_dist, all_max  # noqa: E402
import motiondiffusion_moe_b200 as mdm  # noqa: E402
from oracle import cases, motion_oracle as mo  # noqa: E402


def main():
    rank, world, dev, shared = init_dist()
    cfg, p = cases.case_params("small_b4")            # 4 layers x 2 scales, D256, 4 experts (divisible by 2 / 4 ranks)
    B, T = 3, 60

    def make():
        net = mdm.MotionTransformer(precision="bf16", **cfg)
        net.load_state_dict({k: p[k] for k in net.state_dict()})
        net.load_extras(p)
        return net.to(dev)

    ref, epn = make(), make()
    epn.enable_expert_parallel()
    x, t, length, xf_proj, xf_out = cases.make_inputs(cfg, B, T, seed=20 + rank, device=dev)   # every rank: own batch
    ref.record_routing = epn.record_routing = True
    ok = True
    for it in range(2):
        y0 = ref(x, t, length, None, xf_proj, xf_out)
        y1 = epn(x, t, length, None, xf_proj, xf_out)
        same_y = torch.equal(y0, y1)
        same_r = all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(ref.last_routing, epn.last_routing))
        if not (same_y and same_r):
            print("rank %d forward %d: outputs equal %s (rel %.3e), routing equal %s" %
                  (rank, it, same_y, ((y0 - y1).norm() / y0.norm()).item(), same_r), flush=True)
        ok &= same_y and same_r
    sd0, sd1 = ref.state_dict(), epn.state_dict()
    same_c = all(torch.equal(sd0[k], sd1[k]) for k in sd0 if "expert_usage" in k)
    if not same_c:
        print("rank %d: usage counters differ" % rank, flush=True)
    ok &= same_c
    for ep in epn._ep_inst.values():
        ep.check_health()
    # CFG sampling steps with the step captured in a CUDA graph
    ref.record_routing = epn.record_routing = False
    stub = lambda text, device: mo.stub_text(text, cfg.text_latent_dim, device)
    ref.encode_text = epn.encode_text = stub
    d = mdm.GaussianDiffusion(betas=mdm.get_named_beta_schedule("linear", 1000))
    kw = {"text": ["a person walks"] * B, "length": length, "xf_proj": xf_proj, "xf_out": xf_out}
    g = torch.Generator().manual_seed(5 + rank)
    x_T = torch.randn(B, T, cfg.input_feats, generator=g).to(dev)
    noises = torch.randn(8, B, T, cfg.input_feats, generator=g).to(dev)
    outs = []
    for net in (ref, epn):
        outs.append(d.p_sample_loop_with_cfg(net, (B, T, cfg.input_feats), noise=x_T, clip_denoised=False, model_kwargs=kw,
                                             cfg_scale=7.5, num_steps=8, step_noise=lambda ts: noises[999 - ts]))
    if not torch.equal(outs[0], outs[1]):
        print("rank %d: graph-captured CFG steps differ (rel %.3e)" % (rank, ((outs[0] - outs[1]).norm() / outs[0].norm()).item()), flush=True)
    ok &= torch.equal(outs[0], outs[1])
    for ep in epn._ep_inst.values():
        ep.check_health()
    flag = all_max([0.0 if ok else 1.0], dev, shared)[0]
    if rank == 0:
        print("EP_MODEL_OK" if float(flag) == 0.0 else "EP_MODEL_MISMATCH")
    for ep in epn._ep_inst.values():
        ep.close()
    dist.destroy_process_group()
    sys.exit(0 if float(flag) == 0.0 else 1)


if __name__ == "__main__":
    main()
