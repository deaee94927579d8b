"""CFG step time of the default model as a function of the per-GPU batch (what each rank of the STRONG-scaled, batch-sharded
sampler of BASELINE configs[2] sees: global batch 64 over N GPUs = 64/N sequences per rank, cond + uncond batched).
Graph replay, device-resident inputs.  python tools/step_vs_batch.py [B ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B_  # noqa: E402
import motiondiffusion_moe_b200 as mdm  # noqa: E402


def main():
    batches = [int(a) for a in sys.argv[1:]] or [64, 32, 16, 8]
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    net = mdm.MotionTransformer(precision="bf16", **B_.CFG)
    B_.randomize_zero_init(net)
    net.to(dev)
    d = mdm.GaussianDiffusion(betas=mdm.get_named_beta_schedule("linear", 1000))
    out = []
    for B in batches:
        x0, length, xf_c, xf_u = B_.synth_inputs(B, 1000, dev)
        net.encode_text = lambda text, device: (xf_u.mean(1), xf_u) if text[0] == "" else (xf_c.mean(1), xf_c)
        kw = {"text": ["a"] * B, "length": length, "xf_proj": xf_c.mean(1), "xf_out": xf_c}
        st = d.make_cfg_stepper(net, (B, B_.T, 263), kw, cfg_scale=7.5, clip_denoised=False, device=dev)
        st.x.copy_(x0)
        for i in range(5):
            st.step(999 - i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        torch.cuda.synchronize()
        e0.record()
        for i in range(n):
            st.step(900 - i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out.append({"batch": B, "ms_per_step": ms, "frames_per_s": B * B_.T / (ms * 1e-3), "us_per_sequence": 1e3 * ms / B})
        print(json.dumps(out[-1]))
        del st
        net._ws = {}
        torch.cuda.empty_cache()
    b64 = [o for o in out if o["batch"] == 64]
    if b64:
        for o in out:
            print("B=%d: %.3f ms/step; strong-scaling efficiency vs B=64 on %d GPUs: %.2f" %
                  (o["batch"], o["ms_per_step"], 64 // o["batch"], b64[0]["ms_per_step"] / (64 / o["batch"]) / o["ms_per_step"]))


if __name__ == "__main__":
    main()
