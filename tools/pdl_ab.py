"""A/B of programmatic dependent launch (mdm_set_pdl) on the CUDA-graph CFG step, interleaved in ONE process so that
clock / thermal state is shared: for each batch, capture a stepper with the launch attribute off, one with it on,
and time them alternately.  python tools/pdl_ab.py [B ...]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B_  # noqa: E402
import motiondiffusion_moe_b200 as mdm  # noqa: E402
from motiondiffusion_moe_b200 import _lib  # noqa: E402


def main():
    batches = [int(a) for a in sys.argv[1:]] or [64, 8]
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    lib = _lib.load()
    net = mdm.MotionTransformer(precision="bf16", **B_.CFG)
    B_.randomize_zero_init(net)
    net.to(dev)
    d = mdm.GaussianDiffusion(betas=mdm.get_named_beta_schedule("linear", 1000))
    for B in batches:
        x0, length, xf_c, xf_u = B_.synth_inputs(B, 1000, dev)
        net.encode_text = lambda text, device: (xf_u.mean(1), xf_u) if text[0] == "" else (xf_c.mean(1), xf_c)
        kw = {"text": ["a"] * B, "length": length, "xf_proj": xf_c.mean(1), "xf_out": xf_c}
        steppers = {}
        for on in (0, 1):
            lib.mdm_set_pdl(on)
            st = d.make_cfg_stepper(net, (B, B_.T, 263), kw, cfg_scale=7.5, clip_denoised=False, device=dev)
            st.x.copy_(x0)
            for i in range(4):
                st.step(999 - i)
            steppers[on] = st
        res = {0: [], 1: []}
        for rnd in range(4):
            for on in (0, 1):
                st = steppers[on]
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for i in range(40):
                    st.step(900 - i)
                e1.record()
                torch.cuda.synchronize()
                res[on].append(e0.elapsed_time(e1) / 40)
        # same inputs, eager forward with the attribute off / on: the outputs must be bit-identical (the steppers above
        # draw their own noise, their states are not comparable)
        outs = []
        for on in (0, 1):
            lib.mdm_set_pdl(on)
            outs.append(net(x0.to(dev), torch.full((B,), 500, dtype=torch.long, device=dev), length, None, xf_c.mean(1), xf_c).clone())
        same = bool(torch.equal(outs[0], outs[1]))
        print(json.dumps({"batch": B, "ms_per_step_pdl_off": res[0], "ms_per_step_pdl_on": res[1],
                          "median_off": sorted(res[0])[2], "median_on": sorted(res[1])[2], "forward_bit_identical_off_vs_on": same}))
        del steppers
        net._ws = {}
        torch.cuda.empty_cache()
    lib.mdm_set_pdl(1)


if __name__ == "__main__":
    main()
