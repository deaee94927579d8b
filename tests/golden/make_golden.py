"""Generates tests/golden/*.npz from the UNMODIFIED reference (run in the build container only:
`python tests/golden/make_golden.py`; /root/reference does not exist on the GPU box).

For every case the reference MotionTransformer (with the DeBERTa text encoder replaced by the
parameter-free stub of SURVEY.md Appendix C, because the hub weights cannot be downloaded) is loaded
with oracle.make_params weights, its lazy projection matrices are set from the same dict, and it is
run under torch.manual_seed(EPH_SEED) so that its per-forward ephemeral Linears (H1/H2) equal
oracle.draw_ephemerals.  Stored: forward output, routing indices of every SwitchMoELayer call
(captured by wrapping torch.topk), usage/importance counters, and one p_sample_with_cfg step.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/text2motion")

import models.transformer as mt  # noqa: E402  (reference)
from models.gaussian_diffusion import (GaussianDiffusion, get_named_beta_schedule, ModelMeanType,  # noqa: E402
                                       ModelVarType, LossType)
from oracle import cases, motion_oracle as mo  # noqa: E402


class StubTextEncoder(nn.Module):
    def __init__(self, output_dim, dropout=0.1):
        super().__init__()
        self.output_dim = output_dim

    def forward(self, text, device):
        return mo.stub_text(text, self.output_dim, device)


mt.EnhancedTextEncoder = StubTextEncoder


def build_reference(cfg, params):
    m = mt.MotionTransformer(dropout=0.1, **cfg).eval()
    sd = {k: v for k, v in params.items() if k in m.state_dict()}
    missing = set(m.state_dict()) - set(sd)
    assert not missing, sorted(missing)[:5]
    m.load_state_dict(sd)
    for name, mod in m.named_modules():
        if name.endswith("fast_attention"):
            mod.projection_matrix = params[name + ".projection_matrix"].clone()
    return m


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for case, (cfg_name, B, T) in cases.CASES.items():
        cfg, params = cases.case_params(case)
        model = build_reference(cfg, params)
        x, t, length, xf_proj, xf_out = cases.make_inputs(cfg, B, T, seed=3)
        routing = []
        real_topk = torch.topk

        def spy(inp, k, dim=-1, **kw):
            r = real_topk(inp, k, dim=dim, **kw)
            routing.append(r[1].clone())
            return r

        torch.topk = spy
        try:
            torch.manual_seed(cases.EPH_SEED)
            with torch.no_grad():
                y = model(x, t, length, None, xf_proj, xf_out)
        finally:
            torch.topk = real_topk
        n_low = cfg.num_layers * 2
        usage = torch.stack([m.expert_usage for m in model.modules() if isinstance(m, mt.SwitchMoELayer)])
        imp = torch.stack([m.expert_importance for m in model.modules() if isinstance(m, mt.SwitchMoELayer)])
        moe_loss = float(model.get_moe_loss(model))

        # one CFG step through the reference sampler (ephemerals differ per forward: replayed stream)
        diff = GaussianDiffusion(betas=get_named_beta_schedule("linear", 1000),
                                 model_mean_type=ModelMeanType.EPSILON,
                                 model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
        ts = torch.full((B,), 500, dtype=torch.long)
        torch.manual_seed(cases.EPH_SEED + 1)
        with torch.no_grad():
            step = diff.p_sample_with_cfg(model, x, ts, clip_denoised=False,
                                          model_kwargs={"text": ["a person walks"] * B, "length": length,
                                                        "xf_proj": xf_proj, "xf_out": xf_out},
                                          cfg_scale=7.5)
        np.savez_compressed(
            os.path.join(out_dir, case + ".npz"),
            y=y.numpy(), routing_low=torch.stack(routing[:n_low]).numpy().astype(np.int8),
            routing_high=torch.stack(routing[n_low:]).numpy().astype(np.int8),
            usage=usage.numpy(), importance=imp.numpy(), moe_loss=np.float32(moe_loss),
            cfg_sample=step["sample"].numpy(), cfg_x0=step["pred_xstart"].numpy(),
            torch_version=np.array(torch.__version__))
        print(case, "y", tuple(y.shape), "abs-mean %.4f" % y.abs().mean().item(), "routing calls", len(routing),
              "moe_loss %.4f" % moe_loss)


if __name__ == "__main__":
    main()
