"""Generates tests/golden/train_tiny.npz from the UNMODIFIED reference: one DDPM training-step evaluation
(GaussianDiffusion.training_losses, models/gaussian_diffusion.py:923-992, + the masked loss of
DDPMTrainer.backward_G, trainers/ddpm_trainer.py:201-214) with autograd gradients, in eval() mode (dropout and
StochasticDepth are then deterministic pass-throughs; BASELINE.json configs[4] trains with dropout = 0).

This pins the oracle for the NOT-YET-BUILT backward kernels of the training step:
tests/test_oracle_golden.py checks that autograd through oracle/motion_oracle.py reproduces these gradients, so the
oracle can serve as the parity reference when the backward kernels are written.
Build container only: `python tests/golden/make_train_golden.py`."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import build_reference, GaussianDiffusion, get_named_beta_schedule, ModelMeanType, ModelVarType, LossType  # noqa: E402
from oracle import cases  # noqa: E402


def main():
    case = "tiny_b3"
    cfg_name, B, T = cases.CASES[case]
    cfg, params = cases.case_params(case)
    model = build_reference(cfg, params)          # eval()
    x0, t, length, xf_proj, xf_out = cases.make_inputs(cfg, B, T, seed=5)
    noise = torch.randn(x0.shape, generator=torch.Generator().manual_seed(6))
    diff = GaussianDiffusion(betas=get_named_beta_schedule("linear", 1000), model_mean_type=ModelMeanType.EPSILON,
                             model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
    torch.manual_seed(cases.EPH_SEED)             # ephemeral Linears of the forward (H1/H2) == oracle.draw_ephemerals
    out = diff.training_losses(model, x0, t, model_kwargs={"length": length, "xf_proj": xf_proj, "xf_out": xf_out},
                               noise=noise)
    per_frame = ((out["pred"] - out["target"]) ** 2).mean(dim=-1)
    mask = model.generate_src_mask(T, length).to(per_frame.device).view(per_frame.shape)
    loss_rec = (per_frame * mask).sum() / mask.sum()
    total = loss_rec + out["moe_loss"]
    total.backward()
    names, norms = [], []
    for n, p in model.named_parameters():
        if n.startswith("text_encoder."):
            continue
        names.append(n)
        norms.append(0.0 if p.grad is None else float(p.grad.norm()))
    sd = dict(model.named_parameters())
    full = {k: sd[k].grad.numpy() for k in ("joint_embed.weight", "out.bias",
                                            "decoder_blocks_high.0.module.ffn.branches.0.moe.gate.weight",
                                            "decoder_blocks_low.0.module.dual_self_attn.local_attn.query.weight")}
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "train_tiny.npz"),
                        loss_rec=np.float32(loss_rec.item()), moe_loss=np.float32(float(out["moe_loss"])),
                        mse=out["mse"].detach().numpy(), names=np.array(names), grad_norms=np.array(norms, dtype=np.float32),
                        **{"grad::" + k: v for k, v in full.items()})
    print("loss_rec %.6f moe %.4f params %d nonzero grads %d" % (loss_rec.item(), float(out["moe_loss"]), len(names),
                                                                 sum(1 for v in norms if v > 0)))


if __name__ == "__main__":
    main()
