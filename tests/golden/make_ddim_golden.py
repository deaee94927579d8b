"""Generates tests/golden/ddim_step.npz from the UNMODIFIED reference GaussianDiffusion.ddim_sample
(models/gaussian_diffusion.py:699-742) with a stand-in model that returns a fixed eps tensor.  Run in the build
container only (`python tests/golden/make_ddim_golden.py`): /root/reference does not exist on the GPU box."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference/text2motion")
from models.gaussian_diffusion import (GaussianDiffusion, get_named_beta_schedule, ModelMeanType,  # noqa: E402
                                       ModelVarType, LossType)


def main():
    d = GaussianDiffusion(betas=get_named_beta_schedule("linear", 1000), model_mean_type=ModelMeanType.EPSILON,
                          model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
    g = torch.Generator().manual_seed(11)
    x, eps = torch.randn(6, 5, 33, generator=g), torch.randn(6, 5, 33, generator=g)
    t = torch.tensor([0, 1, 17, 500, 998, 999])
    out = {"x": x.numpy(), "eps": eps.numpy(), "t": t.numpy()}
    for eta in (0.0, 0.5, 1.0):
        for clip in (0, 1):
            torch.manual_seed(5)
            r = d.ddim_sample(lambda xx, tt, **kw: eps, x, t, clip_denoised=bool(clip), eta=eta, model_kwargs={})
            out["sample_eta%g_clip%d" % (eta, clip)] = r["sample"].numpy()
            out["x0_eta%g_clip%d" % (eta, clip)] = r["pred_xstart"].numpy()
    torch.manual_seed(5)
    out["noise"] = torch.randn_like(x).numpy()          # what ddim_sample drew (:735)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ddim_step.npz"), **out)


if __name__ == "__main__":
    main()
