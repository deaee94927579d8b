"""ORACLE support — test infrastructure only.  Runs the UNMODIFIED reference (oracle/_ref, built by oracle/build_ref.py
from /root/reference/text2motion/models) for baselines and cross-checks: the reference's own MotionTransformer and
GaussianDiffusion objects, its own stock code path (per-forward ephemeral Linears on the CPU RNG and all),
with one substitution that the offline sandbox forces: EnhancedTextEncoder (models/text_encoder.py:6-43 downloads
DeBERTa-v3-large from the hub) is replaced by the parameter-free stub of SURVEY.md Appendix C, and text embeddings are
passed explicitly.  Only bench.py (baseline legs) and tests may import this."""
import os
import sys

import torch
import torch.nn as nn

from . import build_ref
from . import motion_oracle as mo

_mods = {}


def available():
    return build_ref.available()


def _import():
    if "t" in _mods:
        return _mods["t"], _mods["g"]
    if not available():
        raise RuntimeError("oracle/_ref is not built: run `python -m oracle.build_ref` where /root/reference exists")
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "models")
    import importlib.abc
    import importlib.machinery
    import importlib.util

    class RefFinder(importlib.abc.MetaPathFinder):
        """Resolves the package `models` (the reference's text2motion/models) to the byte code under oracle/_ref."""
        def find_spec(self, fullname, path=None, target=None):
            if fullname == "models":
                f = os.path.join(root, "__init__.refbc")
                return importlib.util.spec_from_file_location(fullname, f, loader=importlib.machinery.SourcelessFileLoader(fullname, f),
                                                              submodule_search_locations=[root])
            if fullname.startswith("models."):
                f = os.path.join(root, fullname.split(".", 1)[1] + ".refbc")
                if os.path.exists(f):
                    return importlib.util.spec_from_file_location(fullname, f,
                                                                  loader=importlib.machinery.SourcelessFileLoader(fullname, f))
            return None

    if not any(type(f).__name__ == "RefFinder" for f in sys.meta_path):
        sys.meta_path.insert(0, RefFinder())
    import models.transformer as mt            # noqa: E402  (the reference, sourceless)
    import models.gaussian_diffusion as gd     # noqa: E402

    class StubTextEncoder(nn.Module):
        def __init__(self, output_dim, dropout=0.1):
            super().__init__()
            self.output_dim = output_dim

        def forward(self, text, device):
            return mo.stub_text(text, self.output_dim, device)

    mt.EnhancedTextEncoder = StubTextEncoder
    _mods["t"], _mods["g"] = mt, gd
    return mt, gd


def build_model(cfg, params, device="cpu"):
    """The reference MotionTransformer(**cfg).eval() holding `params` (state_dict entries + projection matrices)."""
    mt, _ = _import()
    m = mt.MotionTransformer(dropout=0.1, **cfg).eval()
    sd = {k: v for k, v in params.items() if k in m.state_dict()}
    missing = set(m.state_dict()) - set(sd)
    assert not missing, sorted(missing)[:5]
    m.load_state_dict(sd)
    m.to(device)
    for name, mod in m.named_modules():
        if name.endswith("fast_attention"):
            mod.projection_matrix = params[name + ".projection_matrix"].clone().to(device)
    return m


def diffusion(steps=1000):
    _, gd = _import()
    return gd.GaussianDiffusion(betas=gd.get_named_beta_schedule("linear", steps),
                                model_mean_type=gd.ModelMeanType.EPSILON, model_var_type=gd.ModelVarType.FIXED_SMALL,
                                loss_type=gd.LossType.MSE)


def cfg_step(model, diff, x, t, length, xf_proj, xf_out, cfg_scale=7.5):
    """One p_sample_with_cfg call of the reference (models/gaussian_diffusion.py:1042-1098): two sequential forwards
    (the unconditional one re-encodes text="" through the stub) + the guided DDPM update."""
    with torch.no_grad():
        return diff.p_sample_with_cfg(model, x, t, clip_denoised=False, cfg_scale=cfg_scale,
                                      model_kwargs={"text": ["a person walks forward"] * x.shape[0], "length": length,
                                                    "xf_proj": xf_proj, "xf_out": xf_out})
