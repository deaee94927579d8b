"""ORACLE support — deterministic synthetic inputs shared by the golden generator, the tests, smoke()
and bench.py (SURVEY.md §8(d) 'Synthetic inputs').  Test infrastructure only."""
import torch

from .motion_oracle import CONFIGS, make_params, draw_ephemerals, stub_text  # noqa: F401

# name -> (config name, batch, frames)
CASES = {
    "tiny_b3": ("tiny", 3, 8),
    "small_b4": ("small", 4, 196),      # BASELINE.json configs[0]
    "default_b2": ("default", 2, 196),  # BASELINE.json configs[1] architecture at a CPU-sized batch
}
PARAM_SEED = 0
EPH_SEED = 11


def make_inputs(cfg, B, T, seed=0, device="cpu", min_len=None):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, cfg.input_feats, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    lo = max(2, T // 5) if min_len is None else min_len
    length = torch.randint(lo, T + 1, (B,), generator=g)
    length[0] = T
    xf_out = torch.nn.functional.gelu(torch.randn(B, 20, cfg.text_latent_dim, generator=g))
    xf_proj = xf_out.mean(1)
    return tuple(v.to(device) for v in (x, t, length, xf_proj, xf_out))


def case_params(case, device="cpu"):
    cfg_name, _, _ = CASES[case]
    cfg = CONFIGS[cfg_name]
    p = make_params(cfg, PARAM_SEED)
    p.update(draw_ephemerals(cfg, EPH_SEED))
    return cfg, {k: v.to(device) for k, v in p.items()}
