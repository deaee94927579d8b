"""Generates tests/golden/ric.npz from the UNMODIFIED reference recover_from_ric (utils/motion_process.py:401-417)
on seeded HumanML3D- and KIT-shaped features.  Build container only (`python tests/golden/make_ric_golden.py`)."""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, "/root/reference/text2motion")
# utils/motion_process.py imports the skeleton / paramUtil helpers at module level; they are not needed here
try:
    from utils.motion_process import recover_from_ric
except Exception:                                       # pragma: no cover
    for name in ("utils.skeleton", "utils.paramUtil"):
        sys.modules.setdefault(name, types.ModuleType(name))
    from utils.motion_process import recover_from_ric


def main():
    out = {}
    for tag, (B, T, F, J) in {"t2m": (1, 196, 263, 22), "kit": (2, 40, 251, 21)}.items():
        g = torch.Generator().manual_seed(21 + J)
        x = torch.randn(B, T, F, generator=g)
        mean, std = torch.randn(F, generator=g) * 0.3, torch.rand(F, generator=g) * 0.5 + 0.05
        out[tag + "_x"], out[tag + "_mean"], out[tag + "_std"] = x.numpy(), mean.numpy(), std.numpy()
        out[tag + "_joints"] = recover_from_ric((x * std + mean).float(), J).numpy()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ric.npz"), **out)


if __name__ == "__main__":
    main()
