"""The data formats either side of the hot path: reference checkpoints in, joint positions out.

* `load_reference_checkpoint` reads the trainer's `.tar` (trainers/ddpm_trainer.py:260-289: a torch.save'd dict
  {"encoder": state_dict, "opt_encoder": ..., "ep": ..., "total_it": ...}) into a MotionTransformer with the
  reference's `strict=False` semantics and returns (ep, total_it) like DDPMTrainer.load.
* `recover_from_ric` de-normalises generated features (tools/visualization.py:72-91) and recovers the joint
  positions (utils/motion_process.py:401-417) with one CUDA kernel (mdm_recover_from_ric); no CPU path.
"""
import torch

from . import ops
from ._lib import MdmError


def load_reference_checkpoint(model, path_or_state, map_location="cpu"):
    """DDPMTrainer.load (trainers/ddpm_trainer.py:277-289) for the denoiser: `checkpoint["encoder"]` is loaded with
    strict=False (the reference's call); DeBERTa weights (`text_encoder.*`) go to an attached text encoder if the
    model has one and are skipped otherwise.  Returns (ep, total_it, missing_keys, unexpected_keys)."""
    ckpt = path_or_state if isinstance(path_or_state, dict) else torch.load(path_or_state, map_location=map_location,
                                                                         weights_only=False)
    if "encoder" not in ckpt:
        raise KeyError("not a DDPMTrainer checkpoint: no 'encoder' entry (keys: %s)" % sorted(ckpt)[:6])
    state = {k[7:] if k.startswith("module.") else k: v for k, v in ckpt["encoder"].items()}
    r = model.load_state_dict(state, strict=False)
    return ckpt.get("ep", 0), ckpt.get("total_it", 0), list(r.missing_keys), list(r.unexpected_keys)


def save_reference_checkpoint(model, path, ep=0, total_it=0, opt_state=None):
    """DDPMTrainer.save (trainers/ddpm_trainer.py:260-275): same dict layout, loadable by the reference."""
    torch.save({"opt_encoder": opt_state if opt_state is not None else {}, "ep": ep, "total_it": total_it,
                "encoder": {k: v.detach().cpu() for k, v in model.state_dict().items()}}, path)


def recover_from_ric(data, joints_num, mean=None, std=None):
    """data [..., T, F] normalised (or raw when mean/std are None) features on a CUDA device ->
    [..., T, joints_num, 3] joint positions (utils/motion_process.py:401-417)."""
    if not data.is_cuda:
        raise MdmError("recover_from_ric needs a CUDA tensor: there is no CPU path")
    if (mean is None) != (std is None):
        raise ValueError("mean and std go together")
    lead, (T, F) = data.shape[:-2], data.shape[-2:]
    x = data.reshape(-1, T, F).float().contiguous()
    dev = x.device
    if mean is not None:
        mean = torch.as_tensor(mean, dtype=torch.float32).to(dev).contiguous()
        std = torch.as_tensor(std, dtype=torch.float32).to(dev).contiguous()
        if mean.numel() != F or std.numel() != F:
            raise ValueError("mean / std must have %d entries" % F)
    out = torch.empty(x.shape[0], T, joints_num, 3, dtype=torch.float32, device=dev)
    ops.recover_from_ric(x, joints_num, out, mean, std)
    return out.reshape(*lead, T, joints_num, 3)
