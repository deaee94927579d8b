"""motiondiffusion_moe_b200 — B200-native MoE motion denoiser + CFG sampler (drop-in for the hot path of
ltdoanh2004/MotionDiffusion-MoE).  Python host code over a C-ABI CUDA library (include/mdm_b200.h)."""
from ._lib import MdmError, load as load_library, LIB_PATH  # noqa: F401
from .transformer import MotionTransformer, TextContext  # noqa: F401
from .gaussian_diffusion import (GaussianDiffusion, CFGStepper, get_named_beta_schedule, ModelMeanType,  # noqa: F401
                                 ModelVarType, LossType)

from .trainer import DDPMTrainer  # noqa: F401
from .text_encoder import EnhancedTextEncoder  # noqa: F401
from .postprocess import load_reference_checkpoint, save_reference_checkpoint, recover_from_ric  # noqa: F401

__all__ = ["DDPMTrainer", "EnhancedTextEncoder", "load_reference_checkpoint", "save_reference_checkpoint", "recover_from_ric", "MotionTransformer", "TextContext", "GaussianDiffusion", "CFGStepper", "get_named_beta_schedule", "ModelMeanType",
           "ModelVarType", "LossType", "MdmError", "load_library"]
