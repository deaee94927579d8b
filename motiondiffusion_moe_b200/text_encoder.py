"""EnhancedTextEncoder (reference: models/text_encoder.py:6-43), SURVEY.md section 8 row (f)-2: the producer of the
`xf_proj` / `xf_out` inputs of MotionTransformer.forward.

    tokens -> DeBERTa-v3 backbone -> last_hidden_state [B, L, H]
    hidden = cat(prompt_tokens [1, 8, H] repeated, last_hidden_state)           (text_encoder.py:24-37)
    projected = GELU(Dropout(Linear(LayerNorm(hidden))))  [B, 8 + L, Dt]          (:14-19, :38)
    pooled = mean(projected, dim=1)                                               (:40)

What is built here: the module surface (same constructor arguments, same state_dict keys: `prompt_tokens`, `proj.0.*`
(LayerNorm), `proj.1.*` (Linear), `bert.*`), and the part of the computation that is the reference's OWN code - prompt
concatenation, LayerNorm, Linear + GELU, mean pooling - on the kernels of libmdm_b200.so (mdm_rowop, mdm_gemm_f32 with the
exact-erf GELU, mdm_colsum, mdm_axpby) in fp32: it runs once per sampling loop (MotionTransformer.prepare_text hoists
everything text-side out of the denoising loop), so precision, not speed, is what matters.

What is NOT rebuilt: the backbone itself.  `AutoModel.from_pretrained("microsoft/deberta-v3-large")` is a third-party
model (Hugging Face transformers, DebertaV2Model: 24 layers, hidden 1024, disentangled attention) whose weights and
sentencepiece vocabulary are downloaded at construction time - neither is part of the reference repository, and there is
no network here.  The backbone and the tokenizer are therefore INJECTED (`bert=`, `tokenizer=`): any module returning an
object with `.last_hidden_state` [B, L, H] for (`input_ids`, `attention_mask`), e.g. the Hugging Face model a user has
on disk.  With neither given, construction tries the reference's own `from_pretrained` call and raises MdmError if the
files are not available - never a silent substitute.

Dropout (p = 0.1 between the Linear and the GELU) is the identity in eval(); in train() mode the reference draws a
mask there - this head is inference-only and refuses train() with p > 0 (the DDPM trainer of this package takes the
text features as inputs).
"""
from typing import List, Optional

import torch
import torch.nn as nn

from . import ops
from ._lib import ACT_GELU, MDM_F32, MdmError

MODEL_NAME = "microsoft/deberta-v3-large"      # text_encoder.py:9


class EnhancedTextEncoder(nn.Module):
    def __init__(self, output_dim: int, dropout: float = 0.1, *, bert: Optional[nn.Module] = None, tokenizer=None,
                 hidden_size: Optional[int] = None):
        super().__init__()
        self.model_name = MODEL_NAME
        if bert is None or tokenizer is None:
            try:                                                      # the reference's own construction (:10-11)
                from transformers import AutoModel, AutoTokenizer
                bert = bert if bert is not None else AutoModel.from_pretrained(self.model_name)
                tokenizer = tokenizer if tokenizer is not None else AutoTokenizer.from_pretrained(self.model_name)
            except Exception as e:                                    # no network / no cached files
                raise MdmError("EnhancedTextEncoder needs the DeBERTa-v3 backbone and tokenizer (%s): pass bert= and "
                               "tokenizer= (e.g. loaded from a local directory); there is no substitute encoder. (%s)"
                               % (self.model_name, type(e).__name__))
        self.bert = bert
        self.tokenizer = tokenizer
        if hidden_size is None:
            hidden_size = bert.config.hidden_size
        self.hidden_size = hidden_size
        self.output_dim = output_dim
        # parameter holders with the reference's state_dict keys; the Sequential itself is never called
        self.proj = nn.Sequential(nn.LayerNorm(hidden_size), nn.Linear(hidden_size, output_dim), nn.Dropout(dropout),
                                  nn.GELU())
        self.num_prompt_tokens = 8
        self.prompt_tokens = nn.Parameter(torch.randn(1, self.num_prompt_tokens, hidden_size))

    def tokenize(self, text: List[str], device):
        return self.tokenizer(text, padding=True, truncation=True, max_length=77, return_tensors="pt").to(device)

    @torch.no_grad()
    def project(self, last_hidden_state: torch.Tensor):
        """The reference's own part of the forward (text_encoder.py:36-41) on libmdm_b200.so: [B, L, H] -> (pooled
        [B, Dt], projected [B, 8 + L, Dt]), fp32."""
        if self.training and self.proj[2].p > 0:
            raise MdmError("EnhancedTextEncoder is inference-only here (the reference draws a dropout mask between its "
                           "Linear and its GELU in train() mode): keep it frozen with text_encoder.eval(), or construct "
                           "it with dropout=0.0")
        hs = last_hidden_state
        if not hs.is_cuda:
            raise MdmError("EnhancedTextEncoder.project needs CUDA tensors: there is no CPU fallback")
        B, Lt, H = hs.shape
        if H != self.hidden_size:
            raise MdmError("backbone hidden size %d != %d" % (H, self.hidden_size))
        dev, f32 = hs.device, torch.float32
        with torch.cuda.device(dev):
            P, Nt, Dt = self.num_prompt_tokens, self.num_prompt_tokens + Lt, self.output_dim
            hidden = torch.empty(B, Nt, H, dtype=f32, device=dev)
            hidden[:, :P] = self.prompt_tokens.detach().to(device=dev, dtype=f32)      # prompts.repeat + cat (:27, :37)
            hidden[:, P:] = hs.to(f32)
            rows = B * Nt
            ln = self.proj[0]
            normed = torch.empty(rows, H, dtype=f32, device=dev)
            ops.rowop(hidden.view(rows, H), rows, H, MDM_F32,
                      ln1=(ln.weight.detach().float().contiguous(), ln.bias.detach().float().contiguous()), out1_f32=normed)
            lin = self.proj[1]
            projected = torch.empty(B, Nt, Dt, dtype=f32, device=dev)
            ops.gemm(normed, lin.weight.detach().float().contiguous(), lin.bias.detach().float().contiguous(), act=ACT_GELU,
                     out_f32=projected.view(rows, Dt))
            # mean over ALL positions, padding included, like torch.mean(projected, dim=1) (:40)
            sums = torch.zeros(B, Dt, dtype=f32, device=dev)
            from . import train_ops
            for b in range(B):
                train_ops.colsum_into(projected[b], Nt, Dt, sums[b], slabs=1)
            pooled = torch.empty(B, Dt, dtype=f32, device=dev)
            train_ops.axpby(sums, 1.0 / Nt, None, 0.0, pooled)
        return pooled, projected

    @torch.no_grad()
    def forward(self, text: List[str], device):
        inputs = self.tokenize(text, device)
        outputs = self.bert(input_ids=inputs.input_ids, attention_mask=inputs.attention_mask, return_dict=True)
        return self.project(outputs.last_hidden_state)
