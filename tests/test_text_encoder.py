"""EnhancedTextEncoder (reference models/text_encoder.py:6-43): module surface on CPU, the reference's own
part of the forward (prompt tokens, LayerNorm, Linear, GELU, mean pooling) on the library's kernels against torch on GPU,
with a Hugging Face DebertaV2Model (random weights, built from a config: no network) as the injected backbone."""
import os
import types

import pytest
import torch

import motiondiffusion_moe_b200 as mdm


class StubTokenizer:
    """What the reference asks of its tokenizer: padding=True, truncation, max_length=77, return_tensors='pt'."""
    def __call__(self, text, padding=True, truncation=True, max_length=77, return_tensors="pt"):
        ids = [[1] + [3 + (ord(c) % 90) for c in t][:max_length - 2] + [2] for t in text]
        L = max(len(i) for i in ids)
        input_ids = torch.tensor([i + [0] * (L - len(i)) for i in ids], dtype=torch.long)
        mask = torch.tensor([[1] * len(i) + [0] * (L - len(i)) for i in ids], dtype=torch.long)
        out = types.SimpleNamespace(input_ids=input_ids, attention_mask=mask)
        out.to = lambda device: types.SimpleNamespace(input_ids=input_ids.to(device), attention_mask=mask.to(device))
        return out


class StubBackbone(torch.nn.Module):
    """Any module with .config.hidden_size returning .last_hidden_state works as the injected backbone."""
    def __init__(self, hidden):
        super().__init__()
        self.config = types.SimpleNamespace(hidden_size=hidden)
        self.emb = torch.nn.Embedding(128, hidden)

    def forward(self, input_ids, attention_mask, return_dict=True):
        return types.SimpleNamespace(last_hidden_state=self.emb(input_ids) * 3.0)


def reference_forward(enc, text, device):
    """models/text_encoder.py:22-43 with the module's own `proj` Sequential (eval mode) as plain torch ops."""
    inputs = enc.tokenize(text, device)
    hs = enc.bert(input_ids=inputs.input_ids, attention_mask=inputs.attention_mask, return_dict=True).last_hidden_state
    hidden = torch.cat([enc.prompt_tokens.repeat(len(text), 1, 1).to(device), hs], dim=1)
    projected = enc.proj(hidden)
    return projected.mean(dim=1), projected


def test_state_dict_keys_and_loud_failures(monkeypatch):
    enc = mdm.EnhancedTextEncoder(output_dim=256, bert=StubBackbone(1024), tokenizer=StubTokenizer())
    keys = set(enc.state_dict().keys())
    # the reference's keys: prompt_tokens, proj.0 (LayerNorm), proj.1 (Linear), bert.*
    assert {"prompt_tokens", "proj.0.weight", "proj.0.bias", "proj.1.weight", "proj.1.bias"} <= keys
    assert all(k.startswith("bert.") for k in keys - {"prompt_tokens", "proj.0.weight", "proj.0.bias", "proj.1.weight", "proj.1.bias"})
    assert enc.prompt_tokens.shape == (1, 8, 1024) and enc.model_name == "microsoft/deberta-v3-large"
    enc.eval()
    with pytest.raises(mdm.MdmError):           # no CPU path
        enc(["a person walks"], "cpu")
    enc.train()
    with pytest.raises(mdm.MdmError):           # the dropout of train() mode is not built: refuse, do not skip it silently
        enc.project(torch.zeros(1, 3, 1024))
    # without an injected backbone the reference's from_pretrained call is attempted; offline it must fail loudly
    monkeypatch.setenv("HF_HUB_OFFLINE", "1")
    monkeypatch.setenv("TRANSFORMERS_OFFLINE", "1")
    with pytest.raises(mdm.MdmError):
        mdm.EnhancedTextEncoder(output_dim=256)
    # attached to the model, encode_text goes through it and its parameters are part of the model's state_dict
    net = mdm.MotionTransformer(input_feats=12, num_frames=8, latent_dim=128, ff_size=128, num_layers=1, num_heads=2,
                                text_latent_dim=128, moe_num_experts=2, text_encoder=enc, precision="fp32")
    assert any(k.startswith("text_encoder.proj.1.") for k in net.state_dict())


@pytest.mark.gpu
@pytest.mark.parametrize("hidden,out_dim", [(1024, 256), (128, 512)])
def test_projection_head_matches_torch(hidden, out_dim):
    dev = torch.device("cuda")
    torch.manual_seed(hidden)
    enc = mdm.EnhancedTextEncoder(output_dim=out_dim, bert=StubBackbone(hidden), tokenizer=StubTokenizer()).to(dev).eval()
    with torch.no_grad():
        enc.proj[0].weight.uniform_(0.5, 1.5)
        enc.proj[0].bias.normal_(0, 0.1)
    text = ["a person walks forward", "jump", "someone waves the left hand and then sits down slowly on a chair"]
    pooled, projected = enc(text, dev)
    with torch.no_grad():
        rp, rj = reference_forward(enc, text, dev)
    assert projected.shape == rj.shape and pooled.shape == rp.shape
    assert ((projected - rj).norm() / rj.norm()).item() < 1e-5
    assert ((pooled - rp).norm() / rp.norm()).item() < 1e-5


@pytest.mark.gpu
def test_with_a_huggingface_deberta_v2_backbone_and_the_model():
    """The injected backbone is the class the reference loads (DebertaV2Model: disentangled attention, relative position
    buckets), here with random weights from a small config; the encoder then feeds MotionTransformer.forward."""
    transformers = pytest.importorskip("transformers")
    dev = torch.device("cuda")
    torch.manual_seed(0)
    cfg = transformers.DebertaV2Config(vocab_size=128, hidden_size=128, num_hidden_layers=2, num_attention_heads=4,
                                       intermediate_size=256, relative_attention=True, position_buckets=16,
                                       norm_rel_ebd="layer_norm", share_att_key=True, pos_att_type=["p2c", "c2p"],
                                       max_position_embeddings=128)
    bert = transformers.DebertaV2Model(cfg).eval()
    enc = mdm.EnhancedTextEncoder(output_dim=128, bert=bert, tokenizer=StubTokenizer()).to(dev).eval()
    text = ["a person walks", "a person jumps over an obstacle"]
    pooled, projected = enc(text, dev)
    with torch.no_grad():
        rp, rj = reference_forward(enc, text, dev)
    assert ((projected - rj).norm() / rj.norm()).item() < 1e-5 and ((pooled - rp).norm() / rp.norm()).item() < 1e-5
    net = mdm.MotionTransformer(input_feats=12, num_frames=8, latent_dim=128, ff_size=128, num_layers=1, num_heads=2,
                                text_latent_dim=128, moe_num_experts=2, text_encoder=enc, precision="fp32").to(dev)
    x = torch.randn(2, 8, 12, device=dev)
    out = net(x, torch.tensor([10, 500], device=dev), torch.tensor([8, 6], device=dev), text=text)
    assert out.shape == x.shape and torch.isfinite(out).all()
    xf_proj, xf_out = net.encode_text(text, dev)
    assert torch.equal(out, net(x, torch.tensor([10, 500], device=dev), torch.tensor([8, 6], device=dev), None, xf_proj, xf_out))
