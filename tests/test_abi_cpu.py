"""CPU: the C-ABI library builds/loads and exports every symbol include/mdm_b200.h declares; host-side
logic that needs no GPU (state-dict naming, RNG replay, argument errors, sampler tables)."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from motiondiffusion_moe_b200 import build, _lib
    build.build()
    hdr = open(os.path.join(ROOT, "include", "mdm_b200.h")).read()
    declared = set(re.findall(r"MDM_API\s+[\w\s\*]+?\b(mdm_\w+)\s*\(", hdr))
    assert len(declared) >= 18
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.exported_symbols())
    assert b"sm_100a" in lib.mdm_version()


def test_struct_mirrors_match_the_library_and_the_docs():
    """a binding that mirrors MdmGemmEpi without `tile_k` passes a short struct and the kernel reads
    garbage as a device pointer.  The library exports the sizeof of every struct of the C-ABI; _lib.py checks them at
    load time, and the stub printed in INTEGRATION.md must declare the same fields as _lib.GemmEpi."""
    import ctypes as C
    from motiondiffusion_moe_b200 import _lib
    lib = _lib.load()
    assert lib.mdm_sizeof_gemm_epi() == C.sizeof(_lib.GemmEpi)
    assert lib.mdm_sizeof_rowop() == C.sizeof(_lib.RowOp)
    assert lib.mdm_sizeof_ep_peers() == C.sizeof(_lib.EpPeers)
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    stub = doc[doc.index("class MdmGemmEpi(C.Structure)"):]
    stub = stub[:stub.index("lib.mdm_gemm_bf16.restype")]
    doc_fields = re.findall(r'\("(\w+)", C\.c_\w+\)', stub)
    assert doc_fields == [f[0] for f in _lib.GemmEpi._fields_]
    hdr = open(os.path.join(ROOT, "include", "mdm_b200.h")).read()
    body = hdr[hdr.index("typedef struct MdmGemmEpi {"):hdr.index("} MdmGemmEpi;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    hdr_fields = [n for decl in re.findall(r"[\w\* ]+?([\w ,\*]+);", body) for n in re.findall(r"(\w+)\s*(?:,|$)", decl)]
    assert hdr_fields == doc_fields, (hdr_fields, doc_fields)


def test_sass_is_blackwell_native():
    """tcgen05 / TMA must be in the shipped cubin (UTCHMMA, UTMALDG, LDTM) — no legacy-only build."""
    import shutil
    import subprocess
    from motiondiffusion_moe_b200 import _lib
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
    assert "sm_100a" in sass


def test_state_dict_keys_and_reference_helpers():
    import motiondiffusion_moe_b200 as mdm
    from oracle import cases, motion_oracle as mo
    cfg, p = cases.case_params("tiny_b3")
    net = mdm.MotionTransformer(**cfg)
    assert set(net.state_dict()) == set(mo.param_shapes(cfg))
    for k, shape in mo.param_shapes(cfg).items():
        assert tuple(net.state_dict()[k].shape) == tuple(shape), k
    # zero-init rules of the reference (H4) and reference helper API
    sd = net.state_dict()
    assert sd["out.weight"].abs().sum() == 0
    assert sd["decoder_blocks_low.0.module.ffn.branches.0.moe.gate.weight"].abs().sum() == 0
    assert sd["decoder_blocks_low.0.module.cross_attn.base_ca.proj_out.out_layers.2.weight"].abs().sum() == 0
    assert sd["decoder_blocks_low.0.module.dual_self_attn.local_attn.style_block.out_layers.2.weight"].abs().sum() > 0
    m = net.generate_src_mask(6, torch.tensor([6, 2, 0]))
    assert m.tolist() == [[1] * 6, [1, 1, 0, 0, 0, 0], [0] * 6]
    assert float(net.get_moe_loss(net)) == pytest.approx(2 * cfg.num_layers * 2 * cfg.moe_num_experts)
    # big model doubles the widths (transformer.py:188-192)
    big = mdm.MotionTransformer(input_feats=12, num_frames=8, latent_dim=64, ff_size=128, num_layers=1, num_heads=4,
                                text_latent_dim=64, moe_num_experts=4, model_size="big")
    assert big.latent_dim == 128 and big.ff_size == 256 and big.text_latent_dim == 128
    # ephemerals replay the reference RNG stream
    net.redraw_ephemerals(cases.EPH_SEED)
    for k, v in mo.draw_ephemerals(cfg, cases.EPH_SEED).items():
        assert torch.equal(net.extras_state()[k], v), k


def test_no_cpu_fallback():
    import motiondiffusion_moe_b200 as mdm
    from oracle import cases
    cfg, p = cases.case_params("tiny_b3")
    net = mdm.MotionTransformer(**cfg)
    x, t, length, xf_proj, xf_out = cases.make_inputs(cfg, 2, 8, seed=1)
    with pytest.raises(mdm.MdmError):
        net(x, t, length, None, xf_proj, xf_out)


def test_diffusion_tables_match_oracle():
    import motiondiffusion_moe_b200 as mdm
    from oracle import motion_oracle as mo
    d = mdm.GaussianDiffusion(betas=mdm.get_named_beta_schedule("linear", 1000))
    tab = mo.diffusion_tables(1000)
    for k in ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1",
              "posterior_mean_coef2", "posterior_log_variance_clipped", "sqrt_alphas_cumprod",
              "sqrt_one_minus_alphas_cumprod"):
        assert np.array_equal(getattr(d, k), tab[k]), k
    assert d.num_timesteps == 1000
    with pytest.raises(NotImplementedError):
        mdm.get_named_beta_schedule("nope", 10)
    d2 = mdm.GaussianDiffusion(betas=mdm.get_named_beta_schedule("linear", 1000),
                               model_mean_type=mdm.ModelMeanType.START_X)
    with pytest.raises(NotImplementedError):
        d2._check_supported()


def test_reference_checkpoint_roundtrip(tmp_path):
    """DDPMTrainer.save / load layout (trainers/ddpm_trainer.py:260-289): {"encoder", "opt_encoder", "ep",
    "total_it"}; DDP "module." prefixes and DeBERTa "text_encoder.*" entries of a real checkpoint are tolerated
    (the reference loads with strict=False)."""
    import torch
    import motiondiffusion_moe_b200 as mdm
    from oracle import cases
    cfg, p = cases.case_params("tiny_b3")
    a = mdm.MotionTransformer(precision="fp32", **cfg)
    a.load_state_dict({k: p[k] for k in a.state_dict()})
    path = str(tmp_path / "ckpt_e003.tar")
    mdm.save_reference_checkpoint(a, path, ep=3, total_it=1234)
    raw = torch.load(path, weights_only=False)
    assert set(raw) == {"opt_encoder", "ep", "total_it", "encoder"}
    raw["encoder"] = {"module." + k: v for k, v in raw["encoder"].items()}
    raw["encoder"]["module.text_encoder.model.embeddings.word_embeddings.weight"] = torch.zeros(4, 4)
    torch.save(raw, path)
    b = mdm.MotionTransformer(precision="fp32", **cfg)
    ep, it, missing, unexpected = mdm.load_reference_checkpoint(b, path)
    assert (ep, it) == (3, 1234) and not missing and not unexpected
    sa, sb = a.state_dict(), b.state_dict()
    assert set(sa) == set(sb) and all(torch.equal(sa[k], sb[k]) for k in sa)
    import pytest
    with pytest.raises(KeyError):
        mdm.load_reference_checkpoint(b, {"model": {}})
    with pytest.raises(mdm.MdmError):
        mdm.recover_from_ric(torch.zeros(1, 4, 263), 22)       # CPU tensor: no CPU path


def test_ddim_schedule_and_trainer_host_logic():
    """Host-side logic that needs no GPU: the strided DDIM schedule, the diffusion tables against the oracle's, and the
    training half of DDPMTrainer refusing a CPU model loudly."""
    import types
    import motiondiffusion_moe_b200 as mdm
    from oracle import motion_oracle as mo
    d = mdm.GaussianDiffusion(betas=mdm.get_named_beta_schedule("linear", 1000))
    for S in (1, 2, 7, 50, 333, 1000):
        order, prev = d.ddim_timesteps(S)
        assert order == sorted(set(order), reverse=True) and order[-1] == 0 and 0 <= max(order) < 1000
        assert len(order) <= S and (S > 500 or len(order) == S)
        assert prev == order[1:] + [-1]
    assert d.ddim_timesteps(1000)[0] == list(range(999, -1, -1))          # the reference's full-length loop
    with pytest.raises(ValueError):
        d.ddim_timesteps(0)
    tab = mo.diffusion_tables(1000)
    for k in ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2",
              "posterior_log_variance_clipped", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "alphas_cumprod"):
        assert np.array_equal(getattr(d, k), tab[k]), k
    cfg = dict(input_feats=12, num_frames=8, latent_dim=128, ff_size=256, num_layers=1, num_heads=4, text_latent_dim=128,
               moe_num_experts=4)
    net = mdm.MotionTransformer(**cfg)
    with pytest.raises(mdm.MdmError):           # the training engine has no CPU path: a CPU model is refused loudly
        mdm.DDPMTrainer(types.SimpleNamespace(device="cpu", diffusion_steps=1000, is_train=True), net)
    tr = mdm.DDPMTrainer(types.SimpleNamespace(device="cpu", diffusion_steps=100, is_train=False), net, sampler="ddim")
    assert tr.diffusion.num_timesteps == 100 and tr.cfg_scale == 7.5
    with pytest.raises(RuntimeError):           # a sampling-only trainer (is_train=False) has no optimizer state
        tr.update()
    with pytest.raises(ValueError):
        mdm.DDPMTrainer(types.SimpleNamespace(device="cpu", diffusion_steps=100, is_train=False), net, sampler="euler")


def test_packed_attention_operands_for_head_size_64():
    """Host-side packing of the operands that select the tcgen05 attention kernels at head size 64 (two heads per CTA):
    the block-diagonal diag(P^T, P^T) of FastAttention and the block-diagonal ctx^T of each pair of heads of the linear
    cross-attention (pure torch: no GPU needed)."""
    from motiondiffusion_moe_b200 import ops
    g = torch.Generator().manual_seed(0)
    P = torch.randn(64, 64, generator=g)
    Pt = ops.pack_fastattn_pt(P)
    assert Pt.dtype == torch.bfloat16 and tuple(Pt.shape) == (128, 128)
    ref = P.t().contiguous().to(torch.bfloat16)
    assert torch.equal(Pt[:64, :64], ref) and torch.equal(Pt[64:, 64:], ref)
    assert float(Pt[:64, 64:].abs().max()) == 0.0 and float(Pt[64:, :64].abs().max()) == 0.0
    P128 = torch.randn(128, 128, generator=g)
    assert torch.equal(ops.pack_fastattn_pt(P128), P128.t().contiguous().to(torch.bfloat16))
    assert ops.pack_fastattn_pt(torch.randn(32, 32, generator=g)) is None
    ctx = torch.randn(3, 6, 64, 64, generator=g)                       # [B, H, d, l]
    cT = ops.pack_lincross_ctxT(ctx)
    assert tuple(cT.shape) == (3, 3, 128, 128) and cT.dtype == torch.bfloat16
    for hp in range(3):
        assert torch.equal(cT[:, hp, :64, :64], ctx[:, 2 * hp].transpose(-1, -2).to(torch.bfloat16))
        assert torch.equal(cT[:, hp, 64:, 64:], ctx[:, 2 * hp + 1].transpose(-1, -2).to(torch.bfloat16))
    assert float(cT[:, :, :64, 64:].abs().max()) == 0.0 and float(cT[:, :, 64:, :64].abs().max()) == 0.0
    assert ops.pack_lincross_ctxT(torch.randn(2, 3, 64, 64, generator=g)) is None       # odd head count: mma.sync kernels
    assert ops.pack_lincross_ctxT(torch.randn(2, 2, 32, 32, generator=g)) is None
