"""torchrun worker: data-parallel DDPM training (BASELINE.json configs[4] at N > 1): every rank
trains on its own batch, DDPMTrainer.update all-reduces (averages) the flat gradient buffer before clip + Adam.  Checks:
the averaged gradient equals the mean of the ranks' local gradients, and after three updates every rank holds bit-identical
parameters (the replicas never drift).  Prints DP_TRAIN_OK."""
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from dist_common import init_dist, all_max  # noqa: E402
import motiondiffusion_moe_b200 as mdm  # noqa: E402
from oracle import cases, motion_oracle as mo  # noqa: E402


def main():
    rank, world, dev, shared = init_dist()
    cfg, p = cases.case_params("tiny_b3")
    net = mdm.MotionTransformer(precision="bf16", dropout=0.0, **cfg)
    net.load_state_dict({k: p[k] for k in net.state_dict()})
    net.load_extras(p)
    net.encode_text = lambda text, device: mo.stub_text(text, cfg.text_latent_dim, device)
    tr = mdm.DDPMTrainer(types.SimpleNamespace(device=dev, diffusion_steps=1000, is_train=True, lr=1e-3), net)
    eng = tr.engine
    g = torch.Generator().manual_seed(100 + rank)                       # every rank: its own data
    motions = torch.randn(3, 8, cfg.input_feats, generator=g)
    caps = ["a person walks", "a person jumps", "a person sits"]
    ok = True
    # (1) the all-reduced gradient is the mean of the local gradients
    np.random.seed(5 + rank)
    torch.manual_seed(5 + rank)
    tr.forward((caps, motions, [8, 6, 4]))
    from motiondiffusion_moe_b200 import train_ops as T
    eng.zero_grad()
    tr.backward_G()
    eng.backward(tr._saved, T.masked_mse_grad(tr.fake_noise.contiguous(), tr.real_noise.contiguous(), tr.cur_len.contiguous()))
    local = eng.grad.clone()
    tr._all_reduce_gradients()
    gather_dev = "cpu" if shared else dev                     # gloo (shared GPU) gathers host tensors, NCCL device tensors
    mine = local.to(gather_dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    want = torch.stack(parts).sum(0) / world
    ok &= bool(torch.allclose(eng.grad.to(gather_dev), want, rtol=1e-5, atol=1e-8))
    ok &= not torch.equal(local, eng.grad)
    # (1b) the bucketed path of update(): every layer's bucket all-reduced on a side stream while the backward runs
    eng.zero_grad()
    eng.backward(tr._saved, T.masked_mse_grad(tr.fake_noise.contiguous(), tr.real_noise.contiguous(), tr.cur_len.contiguous()),
                 grad_ready=tr._bucket_all_reduce)
    tr._finish_all_reduce()
    torch.cuda.synchronize()
    ok_b = bool(torch.allclose(eng.grad.to(gather_dev), want, rtol=1e-5, atol=1e-8))
    if not ok_b and rank == 0:
        print("bucketed all-reduce differs from the single all-reduce: max abs %.3e" % (eng.grad.to(gather_dev) - want).abs().max().item(), flush=True)
    ok &= ok_b
    # (2) replicas stay bit-identical through updates
    for it in range(3):
        np.random.seed(10 * it + rank)
        torch.manual_seed(10 * it + rank)
        tr.forward((caps, motions, [8, 6, 4]))
        logs = tr.update()
        ok &= bool(np.isfinite(logs["loss_total"]))
    flat = eng.flat.to(gather_dev)
    parts = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(parts, flat)
    ok &= all(torch.equal(parts[0], q) for q in parts[1:])
    flag = all_max([0.0 if ok else 1.0], dev, shared)[0]
    if rank == 0:
        print("DP_TRAIN_OK" if flag == 0.0 else "DP_TRAIN_MISMATCH")
    dist.destroy_process_group()
    sys.exit(0 if flag == 0.0 else 1)


if __name__ == "__main__":
    main()
