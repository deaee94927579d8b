"""bench.py --workload train: the DDPM training step of BASELINE.json configs[4] - forward + backward (incl. the MoE
load-balance loss value) + clip_grad_norm_(1.0) + Adam on KIT-shaped synthetic data (251 features x 196 frames), batch 256
per GPU, default MoE MotionTransformer (L8 D512 F1024 E8 top-2 H4 Dt256), bf16 operands with fp32 master weights,
dropout 0.  N > 1: data-parallel, the flat gradient buffer all-reduced (averaged) every step.

One JSON line from rank 0: metric = training frames/s (B * T * steps / time), whole job."""
import json
import os
import time

import torch

import bench as B_

CFG = dict(B_.CFG, input_feats=251)
BATCH, T = 256, 196
METRIC, UNIT = "training_frames_per_sec", "frames/s"
WORKLOAD = ("DDPM training step (q_sample, forward, masked noise-prediction loss + MoE balance loss, backward, clip_grad_norm_(1.0), "
            "Adam lr 2e-4) of the default MoE MotionTransformer on KIT-shaped synthetic data: 251 feats x 196 frames, batch 256 per GPU, "
            "bf16 operands / fp32 master weights, dropout 0")
FLOPS_PER_SEQ_FWD = 55.80e9           # BASELINE.md / SURVEY.md section 8(d): KIT-251 forward, 2MNK per matmul


def synth(batch, seed, device):
    g = torch.Generator().manual_seed(seed)
    x0 = torch.randn(batch, T, CFG["input_feats"], generator=g)
    length = torch.randint(40, T + 1, (batch,), generator=g)
    xf = torch.nn.functional.gelu(torch.randn(batch, 20, CFG["text_latent_dim"], generator=g))
    t = torch.randint(0, 1000, (batch,), generator=g)
    noise = torch.randn(batch, T, CFG["input_feats"], generator=g)
    return [v.to(device) for v in (x0, t, length, xf, noise)]


def cpu_train_leg(state, extras, batch=2, steps=1):
    """The reference's training-step evaluation on the host cores: autograd through the oracle port (pinned to the
    unmodified reference's gradients by tests/golden/train_tiny.npz), fp32, all threads, bounded sample."""
    from oracle import motion_oracle as mo
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = mo.Config(**CFG)
    p = {k: v.float().cpu() for k, v in state.items()}
    p.update({k: v.float().cpu() for k, v in extras.items()})
    names = [k for k in mo.param_shapes(cfg) if "expert_" not in k.rsplit(".", 1)[-1]]
    p = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in p.items()}
    x0, t, length, xf, noise = synth(batch, 77, "cpu")
    tab = mo.diffusion_tables(1000)
    times = []
    for i in range(steps + 1):
        t0 = time.perf_counter()
        x_t = mo.q_sample(tab, x0, t, noise)
        pred = mo.forward(p, cfg, x_t, t, length, xf.mean(1), xf)
        per = ((pred - noise) ** 2).mean(-1)
        mask = mo.src_mask(T, length).view(per.shape)
        ((per * mask).sum() / mask.sum()).backward()
        if i:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"value": batch * T / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "ms_per_step": sec * 1e3,
            "sample": "forward + autograd backward of the oracle port (fp32, no optimizer step), batch %d x 196 x 251, %d timed step(s) "
                      "after 1 warm-up, torch %s on the host cores" % (batch, steps, torch.__version__)}


def main(args, rank, world, local):
    import motiondiffusion_moe_b200 as mdm
    from motiondiffusion_moe_b200 import train_ops as T_, _lib
    from motiondiffusion_moe_b200.training import TrainEngine
    W = max(3, args.warmup)
    if args.impl == "reference":
        if rank != 0:
            return
        torch.manual_seed(0)
        net = mdm.MotionTransformer(precision="bf16", dropout=0.0, **CFG)
        B_.randomize_zero_init(net)
        cb = cpu_train_leg(net.state_dict(), net.extras_state(), batch=2, steps=max(1, min(args.steps, 2)))
        B_.emit({"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                 "steps": max(1, min(args.steps, 2)), "warmup": 1, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
                 "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                 "config": {"workload": WORKLOAD, "reference_sample": cb["sample"]}, "cpu_baseline": cb,
                 "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    torch.manual_seed(0)
    net = mdm.MotionTransformer(precision="bf16", dropout=0.0, **CFG)
    B_.randomize_zero_init(net)
    state_cpu, extras_cpu = net.state_dict(), net.extras_state()
    net.to(dev)
    import types
    opt = types.SimpleNamespace(device=dev, diffusion_steps=1000, is_train=True, lr=2e-4)
    xf_holder = {}
    net.encode_text = lambda text, device: (xf_holder["xf"].mean(1), xf_holder["xf"])
    tr = mdm.DDPMTrainer(opt, net)
    eng = tr.engine
    x0, t, length, xf, noise = synth(BATCH, 2000 + rank, dev)
    xf_holder["xf"] = xf
    d = tr.diffusion

    def step():
        eng.zero_grad()
        x_t = d.q_sample(x0, t, noise=noise)
        net.reset_all_moe_counters(net)
        pred, S = eng.forward_train(x_t, t, length, xf.mean(1), xf)
        d_pred = T_.masked_mse_grad(pred, noise, length.clamp(max=T).contiguous())
        eng.backward(S, d_pred, grad_ready=tr._bucket_all_reduce if world > 1 else None)    # buckets all-reduced during the backward
        tr._finish_all_reduce()
        eng.optimizer_step()

    for _ in range(W):
        step()
    clocks = B_.ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    n0 = _lib.LAUNCHES[0]
    clocks.start()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    launches = (_lib.LAUNCHES[0] - n0) // max(1, args.steps)
    finite = bool(torch.isfinite(eng.flat).all())
    mem = torch.cuda.max_memory_allocated(dev) / 2 ** 30

    # ---- end to end through the public trainer API: host motions in, loss value out, every step
    caps = ["a person walks forward"] * BATCH
    motions_h = x0.cpu().pin_memory()
    lens = length.cpu().tolist()
    e2e_steps = max(2, min(args.steps, 5))
    tr.forward((caps, motions_h, lens)); tr.update()
    barrier()
    ev0.record()
    for _ in range(e2e_steps):
        tr.forward((caps, motions_h, lens))
        logs = tr.update()                     # reads the loss back (.item()): D2H every step
    ev1.record()
    barrier()
    ms_e2e = ev0.elapsed_time(ev1)

    if world > 1:
        tt = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(tt[0]), float(tt[1])
    pk, src = B_.peaks()
    per_step = ms / args.steps
    value = world * BATCH * T * args.steps / (ms * 1e-3)
    flops = 3 * FLOPS_PER_SEQ_FWD * BATCH                       # forward + ~2x backward
    ach = flops / (per_step * 1e-3) / 1e12
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * BATCH, "frames": T,
                       "parallelism": "dp%d (per-layer gradient buckets of the flat buffer all-reduced (averaged) on a side stream while the backward runs)" % world,
                       "l2_policy": "working set (2.1 GB fp32 masters + 1 GB bf16 mirror + tens of GB of kept activations) exceeds the 126 MB L2",
                       "timed_region": "zero_grad + q_sample + forward (activations kept) + backward + all-reduce + clip + Adam + bf16 refresh, eager launches"},
            "roofline": {"bound": "tensor", "kernel": "whole training step (99% of the FLOPs are GEMM-shaped)", "achieved": ach,
                         "peak": pk.get("bf16_tflops_sustained", pk["bf16_tflops"]), "unit": "TFLOP/s",
                         "frac": ach / pk.get("bf16_tflops_sustained", pk["bf16_tflops"]), "peak_source": src + " sustained (kernels timed inside a long step)",
                         "traffic": None, "flops_per_step": flops},
            "e2e": {"value": world * BATCH * T * e2e_steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": motions_h.numel() * 4,
                    "d2h_bytes_per_step": 12, "steps": e2e_steps,
                    "path": "DDPMTrainer.forward((captions, pinned host motions, lengths)) + DDPMTrainer.update() -> loss logs (.item())"},
            "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches, "clocks": clk, "finite": finite,
            "max_memory_gib": mem, "loss_total": logs["loss_total"], "grad_norm": float(eng.norm_coef[0])}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_train_leg(state_cpu, extras_cpu, batch=2, steps=1)
        B_.emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
