#!/usr/bin/env python
"""bench.py — denoising frames/s of the default MoE MotionTransformer under bf16 CFG sampling.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): L8 / D512 / F1024 / 8 experts top-2 / 4 heads / Dt256, 196 frames x 263
features, batch 64 per GPU, classifier-free guidance (conditional + unconditional = 128 sequences per
forward), synthetic inputs and random-init weights (zero-initialised tensors re-randomised).
A "step" is one full CFG reverse-diffusion step: batched forward + guided DDPM update.
Scaling is weak: every GPU denoises its own batch of 64 (independent sequences, no data-path collective).

One JSON line is printed by rank 0; see README/DESIGN.md for the meaning of every key.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(input_feats=263, num_frames=196, latent_dim=512, ff_size=1024, num_layers=8, num_heads=4,
           text_latent_dim=256, moe_num_experts=8)
B_PER_GPU, T, CFG_SCALE = 64, 196, 7.5
METRIC, UNIT = "denoising_frames_per_sec", "frames/s"
WORKLOAD = ("default MoE MotionTransformer (L8 D512 F1024 E8 top-2 H4 Dt256), bf16 CFG sampling, batch 64 per GPU, "
            "196 frames x 263 feats, cond+uncond batched as 128 sequences per forward")


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def randomize_zero_init(model, seed=7):
    """gates, zero_module'd Linears and the final `out` start at exactly 0, which would make
    the benchmark degenerate (all-tie routing, output == 0): re-randomise them N(0, 0.05^2)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.abs().sum() == 0 and not n.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    model.repack()


def synth_inputs(B, seed, device):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, CFG["input_feats"], generator=g)
    length = torch.randint(40, T + 1, (B,), generator=g)
    xf_c = torch.nn.functional.gelu(torch.randn(B, 20, CFG["text_latent_dim"], generator=g))
    xf_u = torch.nn.functional.gelu(torch.randn(1, 10, CFG["text_latent_dim"], generator=g)).expand(B, -1, -1)
    return x, length.to(device), xf_c.to(device), xf_u.contiguous().to(device)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for n, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_leg(state, extras, steps, warmup, batch=8):
    """The reference's own CPU path for this metric: the oracle port (oracle/motion_oracle.py, pinned to the
    unmodified reference by tests/golden) timed on the host cores, fp32, on a bounded sample: CFG steps of the
    default model at batch `batch` (two sequential forwards + DDPM update, as the reference does)."""
    from oracle import motion_oracle as mo
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = mo.CONFIGS["default"]
    p = {k: v.float().cpu() for k, v in state.items()}
    p.update({k: v.float().cpu() for k, v in extras.items()})
    x, length, xf_c, xf_u = synth_inputs(batch, 123, "cpu")
    tab = mo.diffusion_tables(1000)
    t = torch.full((batch,), 500, dtype=torch.long)
    noise = torch.randn(x.shape, generator=torch.Generator().manual_seed(1))
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            x, _ = mo.cfg_step(p, cfg, tab, x, t, length, (xf_c.mean(1), xf_c), (xf_u.mean(1), xf_u), noise, CFG_SCALE)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"value": batch * T / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "default model fp32, batch %d x 196 frames, %d CFG steps (2 forwards + update each) after %d "
                      "warm-up, oracle port of the reference on host cores" % (batch, steps, warmup),
            "ms_per_step": sec * 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--expert-parallel", action="store_true",
                    help="N > 1 only: shard the experts over the ranks (NVLink dispatch/combine) instead of replicating them")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(3, args.warmup)

    import motiondiffusion_moe_b200 as mdm

    if args.impl == "reference":
        if rank != 0:
            return
        torch.manual_seed(0)
        net = mdm.MotionTransformer(precision="bf16", **CFG)
        randomize_zero_init(net)
        k = max(1, min(args.steps, 3))
        cb = cpu_reference_leg(net.state_dict(), net.extras_state(), steps=k, warmup=1)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT,
                          "n_gpus": args.gpus, "steps": k, "warmup": 1, "ms_per_step": cb["ms_per_step"],
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": {"workload": WORKLOAD, "reference_sample": cb["sample"]},
                          "cpu_baseline": cb,
                          "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                                  "d2h_bytes_per_step": 0}}))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    torch.manual_seed(0)                       # identical weights on every rank (replicated model)
    net = mdm.MotionTransformer(precision="bf16", **CFG)
    randomize_zero_init(net)
    state_cpu, extras_cpu = net.state_dict(), net.extras_state()
    net.to(dev)
    if args.expert_parallel and world > 1:
        net.enable_expert_parallel()
    B = B_PER_GPU
    x0, length, xf_c, xf_u = synth_inputs(B, 1000 + rank, dev)     # each rank denoises its own batch
    text_stub = {"c": (xf_c.mean(1), xf_c), "u": (xf_u.mean(1), xf_u)}
    net.encode_text = lambda text, device: text_stub["u"] if text[0] == "" else text_stub["c"]
    d = mdm.GaussianDiffusion(betas=mdm.get_named_beta_schedule("linear", 1000))
    kw = {"text": ["a person walks forward"] * B, "length": length, "xf_proj": text_stub["c"][0],
          "xf_out": text_stub["c"][1]}
    shape = (B, T, CFG["input_feats"])

    # ---- device-resident throughput: K graph-replayed CFG steps
    st = d.make_cfg_stepper(net, shape, kw, cfg_scale=CFG_SCALE, clip_denoised=False, device=dev)
    st.x.copy_(x0)
    ts = 999
    for _ in range(W):
        st.step(ts); ts -= 1
    clocks = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    clocks.start()
    ev0.record()
    for _ in range(args.steps):
        st.step(ts); ts = ts - 1 if ts > 0 else 999
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    finite = bool(torch.isfinite(st.x).all())

    # ---- launches per step (eager pass through the same code, counted at the C-ABI binding)
    from motiondiffusion_moe_b200 import _lib
    eager = d.make_cfg_stepper(net, shape, kw, cfg_scale=CFG_SCALE, clip_denoised=False, device=dev, use_cuda_graph=False)
    eager.x.copy_(x0)
    eager.step(500)
    n0 = _lib.LAUNCHES[0]
    eager.step(499)
    torch.cuda.synchronize(dev)
    launches = _lib.LAUNCHES[0] - n0

    # ---- end to end through the public API with host buffers: H2D of x_t, CFG step, D2H of x_{t-1}
    xh = x0.cpu().pin_memory()
    outh = torch.empty_like(xh).pin_memory()
    th = torch.full((B,), 700, dtype=torch.int64).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))

    outs = [outh, torch.empty_like(xh).pin_memory()]

    def e2e_step(i):
        # public API with HOST buffers: upload of x_t, CFG step, download of x_{t-1}, every step; the copies of
        # neighbouring steps overlap the compute (CFGStepper.step_host: side streams, double-buffered staging)
        st.step_host(xh, int(th[0]), outs[i & 1])

    for i in range(2):
        e2e_step(i)
    st.flush()
    barrier()
    ev0.record()
    for i in range(e2e_steps):
        e2e_step(i)
    st.flush()                                    # the timed region ends after the last download has landed
    ev1.record()
    barrier()
    ms_e2e = ev0.elapsed_time(ev1)
    e2e_ok = bool(torch.isfinite(outs[0]).all() and torch.isfinite(outs[1]).all())

    # ---- roofline of the dominant kernel: the grouped expert GEMMs (tcgen05), timed alone on live buffers
    from motiondiffusion_moe_b200 import ops
    from motiondiffusion_moe_b200._lib import ACT_GELU
    pk, Lr = net._packed, net._packed["layers"][-1]
    N2, D, Fd, E = 2 * B * T, CFG["latent_dim"], CFG["ff_size"], CFG["moe_num_experts"]
    reps = 10
    ep_mode = args.expert_parallel and world > 1
    if ep_mode:          # the expert GEMMs of this rank on the rows it received in the last step
        ep = net._ep_for(N2)
        ep.use_weights(Lr["ep_w"])
        run_ffn = ep.phase_experts
        moe_rows = int(ep.ntile.item()) * 128
    else:
        cap = 4 * N2 + 2 * E * 128
        bufs = {k: net._ws[(k, s, dt)] for (k, s, dt) in net._ws if k.startswith("moe_") and (s[0] in (cap, cap // 128, 1))}

        def run_ffn():
            ops.gemm(bufs["moe_xp"], Lr["w1"], Lr["b1"], act=ACT_GELU, out_a=bufs["moe_hp"], N=Fd, M=cap,
                     tiles=bufs["moe_tup"], num_tiles=cap // 128, num_tiles_dev=bufs["moe_ntile"], a_rows=cap, w_rows=2 * E * Fd)
            ops.gemm(bufs["moe_hp"], Lr["w2"], Lr["b2"], out_a=bufs["moe_yp"], N=D, M=cap, rowscale=bufs["moe_rscale"],
                     tiles=bufs["moe_tdn"], num_tiles=cap // 128, num_tiles_dev=bufs["moe_ntile"], a_rows=cap, w_rows=2 * E * D)
        moe_rows = 4 * N2                               # 2 branches x top-2 routed rows per token
    torch.cuda.synchronize(dev)
    torch.cuda.nvtx.range_push("expert_ffn")      # ncu --nvtx --nvtx-include "expert_ffn/" captures exactly these
    ev0.record()
    for _ in range(reps):
        run_ffn()
    ev1.record()
    torch.cuda.synchronize(dev)
    torch.cuda.nvtx.range_pop()
    moe_ms = ev0.elapsed_time(ev1) / reps
    moe_flops = 2 * moe_rows * D * Fd * 2            # up + down
    pkv, src = peaks()
    ach = moe_flops / (moe_ms * 1e-3) / 1e12
    roof = {"bound": "tensor", "kernel": "gemm_tc_kernel (grouped expert FFN up+down of one MoEMultiBranchFFN)",
            "achieved": ach, "peak": pkv["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / pkv["bf16_tflops"],
            "peak_source": src + " burst (kernel timed alone)",
            # dram__bytes_read.sum + dram__bytes_write.sum of the two launches (up: 120.7 + 156.7 MB,
            # down: 224.9 + 75.5 MB) from profiles/expert_ffn_r1f_ncu.txt (ncu --set full of this very loop)
            "traffic": 577.8e6, "traffic_source": "profiles/expert_ffn_r1f_ncu.txt",
            "tensor_pipe_active_pct": {"up": 58.9, "down": 65.0, "source": "ncu sm__pipe_tensor_cycles_active"},
            "flops_per_launch_pair": moe_flops, "ms_per_launch_pair": moe_ms}

    # ---- max over ranks
    if world > 1:
        tt = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(tt[0]), float(tt[1])
    frames = world * B * T * args.steps
    value = frames / (ms * 1e-3)
    e2e_val = world * B * T * e2e_steps / (ms_e2e * 1e-3)
    step_flops = 2 * B * 55.81e9                      # 55.81 GFLOP / sequence / forward
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": world * B, "frames": T, "cfg_scale": CFG_SCALE,
                   "parallelism": ("dp%d x ep%d (batch sharded by sequence, experts sharded over the ranks: NVLink peer-memory "
                                   "dispatch/combine + flag barriers inside the step graph)" % (world, world))
                   if (args.expert_parallel and world > 1) else
                   "dp%d (independent batches, no collective in the loop)" % world,
                   "l2_policy": "per-step working set (1.06 GB bf16 weights + >2 GB activations) exceeds the 126 MB L2",
                   "timed_region": "CUDA-graph replay of the CFG step; inputs resident in HBM"},
        "model_tflops_per_gpu": step_flops / (ms / args.steps * 1e-3) / 1e12,
        "roofline": roof,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": xh.numel() * 4,
                "d2h_bytes_per_step": outh.numel() * 4, "steps": e2e_steps,
                "path": "pinned host x_t -> GaussianDiffusion CFGStepper.step_host (public API; H2D, CFG step, D2H every "
                        "step, copies of neighbouring steps overlapped with compute on side streams) -> pinned host x_{t-1}",
                "finite": e2e_ok},
        "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches,
        "clocks": clk, "finite": finite,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_leg(state_cpu, extras_cpu, steps=2, warmup=1)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
