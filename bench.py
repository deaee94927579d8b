#!/usr/bin/env python
"""bench.py — denoising frames/s of the default MoE MotionTransformer under bf16 CFG sampling.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W
    python bench.py --workload train ...                     (BASELINE configs[4]: DDPM training step, KIT-shaped)

Workload (BASELINE.json configs[1]): L8 / D512 / F1024 / 8 experts top-2 / 4 heads / Dt256, 196 frames x 263
features, batch 64 per GPU, classifier-free guidance (conditional + unconditional = 128 sequences per
forward), synthetic inputs and random-init weights (zero-initialised tensors re-randomised).
A "step" is one full CFG reverse-diffusion step: batched forward + guided DDPM update.

`value` is WEAK scaling: every GPU denoises its own batch of 64 (independent sequences, no data-path collective).
The same line also carries (all measured in this run):
  strong     BASELINE configs[2]: ONE global batch of 64 sharded over the N GPUs (parallel.sample_dp: B_local = 64/N,
             noise drawn on the device from one seed, final all_gather inside the timed region)
  sustained  1000 consecutive graph-replayed CFG steps (a full DDPM sampling loop) with the clocks sampled meanwhile
  e2e        the same metric through the public API with HOST buffers (H2D + D2H every step)
  roofline   the grouped expert FFN (tcgen05) against the measured bf16 peak
  gpu_eager_baseline   the reference's own torch-eager path on the SAME B200 (fp32 and autocast bf16): the real bar
  cpu_baseline         the reference's CPU path on the host cores (N = 1 only)

One JSON line is printed by rank 0; see README/DESIGN.md for the meaning of every key.
"""
import argparse
import contextlib
import json
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


def synth_inputs(B, seed, device, feats=None):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, feats or CFG["input_feats"], generator=g)
    length = torch.randint(40, T + 1, (B,), generator=g)
    xf_c = torch.nn.functional.gelu(torch.randn(B, 20, CFG["text_latent_dim"], generator=g))
    xf_u = torch.nn.functional.gelu(torch.randn(1, 10, CFG["text_latent_dim"], generator=g)).expand(B, -1, -1)
    return x, length.to(device), xf_c.to(device), xf_u.contiguous().to(device)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

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
        pw = []
        for s in self.samples:
            try:
                pw.append(float(s[6]))
            except Exception:
                pass
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw) if pw else None}


# ---------------------------------------------------------------------------------------------------------------
# baselines: the reference's own code (oracle/_ref: the unmodified reference, byte-compiled) or, where that is absent,
# the oracle port.  These legs are the only places bench.py executes anything under oracle/.
# ---------------------------------------------------------------------------------------------------------------
def _reference_cfg_runner(state, extras, device, batch, kind):
    """Returns (step_fn, description).  kind 'reference': models.MotionTransformer + GaussianDiffusion.p_sample_with_cfg
    of the unmodified reference (per-forward ephemeral Linears on the CPU RNG included); 'port': oracle/motion_oracle.py
    (ephemerals pinned: the best case for torch eager)."""
    from oracle import motion_oracle as mo
    x, length, xf_c, xf_u = synth_inputs(batch, 123, device)
    x = x.to(device)
    t = torch.full((batch,), 500, dtype=torch.long, device=device)
    p = {k: v.float().to(device) for k, v in state.items()}
    p.update({k: v.float().to(device) for k, v in extras.items()})
    if kind == "reference":
        from oracle import ref_runner
        with contextlib.redirect_stdout(sys.stderr):        # the reference prints from its constructors
            model = ref_runner.build_model(mo.CONFIGS["default"], p, device)
            diff = ref_runner.diffusion(1000)
        xp = xf_c.mean(1)

        def step():
            with contextlib.redirect_stdout(sys.stderr):
                return ref_runner.cfg_step(model, diff, x, t, length, xp, xf_c, CFG_SCALE)["sample"]
        return step, "unmodified reference (oracle/_ref), stub text encoder"
    cfg = mo.CONFIGS["default"]
    tab = mo.diffusion_tables(1000)
    noise = torch.randn(x.shape, generator=torch.Generator().manual_seed(1)).to(device)

    def step():
        with torch.no_grad():
            return mo.cfg_step(p, cfg, tab, x, t, length, (xf_c.mean(1), xf_c), (xf_u.mean(1), xf_u), noise, CFG_SCALE)[0]
    return step, "oracle port of the reference (ephemeral Linears pinned)"


def cpu_reference_leg(state, extras, steps, warmup, batch=B_PER_GPU):
    """The reference's CPU path for this metric on the host cores, fp32, all threads, on a bounded sample: `steps` CFG
    steps (two sequential forwards + DDPM update, as the reference does) of the default model at batch `batch`."""
    from oracle import ref_runner
    torch.set_num_threads(os.cpu_count() or 1)
    kind = "reference" if ref_runner.available() else "port"
    step, what = _reference_cfg_runner(state, extras, "cpu", batch, kind)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"value": batch * T / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": "default model fp32, batch %d x 196 frames, %d CFG steps (2 forwards + update each) after %d "
                      "warm-up; %s; torch %s on the host cores" % (batch, steps, warmup, what, torch.__version__),
            "ms_per_step": sec * 1e3}


def gpu_eager_leg(state, extras, dev, batch=B_PER_GPU, steps=3):
    """The reference's torch-eager path on the SAME GPU ('the real bar'):
    CFG steps at batch 64 in fp32 and under torch.autocast(bfloat16), for the unmodified reference (if oracle/_ref was
    built) and for the oracle port with pinned ephemerals.  CUDA-event timed, 1 warm-up step."""
    from oracle import ref_runner
    out = {}
    kinds = (["reference"] if ref_runner.available() else []) + ["port"]
    for kind in kinds:
        step, what = _reference_cfg_runner(state, extras, dev, batch, kind)
        for mode in ("fp32", "autocast_bf16"):
            ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode != "fp32" else contextlib.nullcontext()
            try:
                with ctx:
                    step()
                    torch.cuda.synchronize(dev)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(steps):
                        step()
                    e1.record()
                    torch.cuda.synchronize(dev)
                ms = e0.elapsed_time(e1) / steps
                out["%s_%s" % (kind, mode)] = {"value": batch * T / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                                               "steps": steps, "what": what}
            except Exception as exc:      # a baseline that cannot run is reported, never silently dropped
                out["%s_%s" % (kind, mode)] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
        del step
        torch.cuda.empty_cache()
    out["note"] = ("torch %s eager on this GPU, batch %d, two sequential forwards + update per step; TF32 off "
                   "(torch default for matmul)" % (torch.__version__, batch))
    return out


def ncu_side_data():
    """dram traffic / tensor-pipe utilisation of the roofline kernel from the latest ncu capture committed under profiles/
    (tools/ncu_summaries.py writes profiles/expert_ffn_ncu_latest.json with the commit it was taken at).  Nothing is
    hard-coded here: without that file the keys are null."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "expert_ffn_ncu_latest.json")))
    except Exception:
        return None


_REAL_STDOUT = [None]


def own_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner, the reference's
    constructors): from here on fd 1 goes to stderr and the line is written to the saved descriptor by emit()."""
    if _REAL_STDOUT[0] is None:
        sys.stdout.flush()
        _REAL_STDOUT[0] = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        sys.stdout = sys.stderr


def emit(line):
    out = _REAL_STDOUT[0] or sys.__stdout__
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    own_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="sample", choices=["sample", "train"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--sustained-steps", type=int, default=1000)
    ap.add_argument("--strong-steps", type=int, default=100)
    ap.add_argument("--expert-parallel", action="store_true",
                    help="N > 1 only: shard the experts over the ranks (NVLink dispatch/combine) instead of replicating them")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(3, args.warmup)

    import motiondiffusion_moe_b200 as mdm

    if args.workload == "train":
        import bench_train
        bench_train.B_._REAL_STDOUT = _REAL_STDOUT      # `import bench` inside bench_train is a second module instance: share the real stdout
        return bench_train.main(args, rank, world, local)

    if args.impl == "reference":
        if rank != 0:
            return
        torch.manual_seed(0)
        net = mdm.MotionTransformer(precision="bf16", **CFG)
        randomize_zero_init(net)
        k = max(1, min(args.steps, 2))
        cb = cpu_reference_leg(net.state_dict(), net.extras_state(), steps=k, warmup=1)
        emit({"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT,
              "n_gpus": args.gpus, "steps": k, "warmup": 1, "ms_per_step": cb["ms_per_step"],
              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
              "data": "synthetic", "config": {"workload": WORKLOAD, "reference_sample": cb["sample"]},
              "cpu_baseline": cb,
              "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
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

    def max_over_ranks(vals):
        if world == 1:
            return [float(v) for v in vals]
        tt = torch.tensor(vals, device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return [float(v) for v in tt]

    torch.manual_seed(0)                       # identical weights on every rank (replicated model)
    net = mdm.MotionTransformer(precision="bf16", **CFG)
    randomize_zero_init(net)
    state_cpu, extras_cpu = net.state_dict(), net.extras_state()
    net.to(dev)
    ep_mode = args.expert_parallel and world > 1
    if ep_mode:
        net.enable_expert_parallel()
    B = B_PER_GPU
    x0, length, xf_c, xf_u = synth_inputs(B, 1000 + rank, dev)     # each rank denoises its own batch
    text_stub = {"c": (xf_c.mean(1), xf_c), "u": (xf_u.mean(1), xf_u)}
    net.encode_text = lambda text, device: (text_stub["u"][0][:len(text)], text_stub["u"][1][:len(text)]) \
        if text[0] == "" else (text_stub["c"][0][:len(text)], text_stub["c"][1][:len(text)])
    d = mdm.GaussianDiffusion(betas=mdm.get_named_beta_schedule("linear", 1000))
    kw = {"text": ["a person walks forward"] * B, "length": length, "xf_proj": text_stub["c"][0],
          "xf_out": text_stub["c"][1]}
    shape = (B, T, CFG["input_feats"])

    # ---- device-resident throughput (weak scaling): K graph-replayed CFG steps
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
    net.check_health()

    # ---- sustained: a full 1000-step DDPM sampling loop (BASELINE configs[2] is "1000-step CFG sampling"), clocks sampled
    sustained = None
    if not args.no_sustained and args.sustained_steps > 0:
        n_sus = args.sustained_steps
        st.x.copy_(x0)
        sclk = ClockSampler(local)
        barrier()
        sclk.start()
        ev0.record()
        for i in range(n_sus):
            st.step(999 - (i % 1000))
        ev1.record()
        barrier()
        ms_sus = max_over_ranks([ev0.elapsed_time(ev1)])[0]
        sc = sclk.stop()
        sustained = {"steps": n_sus, "ms_per_step": ms_sus / n_sus, "seconds": ms_sus * 1e-3,
                     "value": world * B * T * n_sus / (ms_sus * 1e-3), "unit": UNIT, "clocks": sc,
                     "finite": bool(torch.isfinite(st.x).all()),
                     "what": "%d consecutive graph-replayed CFG steps t = 999..0 (one full DDPM sampling loop of the batch), "
                             "fresh noise every step, max over ranks" % n_sus}

    # ---- launches per step (eager pass through the same code, counted at the C-ABI binding)
    from motiondiffusion_moe_b200 import _lib
    eager = d.make_cfg_stepper(net, shape, kw, cfg_scale=CFG_SCALE, clip_denoised=False, device=dev, use_cuda_graph=False)
    eager.x.copy_(x0)
    eager.step(500)
    n0 = _lib.LAUNCHES[0]
    eager.step(499)
    torch.cuda.synchronize(dev)
    launches = _lib.LAUNCHES[0] - n0
    del eager

    # ---- end to end through the public API with host buffers: H2D of x_t, CFG step, D2H of x_{t-1}
    xh = x0.cpu().pin_memory()
    outh = torch.empty_like(xh).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    outs = [outh, torch.empty_like(xh).pin_memory()]

    def e2e_step(i):
        # public API with HOST buffers: upload of x_t, CFG step, download of x_{t-1}, every step; the copies of
        # neighbouring steps overlap the compute (CFGStepper.step_host: side streams, double-buffered staging)
        st.step_host(xh, 700, outs[i & 1])

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

    # ---- strong scaling: ONE global batch of 64 sharded over the ranks (parallel.sample_dp), all_gather included
    strong = None
    if not ep_mode:
        from motiondiffusion_moe_b200 import parallel
        gB = B_PER_GPU
        if gB % world == 0:
            gx, glen, gxf_c, gxf_u = synth_inputs(gB, 4242, dev)          # the same global batch on every rank
            gstub = {"c": (gxf_c.mean(1), gxf_c), "u": (gxf_u.mean(1), gxf_u)}
            lo, hi = parallel.shard_range(gB, world, rank)
            net.encode_text = lambda text, device: (gstub["u"][0][lo:hi], gstub["u"][1][lo:hi]) if text[0] == "" \
                else (gstub["c"][0][lo:hi], gstub["c"][1][lo:hi])
            gkw = {"text": ["a person walks forward"] * gB, "length": glen, "xf_proj": gstub["c"][0], "xf_out": gstub["c"][1]}
            gshape = (gB, T, CFG["input_feats"])
            lkw = parallel.slice_kwargs(gkw, lo, hi)
            lst = d.make_cfg_stepper(net, (hi - lo, T, CFG["input_feats"]), lkw, cfg_scale=CFG_SCALE, clip_denoised=False,
                                     device=dev)
            n_str = max(10, args.strong_steps)
            run = lambda n: parallel.sample_dp(None, gshape, gkw, 1000, seed=5, num_steps=n, noise="device", stepper=lst)
            run(3)                                                        # warm-up: capture + NCCL all_gather
            barrier()
            ev0.record()
            sample = run(n_str)
            ev1.record()
            barrier()
            ms_str, = max_over_ranks([ev0.elapsed_time(ev1)])
            strong = {"value": gB * T * n_str / (ms_str * 1e-3), "unit": UNIT, "ms_per_step": ms_str / n_str,
                      "steps": n_str, "global_batch": gB, "batch_per_gpu": hi - lo, "scaling": "strong",
                      "finite": bool(torch.isfinite(sample).all()), "gathered_shape": list(sample.shape),
                      "what": "parallel.sample_dp: global batch 64 sharded by sequence, per-step noise drawn on the device "
                              "from one seed (N-GPU sample == 1-GPU sample), CUDA-graph step per rank, final all_gather "
                              "inside the timed region; max over ranks"}
            del lst

    # ---- roofline of the dominant kernel: the grouped expert GEMMs (tcgen05), timed alone on live buffers
    from motiondiffusion_moe_b200 import ops
    from motiondiffusion_moe_b200._lib import ACT_GELU
    pk, Lr = net._packed, net._packed["layers"][-1]
    N2, D, Fd, E = 2 * B * T, CFG["latent_dim"], CFG["ff_size"], CFG["moe_num_experts"]
    reps = 10
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
    with torch.cuda.device(dev):
        run_ffn()
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
    side = ncu_side_data()
    roof = {"bound": "tensor", "kernel": "gemm_tc_kernel (grouped expert FFN up+down of one MoEMultiBranchFFN)",
            "achieved": ach, "peak": pkv["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / pkv["bf16_tflops"],
            "peak_source": src + " burst (kernel timed alone)",
            "frac_of_sustained_peak": ach / pkv.get("bf16_tflops_sustained", pkv["bf16_tflops"]),
            "traffic": side.get("traffic") if side else None,
            "traffic_source": side.get("source") if side else "no ncu capture at this commit (profiles/expert_ffn_ncu_latest.json absent)",
            "tensor_pipe_active_pct": side.get("tensor_pipe_active_pct") if side else None,
            "ncu_captured_at_commit": side.get("commit") if side else None,
            "flops_per_launch_pair": moe_flops, "ms_per_launch_pair": moe_ms}

    # ---- max over ranks
    ms, ms_e2e = max_over_ranks([ms, ms_e2e])
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
                   if ep_mode else "dp%d (independent batches, no collective in the loop)" % world,
                   "l2_policy": "per-step working set (1.06 GB bf16 weights + >2 GB activations) exceeds the 126 MB L2",
                   "timed_region": "CUDA-graph replay of the CFG step; inputs resident in HBM"},
        "model_tflops_per_gpu": step_flops / (ms / args.steps * 1e-3) / 1e12,
        "model_frac_of_sustained_peak": step_flops / (ms / args.steps * 1e-3) / 1e12 / pkv.get("bf16_tflops_sustained", 1376.8),
        "roofline": roof,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": xh.numel() * 4,
                "d2h_bytes_per_step": outh.numel() * 4, "steps": e2e_steps,
                "path": "pinned host x_t -> GaussianDiffusion CFGStepper.step_host (public API; H2D, CFG step, D2H every "
                        "step, copies of neighbouring steps overlapped with compute on side streams) -> pinned host x_{t-1}",
                "finite": e2e_ok},
        "strong": strong, "sustained": sustained,
        "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches,
        "clocks": clk, "finite": finite,
    }
    if rank == 0:
        if world == 1 and not args.no_gpu_eager_baseline:
            del st
            net._ws = {}
            torch.cuda.empty_cache()
            line["gpu_eager_baseline"] = gpu_eager_leg(state_cpu, extras_cpu, dev)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_leg(state_cpu, extras_cpu, steps=1, warmup=1)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
