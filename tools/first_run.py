"""First end-to-end GPU run: forward parity vs the reference-generated goldens, then a timing probe."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import motiondiffusion_moe_b200 as m
from oracle import cases, motion_oracle as mo

dev = torch.device("cuda")
def rel(a, b): return ((a - b).norm() / b.norm()).item()

def build(case, precision):
    cfg, p = cases.case_params(case)
    net = m.MotionTransformer(precision=precision, **cfg)
    net.load_state_dict({k: p[k] for k in net.state_dict()})
    net.load_extras(p)
    return cfg, p, net.cuda()

which = sys.argv[1:] or ["tiny_b3", "small_b4", "default_b2"]
for case in which:
    cfg_name, B, T = cases.CASES[case]
    g = np.load("tests/golden/%s.npz" % case)
    for prec in ("fp32", "bf16"):
        cfg, p, net = build(case, prec)
        x, t, length, xf_proj, xf_out = cases.make_inputs(cfg, B, T, seed=3, device=dev)
        net.record_routing = True
        y = net(x, t, length, None, xf_proj, xf_out)
        torch.cuda.synchronize()
        yr = torch.from_numpy(g["y"]).to(dev)
        n_low = cfg.num_layers
        rl = torch.stack([r[0] for r in net.last_routing[:n_low]]).cpu().numpy()   # [L, N, NB, 2]
        rh = torch.stack([r[0] for r in net.last_routing[n_low:]]).cpu().numpy()
        gl = g["routing_low"].reshape(n_low, 2, -1, 2).transpose(0, 2, 1, 3)
        gh = g["routing_high"].reshape(n_low, 2, -1, 2).transpose(0, 2, 1, 3)
        mis = ((rl != gl).any(-1).sum() + (rh != gh).any(-1).sum()) / float(rl[..., 0].size + rh[..., 0].size)
        print(case, prec, "rel_err %.3e" % rel(y, yr), "nan", bool(torch.isnan(y).any()), "routing mismatch frac %.2e" % mis)
        sys.stdout.flush()
        del net
        torch.cuda.empty_cache()

if "--time" in sys.argv or len(sys.argv) == 1:
    cfg = mo.CONFIGS["default"]
    p = mo.make_params(cfg, 0)
    net = m.MotionTransformer(precision="bf16", **cfg)
    net.load_state_dict({k: p[k] for k in net.state_dict()}); net.load_extras(p); net.cuda()
    B = 64
    x, t, length, xf_proj, xf_out = cases.make_inputs(cfg, 2 * B, 196, seed=5, device=dev)
    ctx = net.prepare_text(xf_proj, xf_out)
    for _ in range(2): y = net(x, t, length, text_ctx=ctx)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3): y = net(x, t, length, text_ctx=ctx)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 3
    print("default bf16 forward of 2B=128 seqs: %.2f ms -> %.0f denoising frames/s" % (ms, B * 196 / ms * 1e3))
print("FIRST_RUN_DONE")
