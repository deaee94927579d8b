"""Per-kernel GPU time of one default-config forward (2B=128 sequences, bf16) via torch.profiler."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import motiondiffusion_moe_b200 as m
from oracle import cases, motion_oracle as mo
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda")
cfg = mo.CONFIGS["default"]
p = mo.make_params(cfg, 0)
net = m.MotionTransformer(precision="bf16", **cfg)
net.load_state_dict({k: p[k] for k in net.state_dict()}); net.load_extras(p); net.cuda()
B = int(os.environ.get("B", "64"))
x, t, length, xf_proj, xf_out = cases.make_inputs(cfg, 2 * B, 196, seed=5, device=dev)
ctx = net.prepare_text(xf_proj, xf_out)
for _ in range(2): net(x, t, length, text_ctx=ctx)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    net(x, t, length, text_ctx=ctx)
    torch.cuda.synchronize()
import json
evs = [(e.time_range.start, e.name, e.time_range.elapsed_us()) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort()
json.dump([(n[:80], d) for _, n, d in evs], open("gpurun_out/timeline.json", "w"))
rows = []
for ev in prof.key_averages():
    dt = getattr(ev, "device_time_total", None) or getattr(ev, "cuda_time_total", 0)
    rows.append((dt, ev.count, ev.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("total device time %.2f ms" % (tot / 1e3))
for dt, n, k in rows[:25]:
    print("%8.2f ms %5.1f%% x%-4d %s" % (dt / 1e3, 100 * dt / tot, n, k[:110]))
