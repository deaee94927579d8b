"""Per-kernel GPU time of one default-config forward (2B=128 sequences, bf16) via torch.profiler."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _model import build
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda")
B = int(os.environ.get("B", "64"))
net, x, t, length, xf_proj, xf_out = build(dev, 2 * B)
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
