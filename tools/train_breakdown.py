"""Per-kernel GPU time of one DDPM training step (bench.py --workload train shapes; B from env, default 256) via
torch.profiler, plus wall-clock vs device-busy time (how launch-bound the eager step is)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B_  # noqa: E402
import bench_train as BT  # noqa: E402
import motiondiffusion_moe_b200 as mdm  # noqa: E402
from motiondiffusion_moe_b200 import train_ops as T_  # noqa: E402
from motiondiffusion_moe_b200.training import TrainEngine  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402

dev = torch.device("cuda")
B = int(os.environ.get("B", "256"))
torch.manual_seed(0)
net = mdm.MotionTransformer(precision="bf16", dropout=0.0, **BT.CFG)
B_.randomize_zero_init(net)
net.to(dev).eval()
eng = TrainEngine(net)
x0, t, length, xf, noise = BT.synth(B, 2000, dev)
d = mdm.GaussianDiffusion(betas=mdm.get_named_beta_schedule("linear", 1000))


def step():
    eng.zero_grad()
    x_t = d.q_sample(x0, t, noise=noise)
    net.reset_all_moe_counters(net)
    pred, S = eng.forward_train(x_t, t, length, xf.mean(1), xf)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    eng.backward(S, T_.masked_mse_grad(pred, noise, length.clamp(max=BT.T).contiguous()))
    torch.cuda.synchronize(); t2 = time.perf_counter()
    eng.optimizer_step()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    return t1, t2, t3


for _ in range(2):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
t1, t2, t3 = step()
print("B=%d wall: forward %.1f ms, backward %.1f ms, optimizer+refresh %.1f ms" % (B, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = []
for ev in prof.key_averages():
    dt = getattr(ev, "device_time_total", None) or getattr(ev, "cuda_time_total", 0)
    rows.append((dt, ev.count, ev.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("total device-busy time %.2f ms over %d kernels" % (tot / 1e3, sum(r[1] for r in rows)))
for dt, n, k in rows[:40]:
    print("%8.2f ms %5.1f%% x%-5d %s" % (dt / 1e3, 100 * dt / tot, n, k[:120]))
