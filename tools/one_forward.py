"""One default-config bf16 forward of 2B sequences (B from env, default 64) after one warm-up forward.
Used under `ncu --metrics gpu__time_duration.sum` for the per-launch list in profiles/."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _model import build
dev = torch.device("cuda")
B = int(os.environ.get("B", "64"))
net, x, t, length, xf_proj, xf_out = build(dev, 2 * B)
ctx = net.prepare_text(xf_proj, xf_out)
for _ in range(int(os.environ.get("FWD", "2"))):
    y = net(x, t, length, text_ctx=ctx)
torch.cuda.synchronize()
print("ONE_FORWARD_DONE", float(y.abs().mean()))
