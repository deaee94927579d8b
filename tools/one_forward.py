"""One default-config bf16 forward of 2B sequences (B from env, default 64) after one warm-up forward.
Used under `ncu --metrics gpu__time_duration.sum` for the per-launch list in profiles/."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import motiondiffusion_moe_b200 as m
from oracle import cases, motion_oracle as mo
dev = torch.device("cuda")
cfg = mo.CONFIGS["default"]
p = mo.make_params(cfg, 0)
net = m.MotionTransformer(precision="bf16", **cfg)
net.load_state_dict({k: p[k] for k in net.state_dict()}); net.load_extras(p); net.cuda()
B = int(os.environ.get("B", "64"))
x, t, length, xf_proj, xf_out = cases.make_inputs(cfg, 2 * B, 196, seed=5, device=dev)
ctx = net.prepare_text(xf_proj, xf_out)
for _ in range(int(os.environ.get("FWD", "2"))):
    y = net(x, t, length, text_ctx=ctx)
torch.cuda.synchronize()
print("ONE_FORWARD_DONE", float(y.abs().mean()))
