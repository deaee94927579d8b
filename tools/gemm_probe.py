"""GEMM-only probe for ncu: (a) bf16-out epilogue, (b) fp32-out + fp32 residual epilogue, (c) GELU bf16-out,
M=25088 N=512 K=512 (the D x D projections of a full-resolution layer at CFG batch 128)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motiondiffusion_moe_b200 import ops
from motiondiffusion_moe_b200._lib import ACT_GELU
dev = "cuda"
M, N, K = 25088, 512, int(os.environ.get("K", "512"))
A = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
b = torch.randn(N, device=dev); R = torch.randn(M, N, device=dev)
ob = torch.empty(M, N, device=dev, dtype=torch.bfloat16); of = torch.empty(M, N, device=dev)
for it in range(int(os.environ.get("IT", "3"))):
    ops.gemm(A, W, b, out_a=ob)
    ops.gemm(A, W, b, out_f32=of, resid=R, alpha=0.1, beta=1.0)
    ops.gemm(A, W, b, out_a=ob, act=ACT_GELU)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("bf16", lambda: ops.gemm(A, W, b, out_a=ob)), ("f32+resid", lambda: ops.gemm(A, W, b, out_f32=of, resid=R, alpha=0.1, beta=1.0)), ("gelu", lambda: ops.gemm(A, W, b, out_a=ob, act=ACT_GELU))):
    s.record()
    for _ in range(10): fn()
    e.record(); torch.cuda.synchronize()
    print(name, "%.1f us" % (s.elapsed_time(e) * 100))
print("GEMM_PROBE_DONE")
