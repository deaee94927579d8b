"""Two launches of the fused Linear + row-pipeline kernel (mdm_gemm_ln) at the config-2 shape, for ncu:
variant 1 (p3 -> LN, L2, LN, FiLM, SiLU) and variant 2 (s_out + residual -> y fp32, LN bf16)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motiondiffusion_moe_b200 import ops
dev = torch.device("cuda")
torch.manual_seed(0)
bf = torch.bfloat16
NSEQ, T, D = int(os.environ.get("NSEQ", "128")), 196, 512
N = NSEQ * T
x = torch.randn(N, D, device=dev).to(bf)
W = (torch.randn(D, D, device=dev) / D ** 0.5).to(bf)
b = torch.randn(D, device=dev)
R = torch.randn(N, D, device=dev)
ln = (torch.rand(D, device=dev) + 0.5, torch.randn(D, device=dev))
ln2 = (torch.rand(D, device=dev) + 0.5, torch.randn(D, device=dev))
film = torch.randn(NSEQ, 2 * D, device=dev)
o = torch.empty(N, D, device=dev, dtype=bf)
y = torch.empty(N, D, device=dev)
for _ in range(int(os.environ.get("REPS", "2"))):
    assert ops.gemm_ln(x, W, b, ln1=ln, l2norm=True, ln2=ln2, film=film, rows_per_seq=T, silu=True, out2_a=o)
    assert ops.gemm_ln(x, W, b, ln1=ln, alpha=0.1, beta=1.0, resid=R, out_f32=y, out1_a=o)
torch.cuda.synchronize()
print("ok")
