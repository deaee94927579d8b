"""Bring-up tool: per-tile cycle accounting of gemm_rowop_kernel (library built with -DMDM_GEMM_PROFILE)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motiondiffusion_moe_b200 import ops, _lib
dev = torch.device("cuda")
lib = _lib.load()
lib.mdm_debug_read_gemm_prof.argtypes = [C.c_void_p, C.c_int]
bf = torch.bfloat16
T, D = 196, 512
N = 128 * T
x = torch.randn(N, D, device=dev).to(bf)
W = (torch.randn(D, D, device=dev) / D ** 0.5).to(bf)
b = torch.randn(D, device=dev)
R, O = torch.randn(N, D, device=dev), torch.empty(N, D, device=dev)
ln = (torch.rand(D, device=dev) + 0.5, torch.randn(D, device=dev))
ln2 = (torch.rand(D, device=dev) + 0.5, torch.randn(D, device=dev))
film = torch.randn(128, 2 * D, device=dev)
for _ in range(3):
    ops.gemm_rowop(x, N, D, W, b, ln1=ln, l2norm=True, ln2=ln2, film=film, rows_per_seq=T, silu=True, out_f32=O, resid=R, alpha=0.1)
torch.cuda.synchronize()
buf = (C.c_ulonglong * (148 * 8))()
assert lib.mdm_debug_read_gemm_prof(buf, 148 * 8) == 0
v = torch.tensor(list(buf), dtype=torch.float64).view(148, 8)
tiles = v[:, 5].clamp_min(1)
print("per tile (cycles, mean over CTAs): A build %6.0f  wait accumulators %6.0f  epilogue %6.0f | MMA thread: wait A %6.0f  wait weights %6.0f  total %6.0f  (tiles/CTA %.2f)"
      % ((v[:, 0] / tiles).mean(), (v[:, 1] / tiles).mean(), (v[:, 6] / tiles).mean(), (v[:, 2] / tiles).mean(), (v[:, 3] / tiles).mean(), (v[:, 4] / tiles).mean(), tiles.mean()))
