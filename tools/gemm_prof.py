"""Bring-up tool: per-CTA cycle accounting of the single-CTA tcgen05 GEMM (library built with
-DMDM_GEMM_PROFILE, MDM_GEMM_PAIR=0): who waits for whom - epilogue for accumulators, MMA thread for a
free accumulator buffer (= epilogue too slow) or for operands (= TMA feed too slow)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MDM_GEMM_PAIR", "0")
from motiondiffusion_moe_b200 import ops, _lib
from motiondiffusion_moe_b200._lib import ACT_GELU, ACT_NONE
dev = torch.device("cuda")
lib = _lib.load()
bf = torch.bfloat16
N = 128 * 196
def run(name, M, Nn, K, act=ACT_NONE, f32=False, rowscale=False):
    A = torch.randn(M, K, device=dev).to(bf); W = (torch.randn(Nn, K, device=dev) / K ** 0.5).to(bf)
    b = torch.randn(Nn, device=dev)
    R = torch.randn(M, Nn, device=dev) if f32 else None
    O = torch.empty(M, Nn, device=dev, dtype=torch.float32 if f32 else bf)
    rs = torch.rand(M, device=dev) if rowscale else None
    for _ in range(3):
        ops.gemm(A, W, b, act=act, out_f32=O if f32 else None, out_a=None if f32 else O, resid=R, beta=1.0 if f32 else 0.0, rowscale=rs)
    torch.cuda.synchronize()
    ph = (C.c_ulonglong * 8)()
    lib.mdm_debug_read_epi_phase(ph, 1)
    for _ in range(1):
        ops.gemm(A, W, b, act=act, out_f32=O if f32 else None, out_a=None if f32 else O, resid=R, beta=1.0 if f32 else 0.0, rowscale=rs)
    torch.cuda.synchronize()
    lib.mdm_debug_read_epi_phase(ph, 1)
    buf = (C.c_ulonglong * (148 * 8))()
    assert lib.mdm_debug_read_gemm_prof(buf, 148 * 8) == 0
    v = torch.tensor(list(buf), dtype=torch.float64).view(148, 8)
    if os.environ["MDM_GEMM_PAIR"] != "0":      # MMA-thread counters live in the leader (even) CTA of a pair
        v[1::2, 2:5] = v[0::2, 2:5]
    tiles = v[:, 5].clamp_min(1)
    print("%-34s per tile (cycles, mean over CTAs): epilogue wait %6.0f  epilogue work %6.0f | MMA thread: wait-accumulator %6.0f  wait-operands %6.0f  total %6.0f  (tiles/CTA %.1f)"
          "\n      epilogue phases per tile: tmem-load %5.0f  bias/act/pack/STS %5.0f  sync+LDS %5.0f  stores(+resid) %5.0f  pre-wait(resid issue) %5.0f  other %5.0f"
          % (name, (v[:, 0] / tiles).mean(), (v[:, 1] / tiles).mean(), (v[:, 2] / tiles).mean(), (v[:, 3] / tiles).mean(), (v[:, 4] / tiles).mean(), tiles.mean(),
             ph[0] / tiles.sum(), ph[1] / tiles.sum(), ph[2] / tiles.sum(), ph[3] / tiles.sum(), ph[4] / tiles.sum(), ph[7] / tiles.sum()))
lib.mdm_debug_read_gemm_prof.argtypes = [C.c_void_p, C.c_int]
lib.mdm_debug_read_epi_phase.argtypes = [C.c_void_p, C.c_int]
run("up   4N x1024x512 gelu bf16", 4 * N, 1024, 512, act=ACT_GELU)
run("down 4N x512x1024 rowscale bf16", 4 * N, 512, 1024, rowscale=True)
run("qkv  N x1536x512 bf16", N, 1536, 512)
run("p3   N x512x512 bf16", N, 512, 512)
run("s_out N x512x512 f32+resid", N, 512, 512, f32=True)
run("f3   N x512x2048 f32+resid", N, 512, 2048, f32=True)
