"""Bring-up tool: per-phase cycles of fastattn_tc_kernel (library built with -DMDM_ATTN_PROFILE)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motiondiffusion_moe_b200 import ops, _lib
dev = torch.device("cuda")
lib = _lib.load()
lib.mdm_debug_read_fa_phase.argtypes = [C.c_void_p, C.c_int]
for T in (196, 98):
    NSEQ, D, H = 128, 512, 4
    hd = D // H
    N = NSEQ * T
    qkv = torch.randn(N, 3 * D, device=dev).to(torch.bfloat16)
    P = torch.randn(hd, hd, device=dev) * hd ** -0.5
    nw, nb = torch.rand(hd, device=dev) + 0.5, torch.randn(hd, device=dev) * 0.1
    length = torch.randint(40, T + 1, (NSEQ,), device=dev, dtype=torch.int64)
    out = torch.empty(N, D, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.fastattn(qkv, P, nw, nb, length, 0, NSEQ, H, T, hd, out)
    torch.cuda.synchronize()
    ph = (C.c_ulonglong * 16)()
    lib.mdm_debug_read_fa_phase(ph, 1)
    ops.fastattn(qkv, P, nw, nb, length, 0, NSEQ, H, T, hd, out)
    torch.cuda.synchronize()
    lib.mdm_debug_read_fa_phase(ph, 1)
    n = NSEQ * H
    names = (["S0 load q + P^T", "LN q", "S2 q features", "k/v window loop", "kv -> smem", "", "", "S4 apply + LN + store"] if os.environ.get("MDM_FA_STREAM", "1") != "0" else ["S0 load k,v + P^T", "S1 LN k,v (+wait q)", "S1 LN q", "S2 features (+den)", "S3 kv", "", "", "S4 apply + LN + store"])
    tot = sum(ph[i] for i in range(8))
    print("T=%d: cycles per CTA %.0f" % (T, tot / n))
    for i, nm in enumerate(names):
        if nm:
            print("   %-24s %7.0f  %5.1f%%" % (nm, ph[i] / n, 100.0 * ph[i] / tot))

# ---- tcgen05 kernel (attention_umma.cu): phases P0 loads issued / v pass / q,k pass / MMA1 + K'^T epilogue / kv MMA + Q' epilogue / kv epilogue / apply MMA + out epilogue
if hasattr(lib, "mdm_debug_read_fau_phase"):
    lib.mdm_debug_read_fau_phase.argtypes = [C.c_void_p, C.c_int]
    for T in (196, 98):
        NSEQ, D, H = 128, 512, 4
        hd = D // H
        N = NSEQ * T
        qkv = torch.randn(N, 3 * D, device=dev).to(torch.bfloat16)
        P = torch.randn(hd, hd, device=dev) * hd ** -0.5
        Pt = P.t().contiguous().to(torch.bfloat16)
        nw, nb = torch.rand(hd, device=dev) + 0.5, torch.randn(hd, device=dev) * 0.1
        length = torch.randint(40, T + 1, (NSEQ,), device=dev, dtype=torch.int64)
        out = torch.empty(N, D, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            ops.fastattn(qkv, P, nw, nb, length, 0, NSEQ, H, T, hd, out, Pt=Pt)
        torch.cuda.synchronize()
        ph = (C.c_ulonglong * 16)()
        lib.mdm_debug_read_fau_phase(ph, 1)
        ops.fastattn(qkv, P, nw, nb, length, 0, NSEQ, H, T, hd, out, Pt=Pt)
        torch.cuda.synchronize()
        lib.mdm_debug_read_fau_phase(ph, 1)
        n = NSEQ * H
        names = ["P0 issue loads + Pt", "v LN + transpose", "q,k LN in place", "MMA K'T,Q'' + K'T epilogue", "kv MMA + Q' epilogue + den",
                 "kv epilogue", "apply MMA + out epilogue"]
        tot = sum(ph[i] for i in range(8))
        print("tcgen05 T=%d: cycles per CTA %.0f" % (T, tot / n))
        for i, nm in enumerate(names):
            print("   %-30s %7.0f  %5.1f%%" % (nm, ph[i] / n, 100.0 * ph[i] / tot))
