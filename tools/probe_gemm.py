"""GPU probe (round 1): torch.topk CUDA tie order + tcgen05 GEMM correctness/timing vs torch.matmul."""
import ctypes as C, os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motiondiffusion_moe_b200._lib import GemmEpi, LIB_PATH

dev = torch.device("cuda")
out = {}

# ---- 1. topk ties on CUDA
def tk(row, k=2):
    t = torch.tensor([row], dtype=torch.float32, device=dev)
    v, i = torch.topk(t, k, dim=1)
    return i[0].tolist()
cases = {
    "all_equal_8": [0.125]*8, "all_equal_4": [0.25]*4,
    "tie_top_3way": [.1,.3,.3,.1,.3,.0,.0,.0], "tie_second": [.3,.1,.1,.1,.1,.1,.1,.1],
    "tie_second_4": [.4,.2,.2,.2], "tie_top_pair_late": [.1,.1,.1,.1,.1,.1,.3,.3],
    "tie_top_pair_mid": [.1,.3,.1,.1,.3,.1,.1,.1], "distinct": [.1,.5,.05,.2,.03,.02,.06,.04],
}
out["topk_cuda"] = {k: tk(v) for k, v in cases.items()}
out["topk_cpu"] = {k: torch.topk(torch.tensor([v]), 2, dim=1)[1][0].tolist() for k, v in cases.items()}
# batched (many rows) may take a different kernel path
big = torch.tensor([cases["all_equal_8"]]*4096 + [cases["tie_second"]]*4096, device=dev)
bi = torch.topk(big, 2, dim=1)[1]
out["topk_cuda_batched_all_equal"] = bi[:4096].unique(dim=0).tolist()
out["topk_cuda_batched_tie_second"] = bi[4096:].unique(dim=0).tolist()
print(json.dumps(out, indent=1)); sys.stdout.flush()

# ---- 2. GEMM
lib = C.CDLL(LIB_PATH)
P, I, L = C.c_void_p, C.c_int, C.c_long
lib.mdm_gemm_bf16.argtypes = [P, I, L, P, I, L, I, I, I, P, I, P, C.POINTER(GemmEpi), I, P]
lib.mdm_gemm_f32.argtypes = [P, I, L, P, I, L, I, I, I, P, I, P, C.POINTER(GemmEpi), P]

def run_bf16(A, W, bias=None, act=0, resid=None, alpha=1.0, beta=0.0, out_dtype=torch.float32):
    M, K = A.shape; N = W.shape[0]
    o = torch.empty(M, N, device=dev, dtype=out_dtype)
    e = GemmEpi()
    e.bias = bias.data_ptr() if bias is not None else None
    e.resid = resid.data_ptr() if resid is not None else None
    e.ld_resid = N; e.alpha = alpha; e.beta = beta; e.act = act
    if out_dtype == torch.float32: e.out_f32 = o.data_ptr(); e.ld_f32 = N
    else: e.out_bf16 = o.data_ptr(); e.ld_bf16 = N
    st = torch.cuda.current_stream().cuda_stream
    r = lib.mdm_gemm_bf16(A.data_ptr(), A.stride(0), M, W.data_ptr(), W.stride(0), N, M, N, K, None, 0, None, C.byref(e), 0, st)
    assert r == 0, r
    return o

torch.manual_seed(0)
res = []
for (M, N, K) in [(128, 256, 64), (128, 256, 512), (256, 512, 512), (25088, 512, 512), (25088, 1024, 512), (25088, 512, 1024), (25088, 2048, 512), (300, 263, 512), (1000, 128, 128), (12544, 1536, 512), (392, 512, 264)]:
    A = (torch.randn(M, K, device=dev)).bfloat16(); W = (torch.randn(N, K, device=dev) / K**0.5).bfloat16()
    bias = torch.randn(N, device=dev)
    o = run_bf16(A, W, bias)
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t() + bias
    err = ((o - ref).norm() / ref.norm()).item()
    maxerr = (o - ref).abs().max().item()
    # timing
    for _ in range(3): run_bf16(A, W, bias, out_dtype=torch.bfloat16)
    s, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): run_bf16(A, W, bias, out_dtype=torch.bfloat16)
    e_.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e_) / 20
    tf = 2.0 * M * N * K / ms / 1e9
    s.record()
    for _ in range(20): torch.nn.functional.linear(A, W, bias.bfloat16())
    e_.record(); torch.cuda.synchronize()
    ms_t = s.elapsed_time(e_) / 20
    res.append(dict(M=M, N=N, K=K, rel=err, maxabs=maxerr, ms=ms, tflops=tf, torch_ms=ms_t, torch_tflops=2.0*M*N*K/ms_t/1e9))
    print(res[-1]); sys.stdout.flush()
# epilogue variants
M, N, K = 1024, 512, 512
A = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(N, K, device=dev) / K**0.5).bfloat16()
bias = torch.randn(N, device=dev); R = torch.randn(M, N, device=dev)
o = run_bf16(A, W, bias, act=1, resid=R, alpha=0.1, beta=1.0)
ref = 0.1 * torch.nn.functional.gelu(A.float() @ W.float().t() + bias) + R
print("gelu+resid rel", ((o - ref).norm() / ref.norm()).item())
o = run_bf16(A, W, bias, act=2, out_dtype=torch.bfloat16)
ref = torch.nn.functional.silu(A.float() @ W.float().t() + bias)
print("silu bf16 rel", ((o.float() - ref).norm() / ref.norm()).item())
# fp32 simt
def run_f32(A, W, bias):
    M, K = A.shape; N = W.shape[0]
    o = torch.empty(M, N, device=dev)
    e = GemmEpi(); e.bias = bias.data_ptr(); e.alpha = 1.0; e.out_f32 = o.data_ptr(); e.ld_f32 = N
    r = lib.mdm_gemm_f32(A.data_ptr(), K, M, W.data_ptr(), K, N, M, N, K, None, 0, None, C.byref(e), torch.cuda.current_stream().cuda_stream)
    assert r == 0
    return o
for (M, N, K) in [(300, 263, 263), (2048, 512, 512)]:
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / K**0.5; bias = torch.randn(N, device=dev)
    o = run_f32(A, W, bias); ref = (A.double() @ W.double().t() + bias.double()).float()
    print("f32 simt", M, N, K, "rel", ((o - ref).norm() / ref.norm()).item())
json.dump(dict(out, gemm=res), open("gpurun_out/probe_gemm.json", "w"), indent=1)
print("PROBE_DONE")
