"""torchrun worker: expert-parallel MoE with one process per GPU (CUDA IPC peer buffers, flag barriers)
against the local MoE kernels on the same tokens.  Prints EP_DIST_OK from rank 0 on success, and the
time per expert-parallel MoE call."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from motiondiffusion_moe_b200.expert_parallel import ExpertParallelFFN  # noqa: E402
from ep_common import make_weights, make_tokens, local_moe  # noqa: E402
from dist_common import init_dist, all_max  # noqa: E402


def main():
    rank, world, dev, shared = init_dist()
    D, Fd, E = 512, 1024, 8
    n_seq, T = int(os.environ.get("EP_NSEQ", "16")), 196
    dtype = torch.bfloat16
    w = make_weights(D, Fd, E, dtype, dev)
    x, film = make_tokens(n_seq, T, D, dev, 10 + rank)
    ref, idx, vals, usage, _ = local_moe(w, x, film, T, D, Fd, E, dtype)
    ep = ExpertParallelFFN.create_distributed(D, Fd, E, 2, n_seq * T, dtype, dev)
    ep.set_weights(w["ln_w"], w["ln_b"], w["gate_w"], w["gate_b"], w["w1"], w["b1"], w["w2"], w["b2"])
    out = torch.empty(n_seq * T, D, device=dev, dtype=dtype)
    for _ in range(3):
        ep.forward(x, w["s_w"], w["s_b"], film, T, out)
    torch.cuda.synchronize()
    ep.check_health()
    ok = torch.equal(out, ref) and torch.equal(ep.idx, idx) and torch.equal(ep.usage, 3 * usage)
    # timing: max over ranks of the device time per call
    iters = 3 if shared else 20
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ep.forward(x, w["s_w"], w["s_b"], film, T, out)
    e1.record()
    torch.cuda.synchronize()
    ep.check_health()
    t = all_max([e0.elapsed_time(e1) / iters, 0.0 if ok else 1.0], dev, shared)
    if rank == 0:
        rows = n_seq * T * 4
        print("ep world=%d tokens/rank=%d: %.3f ms per expert-parallel MoE call (max over ranks); dispatch+combine "
              "NVLink bytes per rank ~ %.1f MB" % (world, n_seq * T, float(t[0]), 2 * rows * D * 2 * (world - 1) / world / 1e6))
        print("EP_DIST_OK" if float(t[1]) == 0.0 else "EP_DIST_MISMATCH")
    ep.close()
    dist.destroy_process_group()
    sys.exit(0 if float(t[1]) == 0.0 else 1)


if __name__ == "__main__":
    main()
