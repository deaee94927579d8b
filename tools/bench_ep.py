"""BASELINE.json configs[3]: expert-parallel MoE FFN (8 experts x 2 branches, top-2) with NVLink token
dispatch, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_ep.py

Every rank owns `--seqs` sequences x 196 frames (weak scaling); one call = gate + counts exchange + dispatch
(peer stores) + grouped expert GEMMs on the owners + combine (peer loads) + FiLM.  Device time, max over
ranks.  Prints one JSON line from rank 0."""
import argparse, json, os, sys
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from motiondiffusion_moe_b200.expert_parallel import ExpertParallelFFN  # noqa: E402
from ep_common import make_weights, make_tokens  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seqs", type=int, default=16)      # CFG batch 8 per GPU = 16 sequences per forward
    ap.add_argument("--iters", type=int, default=50)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    D, Fd, E, T = 512, 1024, 8, 196
    N = a.seqs * T
    w = make_weights(D, Fd, E, torch.bfloat16, dev)
    x, film = make_tokens(a.seqs, T, D, dev, 10 + rank)
    ep = ExpertParallelFFN.create_distributed(D, Fd, E, 2, N, torch.bfloat16, dev)
    ep.set_weights(w["ln_w"], w["ln_b"], w["gate_w"], w["gate_b"], w["w1"], w["b1"], w["w2"], w["b2"])
    out = torch.empty(N, D, device=dev, dtype=torch.bfloat16)
    for _ in range(5):
        ep.forward(x, w["s_w"], w["s_b"], film, T, out)
    torch.cuda.synchronize()
    ep.check_health()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    acc = [0.0] * 5
    dist.barrier()
    torch.cuda.synchronize()
    for _ in range(a.iters):
        ev[0].record(); ep.phase_gate(x); ep.barrier()
        ev[1].record(); ep.phase_dispatch(); ep.barrier()
        ev[2].record(); ep.phase_experts(); ep.barrier()
        ev[3].record(); ep.phase_combine(w["s_w"], w["s_b"], film, T, out)
        ev[4].record()
        torch.cuda.synchronize()
        for i in range(4):
            acc[i] += ev[i].elapsed_time(ev[i + 1])
        acc[4] += ev[0].elapsed_time(ev[4])
    ep.check_health()
    t = torch.tensor([v / a.iters for v in acc], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    remote = (world - 1) / world                       # expected share of rows that leave the GPU
    rows = N * 4
    if rank == 0:
        ms = [float(v) for v in t]
        print(json.dumps({
            "metric": "expert_parallel_moe_tokens_per_sec", "value": world * N / (ms[4] * 1e-3), "unit": "tokens/s",
            "n_gpus": world, "tokens_per_gpu": N, "ms_per_call": ms[4],
            "phases_ms": {"gate+counts+barrier": ms[0], "scan+dispatch+barrier": ms[1], "expert_gemms+barrier": ms[2],
                          "combine+film": ms[3]},
            "nvlink_bytes_per_gpu_per_call": {"dispatch": rows * D * 2 * remote, "combine": rows * D * 2 * remote},
            "dispatch_gbs_per_gpu": rows * D * 2 * remote / (ms[1] * 1e-3) / 1e9,
            "combine_gbs_per_gpu": rows * D * 2 * remote / (ms[3] * 1e-3) / 1e9,
            "expert_tflops_per_gpu": 2 * rows * D * Fd * 2 / (ms[2] * 1e-3) / 1e12,
            "config": {"workload": "MoEMultiBranchFFN: 2 branches x 8 experts top-2, D512 F1024, bf16, %d sequences x 196 "
                                   "frames per GPU; experts sharded %d per GPU per branch" % (a.seqs, E // world)},
            "scaling": "weak", "dtype": "bf16", "data": "synthetic"}))
    ep.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
