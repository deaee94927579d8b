"""Shared set-up of the torchrun workers of the multi-GPU tests.

One process per GPU over NCCL when the box has at least WORLD_SIZE GPUs.  On a single-GPU box (the driver's GPU test
tier) MDM_TEST_SHARED_GPU=1 runs all ranks on cuda:0 with the gloo backend for the plumbing: the CUDA IPC peer buffers
of the expert-parallel path and the sharded sampling loop work the same way between two processes that share a GPU (the
GPU time-slices the two contexts; a spinning flag barrier is pre-empted at the end of its slice), so the multi-rank code
paths are exercised by the driver as well, just without NVLink in between."""
import os

import torch
import torch.distributed as dist


def init_dist():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    shared = os.environ.get("MDM_TEST_SHARED_GPU") == "1"
    dev = torch.device("cuda", 0 if shared else local)
    torch.cuda.set_device(dev)
    if shared:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)
    return rank, world, dev, shared


def all_max(values, dev, shared):
    """max over ranks of a list of floats."""
    t = torch.tensor(values, dtype=torch.float32, device="cpu" if shared else dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def barrier(shared):
    dist.barrier()
    torch.cuda.synchronize()
