"""Host-side logic of the multi-GPU paths, exercised with world_size-2 gloo process groups on the CPU:
  * data-parallel CFG sampling (motiondiffusion_moe_b200/parallel.py): sharded loop == unsharded loop
  * expert-parallel row placement (expert_parallel.plan_segments, the host mirror of ep_scan_kernel):
    every routed row of every rank gets a unique slot inside the right owner's segment.
No CUDA kernel runs here (the compute path has no CPU fallback); the GPU tests cover the kernels."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from motiondiffusion_moe_b200 import parallel
from motiondiffusion_moe_b200.expert_parallel import plan_segments, owner_of


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class ToyStepper:
    """Stand-in for CFGStepper with per-sequence arithmetic only (what makes sequence sharding exact)."""
    def __init__(self, shape, kw):
        self.x = torch.zeros(*shape)
        self.length = kw["length"].float().view(-1, 1, 1)

    def step(self, t, noise, ts_prev=None):
        jump = 1.0 if ts_prev is None else float(t - ts_prev)          # strided (DDIM-style) schedules pass ts_prev
        self.x.copy_(0.9 * self.x + 0.01 * (t % 7) * jump * torch.tanh(self.x) / self.length + 0.1 * noise)
        return self.x


SCHEDULE = [(11, 7), (7, 4), (4, 0), (0, -1)]        # a strided schedule as GaussianDiffusion.ddim_timesteps returns it


def _dp_worker(rank, world, port, B, ret, schedule=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shape = (B, 6, 5)
    kw = {"length": torch.arange(1, B + 1), "text": ["t%d" % i for i in range(B)]}
    full = parallel.sample_dp(lambda s, k: ToyStepper(s, k), shape, kw, num_timesteps=12, seed=3, schedule=schedule)
    if rank == 0:
        ret.put(full)
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [4, 5])
def test_dp_sampling_equals_single_rank(B):
    shape = (B, 6, 5)
    kw = {"length": torch.arange(1, B + 1), "text": ["t%d" % i for i in range(B)]}
    single = parallel.sample_dp(lambda s, k: ToyStepper(s, k), shape, kw, num_timesteps=12, seed=3)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert torch.equal(got, single)


def test_dp_ddim_schedule_equals_single_rank():
    """The strided (DDIM) schedule through sample_dp: sharded == unsharded, and different from the full-length loop."""
    B = 5
    shape = (B, 6, 5)
    kw = {"length": torch.arange(1, B + 1), "text": ["t%d" % i for i in range(B)]}
    single = parallel.sample_dp(lambda s, k: ToyStepper(s, k), shape, kw, num_timesteps=12, seed=3, schedule=SCHEDULE)
    full_len = parallel.sample_dp(lambda s, k: ToyStepper(s, k), shape, kw, num_timesteps=12, seed=3)
    assert not torch.equal(single, full_len)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, B, q, SCHEDULE)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert torch.equal(got, single)


def test_shard_range_covers_batch():
    for B in (1, 7, 64, 65):
        for W in (1, 2, 4, 8):
            spans = [parallel.shard_range(B, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def _ep_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    NB, E = 2, 8
    g = torch.Generator().manual_seed(100 + rank)
    mine = torch.randint(0, 400, (NB * E,), generator=g, dtype=torch.int32)
    table = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(table, mine)                       # what mdm_ep_counts does with peer stores
    cnt = [t.tolist() for t in table]
    ret.put((rank, cnt, plan_segments(cnt, NB, E, world)))
    dist.destroy_process_group()


def test_ep_plan_consistent_across_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ep_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, cnt0, plan0), (_, cnt1, plan1) = res
    assert cnt0 == cnt1 and plan0 == plan1               # every rank derives the same placement
    dest, seg_off, rows, tiles = plan0
    NB, E, R = 2, 8, 2
    for o in range(R):                                   # slots on owner o: unique, inside their segment
        taken = set()
        for g in range(NB * E):
            if owner_of(g % E, E, R) != o:
                continue
            for s in range(R):
                span = range(dest[s][g], dest[s][g] + cnt0[s][g])
                assert not taken.intersection(span)
                taken.update(span)
                assert span.start >= seg_off[g] and span.stop <= seg_off[g] + rows[g]
            assert seg_off[g] % 128 == 0
        assert tiles[o] * 128 >= len(taken)


@pytest.mark.parametrize("R", [1, 2, 4, 8])
def test_ep_plan_matches_single_gpu_layout(R):
    """R = 1 must reproduce the single-GPU segment layout (groups in order, each padded to 128)."""
    NB, E = 2, 8
    g = torch.Generator().manual_seed(R)
    cnt = torch.randint(0, 300, (R, NB * E), generator=g).tolist()
    dest, seg_off, rows, tiles = plan_segments(cnt, NB, E, R)
    if R == 1:
        off = 0
        for grp in range(NB * E):
            assert seg_off[grp] == off and dest[0][grp] == off
            off += (cnt[0][grp] + 127) // 128 * 128
    assert sum(tiles) * 128 == sum((r + 127) // 128 * 128 for r in rows)


@pytest.mark.parametrize("case", ["tiny_b3", "small_b4"])
def test_training_gradient_buckets_partition_the_flat_buffer(case):
    """Host logic of data-parallel training (training.TrainEngine.layout): the per-layer gradient buckets that
    DDPMTrainer.update all-reduces while the backward is still running, plus the rest, cover the flat gradient buffer exactly
    once; every parameter lies inside exactly one bucket; the stacked groups the kernels consume are contiguous."""
    import motiondiffusion_moe_b200 as mdm
    from motiondiffusion_moe_b200.training import TrainEngine
    from oracle import cases
    cfg, _ = cases.case_params(case)
    net = mdm.MotionTransformer(**cfg)
    eng = TrainEngine.__new__(TrainEngine)          # layout() is pure host logic: no CUDA device needed
    eng.m = net
    order, offset, numel, block_ranges, rest = eng.layout()
    assert sorted(order) == sorted(net._param_names) and len(block_ranges) == 2 * cfg.num_layers
    spans = sorted([r for rs in block_ranges for r in rs] + rest)
    assert spans[0][0] == 0 and spans[-1][1] == numel
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))                # disjoint, no gaps
    for li, blk in enumerate(net.block_prefixes()):
        for n in order:
            inside = [any(lo <= offset[n] and offset[n] + net._t(n).numel() <= hi for lo, hi in rs) for rs in block_ranges]
            if n.startswith(blk + ".") and ".emb_layers.1." not in n:
                assert inside[li] and sum(inside) == 1, n
    for n in order:                                                             # FiLM MLPs and globals: in the rest
        if ".emb_layers.1." in n or not n.startswith("decoder_blocks"):
            assert any(lo <= offset[n] and offset[n] + net._t(n).numel() <= hi for lo, hi in rest), n
    for g in eng._groups():                                                     # stacked groups are contiguous
        o = offset[g[0]]
        for n in g:
            assert offset[n] == o, n
            o += net._t(n).numel()
