"""Expert-parallel MoE (csrc/ep.cu, expert_parallel.py).  On one GPU the R ranks are emulated inside one
process (all peer pointers local, phases run rank by rank): the result must equal the local MoE kernels
BIT FOR BIT (same gate arithmetic, row-wise independent GEMMs, same combine order).  With >= 2 GPUs the
same check runs with one process per GPU over CUDA IPC / NVLink (tests/ep_dist_worker.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0) if torch.cuda.is_available() else None

from motiondiffusion_moe_b200.expert_parallel import ExpertParallelFFN, run_emulated  # noqa: E402
from ep_common import make_weights, make_tokens, local_moe  # noqa: E402


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("R,E,n_seq,T", [(2, 8, 3, 98), (4, 8, 2, 60), (8, 8, 1, 196), (2, 4, 5, 40), (1, 8, 2, 98)])
def test_emulated_ranks_match_local_moe_bit_exact(dtype, R, E, n_seq, T):
    D, Fd = 256, 512
    w = make_weights(D, Fd, E, dtype, DEV)
    toks = [make_tokens(n_seq, T, D, DEV, 10 + r) for r in range(R)]
    # reference: each rank's tokens through the local (all experts on one GPU) path
    refs = [local_moe(w, x, film, T, D, Fd, E, dtype) for x, film in toks]
    inst = ExpertParallelFFN.create_emulated(R, D, Fd, E, 2, n_seq * T, dtype, DEV)
    for i in inst:
        i.set_weights(w["ln_w"], w["ln_b"], w["gate_w"], w["gate_b"], w["w1"], w["b1"], w["w2"], w["b2"])
    outs = [torch.empty(n_seq * T, D, device=DEV, dtype=dtype) for _ in range(R)]
    for _ in range(2):      # twice: buffers and counts tables are reused across calls
        run_emulated(inst, [x for x, _ in toks], w["s_w"], w["s_b"], [f for _, f in toks], T, outs)
    torch.cuda.synchronize()
    for r in range(R):
        inst[r].check_health()
        res, idx, vals, usage, importance = refs[r]
        assert torch.equal(inst[r].idx, idx) and torch.equal(inst[r].vals, vals)      # routing bit-exact
        assert torch.equal(outs[r], res)                                               # output bit-exact
        assert torch.equal(inst[r].usage, 2 * usage)                                   # two calls accumulated
        assert torch.allclose(inst[r].importance, 2 * importance, rtol=1e-5)   # different (fixed) summation orders
    # every owner received exactly the rows routed to its experts
    total_rows = sum(int(i.cnt[i.me].sum()) for i in inst)
    assert total_rows == R * n_seq * T * 4


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_distributed_ranks_match_local_moe(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
           "--master-addr", "127.0.0.1", "--master-port", "29%03d" % (500 + world), os.path.join(here, "ep_dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "EP_DIST_OK" in r.stdout


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("world", [2, 4])
def test_model_with_expert_parallelism_is_bit_identical(world):
    """MotionTransformer.enable_expert_parallel(): forward, routing, counters and graph-captured CFG steps."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
           "--master-addr", "127.0.0.1", "--master-port", "29%03d" % (600 + world), os.path.join(here, "ep_model_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "EP_MODEL_OK" in r.stdout
