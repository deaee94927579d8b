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


def _run_workers(script, world, port, timeout, token):
    """One process per GPU (NCCL) when the box has `world` GPUs; world == 2 on a single-GPU box: both ranks share
    cuda:0 (gloo plumbing, CUDA IPC between the two processes: tests/dist_common.py), so that the multi-rank paths run
    under the driver's 1-GPU test tier too.  Larger worlds need the GPUs."""
    ngpu = torch.cuda.device_count()
    env = dict(os.environ)
    if ngpu < world:
        if world != 2:
            pytest.skip("needs %d GPUs" % world)
        env["MDM_TEST_SHARED_GPU"] = "1"
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(here, script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    print(r.stdout[-2000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert token in r.stdout


@pytest.mark.parametrize("world", [2, 4, 8])
def test_distributed_ranks_match_local_moe(world):
    """EP parity compares the expert-parallel path with the repo's own single-GPU MoE kernels on the same tokens; that
    is a valid oracle check because the single-GPU MoE path is itself checked against oracle.switch_moe
    (tests/test_ops_gpu.py::test_moe_multibranch) and the whole model against the reference goldens."""
    _run_workers("ep_dist_worker.py", world, 29500 + world, 600, "EP_DIST_OK")


@pytest.mark.parametrize("world", [2, 4])
def test_model_with_expert_parallelism_is_bit_identical(world):
    """MotionTransformer.enable_expert_parallel(): forward, routing, counters and graph-captured CFG steps."""
    _run_workers("ep_model_worker.py", world, 29600 + world, 900, "EP_MODEL_OK")


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_cfg_sampling_equals_single_gpu_bit_for_bit(world):
    """parallel.sample_dp on the real model: N-rank sharded CFG sampling == the unsharded loop, bit for bit."""
    _run_workers("dp_model_worker.py", world, 29700 + world, 900, "DP_MODEL_OK")


@pytest.mark.parametrize("world", [2, 4])
def test_data_parallel_training_keeps_replicas_identical(world):
    """DDPMTrainer.update with torch.distributed initialised: flat gradient all-reduce (average) before clip + Adam."""
    _run_workers("dp_train_worker.py", world, 29800 + world, 600, "DP_TRAIN_OK")
