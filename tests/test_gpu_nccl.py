"""The exchange step of training on real GPUs: two ranks, one process per GPU, NCCL.  Each rank runs ONE insta_train_nerf step with frozen
parameters on its own ray batch (its own RNG seed); the gradient vector both ranks hold afterwards must equal the mean of the gradients two single-rank steps
produce for the same two batches -- i.e. the 1-rank step on the union batch (every rank pre-scales its loss by 1 / world, trainer.cu).  Skipped with fewer than
two GPUs; the CPU logic of the sharding rule is covered by tests/test_shard_gloo.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
WORKER = os.path.join(HERE, "nccl_train_worker.py")


def _spawn(rank, world, seed, out_dir, exchange=None):
    env = dict(os.environ)
    env.pop("INSTA_GRAD_EXCHANGE", None)
    if exchange:
        env["INSTA_GRAD_EXCHANGE"] = exchange
        env["EXPECT_EXCHANGE_MODE"] = "2" if exchange == "p2p" else "1"
    return subprocess.Popen([sys.executable, WORKER, str(rank), str(world), str(seed), str(out_dir)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)


@pytest.mark.parametrize("exchange", ["serial", "p2p"])
def test_two_rank_gradient_sum_equals_single_rank_on_the_union_batch(tmp_path, exchange):
    """serial = one ncclAllReduce after the backward pass; p2p = this library's own reduce kernel over peer memory (trainer.cu peer_exchange_run)"""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    seeds = (101, 202)
    procs = [_spawn(r, 2, seeds[r], tmp_path, exchange) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    g2 = [np.load(tmp_path / ("grads_rank%d_world2_seed%d.npy" % (r, seeds[r]))).astype(np.float32) for r in range(2)]
    assert np.array_equal(g2[0], g2[1]), "the ranks disagree after the allreduce"
    singles = []
    for sd in seeds:
        p = _spawn(0, 1, sd, tmp_path)
        out = p.communicate(timeout=600)[0]
        assert p.returncode == 0, out
        singles.append(np.load(tmp_path / ("grads_rank0_world1_seed%d.npy" % sd)).astype(np.float32))
    want = 0.5 * (singles[0] + singles[1])
    assert np.abs(want).max() > 0 and not np.array_equal(singles[0], singles[1])
    # fp16 gradients: each rank rounds its pre-scaled contribution, NCCL adds in fp16 (the peer-memory kernel in fp32, rounding once); scatter-adds are order-dependent within a rank
    mlp = slice(0, 25600)
    cos = float((g2[0][mlp] * want[mlp]).sum() / (np.linalg.norm(g2[0][mlp]) * np.linalg.norm(want[mlp])))
    assert cos > 0.9995, cos
    assert abs(np.linalg.norm(g2[0][mlp]) / np.linalg.norm(want[mlp]) - 1) < 0.01
    grid_got, grid_want = g2[0][25600:], want[25600:]
    # the union of the two batches' entries -- except where a rank's halved contribution underflows fp16 (pre-scaling by 1 / world): those are tiny
    touched = (grid_got != 0) | (grid_want != 0)
    lost = (grid_got == 0) & (grid_want != 0)
    assert lost.sum() < 0.05 * touched.sum() and (not lost.any() or np.abs(grid_want[lost]).max() < 1e-3 * np.abs(grid_want).max())
    assert ((grid_got != 0) & (grid_want == 0)).sum() < 2e-3 * touched.sum()  # exact cancellation in fp32 that the two fp16 halves miss (0.09 - 0.12 % of the entries, run to run)
    cosg = float((grid_got * grid_want).sum() / (np.linalg.norm(grid_got) * np.linalg.norm(grid_want)))
    assert cosg > 0.995, cosg
