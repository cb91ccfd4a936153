"""Host-side restatement of the trainer's lazily evaluated rounds (insta_b200/csrc/train.cu train_round_kernel): for random per-sample
alphas, the samples the rounds evaluate always contain every sample the loss loop reads (all samples up to and including the last one it consumes before
T < 1e-4), for any round boundaries -- and they are far fewer than `all of them` when rays become opaque early.  The GPU test of the real kernels is
tests/test_gpu_trainer.py::test_lazy_rounds_equal_full_evaluation."""
import numpy as np
import pytest

EPS = np.float32(1e-4)


def loss_loop_reads(alpha):
    """compute_loss_kernel_train_nerf's forward loop (ngp testbed_nerf.cu:1379-1401): the test T < EPSILON comes BEFORE a sample is consumed"""
    T = np.float32(1.0)
    n = 0
    for a in alpha:
        if T < EPS:
            break
        T = np.float32(T * np.float32(np.float32(1.0) - a))
        n += 1
    return n


def rounds_evaluate(alpha, edges):
    """which samples the rounds evaluate: round r covers [edges[r], edges[r+1]); a ray contributes to round r only if the replayed recurrence over the
    samples of the earlier rounds has neither stopped nor dropped below the threshold, and it still has samples"""
    n = len(alpha)
    evaluated = np.zeros(n, bool)
    T = np.float32(1.0)
    for r in range(len(edges) - 1):
        lo, hi = edges[r], min(edges[r + 1], n)
        if r > 0:
            plo, phi = edges[r - 1], min(edges[r], n)
            stop = False
            for j in range(plo, phi):
                assert evaluated[j]  # the replay only reads evaluated samples
                if T < EPS:
                    stop = True
                    break
                T = np.float32(T * np.float32(np.float32(1.0) - alpha[j]))
            if stop or T < EPS or phi >= n:
                break
        elif n == 0:
            break
        evaluated[lo:hi] = True
    return evaluated


@pytest.mark.parametrize("edges", [(0, 8, 32, 1 << 32), (0, 4, 16, 1 << 32), (0, 1, 2, 3, 1 << 32), (0, 16, 1 << 32), (0, 1 << 32)])
def test_rounds_cover_what_the_loss_reads(edges):
    rng = np.random.default_rng(len(edges))
    total = used = evaluated_total = 0
    for _ in range(3000):
        n = int(rng.integers(0, 80))
        opacity = rng.choice([0.0, 0.02, 0.3, 0.9, 1.0])
        alpha = (rng.random(n) * opacity).astype(np.float32)
        if n and rng.random() < 0.2:
            alpha[rng.integers(0, n)] = np.float32(1.0)  # a fully opaque sample: T becomes exactly 0
        need = loss_loop_reads(alpha)
        ev = rounds_evaluate(alpha, edges)
        assert ev[:need].all(), (n, need, np.flatnonzero(~ev[:need]))
        # ... and nothing beyond the end of the round in which the ray stopped
        if need < n:
            r_stop = max(k for k in range(len(edges) - 1) if edges[k] <= max(need - 1, 0))
            last = min(edges[r_stop + 1], n)
            # a ray that drops below the threshold exactly at a round boundary still contributes nothing to the next round
            assert not ev[last:].any()
        total += n; used += need; evaluated_total += int(ev.sum())
    assert used <= evaluated_total <= total
    if edges[:3] == (0, 8, 32):  # the default boundaries skip a good part of this mix of transparent and opaque rays
        assert evaluated_total < 0.8 * total
