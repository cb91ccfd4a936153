"""The CPU oracle, in its DEVICE arithmetic flavour, against golden vectors produced by the REFERENCE's own CUDA kernels on a B200
(tests/golden/make_golden_gpu.py: closest_triangle on the device, `warp`, kernel_grid; nvcc 12.9, sm_100): closest-triangle indices and
distances bit-exact, warp outputs within the stated tolerance, fp16 encodings bit-exact."""
import os

import numpy as np
import pytest

from insta_b200 import synth

POS_TOL, POS_TOL_BLEND, DIR_TOL = 1e-5, 1e-4, 1e-5


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stage12_refcuda.npz"))


@pytest.fixture()
def device_oracle(oracle):
    oracle.set_arith("device")
    yield oracle
    oracle.set_arith("host")


def test_closest_triangle_matches_reference_cuda(device_oracle, scene, gold):
    rec = scene.golden_records()
    for k in range(len(scene.nodes)):
        idx, dist = device_oracle.closest_triangle(rec[:, :3].copy(), scene.nodes[k], scene.tbvh[k])
        assert (idx == gold["closest_idx_f%d" % k]).all(), "frame %d: closest-triangle ids differ from the reference CUDA build" % k
        assert (dist.view(np.uint32) == gold["closest_dist_f%d" % k].view(np.uint32)).all()


def test_host_flavour_differs_somewhere(oracle, scene, gold):
    """the two flavours exist for a reason: the uncontracted host arithmetic picks another triangle for a few tie / near-tie queries"""
    rec = scene.golden_records()
    oracle.set_arith("host")
    diff = 0
    for k in range(len(scene.nodes)):
        idx, dist = oracle.closest_triangle(rec[:, :3].copy(), scene.nodes[k], scene.tbvh[k])
        diff += int((dist.view(np.uint32) != gold["closest_dist_f%d" % k].view(np.uint32)).sum())
    assert diff > 0


def test_warp_matches_reference_cuda(device_oracle, scene, gold):
    rec = scene.golden_records()
    for name, di in (("interp", False), ("nointerp", True)):
        r, tri = scene.oracle_warp(device_oracle, rec, disable_interp=di)
        g = gold["warp_%s_rec" % name]
        tol = POS_TOL if di else POS_TOL_BLEND
        assert np.abs(r[:, 0:3] - g[:, 0:3]).max() < tol
        assert np.abs(r[:, 8:11] - g[:, 3:6]).max() < DIR_TOL
        assert (r[:, 11] == g[:, 6]).all() and (r[:, 26] == g[:, 7]).all()


def test_kernel_grid_encodings_bit_exact(oracle, gold):
    offs, scales, res = oracle.grid_layout()
    dev_scales = gold["grid_scales_device"]
    assert np.abs(scales.view(np.int32) - dev_scales.view(np.int32)).max() <= 1
    n = gold["grid_enc"].shape[1]
    p = np.random.default_rng(2).random((n, 3), dtype=np.float32)
    p[:8] = np.array([[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0.999999, 0.5, 0.25], [1e-7, 1e-7, 1e-7]], np.float32)
    table = synth.make_grid_table(int(offs[-1]), seed=1337)
    enc = oracle.grid_forward(p, 3, (offs, dev_scales, res), table)
    assert (enc.view(np.uint16) == gold["grid_enc"]).all()
