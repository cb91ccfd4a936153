"""The CPU oracle against golden vectors produced by the REFERENCE's own host code (tests/golden/make_golden.py):
BVH topology + leaf order, closest-triangle indices and distances (bit-exact), warp outputs (tolerance)."""
import hashlib

import numpy as np

from common import UNIT_AABB

POS_TOL = 1e-5       # canonical position, non-blended and blended, vs the reference's fp32 Eigen path (unit-cube coords)
POS_TOL_BLEND = 1e-4
DIR_TOL = 1e-5


def sha(a):
    return np.frombuffer(hashlib.sha256(a.tobytes()).digest(), np.uint8)


def test_bvh_build_matches_reference(oracle, scene, golden):
    assert (sha(scene.canon_nodes) == golden["canon_nodes_sha"]).all()
    assert (scene.canon_bvh["id"] == golden["canon_order"]).all()
    for k in range(len(scene.nodes)):
        assert len(scene.nodes[k]) == int(golden["frame%d_n_nodes" % k])
        assert (sha(scene.nodes[k]) == golden["frame%d_nodes_sha" % k]).all()
        assert (scene.tbvh[k]["id"] == golden["frame%d_order" % k]).all()


def test_closest_triangle_bit_exact(oracle, scene, golden):
    rec = scene.golden_records()
    idx, dist = oracle.closest_triangle(rec[:, :3].copy(), scene.canon_nodes, scene.canon_bvh)
    assert (idx == golden["canon_closest_idx"]).all()
    assert (dist.view(np.uint32) == golden["canon_closest_dist"].view(np.uint32)).all()


def test_warp_matches_reference(oracle, scene, golden):
    rec = scene.golden_records()
    for name, di in (("interp", False), ("nointerp", True)):
        r, tri = scene.oracle_warp(oracle, rec, disable_interp=di)
        assert (tri == golden["warp_%s_tri" % name]).all()
        g = golden["warp_%s_rec" % name]
        tol = POS_TOL if di else POS_TOL_BLEND
        assert np.abs(r[:, 0:3] - g[:, 0:3]).max() < tol
        assert np.abs(r[:, 8:11] - g[:, 3:6]).max() < DIR_TOL
        assert (r[:, 11] == g[:, 6]).all() and (r[:, 26] == g[:, 7]).all()  # expression dims forced to 1 outside Lips
        skipped = tri < 0
        assert skipped.sum() > 0 and (r[skipped] == rec[skipped]).all()


def test_render_and_points_paths(oracle, scene, golden):
    rec = scene.golden_records()
    r, tri = scene.oracle_warp(oracle, rec, frame_id=1)
    assert (tri == golden["render_tri"]).all()
    assert np.abs(r[:, [0, 1, 2, 8, 9, 10]] - golden["render_rec"]).max() < POS_TOL_BLEND
    xyz = rec[:, :3].copy()
    fids = (np.arange(len(xyz)) % len(scene.nodes)).astype(np.int32)
    x, tri = scene.oracle_warp(oracle, xyz, stride=3, frame_ids=fids)
    assert (tri == golden["points_tri"]).all()
    assert np.abs(x.reshape(-1, 3) - golden["points_xyz"]).max() < POS_TOL_BLEND


def test_mask_counts(scene):
    # per-region triangle counts of the shipped assets
    u, c = np.unique(scene.masks, return_counts=True)
    assert dict(zip(u.tolist(), c.tolist())) == {0: 276, 1: 896, 2: 2036, 4: 1154, 7: 319, 10: 1149, 15: 2260}
