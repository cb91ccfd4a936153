"""The mesh ray cast of rta::Core::raycast_flame (src/core.cu:350-402 -> TriangleBvh4::ray_intersect, ngp src/triangle_bvh.cu:271-312): the CPU restatement
(oracle.cpp ray_tri) against the REFERENCE's own host-compiled code (oracle/_ref/libref_host.so: bvh.cuh:98-144 with Triangle::ray_intersect and
BoundingBox::ray_intersect) -- first-hit triangle and ray parameter bit for bit, on random, tie-heavy, axis-parallel, inside-out and missing rays."""
import numpy as np
import pytest

from oracle import pyoracle
from raycast_common import make_rays

needs_ref = pytest.mark.skipif(not pyoracle.have_ref_host(), reason="oracle/_ref/libref_host.so not built")


@needs_ref
@pytest.mark.parametrize("frame", [0, 1])
def test_ray_intersect_bit_exact_vs_reference_host_code(oracle, scene, frame):
    ref = pyoracle.RefHost()
    ro, rd = make_rays(scene, frame)
    i1, t1 = oracle.ray_intersect(ro, rd, scene.nodes[frame], scene.tbvh[frame])
    i2, t2 = ref.ray_intersect(ro, rd, scene.nodes[frame], scene.tbvh[frame])
    assert (i1 == i2).all(), int((i1 != i2).sum())
    assert (t1.view(np.uint32) == t2.view(np.uint32)).all()
    hit = i1 >= 0
    assert 0.3 < hit.mean() < 0.95                      # both hits and misses are exercised
    assert (t1[~hit] == np.float32(10.0)).all()         # MAX_DIST when nothing is hit
    assert (t1[hit] >= 0).all() and (t1[hit] < 10).all()
    # the tie-heavy block really produces ties: rays through shared vertices hit several triangles at (nearly) the same parameter
    assert hit[20000:20000 + 2000].mean() > 0.3


@pytest.mark.parametrize("frame", [0, 1])
def test_ray_intersect_against_committed_reference_outputs(oracle, scene, frame):
    """The same pin without the reference build: tests/golden/raycast_ref.npz holds what the reference's host code returned for these rays
    (tests/golden/make_golden_raycast.py) -- topology id of the first hit and the ray parameter, bit for bit."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "raycast_ref.npz"))
    ro, rd = make_rays(scene, frame, n_random=4000)
    idx, t = oracle.ray_intersect(ro, rd, scene.nodes[frame], scene.tbvh[frame])
    ids = np.where(idx >= 0, scene.tbvh[frame]["id"][np.maximum(idx, 0)].astype(np.int64), -1)
    assert (ids == g["tri_id_%d" % frame]).all()
    assert (t.view(np.uint32) == g["t_%d" % frame].view(np.uint32)).all()
