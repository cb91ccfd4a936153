"""Direct comparison of the oracle with the reference's own host code at larger sizes.  Only runs where
oracle/_ref/libref_host.so exists (built from /root/reference by `make -C oracle ref`)."""
import numpy as np
import pytest

from oracle import pyoracle

pytestmark = pytest.mark.skipif(not pyoracle.have_ref_host(), reason="oracle/_ref/libref_host.so not built")


@pytest.fixture(scope="module")
def ref():
    return pyoracle.RefHost()


def test_distance_primitives_bit_exact(oracle, ref, scene):
    rng = np.random.default_rng(3)
    pts = (scene.aabb[:3] + (scene.aabb[3:] - scene.aabb[:3]) * rng.random((2000, 3))).astype(np.float32)
    for t in scene.canon[rng.integers(0, len(scene.canon), 25)]:
        mine = np.array([oracle.triangle_distance_sq(t, p) for p in pts[:200]], np.float32)
        assert (mine.view(np.uint32) == ref.triangle_distance_sq(t, pts[:200]).view(np.uint32)).all()
    for n in scene.canon_nodes[rng.integers(0, len(scene.canon_nodes), 25)]:
        mine = np.array([oracle.bbox_distance_sq(n["mn"], n["mx"], p) for p in pts[:200]], np.float32)
        assert (mine.view(np.uint32) == ref.bbox_distance_sq(n["mn"], n["mx"], pts[:200]).view(np.uint32)).all()


def test_bvh_and_closest_200k(oracle, ref, scene):
    rn, rt = ref.bvh_build(scene.ttopo[0])
    assert rn.tobytes() == scene.nodes[0].tobytes() and rt.tobytes() == scene.tbvh[0].tobytes()
    rng = np.random.default_rng(42)
    pts = (scene.aabb[:3] + (scene.aabb[3:] - scene.aabb[:3]) * rng.random((200000, 3))).astype(np.float32)
    # add points exactly on vertices / edge midpoints / centroids: the tie-heavy cases
    v = scene.frame_verts[0]
    f = scene.faces[:2000]
    special = np.concatenate([v[:3000], (v[f[:, 0]] + v[f[:, 1]]) * np.float32(0.5), (v[f[:, 0]] + v[f[:, 1]] + v[f[:, 2]]) / np.float32(3)])
    pts = np.concatenate([pts, special.astype(np.float32)])
    i1, d1 = oracle.closest_triangle(pts, scene.nodes[0], scene.tbvh[0])
    i2, d2 = ref.closest_triangle(pts, rn, rt)
    assert (i1 == i2).all()
    assert (d1.view(np.uint32) == d2.view(np.uint32)).all()


def test_tri2projection(oracle, ref, scene):
    for t in scene.ttopo[0][[0, 100, 5000, 8089]]:
        assert np.abs(oracle.tri2projection(t) - ref.tri2projection(t)).max() < 1e-6
