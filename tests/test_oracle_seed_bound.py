"""The deformation kernels start a closest-triangle traversal from a finite upper bound (distance to the centroid of a triangle near the query's cell,
insta_b200/csrc/deform.cu seed_bound) where the reference passes max_distance_sq = FLT_MAX (ngp bvh.cuh:48-96; src/core.cu:480).  This checks the
argument of DESIGN.md 3.1 with the CPU restatement of the reference traversal itself: started from that bound it returns the SAME triangle -- for
random points, for points on / next to the surface where equidistant triangles (shared edges, shared vertices) are the rule, in both arithmetic
flavours -- and it returns nothing only if the bound is below the true minimum, which the inflation excludes.  (The GPU kernels are compared with the
reference's own CUDA kernels in tests/test_gpu_vs_ref_cuda.py.)"""
import numpy as np
import pytest

GRID = 32


def seed_grid(canon, aabb):
    """numpy restatement of build_seed_grid_kernel: per cell the canonical triangle whose centroid is nearest to the cell centre"""
    from scipy.spatial import cKDTree

    cent = ((canon["a"] + canon["b"] + canon["c"]) / np.float32(3.0)).astype(np.float32)
    lo, hi = np.asarray(aabb[:3], np.float32), np.asarray(aabb[3:], np.float32)
    g = np.stack(np.meshgrid(*[(np.arange(GRID, dtype=np.float32) + 0.5) / GRID for _ in range(3)], indexing="ij"), -1).reshape(-1, 3) * (hi - lo) + lo
    return cKDTree(cent).query(g)[1].astype(np.int64), lo, hi


def bound(pts, seeds, lo, hi, frame_tris_topo):
    cell = np.clip(((pts - lo) * (np.float32(GRID) / (hi - lo))).astype(np.int64), 0, GRID - 1)
    s = seeds[(cell[:, 0] * GRID + cell[:, 1]) * GRID + cell[:, 2]]
    c = ((frame_tris_topo["a"] + frame_tris_topo["b"] + frame_tris_topo["c"]) / np.float32(3.0)).astype(np.float32)[s]
    d = ((pts - c) ** 2).sum(1).astype(np.float32)
    return (d * np.float32(1.0001) + np.float32(1e-9)).astype(np.float32)


def surface_points(tris, rng, n):
    """vertices, edge midpoints, centroids and points a hair off the surface: distance ~ 0, ties between the triangles of a shared edge / vertex"""
    t = tris[rng.integers(0, len(tris), n)]
    a, b, c = t["a"], t["b"], t["c"]
    kind = rng.integers(0, 5, n)[:, None]
    w = rng.random((n, 3)).astype(np.float32)
    w /= w.sum(1, keepdims=True)
    inside = w[:, :1] * a + w[:, 1:2] * b + w[:, 2:] * c
    p = np.where(kind == 0, a, np.where(kind == 1, (a + b) * np.float32(0.5), np.where(kind == 2, (a + b + c) / np.float32(3.0), inside)))
    off = np.where(kind == 4, rng.normal(size=(n, 3)) * 1e-5, 0.0)
    return np.ascontiguousarray((p + off).astype(np.float32))


@pytest.mark.parametrize("arith", ["host", "device"])
def test_seeded_traversal_returns_the_reference_triangle(oracle, scene, arith):
    from insta_b200 import synth

    oracle.set_arith(arith)
    try:
        seeds, lo, hi = seed_grid(scene.canon, scene.aabb)
        rng = np.random.default_rng(5)
        total = 0
        for f in range(len(scene.frame_verts)):
            rec = synth.make_records(60000, scene.aabb, np.zeros(60000, np.int32), seed=31 + f, pad_fraction=0.0)
            far = np.ascontiguousarray(rec[:, :3])
            outside = (far + (far - 0.5) * np.float32(0.8)).astype(np.float32)[:10000]  # beyond the grid: clamped cells, still a valid bound
            near = surface_points(scene.ttopo[f], rng, 40000)
            for pts in (far, outside, near):
                want, dist = oracle.closest_triangle(pts, scene.nodes[f], scene.tbvh[f])
                U = bound(pts, seeds, lo, hi, scene.ttopo[f])
                assert (U >= dist.astype(np.float64) ** 2).all()  # the bound is never below the minimum ...
                got = oracle.closest_triangle_bounded(pts, scene.nodes[f], scene.tbvh[f], U)
                assert (got >= 0).all()
                assert np.array_equal(got, want)  # ... and the traversal returns the reference's triangle, ties included
                total += len(pts)
        assert total >= 200000
    finally:
        oracle.set_arith("host")


def test_bound_below_the_minimum_finds_nothing(oracle, scene):
    """why the inflation matters: max_distance_sq below the true minimum prunes everything (the reference would return -1 as well)"""
    pts = np.array([[0.5, 0.9, 0.5]], np.float32)
    want, dist = oracle.closest_triangle(pts, scene.nodes[0], scene.tbvh[0])
    assert want[0] >= 0 and dist[0] > 0.05
    got = oracle.closest_triangle_bounded(pts, scene.nodes[0], scene.tbvh[0], np.array([dist[0] ** 2 * 0.5], np.float32))
    assert got[0] == -1
