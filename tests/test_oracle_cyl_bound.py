"""The traversal kernel prunes a quad child whose CYLINDER bound (insta_b200/csrc/deform.cu: build_cyl_kernel, cyl_cull) exceeds the best distance so far,
in addition to the reference's box test (ngp bvh.cuh:48-96).  The argument: such a subtree holds no triangle the reference would accept
(d <= best), so `best` evolves as in the reference and the winner -- ties included -- is unchanged.  This checks the argument with the CPU restatement
of the reference traversal plus the restated bound: (a) the winners equal the plain traversal's for random, out-of-box, on-surface and tie-heavy points in
both arithmetic flavours, (b) every culled subtree is searched exhaustively and never holds an acceptable triangle, (c) the cost model quoted in DESIGN.md.
(The GPU kernel is compared with the reference's own CUDA kernels in tests/test_gpu_vs_ref_cuda.py.)"""
import numpy as np
import pytest

from test_oracle_seed_bound import bound, seed_grid, surface_points


@pytest.mark.parametrize("arith", ["host", "device"])
def test_cylinder_cull_returns_the_reference_triangle(oracle, scene, arith):
    from insta_b200 import synth

    oracle.set_arith(arith)
    try:
        seeds, lo, hi = seed_grid(scene.canon, scene.aabb)
        rng = np.random.default_rng(9)
        total = 0
        for f in range(len(scene.frame_verts)):
            cyl = oracle.cyl_build(scene.nodes[f], scene.tbvh[f])
            assert (cyl[1:, 4] != 0x7C00).mean() > 0.99  # the synthetic frames have no slivers: (almost) every node has a finite cylinder
            rec = synth.make_records(40000, scene.aabb, np.zeros(40000, np.int32), seed=77 + f, pad_fraction=0.0)
            far = np.ascontiguousarray(rec[:, :3])
            outside = (far + (far - 0.5) * np.float32(3.0)).astype(np.float32)[:8000]
            near = surface_points(scene.ttopo[f], rng, 30000)
            for name, pts in (("far", far), ("outside", outside), ("near", near)):
                want, _ = oracle.closest_triangle(pts, scene.nodes[f], scene.tbvh[f])
                U = bound(pts, seeds, lo, hi, scene.ttopo[f])
                for init in (U, None):  # seeded like the kernel, and from FLT_MAX (no cull until a first triangle is found)
                    got, st = oracle.closest_triangle_cyl(pts, scene.nodes[f], scene.tbvh[f], cyl, init, check=True)
                    assert st[3] == 0, (name, st)  # no culled subtree held an acceptable triangle
                    assert np.array_equal(got, want), (name, int((got != want).sum()))
                total += len(pts)
        assert total >= 150000
    finally:
        oracle.set_arith("host")


def test_cylinder_cost_model(oracle, scene):
    """the step counts quoted in DESIGN.md 3.1 for U(aabb) points"""
    from insta_b200 import synth

    seeds, lo, hi = seed_grid(scene.canon, scene.aabb)
    rec = synth.make_records(50000, scene.aabb, np.zeros(50000, np.int32), seed=31, pad_fraction=0.0)
    pts = np.ascontiguousarray(rec[:, :3])
    U = bound(pts, seeds, lo, hi, scene.ttopo[0])
    cyl = oracle.cyl_build(scene.nodes[0], scene.tbvh[0])
    inf = cyl.copy()
    inf[:] = np.array([0, 0, 0x3C00, 0, 0x7C00, 0x7C00, 0, 0], np.uint16)
    _, base = oracle.closest_triangle_cyl(pts, scene.nodes[0], scene.tbvh[0], inf, U)
    _, st = oracle.closest_triangle_cyl(pts, scene.nodes[0], scene.tbvh[0], cyl, U)
    n = len(pts)
    assert base[2] == 0
    assert 21 < base[0] / n < 25 and 14 < base[1] / n < 17
    assert st[0] / n < 19 and st[1] / n < 8.5
