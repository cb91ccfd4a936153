"""Frame-cache turnover (rta::Core::reload_training_data, src/core.cu:663-706 with load_meshes 294-334): the meshes of resident frames
are replaced between training steps.  insta_replace_frames rebuilds on the host in the reference's order, so closest-triangle ids stay bit-exact against
the oracle for the NEW meshes while untouched frames are unaffected; insta_refit_frames (GPU refit, not in the reference) must still find a nearest
triangle (same distance) and may differ only where triangles are exactly equidistant -- the rate is printed and bounded."""
import numpy as np
import pytest

from common import UNIT_AABB

pytestmark = pytest.mark.gpu


def _deform(ctx, rec):
    import torch

    d = torch.from_numpy(rec.copy()).cuda()
    tri = torch.full((rec.shape[0],), -7, dtype=torch.int32, device="cuda")
    ctx.surface_closest_point(d, tri)
    torch.cuda.synchronize()
    return d.cpu().numpy(), tri.cpu().numpy()


def _oracle(oracle, scene, verts_list, rec):
    canon = scene.canon
    nodes, tbvh, ttopo = [], [], []
    for v in verts_list:
        t = oracle.build_triangles(v, scene.faces, scene.masks, canon)
        n, tb = oracle.bvh_build(t)
        nodes.append(n); tbvh.append(tb); ttopo.append(t)
    r = rec.copy()
    tri, _ = oracle.warp(r, 27, -1, nodes, tbvh, ttopo, canon, scene.adj, UNIT_AABB)
    return r, tri, (nodes, tbvh, ttopo)


@pytest.fixture(scope="module")
def many_frames(scene):
    from insta_b200 import synth

    fs = synth.FrameSynth(scene.verts, scene.faces)
    return [fs.frame(f) for f in range(48)], [fs.frame(1000 + f) for f in range(48)]


def test_replace_frames_mid_training_keeps_ids_bit_exact(oracle, scene, many_frames):
    from insta_b200 import capi, synth

    oracle.set_arith("device")
    try:
        old, new = many_frames
        ctx = capi.Context(device=0)
        ctx.set_canonical(scene.verts, scene.faces, scene.masks)
        ctx.add_frames(np.stack(old))
        n = 24000
        fids = synth.train_regime_frame_ids(n, len(old))
        rec = synth.make_records(n, scene.aabb, fids, seed=5, pad_fraction=0.01)
        want0, tri0, _ = _oracle(oracle, scene, old, rec)
        got0, t0 = _deform(ctx, rec)
        assert (t0 == tri0).all()
        # swap a third of the frames (every third one) for new meshes
        idx = np.arange(0, len(old), 3, dtype=np.int32)
        ctx.replace_frames(idx, np.stack([new[i] for i in idx]), n_threads=4)
        assert ctx.n_frames == len(old)
        mixed = [new[i] if i % 3 == 0 else old[i] for i in range(len(old))]
        want1, tri1, _ = _oracle(oracle, scene, mixed, rec)
        got1, t1 = _deform(ctx, rec)
        assert (t1 == tri1).all(), "ids differ after insta_replace_frames at %s" % np.nonzero(t1 != tri1)[0][:10]
        assert np.abs(got1[:, 0:3] - want1[:, 0:3]).max() < 1e-4
        untouched = (fids % 3) != 0
        assert (t1[untouched] == t0[untouched]).all() and (got1[untouched] == got0[untouched]).all()
        assert (t1[~untouched] != t0[~untouched]).any()  # the new meshes really are different
        # errors
        with pytest.raises(capi.InstaError):
            ctx.replace_frames(np.array([0, 0], np.int32), np.stack([new[0], new[1]]))
        with pytest.raises(capi.InstaError):
            ctx.replace_frames(np.array([len(old)], np.int32), np.stack([new[0]]))
        ctx.close()
    finally:
        oracle.set_arith("host")


def test_refit_frames_finds_a_nearest_triangle(oracle, scene, many_frames):
    """GPU refit (non-parity): same minimum distance as the host-built tree of the new mesh; ids differ only between equidistant triangles."""
    import torch
    from insta_b200 import capi, synth

    oracle.set_arith("device")
    try:
        old, new = many_frames
        k = 8
        ctx = capi.Context(device=0)
        ctx.set_canonical(scene.verts, scene.faces, scene.masks)
        ctx.add_frames(np.stack(old[:k]))
        ctx.refit_frames(np.arange(k, dtype=np.int32), np.stack(new[:k]))
        n = 40000
        fids = synth.train_regime_frame_ids(n, k)
        rec = synth.make_records(n, scene.aabb, fids, seed=9, pad_fraction=0.0)
        got, tri = _deform(ctx, rec)
        want, want_tri, (nodes, tbvh, ttopo) = _oracle(oracle, scene, new[:k], rec)
        assert (tri >= 0).all()
        diff = tri != want_tri
        rate = diff.mean()
        print("refit: %d of %d ids differ from the host-built tree (%.4f %%)" % (diff.sum(), n, 100 * rate))
        # For points spread over the render box most nearest features are mesh vertices or edges, i.e. several triangles at the same distance: the
        # winner among them follows the visit order, hence the tree (measured: ~23 % of U(aabb) queries).  What must hold: wherever the ids differ, the
        # two triangles are equidistant from the query (same fp32 distance up to rounding).
        assert rate < 0.5
        for i in np.nonzero(diff)[0][:3000]:
            f = fids[i]
            p = rec[i, :3]
            da = oracle.triangle_distance_sq(ttopo[f][tri[i]], p)
            db = oracle.triangle_distance_sq(ttopo[f][want_tri[i]], p)
            assert abs(da - db) <= 1e-5 * max(da, db) + 1e-12, (i, da, db)
        same = ~diff
        assert np.abs(got[same, 0:3] - want[same, 0:3]).max() < 1e-4
        ctx.close()
    finally:
        oracle.set_arith("host")


def _frame_tris(scene, verts):
    """a,b,c of every triangle of a frame (structured like the oracle's triangle records)"""
    t = np.zeros(len(scene.faces), dtype=[("a", np.float32, 3), ("b", np.float32, 3), ("c", np.float32, 3)])
    t["a"], t["b"], t["c"] = verts[scene.faces[:, 0]], verts[scene.faces[:, 1]], verts[scene.faces[:, 2]]
    return t


def _node_to_quad_child(nodes):
    """(quad, child) of every non-root node of the reference tree: quads are the inner nodes in node order (mesh_host.cpp build_bvh)"""
    inner = np.nonzero(nodes["left_idx"] >= 0)[0]
    where = {}
    for q, i in enumerate(inner):
        for k in range(4):
            where[int(nodes["left_idx"][i]) + k] = (q, k)
    return where


def test_cylinder_bounds_equal_the_checked_restatement(oracle, scene, many_frames):
    """The bounds the traversal kernel prunes with (build_cyl_kernel, GPU, double precision) are the ones tests/test_oracle_cyl_bound.py checks
    exhaustively on the CPU (oracle cyl_build): same fp16 values up to one ulp of the rounded-up slab / radius (fused vs unfused double arithmetic),
    after an ingest, after insta_replace_frames and after insta_refit_frames."""
    from insta_b200 import capi
    from test_oracle_seed_bound import surface_points

    def h2f(a):
        return a.view(np.float16).astype(np.float64)

    old, new = many_frames
    ctx = capi.Context(device=0)
    ctx.set_canonical(scene.verts, scene.faces, scene.masks)
    ctx.add_frames(np.stack(old[:6]))

    def check(frame, verts, rebuilt_tree=True):
        t = oracle.build_triangles(verts, scene.faces, scene.masks, scene.canon)
        nodes, tb = oracle.bvh_build(t)
        got = ctx.frame_cylinders(frame)
        assert got.shape[0] == int((nodes["left_idx"] >= 0).sum())
        if not rebuilt_tree:  # refit: the frame keeps the tree it was built with (not the one a rebuild would give)
            return got
        want = oracle.cyl_build(nodes, tb)
        where = _node_to_quad_child(nodes)
        mine = np.zeros_like(want)
        mine[0] = want[0]  # the root is never tested
        same_axis = 0
        for node, (q, k) in where.items():
            g, w = got[q, k], want[node]
            mine[node] = g
            if (g[:3] == w[:3]).all():  # same axis (the fp32 normal sum is order dependent: a component may differ by one fp16 ulp) -> same slab, radius
                same_axis += 1
                assert abs(int(g[3]) - int(w[3])) <= 1 and abs(int(g[4]) - int(w[4])) <= 1 and abs(int(g[5]) - int(w[5])) <= 1, (frame, node, g, w)
            else:
                assert np.abs(h2f(g[:3]) - h2f(w[:3])).max() < 2e-3, (frame, node, g, w)
        assert same_axis > 0.97 * len(where)
        # ... and the GPU's own values pass the exhaustive check: no culled subtree holds a triangle the reference would accept, winners unchanged
        rng = np.random.default_rng(frame)
        lo, hi = scene.aabb[:3], scene.aabb[3:]
        far = (rng.random((20000, 3), np.float32) * (hi - lo) + lo).astype(np.float32)
        pts = np.concatenate([far, surface_points(tb, rng, 20000)])
        ref_idx, _ = oracle.closest_triangle(pts, nodes, tb)
        idx, st = oracle.closest_triangle_cyl(pts, nodes, tb, mine, None, check=True)
        assert st[3] == 0 and st[2] > 0 and np.array_equal(idx, ref_idx)
        return got

    for f in (0, 5):
        check(f, old[f])
    ctx.replace_frames(np.array([5], np.int32), np.stack([new[5]]))
    check(5, new[5])
    before = ctx.frame_cylinders(2).copy()
    ctx.refit_frames(np.array([2], np.int32), np.stack([new[2]]))
    after = check(2, new[2], rebuilt_tree=False)
    assert (after != before).any()  # the bounds followed the new vertices
    assert np.isfinite(h2f(after[:, :, :4])).all()
    ctx.close()


def test_cylinder_cull_switch_gives_identical_records(scene, many_frames, monkeypatch):
    """INSTA_DEFORM_CYL=0 (traversal with the reference's box test only) and the default produce the same bits: ids and deformed records, on random
    samples and on points on / next to the surface where ties are the rule."""
    import torch

    from insta_b200 import capi, synth
    from test_oracle_seed_bound import surface_points

    old, _ = many_frames
    n = 400000
    fids = synth.train_regime_frame_ids(n, len(old))
    rec = synth.make_records(n, scene.aabb, fids, seed=21, pad_fraction=0.01)
    # a quarter of the batch sits on / next to the surface of its frame (vertices, edge midpoints, centroids: ties between neighbouring triangles)
    rng = np.random.default_rng(4)
    q = n // 4
    for f in np.unique(fids[:q]):
        sel = np.nonzero(fids[:q] == f)[0]
        t = _frame_tris(scene, old[int(f)])
        rec[sel, 0:3] = surface_points(t, rng, len(sel))
    out = []
    for env in ("1", "0"):
        monkeypatch.setenv("INSTA_DEFORM_CYL", env)
        ctx = capi.Context(device=0)
        ctx.set_canonical(scene.verts, scene.faces, scene.masks)
        ctx.add_frames(np.stack(old))
        assert (ctx.frame_cylinders(0).shape[0] > 0) == (env == "1")
        d = torch.from_numpy(rec.copy()).cuda()
        tri = torch.full((n,), -7, dtype=torch.int32, device="cuda")
        ctx.surface_closest_point(d, tri)
        torch.cuda.synchronize()
        out.append((d.cpu().numpy(), tri.cpu().numpy()))
        ctx.close()
    assert (out[0][1] == out[1][1]).all()
    assert (out[0][0].view(np.int32) == out[1][0].view(np.int32)).all()
