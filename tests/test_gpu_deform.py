"""Stage 1 on the GPU, through the C ABI, against the oracle: closest-triangle topology ids bit-exact; canonical positions /
directions within the stated fp32 tolerance.  Every parity test runs in both arithmetic flavours: 'host' (oracle pinned
bit-for-bit to the reference's host code in this repo's CPU tests) and 'device' (oracle pinned to the reference's CUDA
kernels in tests/test_gpu_vs_ref_cuda.py); the product's default is 'device'."""
import numpy as np
import pytest

from common import UNIT_AABB

pytestmark = pytest.mark.gpu

POS_TOL = 1e-5        # abs, unit-cube coordinates, non-blended (closed-form R^T inverse vs Eigen 4x4 inverse)
POS_TOL_BLEND = 1e-4  # abs, blended (inverse of a weighted matrix sum, __expf weights)
DIR_TOL = 2e-5


def _run(ctx, rec, frame_id=-1):
    import torch

    d = torch.from_numpy(rec.copy()).cuda()
    tri = torch.full((rec.shape[0],), -7, dtype=torch.int32, device="cuda")
    if frame_id < 0:
        ctx.surface_closest_point(d, tri)
    else:
        ctx.render_deformed(d, frame_id, tri)
    torch.cuda.synchronize()
    return d.cpu().numpy(), tri.cpu().numpy()


def _check(scene, oracle, rec, got, tri, want, want_tri, blend):
    assert (tri == want_tri).all(), "closest-triangle ids differ at %s" % np.nonzero(tri != want_tri)[0][:10]
    tol = POS_TOL_BLEND if blend else POS_TOL
    assert np.abs(got[:, 0:3] - want[:, 0:3]).max() < tol
    assert np.abs(got[:, 8:11] - want[:, 8:11]).max() < DIR_TOL
    # untouched fields are bit-identical; expression dims are exactly the oracle's (1.0 outside Lips)
    assert (got[:, 3:8] == rec[:, 3:8]).all()
    assert (got[:, 11:] == want[:, 11:]).all()
    sk = want_tri < 0
    assert (got[sk] == rec[sk]).all()


@pytest.mark.parametrize("n", [8192, 1000, 1, 129])
def test_warp_parity_holder_frames(arith_pair, scene, n):
    gpu_ctx, oracle = arith_pair
    rec = scene.golden_records(8192)[:n] if n <= 8192 else None
    got, tri = _run(gpu_ctx, rec)
    want, want_tri = scene.oracle_warp(oracle, rec)
    _check(scene, oracle, rec, got, tri, want, want_tri, blend=True)


def test_warp_matches_reference_golden(gpu_ctx_host, scene, golden):
    gpu_ctx = gpu_ctx_host
    """Straight against the vectors the reference's own code produced (tests/golden/stage1_ref.npz)."""
    rec = scene.golden_records()
    got, tri = _run(gpu_ctx, rec)
    assert (tri == golden["warp_interp_tri"]).all()
    g = golden["warp_interp_rec"]
    assert np.abs(got[:, 0:3] - g[:, 0:3]).max() < POS_TOL_BLEND
    assert np.abs(got[:, 8:11] - g[:, 3:6]).max() < DIR_TOL
    got, tri = _run(gpu_ctx, rec, frame_id=1)
    assert (tri == golden["render_tri"]).all()
    assert np.abs(got[:, [0, 1, 2, 8, 9, 10]] - golden["render_rec"]).max() < POS_TOL_BLEND


def test_tie_heavy_points(arith_pair, scene):
    gpu_ctx, oracle = arith_pair
    """Samples placed exactly on vertices, edge midpoints and centroids of the posed mesh: equal distances everywhere,
    the winner is decided by visit order alone."""
    from insta_b200 import synth

    v = scene.frame_verts[0]
    f = scene.faces
    pts = np.concatenate([v, (v[f[:, 0]] + v[f[:, 1]]) * np.float32(0.5), (v[f[:, 0]] + v[f[:, 1]] + v[f[:, 2]]) / np.float32(3)]).astype(np.float32)
    rec = synth.make_records(len(pts), scene.aabb, np.zeros(len(pts)), seed=3)
    rec[:, 0:3] = pts
    got, tri = _run(gpu_ctx, rec)
    want, want_tri = scene.oracle_warp(oracle, rec)
    assert (tri == want_tri).all()
    # on-surface points are ill-conditioned for nothing: positions still agree
    assert np.abs(got[:, 0:3] - want[:, 0:3]).max() < POS_TOL_BLEND


def test_skip_rules(arith_pair, scene):
    gpu_ctx, oracle = arith_pair
    from insta_b200 import synth

    n = 2048
    rec = synth.make_records(n, scene.aabb, np.zeros(n), seed=5)
    rec[0::4, 4] = -1.0                      # padding
    rec[1::4, 4] = 2.0                       # frame >= n_frames
    rec[2::8, 0] = 1.5                       # outside m_aabb
    rec[6::8, 1] = -0.25
    rec[3, 0:3] = [1.0, 1.0, 1.0]            # on the inclusive boundary: processed
    got, tri = _run(gpu_ctx, rec)
    want, want_tri = scene.oracle_warp(oracle, rec)
    assert (tri == want_tri).all()
    assert (want_tri[0::4] == -1).all() and (want_tri[1::4] == -1).all() and (want_tri[2::8] == -1).all() and want_tri[3] >= 0
    sk = want_tri < 0
    assert (got[sk] == rec[sk]).all()
    assert np.abs(got[~sk][:, 0:3] - want[~sk][:, 0:3]).max() < 2e-3  # far-away corner point: looser, ids are what matter


def test_render_deformed_fixed_frame(arith_pair, scene):
    gpu_ctx, oracle = arith_pair
    rec = scene.golden_records(4096)
    for f in (0, 1):
        got, tri = _run(gpu_ctx, rec, frame_id=f)
        want, want_tri = scene.oracle_warp(oracle, rec, frame_id=f)
        _check(scene, oracle, rec, got, tri, want, want_tri, blend=True)


def test_points_path_explicit_and_rng_frames(arith_pair, scene):
    gpu_ctx, oracle = arith_pair
    import torch

    n = 5000
    rec = scene.golden_records(8192)[:n]
    xyz = rec[:, 0:3].copy()
    fids = (np.arange(n) % 2).astype(np.int32)
    d = torch.from_numpy(xyz.copy()).cuda()
    tri = torch.empty(n, dtype=torch.int32, device="cuda")
    gpu_ctx.deform_points(d, torch.from_numpy(fids).cuda(), tri)
    want, want_tri = scene.oracle_warp(oracle, xyz, stride=3, frame_ids=fids)
    assert (tri.cpu().numpy() == want_tri).all()
    assert np.abs(d.cpu().numpy() - want.reshape(-1, 3)).max() < POS_TOL_BLEND
    # reference frame draw: pcg32(42).advance(i * (n_frames - 1)).next_float() * (n_frames - 1)
    d = torch.from_numpy(xyz.copy()).cuda()
    gpu_ctx.surface_closest_point(d, tri)
    rfids = oracle.random_frame_ids(n, 2)
    assert (rfids == 0).all()  # with two frames the draw is always frame 0 (r * 1 < 1) -- reference behaviour
    want, want_tri = scene.oracle_warp(oracle, xyz, stride=3, frame_ids=rfids)
    assert (tri.cpu().numpy() == want_tri).all()


def test_disable_interp_and_geo_and_strides(scene, oracle):
    import torch
    from insta_b200 import capi, synth

    rec = scene.golden_records(4096)
    oracle.set_arith("device")
    ctx = capi.Context(device=0, disable_interp=1, geo_params=1)
    try:
        ctx.set_canonical(scene.verts, scene.faces, scene.masks)
        for v in scene.frame_verts:
            ctx.add_frame(v)
        r28 = np.concatenate([rec, np.zeros((len(rec), 1), np.float32)], axis=1)  # 16 exp + 1 geo
        d = torch.from_numpy(r28.copy()).cuda()
        tri = torch.empty(len(rec), dtype=torch.int32, device="cuda")
        ctx.surface_closest_point(d, tri)
        want, want_tri = scene.oracle_warp(oracle, r28, stride=28, disable_interp=True, extra_dims=17, use_geo=True)
        got = d.cpu().numpy()
        assert (tri.cpu().numpy() == want_tri).all()
        assert np.abs(got[:, 0:3] - want[:, 0:3]).max() < POS_TOL
        assert np.abs(got[:, 27] - want[:, 27]).max() < 1e-4      # cosine feature
        assert (got[:, 11:27] == want[:, 11:27]).all()
    finally:
        ctx.close()
    ctx = capi.Context(device=0, exp_params=0)
    try:
        ctx.set_canonical(scene.verts, scene.faces, scene.masks)
        ctx.add_frame(scene.frame_verts[0])
        r11 = np.ascontiguousarray(rec[:1001, :11])
        r11[:, 4] = 0
        d = torch.from_numpy(r11.copy()).cuda()
        tri = torch.empty(len(r11), dtype=torch.int32, device="cuda")
        ctx.surface_closest_point(d, tri)
        want, want_tri = scene.oracle_warp(oracle, r11, stride=11, extra_dims=0)
        assert (tri.cpu().numpy() == want_tri).all()
        assert np.abs(d.cpu().numpy()[:, 0:3] - want[:, 0:3]).max() < POS_TOL_BLEND
    finally:
        ctx.close()
        oracle.set_arith("host")


def test_errors_are_status_codes(gpu_ctx):
    import torch
    from insta_b200 import capi

    bad = torch.zeros((4, 5), device="cuda")
    with pytest.raises(capi.InstaError) as e:
        gpu_ctx.surface_closest_point(bad)
    assert e.value.code == -1
    empty = torch.zeros((0, 27), device="cuda")
    gpu_ctx.surface_closest_point(empty)  # n == 0 is a no-op


def test_full_size_batch_properties(arith_pair, scene):
    gpu_ctx, oracle = arith_pair
    """BASELINE config 2 size (4,194,304 records): checked through size-independent properties and an oracle subsample."""
    import torch
    from insta_b200 import synth

    n = 1 << 22
    fids = synth.train_regime_frame_ids(n, 2)
    rec = synth.make_records(n, scene.aabb, fids, seed=99, pad_fraction=0.01)
    d = torch.from_numpy(rec).cuda()
    tri = torch.empty(n, dtype=torch.int32, device="cuda")
    before = d.clone()
    gpu_ctx.surface_closest_point(d, tri)
    torch.cuda.synchronize()
    t = tri.cpu().numpy()
    assert ((t >= 0) == (rec[:, 4] >= 0)).all()                    # every non-padding sample found a triangle
    assert t.max() < len(scene.faces)
    assert torch.equal(d[:, 3:8], before[:, 3:8])                   # dt / holder / deformed_pos untouched
    sk = torch.from_numpy(t < 0).cuda()
    assert torch.equal(d[sk], before[sk])
    lips = torch.from_numpy(scene.masks[np.maximum(t, 0)] == 7).cuda() | sk
    assert bool((d[~lips][:, 11:] == 1.0).all()) and torch.equal(d[lips][:, 11:], before[lips][:, 11:])
    dirs = d[~sk][:, 8:11] * 2 - 1
    assert float((dirs.norm(dim=1) - 1).abs().max()) < 1e-5         # canonical directions are unit vectors
    # deterministic: a second run gives the same bits
    d2 = before.clone()
    tri2 = torch.empty_like(tri)
    gpu_ctx.surface_closest_point(d2, tri2)
    assert torch.equal(tri, tri2) and torch.equal(d, d2)
    # oracle on a strided subsample
    sub = np.arange(0, n, 173)
    want, want_tri = scene.oracle_warp(oracle, rec[sub])
    assert (t[sub] == want_tri).all()
    got = d.cpu().numpy()[sub]
    assert np.abs(got[:, 0:3] - want[:, 0:3]).max() < POS_TOL_BLEND


def test_device_side_count(gpu_ctx, scene):
    """insta_deform_samples_n / insta_network_inference_n: only min(n_max, *d_n_valid) records are touched"""
    import torch

    from insta_b200 import synth

    n, k = 6000, 2500
    rec = scene.golden_records(n, seed=21)
    a, b = torch.from_numpy(rec.copy()).cuda(), torch.from_numpy(rec.copy()).cuda()
    ta, tb = torch.full((n,), -7, dtype=torch.int32, device="cuda"), torch.full((n,), -7, dtype=torch.int32, device="cuda")
    gpu_ctx.surface_closest_point(a[:k], ta[:k])
    gpu_ctx.surface_closest_point_n(b, torch.tensor([k], dtype=torch.int32, device="cuda"), tb)
    torch.cuda.synchronize()
    assert torch.equal(a.view(torch.int32), b.view(torch.int32)) and torch.equal(ta, tb)
    assert bool((tb[k:] == -7).all()) and torch.equal(b[k:].view(torch.int32), torch.from_numpy(rec[k:].copy()).cuda().view(torch.int32))
    c, d = torch.from_numpy(rec.copy()).cuda(), torch.from_numpy(rec.copy()).cuda()
    tc, td = torch.empty(n, dtype=torch.int32, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda")
    gpu_ctx.surface_closest_point_n(c, torch.tensor([10 * n], dtype=torch.int32, device="cuda"), tc)  # count above n_max: clamped
    gpu_ctx.surface_closest_point(d, td)
    torch.cuda.synchronize()
    assert torch.equal(tc, td) and torch.equal(c.view(torch.int32), d.view(torch.int32))
    params = torch.from_numpy(synth.make_network_params(gpu_ctx.network_n_params, seed=2)).cuda()
    o1 = torch.zeros((n, 4), dtype=torch.float16, device="cuda")
    o2 = torch.zeros((n, 4), dtype=torch.float16, device="cuda")
    gpu_ctx.network_inference(a[:k], params, o1[:k])
    gpu_ctx.network_inference_n(a, torch.tensor([k], dtype=torch.int32, device="cuda"), params, o2)
    torch.cuda.synchronize()
    assert torch.equal(o1, o2) and float(o2[k:].abs().sum()) == 0.0 and float(o2[:k].abs().sum()) > 0
