"""Our CUDA path against the REFERENCE's own CUDA kernels (warp, kernel_grid, kernel_grid_backward compiled for sm_100
into oracle/_ref/libref_cuda.so) on identical device inputs.  Hash-slot-dependent outputs and encodings must be
bit-exact; triangle ids are compared and triangle ids and
distances are bit-exact too (the product's default arithmetic reproduces nvcc's FMA contraction of the reference's
distance functions; the 'host' flavour, which follows the reference's host build instead, differs for a few percent)."""
import numpy as np
import pytest

from insta_b200 import synth
from oracle import pyrefcuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not pyrefcuda.available(), reason="oracle/_ref/libref_cuda.so not built")]


@pytest.fixture(scope="module")
def ref(scene):
    r = pyrefcuda.RefCuda()
    r.upload_scene(scene.nodes, scene.tbvh, scene.ttopo, scene.canon, scene.adj)
    return r


def test_grid_forward_bit_exact_vs_reference_kernel(gpu_ctx, ref):
    import torch

    offs, scales, res = gpu_ctx.grid_layout()
    pls = float(gpu_ctx.params.per_level_scale) or float(np.exp(np.log(np.float32(2048.0) / np.float32(16)) / np.float32(15)))
    n = 1 << 20
    rng = np.random.default_rng(31)
    pos = torch.from_numpy(rng.random((n, 3), dtype=np.float32)).cuda()
    table = torch.from_numpy(synth.make_grid_table(int(offs[-1]))).cuda()
    a = torch.empty((128, n), dtype=torch.float16, device="cuda")
    b = torch.empty((128, n), dtype=torch.float16, device="cuda")
    gpu_ctx.hash_encode_fwd(pos, table, a)
    ref.grid_forward(pos, 3, table, b, per_level_scale=pls)
    torch.cuda.synchronize()
    assert torch.equal(a.view(torch.int16), b.view(torch.int16))
    # strided records as the reference feeds them (27-float stride)
    rec = torch.from_numpy(rng.random((100000, 27), dtype=np.float32)).cuda()
    a = torch.empty((128, 100000), dtype=torch.float16, device="cuda")
    b = torch.empty_like(a)
    gpu_ctx.hash_encode_fwd(rec, table, a)
    ref.grid_forward(rec, 27, table, b, per_level_scale=pls)
    assert torch.equal(a.view(torch.int16), b.view(torch.int16))


def test_grid_backward_vs_reference_kernel(gpu_ctx, ref):
    import torch

    offs, _, _ = gpu_ctx.grid_layout()
    pls = float(np.exp(np.log(np.float32(2048.0) / np.float32(16)) / np.float32(15)))
    n = 1 << 18
    rng = np.random.default_rng(32)
    pos = torch.from_numpy(rng.random((n, 3), dtype=np.float32)).cuda()
    dy = torch.from_numpy((rng.standard_normal((128, n)) * 0.01).astype(np.float16)).cuda()
    a = torch.empty((int(offs[-1]), 8), dtype=torch.float16, device="cuda")
    b = torch.empty_like(a)
    gpu_ctx.hash_encode_bwd(pos, dy, a, zero_first=True)
    ref.grid_backward(pos, 3, dy, b, per_level_scale=pls)
    torch.cuda.synchronize()
    a, b = a.float(), b.float()
    # both are fp16 atomic accumulations in nondeterministic order: same support, same values to fp16 accumulation noise
    assert torch.equal(a != 0, b != 0) or float(((a != 0) ^ (b != 0)).float().mean()) < 1e-4
    err = (a - b).abs()
    scale = b.abs().clamp_min(1e-3)
    assert float((err / scale).mean()) < 5e-3
    assert float(err.max()) < 0.05 * float(b.abs().max())


def _tie_heavy(scene, k):
    v = scene.frame_verts[k]
    f = scene.faces
    return np.concatenate([v, (v[f[:, 0]] + v[f[:, 1]]) * np.float32(0.5), (v[f[:, 0]] + v[f[:, 1]] + v[f[:, 2]]) / np.float32(3)]).astype(np.float32)


def test_distance_primitives_bit_exact_vs_reference_device_code(oracle, ref, scene):
    """Triangle::distance_sq and BoundingBox::distance_sq as the reference's CUDA build evaluates them (FMA contraction),
    against the oracle's 'device' flavour, bit for bit."""
    import torch

    rng = np.random.default_rng(5)
    pts = (scene.aabb[:3] + (scene.aabb[3:] - scene.aabb[:3]) * rng.random((4096, 3))).astype(np.float32)
    pts = np.concatenate([pts, _tie_heavy(scene, 0)[:4096]])
    d_pts = torch.from_numpy(pts).cuda()
    oracle.set_arith("device")
    try:
        for ti, ni in ((0, 1), (17, 5), (4000, 300), (8089, 5460), (1234, 77)):
            a, b = ref.distances(d_pts, 0, ti, ni)
            a, b = a.cpu().numpy(), b.cpu().numpy()
            t = scene.tbvh[0][ti]
            nd = scene.nodes[0][ni]
            mine_t = np.array([oracle.triangle_distance_sq(t, p) for p in pts], np.float32)
            mine_b = np.array([oracle.bbox_distance_sq(nd["mn"], nd["mx"], p) for p in pts], np.float32)
            assert (mine_t.view(np.uint32) == a.view(np.uint32)).all()
            assert (mine_b.view(np.uint32) == b.view(np.uint32)).all()
    finally:
        oracle.set_arith("host")


def test_closest_triangle_bit_exact_vs_reference_device_code(gpu_ctx, oracle, ref, scene):
    """2M random + tie-heavy points: the reference's closest_triangle run on the GPU, our kernel (default arithmetic) and
    the oracle's 'device' flavour must name the same triangle for every point, with bit-identical distances."""
    import torch

    rng = np.random.default_rng(6)
    n = 1 << 21
    pts = (scene.aabb[:3] + (scene.aabb[3:] - scene.aabb[:3]) * rng.random((n, 3))).astype(np.float32)
    pts = np.concatenate([pts, _tie_heavy(scene, 1)])
    d_pts = torch.from_numpy(pts).cuda()
    idx, dist = ref.closest_triangle(d_pts, 1)
    ref_ids = torch.from_numpy(scene.tbvh[1]["id"].astype(np.int64)).cuda()[idx.long()].int()
    xyz = d_pts.clone()
    tri = torch.empty(len(pts), dtype=torch.int32, device="cuda")
    gpu_ctx.deform_points(xyz, torch.ones(len(pts), dtype=torch.int32, device="cuda"), tri)
    torch.cuda.synchronize()
    assert torch.equal(tri, ref_ids), "mismatches: %d" % int((tri != ref_ids).sum())
    oracle.set_arith("device")
    try:
        sub = np.concatenate([np.arange(0, n, 37), np.arange(n, len(pts))])
        oi, od = oracle.closest_triangle(pts[sub], scene.nodes[1], scene.tbvh[1])
    finally:
        oracle.set_arith("host")
    assert (oi == idx.cpu().numpy()[sub]).all()
    assert (od.view(np.uint32) == dist.cpu().numpy()[sub].view(np.uint32)).all()


def test_warp_vs_reference_kernel(gpu_ctx, ref, scene):
    """The whole `warp` kernel of the reference against ours on 1M records: same triangle everywhere, so positions and
    directions agree to rounding and the expression dims are identical."""
    import torch

    n = 1 << 20
    fids = synth.train_regime_frame_ids(n, 2)
    rec = synth.make_records(n, scene.aabb, fids, seed=41, pad_fraction=0.01)
    a = torch.from_numpy(rec.copy()).cuda()
    b = torch.from_numpy(rec.copy()).cuda()
    tri = torch.empty(n, dtype=torch.int32, device="cuda")
    gpu_ctx.surface_closest_point(a, tri)
    ref.warp(b, 27)
    torch.cuda.synchronize()
    d = (a[:, 0:3] - b[:, 0:3]).abs().max(dim=1).values
    print("\n[warp vs reference CUDA kernel] max |dpos| = %.3g, samples differing by > 1e-4: %d of %d" % (float(d.max()), int((d > 1e-4).sum()), n))
    assert float(d.max()) < 1e-4
    assert float((a[:, 8:11] - b[:, 8:11]).abs().max()) < 2e-5
    assert torch.equal(a[:, 11:], b[:, 11:])
    assert torch.equal(a[:, 3:8], b[:, 3:8])


def test_host_arithmetic_differs_from_device_build(gpu_ctx_host, ref, scene):
    """Documents why there are two flavours: the reference's host and CUDA builds of the same source disagree on the
    winner for a few percent of far-away samples (near-ties decided by the last bit)."""
    import torch

    n = 1 << 19
    rec = synth.make_records(n, scene.aabb, np.zeros(n), seed=43)
    a = torch.from_numpy(rec.copy()).cuda()
    b = torch.from_numpy(rec.copy()).cuda()
    gpu_ctx_host.surface_closest_point(a)
    ref.warp(b, 27)
    d = (a[:, 0:3] - b[:, 0:3]).abs().max(dim=1).values
    rate = float((d > 1e-4).float().mean())
    print("\n[host-flavour vs reference CUDA kernel] samples with a different winner: %.3f%%" % (100 * rate))
    assert 0.0 < rate < 0.08
