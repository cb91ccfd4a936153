"""Our CUDA path against the REFERENCE's own CUDA kernels (warp, kernel_grid, kernel_grid_backward compiled for sm_100
into oracle/_ref/libref_cuda.so) on identical device inputs.  Hash-slot-dependent outputs and encodings must be
bit-exact; triangle ids are compared and the residual mismatch rate (the reference's default -fmad=true build contracts
the distance arithmetic differently from its own host build, SURVEY.md 7 hard part 1) is bounded and printed."""
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


def test_warp_vs_reference_kernel(gpu_ctx, ref, scene):
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
    same = d < 1e-4
    rate = 1.0 - float(same.float().mean())
    print("\n[warp vs reference CUDA (default -fmad)] samples whose canonical position differs by > 1e-4: %.5f%%" % (100 * rate))
    # a different winner among (near-)tied triangles moves the point by millimetres; the reference's own two builds
    # (host vs device) disagree at this rate, ours follows the host build bit-for-bit
    assert rate < 0.02
    assert float((a[same][:, 8:11] - b[same][:, 8:11]).abs().max()) < 1e-3
    assert torch.equal(a[:, 11:][same], b[:, 11:][same])
