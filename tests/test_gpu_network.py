"""GPU parity of the fused network kernel (hash grid + density MLP + SH + colour MLP, tcgen05) against an fp32 PyTorch
evaluation of the same fp16 weights with fp16-rounded activations between layers (the points where the reference stores
fp16, tcnn fully_fused_mlp.cu:500-634), on top of the grid features of the stand-alone encoder, which is pinned bit-for-bit
to the reference's kernel_grid elsewhere.  Tolerance: |d| <= 2e-3 + 1% (fp32 vs fp16-accumulating reference arithmetic)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_params(n_params, seed=5, grid_scale=1.0):
    from insta_b200 import synth

    return synth.make_network_params(n_params, seed, grid_scale), synth.MLP_SHAPES


def torch_reference(ctx, rec, params, shapes):
    import torch

    n = rec.shape[0]
    P = params
    feats = torch.empty((128, n), dtype=torch.float16, device="cuda")
    ctx.hash_encode_fwd(rec, P[25600:].view(-1, 8), feats)
    x = torch.cat([feats.t().float(), rec[:, -16:].half().float()], dim=1)  # [n, 144]
    off = 0
    W = []
    for o, i in shapes:
        W.append(P[off:off + o * i].view(o, i).float())
        off += o * i

    def h(t):
        return t.half().float()

    a = h(torch.relu(x @ W[0].t()))
    a = h(torch.relu(a @ W[1].t()))
    dens = h(a @ W[2].t())  # [n, 16]
    d = rec[:, 8:11] * 2 - 1
    X, Y, Z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz, x2, y2, z2 = X * Y, X * Z, Y * Z, X * X, Y * Y, Z * Z
    sh = torch.stack([
        torch.full_like(X, 0.28209479177387814), -0.48860251190291987 * Y, 0.48860251190291987 * Z, -0.48860251190291987 * X,
        1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.94617469575755997 * z2 - 0.31539156525251999, -1.0925484305920792 * xz,
        0.54627421529603959 * x2 - 0.54627421529603959 * y2, 0.59004358992664352 * Y * (-3.0 * x2 + y2), 2.8906114426405538 * xy * Z,
        0.45704579946446572 * Y * (1.0 - 5.0 * z2), 0.3731763325901154 * Z * (5.0 * z2 - 3.0), 0.45704579946446572 * X * (1.0 - 5.0 * z2),
        1.4453057213202769 * Z * (x2 - y2), 0.59004358992664352 * X * (-x2 + 3.0 * y2)], dim=1)
    r = torch.cat([dens, h(sh)], dim=1)
    r = h(torch.relu(r @ W[3].t()))
    r = h(torch.relu(r @ W[4].t()))
    r = h(torch.relu(r @ W[5].t()))
    rgb = h(r @ W[6].t())
    return torch.cat([rgb[:, :3], dens[:, :1]], dim=1)


@pytest.mark.parametrize("n", [1000, 128 * 300 + 17])
def test_network_inference_vs_torch(gpu_ctx, scene, n):
    import torch

    from insta_b200 import synth

    rec_np = synth.make_records(n, scene.aabb, np.zeros(n, np.int32), seed=21)
    rec = torch.from_numpy(rec_np).cuda()
    params_np, shapes = make_params(gpu_ctx.network_n_params)
    params = torch.from_numpy(params_np).cuda()
    want = torch_reference(gpu_ctx, rec, params, shapes)
    out = torch.zeros((n, 4), dtype=torch.float16, device="cuda")
    gpu_ctx.network_inference(rec, params, out)
    torch.cuda.synchronize()
    got = out.float()
    err = (got - want).abs()
    tol = 2e-3 + 1e-2 * want.abs()
    assert want.abs().max() > 0.05  # the test is not vacuous
    assert bool((err <= tol).all()), "max err %.4g at %s (want %.4g)" % (float(err.max()), int(err.argmax()), float(want.flatten()[err.argmax()]))
    # the reference's row-major output layout
    out2 = torch.zeros((16, n + 5), dtype=torch.float16, device="cuda")
    gpu_ctx.network_inference(rec, params, out2, out_stride=n + 5)
    torch.cuda.synchronize()
    assert torch.equal(out2[:4, :n].t().contiguous(), out)


def _sh16(d):
    import torch

    X, Y, Z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz, x2, y2, z2 = X * Y, X * Z, Y * Z, X * X, Y * Y, Z * Z
    return torch.stack([
        torch.full_like(X, 0.28209479177387814), -0.48860251190291987 * Y, 0.48860251190291987 * Z, -0.48860251190291987 * X,
        1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.94617469575755997 * z2 - 0.31539156525251999, -1.0925484305920792 * xz,
        0.54627421529603959 * x2 - 0.54627421529603959 * y2, 0.59004358992664352 * Y * (-3.0 * x2 + y2), 2.8906114426405538 * xy * Z,
        0.45704579946446572 * Y * (1.0 - 5.0 * z2), 0.3731763325901154 * Z * (5.0 * z2 - 3.0), 0.45704579946446572 * X * (1.0 - 5.0 * z2),
        1.4453057213202769 * Z * (x2 - y2), 0.59004358992664352 * X * (-x2 + 3.0 * y2)], dim=1)


@pytest.mark.parametrize("n", [128 * 37 + 9, 4096])
def test_network_forward_backward_vs_torch_autograd(gpu_ctx, scene, n):
    """dL/dW of the seven layers and dL/d(grid table) from the tcgen05 forward+backward kernels against PyTorch autograd (fp32
    math on the same fp16 weights, fp16 rounding of the activations as a straight-through op).  The kernels round every layer's
    gradient to fp16 as the reference does (tcnn fully_fused_mlp.cu:151-313), autograd does not: tolerances are relative to each
    matrix's scale (2 %), the grid gradient additionally carries the order-dependent fp16 atomics (3 % of its max)."""
    import torch

    from insta_b200 import synth

    rec_np = synth.make_records(n, scene.aabb, np.zeros(n, np.int32), seed=33)
    rec = torch.from_numpy(rec_np).cuda()
    params_np, shapes = make_params(gpu_ctx.network_n_params, seed=8)
    params = torch.from_numpy(params_np).cuda()
    g = torch.Generator(device="cuda").manual_seed(2)
    dloss = torch.zeros((n, 16), dtype=torch.float16, device="cuda")
    dloss[:, :4] = (torch.randn((n, 4), device="cuda", generator=g) * 0.05).half()
    # ---- ours ----
    grads = torch.full((gpu_ctx.network_n_params,), 7.0, dtype=torch.float16, device="cuda")  # must be overwritten
    out = torch.zeros((n, 4), dtype=torch.float16, device="cuda")
    gpu_ctx.network_forward_backward(rec, params, dloss, grads, out)
    torch.cuda.synchronize()
    # the forward output equals the inference kernel's
    out_inf = torch.zeros((n, 4), dtype=torch.float16, device="cuda")
    gpu_ctx.network_inference(rec, params, out_inf)
    assert torch.equal(out, out_inf)
    # ---- autograd reference ----
    feats16 = torch.empty((128, n), dtype=torch.float16, device="cuda")
    ctx_table = params[25600:].view(-1, 8)
    gpu_ctx.hash_encode_fwd(rec, ctx_table, feats16)
    feats = feats16.t().float().clone().requires_grad_(True)
    W, off = [], 0
    for o, i in shapes:
        W.append(params[off:off + o * i].view(o, i).float().clone().requires_grad_(True))
        off += o * i

    def h(t):
        return t + (t.half().float() - t).detach()

    x = torch.cat([feats, rec[:, -16:].half().float()], dim=1)
    a = h(torch.relu(x @ W[0].t()))
    a = h(torch.relu(a @ W[1].t()))
    dens = h(a @ W[2].t())
    r = torch.cat([dens, h(_sh16(rec[:, 8:11] * 2 - 1))], dim=1)
    r = h(torch.relu(r @ W[3].t()))
    r = h(torch.relu(r @ W[4].t()))
    r = h(torch.relu(r @ W[5].t()))
    rgb = r @ W[6].t()
    dl = dloss.float()
    ((rgb[:, :3] * dl[:, :3]).sum() + (dens[:, 0] * dl[:, 3]).sum()).backward()
    off = 0
    for k, (o, i) in enumerate(shapes):
        want = W[k].grad
        got = grads[off:off + o * i].view(o, i).float()
        off += o * i
        if k == 6:  # only the three colour rows of the padded output layer receive gradient
            assert float(got[3:].abs().max()) == 0.0
        scale = float(want.abs().max())
        assert scale > 0
        err = float((got - want).abs().max())
        assert err <= 2e-2 * scale, "layer %d: max err %.4g vs scale %.4g" % (k, err, scale)
    # grid: scatter autograd's dL/dfeatures with the stand-alone backward kernel (pinned to kernel_grid_backward elsewhere)
    dfeat = feats.grad.t().contiguous().half()
    want_grid = torch.zeros_like(ctx_table)
    gpu_ctx.hash_encode_bwd(rec, dfeat, want_grid)
    torch.cuda.synchronize()
    got_grid = grads[25600:].view(-1, 8).float()
    scale = float(want_grid.float().abs().max())
    assert scale > 0
    assert float((got_grid - want_grid.float()).abs().max()) <= 3e-2 * scale
    assert int((got_grid != 0).sum()) > 1000


def test_network_density_mode(gpu_ctx, scene):
    """NerfNetwork::density (nerf_network.h:385-414): bare positions, expression inputs all 1.  Must equal channel 3 of the full inference on
    records that carry the same positions and all-one expression dims (bit-for-bit: same encoder and density MLP), and the fp32 PyTorch
    density within the network tolerance."""
    import torch

    from insta_b200 import synth

    n = 128 * 21 + 5
    rec_np = synth.make_records(n, scene.aabb, np.zeros(n, np.int32), seed=31)
    rec_np[:, 11:27] = 1.0
    rec = torch.from_numpy(rec_np).cuda()
    p_np, shapes = make_params(gpu_ctx.network_n_params, seed=9)
    params = torch.from_numpy(p_np).cuda()
    full = torch.empty((n, 4), dtype=torch.float16, device="cuda")
    gpu_ctx.network_inference(rec, params, full)
    xyz = rec[:, :3].contiguous()
    dens = torch.full((n,), float("nan"), dtype=torch.float16, device="cuda")
    gpu_ctx.network_density(xyz, params, dens)
    wide = torch.zeros((n, 7), device="cuda")  # a wider stride than 3 floats
    wide[:, :3] = xyz
    dens2 = torch.empty(n, dtype=torch.float16, device="cuda")
    gpu_ctx.network_density(wide, params, dens2)
    torch.cuda.synchronize()
    assert torch.equal(dens.view(torch.int16), full[:, 3].contiguous().view(torch.int16))
    assert torch.equal(dens, dens2)
    want = torch_reference(gpu_ctx, rec, params, shapes)[:, 3]
    d = (dens.float() - want).abs()
    assert bool((d <= 2e-3 + 0.01 * want.abs()).all()), float(d.max())
