"""GPU parity of the fused tcgen05 network kernels against the REFERENCE's own network -- ngp::NerfNetwork<__half> over tiny-cuda-nn's GridEncoding,
SphericalHarmonics and FullyFusedMLP, compiled for sm_100 (oracle/_ref/libref_net.so, `make -C oracle refnet`) -- on identical parameters and records:
inference (NerfNetwork::inference_mixed_precision_impl), density (NerfNetwork::density) and forward + backward gradients (train_nerf_step).
Tolerances: the reference accumulates its MLP products in fp16 (wmma, fully_fused_mlp.cu:500-634), this library in fp32 tensor memory."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def refnet(gpu_ctx):
    from oracle import pyrefnet

    if not pyrefnet.available():
        pytest.skip("oracle/_ref/libref_net.so not built (`make -C oracle refnet`, ~12 minutes, needs /root/reference)")
    R = pyrefnet.RefNet()
    assert R.n_params == gpu_ctx.network_n_params, "parameter count differs from the reference network"
    assert R.out_width == 16
    return R


def _setup(gpu_ctx, scene, refnet, n, seed):
    import torch

    from insta_b200 import synth

    rec_np = synth.make_records(n, scene.aabb, np.zeros(n, np.int32), seed=seed)
    rec = torch.from_numpy(rec_np).cuda()
    params = torch.from_numpy(synth.make_network_params(gpu_ctx.network_n_params, seed=seed + 1)).cuda()
    refnet.set_params(params)
    return rec, params


def test_inference_vs_reference_network(gpu_ctx, scene, refnet):
    import torch

    n = 128 * 200
    rec, params = _setup(gpu_ctx, scene, refnet, n, 41)
    want = torch.zeros((16, n), dtype=torch.float16, device="cuda")
    refnet.inference(rec, want)
    got = torch.zeros((n, 4), dtype=torch.float16, device="cuda")
    gpu_ctx.network_inference(rec, params, got)
    torch.cuda.synchronize()
    w, g = want[:4].t().float(), got.float()
    assert float(w.abs().max()) > 0.05
    err = (g - w).abs()
    assert bool((err <= 4e-3 + 2e-2 * w.abs()).all()), "max err %.4g (want %.4g)" % (float(err.max()), float(w.flatten()[err.argmax()]))
    assert float(err.mean()) < 1e-3


def test_density_vs_reference_network(gpu_ctx, scene, refnet):
    import torch

    n = 128 * 100
    rec, params = _setup(gpu_ctx, scene, refnet, n, 43)
    xyz = rec[:, :3].contiguous()
    want = torch.zeros((16, n), dtype=torch.float16, device="cuda")
    refnet.density(xyz, want)
    got = torch.zeros(n, dtype=torch.float16, device="cuda")
    gpu_ctx.network_density(xyz, params, got)
    torch.cuda.synchronize()
    w, g = want[0].float(), got.float()
    err = (g - w).abs()
    assert float(w.abs().max()) > 0.05
    assert bool((err <= 4e-3 + 2e-2 * w.abs()).all()), float(err.max())


def test_forward_backward_vs_reference_network(gpu_ctx, scene, refnet):
    import torch

    n = 128 * 128
    rec, params = _setup(gpu_ctx, scene, refnet, n, 45)
    g = torch.Generator(device="cuda").manual_seed(5)
    dl = torch.zeros((n, 16), dtype=torch.float16, device="cuda")
    dl[:, :4] = (torch.randn((n, 4), device="cuda", generator=g) * 0.05).half()
    want_out = torch.zeros((n, 16), dtype=torch.float16, device="cuda")
    want_g = torch.zeros(refnet.n_params, dtype=torch.float16, device="cuda")
    refnet.forward_backward(rec, dl, want_out, want_g)
    got_g = torch.zeros(gpu_ctx.network_n_params, dtype=torch.float16, device="cuda")
    got_out = torch.zeros((n, 4), dtype=torch.float16, device="cuda")
    gpu_ctx.network_forward_backward(rec, params, dl, got_g, out=got_out)
    torch.cuda.synchronize()
    err = (got_out.float() - want_out[:, :4].float()).abs()
    assert bool((err <= 4e-3 + 2e-2 * want_out[:, :4].float().abs()).all())
    cos = torch.nn.functional.cosine_similarity
    wm, gm = want_g[:25600].float(), got_g[:25600].float()
    assert float(wm.abs().max()) > 0 and float(cos(wm, gm, dim=0)) > 0.999, float(cos(wm, gm, dim=0))
    assert float((wm - gm).norm()) < 0.03 * float(wm.norm())
    wg, gg = want_g[25600:].float(), got_g[25600:].float()
    assert float(wg.abs().max()) > 0 and float(cos(wg, gg, dim=0)) > 0.995, float(cos(wg, gg, dim=0))
    # the same entries are touched
    assert float(((wg != 0) != (gg != 0)).float().mean()) < 2e-3


def test_adam_ema_steps_vs_tcnn_optimizer(gpu_ctx, refnet):
    """insta_optimizer_step against tiny-cuda-nn's own EmaOptimizer{AdamOptimizer} (optimizers/ema.h:64-130, adam.h:48-114; the optimizer of
    configs/nerf/main.json:21-32, compiled into libref_net.so): four steps on identical parameters and gradient vectors -- fp32 master weights, fp16 weights
    and EMA weights after every step.  Gradients as in training: dense for the 25,600 MLP weights, ~10 % of the hash-grid entries touched per step, so the
    per-parameter step counts and the skip rule for untouched grid entries (adam.h:79-82) are exercised too."""
    import torch

    n = gpu_ctx.network_n_params
    g = torch.Generator(device="cuda").manual_seed(17)
    params16 = (torch.randn(n, device="cuda", generator=g) * 0.05).half()
    refnet.set_params(params16)
    refnet.reset_optimizer()
    master, w16 = params16.float(), params16.clone()
    m1, m2 = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    steps = torch.zeros(n, dtype=torch.int32, device="cuda")
    ema = torch.zeros(n, dtype=torch.float16, device="cuda")
    r_master, r_w16, r_ema = torch.empty(n, device="cuda"), torch.empty(n, dtype=torch.float16, device="cuda"), torch.empty(n, dtype=torch.float16, device="cuda")
    for k in range(4):
        grads = (torch.randn(n, device="cuda", generator=g) * 2.0).half()
        grads[25600:][torch.rand(n - 25600, device="cuda", generator=g) < 0.9] = 0
        refnet.set_grads(grads)
        refnet.optimizer_step(128.0)
        refnet.optimizer_state(r_master, r_w16, r_ema)
        gpu_ctx.optimizer_step(k, master, w16, grads, m1, m2, steps, ema)
        torch.cuda.synchronize()
        assert torch.allclose(master, r_master, rtol=3e-5, atol=1e-7), "fp32 master weights differ at step %d: %g" % (k, float((master - r_master).abs().max()))
        # fp16 copies: the same fp32 value rounded -- allow the rare last-bit difference of the fp32 input
        dw = (w16.float() - r_w16.float()).abs()
        assert float(dw.max()) <= 1e-3 * float(r_w16.float().abs().max()) and float((dw > 0).float().mean()) < 1e-3, "fp16 weights differ at step %d" % k
        de = (ema.float() - r_ema.float()).abs()
        assert float(de.max()) <= 2e-3 * float(r_ema.float().abs().max() + 1e-6), "EMA weights differ at step %d: %g" % (k, float(de.max()))
    assert int(steps[:25600].min()) == 4 and int(steps[25600:].max()) <= 4 and int(steps[25600:].min()) == 0
