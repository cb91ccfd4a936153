"""GPU parity of the stage-3 render tracer against the REFERENCE's own CUDA kernels (oracle/_ref/libref_nerf.so) on
identical inputs: ray setup + skip to the first occupied cell, per-iteration live-ray counts, generated sample records
(bit-exact) and the composited frame (toleranced).  The network between generate and composite is a fixed analytic
function of the record evaluated with torch for both pipelines."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RES = (160, 120)
FOCAL = 300.0


def _bitfield(scene):
    import torch

    from insta_b200 import synth

    bf = synth.occupancy_bitfield(scene.frame_verts, radius=0.03)
    return torch.from_numpy(bf).cuda()


def _fake_network(records, n):
    """rgb/sigma logits from the record: smooth, deterministic, with densities large enough to saturate some rays."""
    import torch

    p = records[:n, 0:3]
    out = torch.empty((16, records.shape[0]), dtype=torch.float16, device=records.device)
    out.zero_()
    out[0, :n] = (torch.sin(p[:, 0] * 17.0) * 2.0).half()
    out[1, :n] = (torch.cos(p[:, 1] * 13.0) * 2.0).half()
    out[2, :n] = (torch.sin(p[:, 2] * 11.0 + 1.0) * 2.0).half()
    out[3, :n] = (4.0 + 3.0 * torch.sin(p[:, 0] * 40.0 + p[:, 1] * 30.0)).half()
    return out


@pytest.fixture(scope="module")
def refnerf():
    from oracle import pyrefnerf

    if not pyrefnerf.available():
        pytest.skip("oracle/_ref/libref_nerf.so not built (needs /root/reference at build time)")
    return pyrefnerf.RefNerf()


@pytest.mark.parametrize("snap,spp", [(True, 0), (False, 3)])
def test_init_rays_bit_exact(gpu_ctx, scene, refnerf, snap, spp):
    import torch

    from insta_b200 import capi, synth

    cam = synth.random_camera(1)
    bf = _bitfield(scene)
    n = RES[0] * RES[1]
    fr, fd = torch.zeros((n, 4), device="cuda"), torch.zeros(n, device="cuda")
    o_t, d_a = torch.zeros((n, 4), device="cuda"), torch.zeros((n, 4), device="cuda")
    aabb = gpu_ctx.render_aabb()
    n_alive = gpu_ctx.tracer_init_rays(capi.make_camera(cam, FOCAL, RES, sample_index=spp, snap=snap), fr, fd, bitfield=bf, dbg_origin_t=o_t, dbg_dir_alive=d_a)
    rfr, rfd = torch.zeros((n, 4), device="cuda"), torch.zeros(n, device="cuda")
    ro, rd = torch.zeros((n, 4), device="cuda"), torch.zeros((n, 4), device="cuda")
    refnerf.init_rays(cam, FOCAL, (0.5, 0.5), RES, spp, snap, aabb, [0, 0, 0, 1, 1, 1], bf, 0.0, 0.01, rfr, rfd)
    refnerf.export_init(ro, rd)
    torch.cuda.synchronize()
    alive_ref = rd[:, 3] > 0
    assert int(alive_ref.sum()) == n_alive and n_alive > 1000
    assert torch.equal(d_a[:, 3] > 0, alive_ref)
    # origin of every pixel; direction and t of the live ones (the reference never writes the direction of a ray that misses
    # the render box and leaves t of rays that die during the skip wherever the march stopped)
    assert torch.equal(d_a[alive_ref, :3].view(torch.int32), rd[alive_ref, :3].view(torch.int32)), "ray directions differ"
    assert torch.equal(o_t[:, :3].view(torch.int32), ro[:, :3].view(torch.int32))
    assert torch.equal(o_t[alive_ref, 3].view(torch.int32), ro[alive_ref, 3].view(torch.int32)), "t after the skip to the first occupied cell differs"
    assert torch.equal(fd, rfd)


def test_march_and_composite_vs_reference(gpu_ctx, scene, refnerf):
    import torch

    from insta_b200 import capi, synth

    cam = synth.random_camera(2)
    bf = _bitfield(scene)
    n = RES[0] * RES[1]
    aabb = gpu_ctx.render_aabb()
    extra = torch.linspace(-1, 1, 16, device="cuda")
    cap = n * 8
    # ---- ours ----
    fr, fd = torch.zeros((n, 4), device="cuda"), torch.zeros(n, device="cuda")
    rec = torch.zeros((cap, 27), device="cuda")
    n_alive = gpu_ctx.tracer_init_rays(capi.make_camera(cam, FOCAL, RES), fr, fd, bitfield=bf)
    ours_counts, ours_first = [], None
    while n_alive > 0:
        ns = gpu_ctx.tracer_n_steps()
        m = gpu_ctx.tracer_generate_samples(rec, ns, 1, extra)
        if ours_first is None:
            ours_first = rec[:m].clone()
        out = _fake_network(rec, m)
        ours_counts.append((n_alive, ns))
        n_alive = gpu_ctx.tracer_composite(rec, out, rec.shape[0], ns, fr, fd)
    # ---- reference ----
    rfr, rfd = torch.zeros((n, 4), device="cuda"), torch.zeros(n, device="cuda")
    rrec = torch.zeros((cap, 27), device="cuda")
    refnerf.init_rays(cam, FOCAL, (0.5, 0.5), RES, 0, True, aabb, [0, 0, 0, 1, 1, 1], bf, 0.0, 0.01, rfr, rfd)
    ref_counts, ref_first = [], None
    while True:
        na, ns = refnerf.compact_and_generate(rrec, 16, extra, 1)
        if na == 0:
            break
        m = na * ns
        if ref_first is None:
            ref_first = rrec[:m].clone()
        out = _fake_network(rrec, m)
        ref_counts.append((na, ns))
        refnerf.composite(rrec, 16, out, rrec.shape[0], ns)
    n_hit = refnerf.shade(rfr, rfd)
    torch.cuda.synchronize()
    assert ours_counts == ref_counts, "live-ray counts per iteration differ (occupancy-skip decisions)"
    assert n_hit > 1000
    # records of the first iteration, as multisets (the compaction order is atomic-order in both implementations)
    a, b = ours_first.cpu().numpy(), ref_first.cpu().numpy()
    valid_a, valid_b = a[a[:, 4] >= 0], b[b[:, 4] >= 0]
    assert len(valid_a) == len(valid_b)
    ka = np.lexsort(valid_a[:, :11].view(np.uint32).T[::-1])
    kb = np.lexsort(valid_b[:, :11].view(np.uint32).T[::-1])
    assert (valid_a[ka].view(np.uint32) == valid_b[kb].view(np.uint32)).all(), "generated sample records differ"
    # frame buffer: RGBA 1e-3, depth 1e-4
    assert torch.equal(fr[:, 3] > 0, rfr[:, 3] > 0)
    assert float((fr - rfr).abs().max()) < 1e-3
    hit = rfr[:, 3] > 0.2
    assert int(hit.sum()) > 500
    assert float((fd[hit] - rfd[hit]).abs().max()) < 1e-4
    assert torch.equal(fd[~hit], rfd[~hit])


def test_render_frame_vs_reference_tracer(gpu_ctx, scene, refnerf):
    """insta_render_frame (our tracer + deformation + fused network) against the REFERENCE's tracer kernels driving the same
    deformation and network between generate and composite: identical sample sets => frames within the RGBA / depth tolerance."""
    import torch

    from insta_b200 import capi, synth

    cam = synth.random_camera(3)
    bf = _bitfield(scene)
    n = RES[0] * RES[1]
    aabb = gpu_ctx.render_aabb()
    extra = torch.linspace(-1, 1, 16, device="cuda")
    pn = synth.make_network_params(gpu_ctx.network_n_params, seed=9)
    # density logits of a random network are ~0 (sigma = exp(logit) ~ 1, alpha per step ~ 0.002): make row 0 of the density
    # output layer positive so that the logit grows with the (non-negative) hidden activations and rays saturate
    pn[13312:13312 + 64] = np.float16(0.6)
    params = torch.from_numpy(pn).cuda()
    fr, fd = torch.zeros((n, 4), device="cuda"), torch.zeros(n, device="cuda")
    iters, samples = gpu_ctx.render_frame(capi.make_camera(cam, FOCAL, RES), 1, params, fr, fd, bitfield=bf, extra_dims=extra, max_steps=8)  # the reference's schedule
    assert iters > 3 and samples > 10000
    rfr, rfd = torch.zeros((n, 4), device="cuda"), torch.zeros(n, device="cuda")
    rrec = torch.zeros((n * 8, 27), device="cuda")
    rout = torch.zeros((16, n * 8), dtype=torch.float16, device="cuda")
    refnerf.init_rays(cam, FOCAL, (0.5, 0.5), RES, 0, True, aabb, [0, 0, 0, 1, 1, 1], bf, 0.0, 0.01, rfr, rfd)
    ref_iters = 0
    while True:
        na, ns = refnerf.compact_and_generate(rrec, 16, extra, 1)
        if na == 0:
            break
        m = na * ns
        gpu_ctx.render_deformed(rrec[:m], 1)
        gpu_ctx.network_inference(rrec, params, rout, out_stride=rout.shape[1], n=m)
        refnerf.composite(rrec, 16, rout, rout.shape[1], ns)
        ref_iters += 1
    refnerf.shade(rfr, rfd)
    torch.cuda.synchronize()
    assert ref_iters == iters
    assert torch.equal(fr[:, 3] > 0, rfr[:, 3] > 0)
    assert float((fr - rfr).abs().max()) < 1e-3
    hit = rfr[:, 3] > 0.2
    assert int(hit.sum()) > 500
    assert float((fd[hit] - rfd[hit]).abs().max()) < 1e-4


def test_render_frame_schedule_independent(gpu_ctx, scene):
    """The number of march steps per iteration (insta_render_opts.max_steps_per_iteration: 8 = the reference's schedule, 0 = the library's larger
    batches, 1 = one step per iteration) changes how many iterations a frame takes and which deformation kernel its batches use -- not the frame:
    a ray's samples and their compositing order are the same, so RGBA and depth must be bit-identical."""
    import torch

    from insta_b200 import capi, synth

    cam = capi.make_camera(synth.random_camera(5), FOCAL, RES)
    bf = _bitfield(scene)
    n = RES[0] * RES[1]
    extra = torch.linspace(-1, 1, 16, device="cuda")
    pn = synth.make_network_params(gpu_ctx.network_n_params, seed=11)
    pn[13312:13312 + 64] = np.float16(0.6)  # rays saturate (see test_render_frame_vs_reference_tracer)
    params = torch.from_numpy(pn).cuda()
    frames = {}
    for ms in (8, 0, 1, 3, 40):  # 40: more than one 16-step chunk of the march kernel per iteration
        fr, fd = torch.zeros((n, 4), device="cuda"), torch.zeros(n, device="cuda")
        iters, samples = gpu_ctx.render_frame(cam, 1, params, fr, fd, bitfield=bf, extra_dims=extra, max_steps=ms)
        torch.cuda.synchronize()
        frames[ms] = (fr, fd, iters, samples)
    ref = frames[8]
    assert int((ref[0][:, 3] > 0.2).sum()) > 500
    assert frames[0][2] < ref[2] < frames[1][2]  # fewer, larger iterations by default; one step per iteration needs the most
    for ms in (0, 1, 3, 40):
        assert torch.equal(frames[ms][0].view(torch.int32), ref[0].view(torch.int32)), ms
        assert torch.equal(frames[ms][1].view(torch.int32), ref[1].view(torch.int32)), ms


@pytest.mark.parametrize("env", [{"INSTA_DEFORM_COHERENT_MAX": "0"}, {"INSTA_RENDER_RECORDS": "full"}, {"INSTA_RENDER_RECORDS": "full", "INSTA_DEFORM_SEED": "0"}])
def test_render_frame_internal_variants(scene, gpu_ctx, monkeypatch, env):
    """insta_render_frame's internal choices do not change the frame, bit for bit: (a) large tracer batches go through the warp-queue deformation
    kernel, small ones through the thread-per-query kernel -- a context that uses the warp-queue kernel for every batch; (b) between its own kernels
    the records are the 11 NerfCoordinate floats (expression values from the frame's vector, `all ones` flagged in the holder slot) -- a context that
    keeps the reference's 27-float records; (c) additionally without the seed bound of the traversal."""
    import torch

    from conftest import _make_ctx
    from insta_b200 import capi, synth

    cam = capi.make_camera(synth.random_camera(5), FOCAL, RES)
    bf = _bitfield(scene)
    n = RES[0] * RES[1]
    extra = torch.linspace(-1, 1, 16, device="cuda")
    pn = synth.make_network_params(gpu_ctx.network_n_params, seed=11)
    pn[13312:13312 + 64] = np.float16(0.6)
    params = torch.from_numpy(pn).cuda()
    fr, fd = torch.zeros((n, 4), device="cuda"), torch.zeros(n, device="cuda")
    gpu_ctx.render_frame(cam, 1, params, fr, fd, bitfield=bf, extra_dims=extra)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    ctx = _make_ctx(scene)
    try:
        fr2, fd2 = torch.zeros((n, 4), device="cuda"), torch.zeros(n, device="cuda")
        ctx.render_frame(cam, 1, params, fr2, fd2, bitfield=bf, extra_dims=extra)
        torch.cuda.synchronize()
    finally:
        ctx.close()
    assert int((fr[:, 3] > 0.2).sum()) > 500
    assert torch.equal(fr2.view(torch.int32), fr.view(torch.int32))
    assert torch.equal(fd2.view(torch.int32), fd.view(torch.int32))
