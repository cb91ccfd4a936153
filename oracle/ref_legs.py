"""TEST INFRASTRUCTURE ONLY: the legs of bench.py run with the REFERENCE's own CUDA code (oracle/_ref: `warp`, kernel_grid / kernel_grid_backward,
ngp::NerfNetwork<__half> over tiny-cuda-nn, the NeRF tracer / training kernels, tcnn's Ema{Adam}) compiled for sm_100, on the same GPU, buffers and workloads.
Used by tests/bench_ref_cuda.py and by bench.py's `ref_cuda` leg, which runs AFTER the product's timed region, on rank 0 at N = 1, and only reports the
reference beside the product (the north star's ">= 5x the reference's own CUDA path" needs both numbers from one run on one box).  Nothing here is
ever on a product path."""
import numpy as np

N_SAMPLES = 1 << 22
STRIDE = 27
UNIT = "samples/s"


def ref_cuda_leg(verts, faces, masks, fv, rec_np, table, out, args):
    """Reported beside ours: the reference's own CUDA kernels (`warp` + kernel_grid) recompiled for sm_100, same buffers."""
    import torch

    from oracle import pyoracle, pyrefcuda

    if not pyrefcuda.available():
        return {"unavailable": "oracle/_ref/libref_cuda.so not built"}
    O = pyoracle.Oracle()
    canon = O.build_triangles(verts, faces, masks)
    adj = O.adjacency(faces, len(verts))
    nodes, tbvh, ttopo = [], [], []
    for v in fv:
        t = O.build_triangles(v, faces, masks, canon)
        nn, tb = O.bvh_build(t)
        nodes.append(nn); tbvh.append(tb); ttopo.append(t)
    R = pyrefcuda.RefCuda()
    R.upload_scene(nodes, tbvh, ttopo, canon, adj)
    pristine = torch.from_numpy(rec_np).cuda()
    buf = pristine.clone()
    fid = -1 if args.regime == "train" else 0
    ts_w, ts_g = [], []
    for it in range(2 + 5):
        buf.copy_(pristine)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        R.warp(buf, STRIDE, fid)
        e[1].record()
        R.grid_forward(buf, STRIDE, table, out)
        e[2].record()
        torch.cuda.synchronize()
        if it >= 2:
            ts_w.append(e[0].elapsed_time(e[1]))
            ts_g.append(e[1].elapsed_time(e[2]))
    ms = float(np.mean(ts_w) + np.mean(ts_g))
    res = {"value": N_SAMPLES / (ms * 1e-3), "unit": UNIT, "warp_ms": float(np.mean(ts_w)), "kernel_grid_ms": float(np.mean(ts_g)),
           "note": "reference `warp` (src/core.cu:442) + kernel_grid (tcnn grid.h:138), nvcc 12.9 -arch=sm_100, identical device buffers"}
    # stage 2 backward (scatter-add) on the 2^18-sample training batch: reference kernel_grid_backward (+ its memset) vs insta_hash_encode_bwd
    try:
        nb = 1 << 18
        g = torch.Generator(device="cuda").manual_seed(1)
        dy = (torch.randn((128, nb), device="cuda", generator=g) * 0.01).half()
        grad_r = torch.empty((table.shape[0], 8), dtype=torch.float16, device="cuda")
        pos = pristine[:nb].contiguous()
        tr = []
        for it in range(2 + 5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            grad_r.zero_()
            R.grid_backward(pos, STRIDE, dy, grad_r)
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                tr.append(e0.elapsed_time(e1))
        res["kernel_grid_backward_ms_2p18"] = float(np.mean(tr))
        from insta_b200 import capi as _capi
        c2 = _capi.Context(device=torch.cuda.current_device())
        to = []
        for it in range(2 + 5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            c2.hash_encode_bwd(pos, dy, grad_r, zero_first=True)
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                to.append(e0.elapsed_time(e1))
        res["ours_hash_bwd_ms_2p18"] = float(np.mean(to))
        c2.close()
        del dy, grad_r
    except Exception as e:
        res["backward_error"] = repr(e)
    # the reference's own network (ngp::NerfNetwork<__half> over tiny-cuda-nn, oracle/_ref/libref_net.so) on the same batch
    try:
        from oracle import pyrefnet

        if pyrefnet.available():
            from insta_b200 import synth

            del buf
            torch.cuda.empty_cache()
            N = pyrefnet.RefNet()
            params = torch.from_numpy(synth.make_network_params(N.n_params, seed=1337)).cuda()
            N.set_params(params)
            out16 = torch.empty((16, N_SAMPLES), dtype=torch.float16, device="cuda")
            ts = []
            for it in range(2 + 4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                N.inference(pristine, out16)
                e1.record()
                torch.cuda.synchronize()
                if it >= 2:
                    ts.append(e0.elapsed_time(e1))
            res["network_inference_ms"] = float(np.mean(ts))
            del out16
            nb = 1 << 18
            dl = torch.zeros((nb, 16), dtype=torch.float16, device="cuda")
            dl[:, :4] = 0.01
            o = torch.empty((nb, 16), dtype=torch.float16, device="cuda")
            gr = torch.empty(N.n_params, dtype=torch.float16, device="cuda")
            ts = []
            for it in range(2 + 4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                N.forward_backward(pristine[:nb], dl, o, gr)
                e1.record()
                torch.cuda.synchronize()
                if it >= 2:
                    ts.append(e0.elapsed_time(e1))
            res["network_forward_backward_ms_2p18"] = float(np.mean(ts))
            res["network_note"] = "ngp::NerfNetwork<__half>::inference_mixed_precision on the 4,194,304-record batch; forward + backward (Overwrite) on 2^18 records"
    except Exception as e:
        res["network_error"] = repr(e)
    return res



def ref_train_leg(verts, frame_verts, dev_images, bitfield, extra, params16, n_rays, mean_density, args):
    """The same training step composed from the REFERENCE's own CUDA code (oracle/_ref: generate_training_samples_nerf, `warp`, NerfNetwork inference,
    compute_loss_kernel_train_nerf + roll-over, NerfNetwork forward + backward, tcnn Ema{Adam} step, counter read-back) in the order and with the record
    counts of Testbed::train_nerf_step (testbed_nerf.cu:3170-3426: deformation and inference over `max_inference` = last step's sample count rounded up),
    on the same images, frames, occupancy grid, ray count and (trained) parameters as our step above.  Reported beside ours; test infrastructure only."""
    import torch

    from oracle import pyoracle, pyrefcuda, pyrefnerf, pyrefnet

    if not (pyrefcuda.available() and pyrefnerf.available() and pyrefnet.available()):
        return {"unavailable": "oracle/_ref libraries (refcuda, refnerf, refnet) not all built"}
    from insta_b200 import synth

    O = pyoracle.Oracle()
    vts, faces, regions = synth.load_canonical()
    masks = synth.mask_ids_from_regions(faces, regions)
    canon = O.build_triangles(vts, faces, masks)
    adj = O.adjacency(faces, len(vts))
    nodes, tbvh, ttopo = [], [], []
    for v in frame_verts:
        t = O.build_triangles(v, faces, masks, canon)
        nn, tb = O.bvh_build(t)
        nodes.append(nn); tbvh.append(tb); ttopo.append(t)
    R = pyrefcuda.RefCuda()
    R.upload_scene(nodes, tbvh, ttopo, canon, adj)
    G = pyrefnerf.RefNerf()
    G.train_set_images(dev_images)
    N = pyrefnet.RefNet()
    N.set_params(params16)
    target, max_samples = 1 << 18, 16 << 18
    coords = torch.zeros((max_samples, STRIDE), device="cuda")
    counters = torch.zeros(4, dtype=torch.int32, device="cuda")
    ray_indices = torch.zeros(1 << 18, dtype=torch.int32, device="cuda")
    rays = torch.zeros((1 << 18, 6), device="cuda")
    numsteps = torch.zeros((1 << 18, 2), dtype=torch.int32, device="cuda")
    out_cm = torch.zeros((max_samples, 16), dtype=torch.float16, device="cuda")
    cc = torch.zeros((target, STRIDE), device="cuda")
    dl = torch.zeros((target, 16), dtype=torch.float16, device="cuda")
    out_cc = torch.zeros((target, 16), dtype=torch.float16, device="cuda")
    loss = torch.zeros(1 << 18, device="cuda")
    grads = torch.zeros(N.n_params, dtype=torch.float16, device="cuda")
    unit = [0, 0, 0, 1, 1, 1]
    max_inference, n_total, ts, gen = max_samples, 0, [], 0
    for it in range(3 + args.ref_train_steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        N.set_params(params16)  # outside the timed region: every step sees the parameters our step saw (the reference's fresh Adam state would drift)
        e0.record()
        c = coords[:max_inference]
        G.train_generate(n_rays, n_total, c, counters, ray_indices, rays, numsteps, bitfield, unit, extra, rng_advances=it)
        R.warp(c, STRIDE, -1)
        N.inference_cm(c, out_cm)
        G.train_loss(n_rays, n_total, target, c, counters, ray_indices, rays, numsteps, out_cm, 16, cc, dl, loss, mean_density, unit, rng_advances=it)
        N.forward_backward(cc, dl, out_cc, grads)
        N.optimizer_step(128.0)
        cnt = counters.cpu()  # Counters::update_after_training reads both counters back every step
        e1.record()
        torch.cuda.synchronize()
        n_total += n_rays
        gen = int(cnt[1])
        if gen == 0:
            break
        max_inference = min((min(gen, max_samples) + 127) // 128 * 128, max_samples)
        if it >= 3:
            ts.append(e0.elapsed_time(e1))
    if not ts:
        return {"error": "the reference step generated no samples"}
    return {"ms_per_step": float(np.mean(ts)), "ms_per_step_min": float(np.min(ts)), "steps_timed": len(ts), "rays_per_step": int(n_rays), "samples_generated": gen,
            "samples_compacted": int(cnt[2]),
            "note": "reference kernels / network / optimizer (nvcc 12.9 -arch=sm_100) composed as Testbed::train_nerf_step does, same scene, rays and parameters"}


def ref_render_leg(ctx, fv, params, bf, extra, focal, res, cams_frames):
    """Testbed::render_nerf composed from the REFERENCE's own CUDA code (oracle/_ref: init_rays_with_payload_kernel_nerf + advance_pos_nerf, compact_kernel_nerf,
    generate_next_nerf_network_inputs, `warp` with the frame fixed, NerfNetwork inference, composite_kernel_nerf, shade_kernel_nerf; one host read-back per march
    iteration as NerfTracer::trace does) for the same cameras, frames, occupancy grid and parameters as our render leg.  Test infrastructure only."""
    import torch

    from oracle import pyoracle, pyrefcuda, pyrefnerf, pyrefnet

    if not (pyrefcuda.available() and pyrefnerf.available() and pyrefnet.available()):
        return {"unavailable": "oracle/_ref libraries (refcuda, refnerf, refnet) not all built"}
    from insta_b200 import synth

    O = pyoracle.Oracle()
    vts, faces, regions = synth.load_canonical()
    masks = synth.mask_ids_from_regions(faces, regions)
    canon = O.build_triangles(vts, faces, masks)
    adj = O.adjacency(faces, len(vts))
    frames = sorted(set(f for _, f in cams_frames))
    nodes, tbvh, ttopo = [], [], []
    for f in range(max(frames) + 1):
        t = O.build_triangles(fv[f], faces, masks, canon)
        nn, tb = O.bvh_build(t)
        nodes.append(nn); tbvh.append(tb); ttopo.append(t)
    R = pyrefcuda.RefCuda()
    R.upload_scene(nodes, tbvh, ttopo, canon, adj)
    G = pyrefnerf.RefNerf()
    N = pyrefnet.RefNet()
    N.set_params(params)
    n = res[0] * res[1]
    aabb = ctx.render_aabb()
    fr, fd = torch.zeros((n, 4), device="cuda"), torch.zeros(n, device="cuda")
    cap = (n + 127) // 128 * 128
    rec = torch.zeros((cap, STRIDE), device="cuda")
    out = torch.zeros((16, cap), dtype=torch.float16, device="cuda")
    ms, iters = [], []
    for k, (cam, f) in enumerate(cams_frames):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fr.zero_()
        G.init_rays(cam, focal, (0.5, 0.5), res, 0, True, aabb, [0, 0, 0, 1, 1, 1], bf, 0.0, 0.01, fr, fd)
        it = 0
        while True:
            na, ns = G.compact_and_generate(rec, 16, extra, f)
            if na == 0:
                break
            m = na * ns
            mp = (m + 127) // 128 * 128
            R.warp(rec[:m], STRIDE, f)
            N.inference(rec[:mp], out[:, :mp])
            G.composite(rec, 16, out[:, :mp], mp, ns)
            it += 1
        G.shade(fr, fd)
        e1.record()
        torch.cuda.synchronize()
        if k >= 2:
            ms.append(e0.elapsed_time(e1)); iters.append(it)
    return {"ms_per_frame": float(np.mean(ms)), "ms_per_frame_min": float(np.min(ms)), "frames_timed": len(ms), "march_iterations_per_frame": float(np.mean(iters)),
            "hit_fraction": float((fr[:, 3] > 0.2).float().mean()),
            "note": "reference tracer kernels + `warp` + NerfNetwork (nvcc 12.9 -arch=sm_100) composed as Testbed::render_nerf / NerfTracer::trace, same cameras / frames / parameters"}
