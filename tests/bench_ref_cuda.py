#!/usr/bin/env python
"""TEST INFRASTRUCTURE (lives under tests/ because it executes oracle/_ref): the legs of bench.py side by side with the REFERENCE's own CUDA code
(oracle/_ref: `warp`, kernel_grid / kernel_grid_backward, ngp::NerfNetwork<__half> over tiny-cuda-nn, the NeRF tracer / training kernels, tcnn's
Ema{Adam}) compiled for sm_100, on the same GPU, buffers and workloads.  bench.py itself never touches oracle/ outside its CPU baseline.

    python tests/bench_ref_cuda.py [--steps K] [--warmup W] > profiles/rNN_vs_ref_cuda.json       (one GPU)

Prints one JSON object: our numbers (bench.py's functions), the reference's, and the ratios the north star asks for."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bench import N_SAMPLES, STRIDE, UNIT  # noqa: E402


from oracle.ref_legs import ref_cuda_leg, ref_render_leg, ref_train_leg  # noqa: E402,F401


def main():
    import torch

    from insta_b200 import capi, synth

    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=1700)
    ap.add_argument("--train-images", type=int, default=128)
    ap.add_argument("--train-steps", type=int, default=20)
    ap.add_argument("--train-warmup", type=int, default=70, help="untimed steps before the timed ones: rays_per_batch grows ~9 %% per step from 4096 and needs ~65 steps to reach its fixed point")
    ap.add_argument("--ref-train-steps", type=int, default=6)
    ap.add_argument("--render-res", type=int, default=512)
    ap.add_argument("--render-frames", type=int, default=8)
    args = ap.parse_args()
    args.regime = "train"
    assert torch.cuda.is_available()
    capi.load_library(build_if_missing=False)
    verts, faces, masks, fv = bench.build_scene(args.frames, 0, 1)
    ctx = capi.Context(device=0)
    ctx.set_canonical(verts, faces, masks)
    ctx.add_frames(fv)
    layout = ctx.grid_layout()
    table = torch.from_numpy(synth.make_grid_table(int(layout[0][-1]))).cuda()
    rec_np, _ = bench.make_batch(verts, N_SAMPLES, ctx.n_frames, "train", seed=7)
    out = torch.empty((128, N_SAMPLES), dtype=torch.float16, device="cuda")
    # ---- headline batch: ours ----
    pristine = torch.from_numpy(rec_np).cuda()
    buf = pristine.clone()
    td, te = [], []
    for it in range(args.warmup + args.steps):
        buf.copy_(pristine)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        ctx.surface_closest_point(buf)
        e[1].record()
        ctx.hash_encode_fwd(buf, table, out)
        e[2].record()
        torch.cuda.synchronize()
        if it >= args.warmup:
            td.append(e[0].elapsed_time(e[1])); te.append(e[1].elapsed_time(e[2]))
    del buf, pristine
    ours = {"deform_ms": float(np.mean(td)), "hash_fwd_ms": float(np.mean(te)), "value": N_SAMPLES / ((np.mean(td) + np.mean(te)) * 1e-3), "unit": UNIT}
    ours["network"] = bench.network_leg(ctx, rec_np, args)
    ref = ref_cuda_leg(verts, faces, masks, fv, rec_np, table, out, args)
    del out
    torch.cuda.empty_cache()
    train = bench.train_leg(ctx, verts, fv, args, 0, 1, ref_hook=ref_train_leg)
    render = bench.render_leg(ctx, verts, fv, args, 0, 1, ref_hook=ref_render_leg)
    ref["train_step"] = train.pop("ref_cuda")
    ref["render"] = render.pop("ref_cuda")
    ours["train_step"], ours["render"] = train, render

    def ratio(a, b):
        return None if not a or not b else a / b

    res = {"what": "this library against the reference's own CUDA code (nvcc 12.9, -arch=sm_100) on the same B200, buffers and workloads", "ours": ours, "ref_cuda": ref,
           "speedup": {"deformed+encoded samples/s": ratio(ours["value"], ref.get("value")),
                       "network inference": ratio(ref.get("network_inference_ms"), ours["network"]["ms"]),
                       "network forward+backward (2^18)": ratio(ref.get("network_forward_backward_ms_2p18"), ours["network"]["forward_backward_ms_2p18"]),
                       "hash backward (2^18)": ratio(ref.get("kernel_grid_backward_ms_2p18"), ref.get("ours_hash_bwd_ms_2p18")),
                       # the ray count still drifts while the model trains: compare at the ray count the reference leg ran (= our last timed steps)
                       "training step": ratio((ref["train_step"] or {}).get("ms_per_step"), float(np.mean(train["ms_steps"][-4:]))),
                       "training step: ours ms at the reference leg's ray count": float(np.mean(train["ms_steps"][-4:])),
                       "512^2 frame": ratio((ref["render"] or {}).get("ms_per_frame"), render["ms_per_frame"])}}
    print(json.dumps(res))
    ctx.close()


if __name__ == "__main__":
    main()
