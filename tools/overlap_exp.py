"""Experiment: stage 1 of chunk k on one stream while stage 2 of chunk k-1 runs on another.  GPU box only (INSTA_TUNE build for the blocks/SM knob)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from insta_b200 import capi, synth

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1700
verts, faces, masks, fv = bench.build_scene(frames, 0, 1)
ctx = capi.Context(device=0)
ctx.set_canonical(verts, faces, masks); ctx.add_frames(fv)
layout = ctx.grid_layout()
table = torch.from_numpy(synth.make_grid_table(int(layout[0][-1]))).cuda()
rec_np, _ = bench.make_batch(verts, bench.N_SAMPLES, ctx.n_frames, "train", seed=7)
pristine = torch.from_numpy(rec_np).cuda()
N = bench.N_SAMPLES
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()

def run(chunks, blocks):
    if blocks: os.environ["INSTA_DEFORM_BLOCKS_PER_SM"] = str(blocks)
    else: os.environ.pop("INSTA_DEFORM_BLOCKS_PER_SM", None)
    C = N // chunks
    outs = [torch.empty((128, C), dtype=torch.float16, device="cuda") for _ in range(chunks)]
    ts = []
    for it in range(6):
        buf = pristine.clone()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(sA)
        for k in range(chunks):
            part = buf[k * C:(k + 1) * C]
            ctx.surface_closest_point(part, stream=sA)
            ev = torch.cuda.Event(); ev.record(sA)
            sB.wait_event(ev)
            ctx.hash_encode_fwd(part, table, outs[k], stream=sB)
        evb = torch.cuda.Event(); evb.record(sB); sA.wait_event(evb)
        e1.record(sA)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("chunks=%d deform blocks/SM=%s : %.3f ms (min %.3f)" % (chunks, blocks, np.mean(ts[2:]), np.min(ts)), flush=True)

for v in (None, "2"):
    if v is None: os.environ.pop("INSTA_DEFORM_VARIANT", None)
    else: os.environ["INSTA_DEFORM_VARIANT"] = v
    print("variant", v)
    run(1, None)
    for ch in (4, 8, 16):
        for b in (None, 2, 1):
            run(ch, b)
ctx.close()
