"""Runs the stage-1 kernel alone on the headline batch (train regime) -- the command ncu captures.  GPU box only.
usage: python tools/prof_deform.py [frames=200] [launches=4]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from insta_b200 import capi

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 200
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 4
verts, faces, masks, fv = bench.build_scene(frames, 0, 1)
ctx = capi.Context(device=0)
ctx.set_canonical(verts, faces, masks)
ctx.add_frames(fv)
rec_np, _ = bench.make_batch(verts, bench.N_SAMPLES, ctx.n_frames, "train", seed=7)
pristine = torch.from_numpy(rec_np).cuda()
for k in range(launches):
    buf = pristine.clone()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.surface_closest_point(buf); e1.record(); torch.cuda.synchronize()
    print("launch %d: %.3f ms" % (k, e0.elapsed_time(e1)), flush=True)
ctx.close()
