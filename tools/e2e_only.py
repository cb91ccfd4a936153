"""The e2e leg of bench.py alone: insta_deform_encode_host on pinned host buffers (INSTA_HOST_CHUNK_LOG2 = pipeline chunk, read at first use).  GPU box only."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from insta_b200 import capi, synth
verts, faces, masks, fv = bench.build_scene(int(os.environ.get("FRAMES", 400)), 0, 1)
ctx = capi.Context(device=0)
ctx.set_canonical(verts, faces, masks); ctx.add_frames(fv)
rec_np, _ = bench.make_batch(verts, bench.N_SAMPLES, ctx.n_frames, "train", seed=7); frame_id = -1
table = torch.from_numpy(synth.make_grid_table(int(ctx.grid_layout()[0][-1]))).cuda()
h0 = torch.from_numpy(rec_np).pin_memory(); h = torch.empty_like(h0).pin_memory()
out = torch.empty((128, bench.N_SAMPLES), dtype=torch.float16).pin_memory()
ts = []
for it in range(6):
    h.copy_(h0); torch.cuda.synchronize()
    a = time.perf_counter(); ctx.deform_encode_host(h, frame_id, table, out); b = time.perf_counter()
    if it: ts.append(b - a)
print("chunk 2^%s: %.2f ms per step (min %.2f) = %.1f M samples/s" % (os.environ.get("INSTA_HOST_CHUNK_LOG2", "19"), 1e3 * np.mean(ts), 1e3 * np.min(ts), bench.N_SAMPLES / np.mean(ts) / 1e6))
ctx.close()
