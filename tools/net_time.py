"""Times the fused network kernel alone (bench.network_leg) and a few 512^2 frames.  GPU box only."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from insta_b200 import capi
class A: pass
a = A(); a.render_res = 512; a.render_frames = 6; a.render_max_steps = 0
verts, faces, masks, fv = bench.build_scene(8, 0, 1)
ctx = capi.Context(device=0)
ctx.set_canonical(verts, faces, masks); ctx.add_frames(fv)
rec_np, _ = bench.make_batch(verts, bench.N_SAMPLES, ctx.n_frames, "train", seed=7)
r = bench.network_leg(ctx, rec_np, a); r.pop("note", None)
print(json.dumps(r))
r = bench.render_leg(ctx, verts, fv, a, 0, 1); r.pop("note", None); r.pop("ref_cuda", None)
print(json.dumps(r))
