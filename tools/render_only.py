"""Renders a few 512^2 frames (the render leg of bench.py alone) -- for ncu captures of the render-regime kernels and for A/B runs of the march
schedule (MAX_STEPS = insta_render_opts.max_steps_per_iteration, comma-separated list; INSTA_DEFORM_COHERENT_MAX = records below which a tracer
batch uses the thread-per-query deformation kernel).  GPU box only."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from insta_b200 import capi

class A: pass
a = A(); a.render_res = int(os.environ.get("RES", 512)); a.render_frames = int(os.environ.get("FRAMES", 3))
verts, faces, masks, fv = bench.build_scene(8, 0, 1)
ctx = capi.Context(device=0)
ctx.set_canonical(verts, faces, masks); ctx.add_frames(fv)
for ms in os.environ.get("MAX_STEPS", "0").split(","):
    a.render_max_steps = int(ms)
    r = bench.render_leg(ctx, verts, fv, a, 0, 1)
    r.pop("note", None)
    print(json.dumps(r), flush=True)
