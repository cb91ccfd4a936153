"""Renders a few 512^2 frames (the render leg of bench.py alone) -- for ncu captures of the render-regime kernels.  GPU box only."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from insta_b200 import capi

class A: pass
a = A(); a.render_res = 512; a.render_frames = int(os.environ.get("FRAMES", 3))
verts, faces, masks, fv = bench.build_scene(8, 0, 1)
ctx = capi.Context(device=0)
ctx.set_canonical(verts, faces, masks); ctx.add_frames(fv)
print(json.dumps(bench.render_leg(ctx, verts, fv, a, 0, 1)))
