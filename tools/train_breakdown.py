"""Runs only the train-step leg of bench.py (for `ncu --metrics gpu__time_duration.sum` launch lists).  GPU box only."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from insta_b200 import capi

class A: pass
a = A(); a.train_images = int(os.environ.get("IMAGES", 128)); a.train_steps = int(os.environ.get("STEPS", 4)); a.train_warmup = int(os.environ.get("WARMUP", 40)); 
verts, faces, masks, fv = bench.build_scene(a.train_images, 0, 1)
ctx = capi.Context(device=0)
ctx.set_canonical(verts, faces, masks); ctx.add_frames(fv)
print(json.dumps(bench.train_leg(ctx, verts, fv, a, 0, 1)))
