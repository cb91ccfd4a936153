"""A/B timing of the deform kernel variants (library built with INSTA_TUNE=1).  GPU box only."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from insta_b200 import capi

class A: pass
def run(regime, frames, variants, blocks=(None,)):
    a = A(); a.regime = regime; a.frames = frames
    verts, faces, masks, fv = bench.build_scene(frames if regime == "train" else 1, 0, 1)
    ctx = capi.Context(device=0)
    ctx.set_canonical(verts, faces, masks); ctx.add_frames(fv)
    rec_np, _ = bench.make_batch(verts, bench.N_SAMPLES, ctx.n_frames, regime, seed=7)
    pristine = torch.from_numpy(rec_np).cuda()
    bufs = [pristine.clone() for _ in range(8)]
    fid = -1 if regime == "train" else 0
    os.environ.pop("INSTA_DEFORM_VARIANT", None)
    ref = pristine.clone()  # the default kernel's output: every variant must reproduce it bit for bit
    if fid < 0: ctx.surface_closest_point(ref)
    else: ctx.render_deformed(ref, fid)
    torch.cuda.synchronize()
    for v in variants:
        for b in blocks:
            if v is None: os.environ.pop("INSTA_DEFORM_VARIANT", None)
            else: os.environ["INSTA_DEFORM_VARIANT"] = str(v)
            if b is None: os.environ.pop("INSTA_DEFORM_BLOCKS_PER_SM", None)
            else: os.environ["INSTA_DEFORM_BLOCKS_PER_SM"] = str(b)
            for x in bufs: x.copy_(pristine)
            ts = []
            for k in range(8):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                if fid < 0: ctx.surface_closest_point(bufs[k])
                else: ctx.render_deformed(bufs[k], fid)
                e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            same = bool(torch.equal(bufs[7].view(torch.int32), ref.view(torch.int32)))
            print("%s frames=%d variant=%s blocks/SM=%s : %.3f ms (min %.3f) identical=%s" % (regime, frames, v, b, np.mean(ts[2:]), np.min(ts), same), flush=True)
    ctx.close()

if __name__ == "__main__":
    vs = [int(x) if x != "d" else None for x in (sys.argv[1] if len(sys.argv) > 1 else "0,1,2,3,4,5,6,d").split(",")]
    bl = [int(x) if x != "d" else None for x in (sys.argv[2] if len(sys.argv) > 2 else "d").split(",")]
    run("render", 1, vs, bl)
    run("train", int(sys.argv[3]) if len(sys.argv) > 3 else 400, vs, bl)
