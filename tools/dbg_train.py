import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import test_gpu_train as T
from insta_b200 import capi, synth
from oracle.pyoracle import Oracle
from oracle import pyrefnerf
from common import Scene
O = Oracle(); sc = Scene(O)
ctx = capi.Context(device=0); ctx.set_canonical(sc.verts, sc.faces, sc.masks)
R = pyrefnerf.RefNerf()
ims = synth.make_training_images(T.N_IMAGES, T.RES, T.FOCAL, seed=3)
dev = []
for im in ims:
    d = dict(im)
    for k in ("pixels", "seg", "depth", "rays"): d[k] = torch.from_numpy(np.ascontiguousarray(im[k])).cuda()
    dev.append(d)
bf = torch.from_numpy(synth.occupancy_bitfield(sc.frame_verts, radius=0.03)).cuda()
extra = torch.randn((T.N_IMAGES, 16), device="cuda")
step, ours, ref = T._generate_both(ctx, R, (dev, bf, extra), 1 << 20)
a, na = T._by_ray(ours); b, nb = T._by_ray(ref)
ca, cb = ours["coords"].cpu().numpy(), ref["coords"].cpu().numpy()
cols = np.zeros(27, np.int64); tot = 0; maxd = np.zeros(27)
bad_rays = 0
for i in a:
    (n, oa), (_, ob) = a[i], b[i]
    x, y = ca[oa:oa+n], cb[ob:ob+n]
    ne = x.view(np.uint32) != y.view(np.uint32)
    cols += ne.sum(0); tot += n
    maxd = np.maximum(maxd, np.abs(x - y).max(0))
    bad_rays += int(ne.any())
print("samples", tot, "bad rays", bad_rays, "of", len(a))
print("mismatching per column", cols.tolist())
print("max abs diff per column", np.round(maxd, 9).tolist())
ra, rb = ours["rays"].cpu().numpy(), ref["rays"].cpu().numpy()
