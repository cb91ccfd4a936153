import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from insta_b200 import capi, synth
from oracle.pyoracle import Oracle
from oracle import pyrefnerf
from common import Scene
O = Oracle(); sc = Scene(O)
ctx = capi.Context(device=0); ctx.set_canonical(sc.verts, sc.faces, sc.masks)
for v in sc.frame_verts: ctx.add_frame(v)
R = pyrefnerf.RefNerf()
RES = (160, 120); FOCAL = 300.0
bf = torch.from_numpy(synth.occupancy_bitfield(sc.frame_verts)).cuda()
n = RES[0] * RES[1]
for snap, spp in ((True, 0), (False, 3)):
    cam = synth.random_camera(1)
    fr, fd = torch.zeros((n, 4), device="cuda"), torch.zeros(n, device="cuda")
    o_t, d_a = torch.zeros((n, 4), device="cuda"), torch.zeros((n, 4), device="cuda")
    na = ctx.tracer_init_rays(capi.make_camera(cam, FOCAL, RES, sample_index=spp, snap=snap), fr, fd, bitfield=bf, dbg_origin_t=o_t, dbg_dir_alive=d_a)
    rfr, rfd = torch.zeros((n, 4), device="cuda"), torch.zeros(n, device="cuda")
    ro, rd = torch.zeros((n, 4), device="cuda"), torch.zeros((n, 4), device="cuda")
    R.init_rays(cam, FOCAL, (0.5, 0.5), RES, spp, snap, ctx.render_aabb(), [0, 0, 0, 1, 1, 1], bf, 0.0, 0.01, rfr, rfd)
    R.export_init(ro, rd); torch.cuda.synchronize()
    al = rd[:, 3] > 0
    print("snap", snap, "alive ours/ref", na, int(al.sum()), "mask equal", bool(torch.equal(d_a[:, 3] > 0, al)))
    for name, a, b in (("dir", d_a[al, :3], rd[al, :3]), ("origin", o_t[:, :3], ro[:, :3]), ("t", o_t[al, 3:4], ro[al, 3:4])):
        ai, bi = a.contiguous().view(torch.int32).long(), b.contiguous().view(torch.int32).long()
        d = (ai - bi).abs()
        print("  %s: differing %d of %d, max ulp %d, per-comp differing %s, max abs %.3g" % (name, int((d > 0).sum()), d.numel(), int(d.max()), (d > 0).sum(0).tolist(), float((a - b).abs().max())))
    print("  depth eq", bool(torch.equal(fd, rfd)))
