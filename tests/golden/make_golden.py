"""Generates tests/golden/stage1_ref.npz by running the REFERENCE's own host code (oracle/_ref/libref_host.so, built by
`make -C oracle ref` from /root/reference) on seeded synthetic inputs.  Run in the build container only; the committed
.npz is what travels.  Inputs are regenerated from the same seeds by the tests (insta_b200.synth), only the
reference's OUTPUTS are stored.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
from insta_b200 import synth  # noqa: E402
from oracle.pyoracle import RefHost, Oracle  # noqa: E402

N_POINTS = 8192
FRAMES = (0, 7)


def inputs():
    verts, faces, regions = synth.load_canonical()
    masks = synth.mask_ids_from_regions(faces, regions)
    fs = synth.FrameSynth(verts, faces)
    aabb = synth.mesh_aabb(verts)
    fids = np.repeat(np.arange(len(FRAMES)), N_POINTS // len(FRAMES)).astype(np.int32)
    rec = synth.make_records(N_POINTS, aabb, fids, seed=11, pad_fraction=0.02)
    return verts, faces, masks, [fs.frame(f) for f in FRAMES], aabb, rec


def main():
    R = RefHost()
    O = Oracle()  # only for the (already reference-verified) triangle assembly helpers
    verts, faces, masks, frame_verts, aabb, rec = inputs()
    canon = O.build_triangles(verts, faces, masks)
    adj = O.adjacency(faces, len(verts))
    out = {}
    cn, ct = R.bvh_build(canon)
    out["canon_nodes_sha"] = np.frombuffer(hashlib.sha256(cn.tobytes()).digest(), np.uint8)
    out["canon_order"] = ct["id"].astype(np.uint32)
    nodes, tbvh, ttopo = [], [], []
    for k, v in enumerate(frame_verts):
        t = O.build_triangles(v, faces, masks, canon)
        n, tb = R.bvh_build(t)
        nodes.append(n); tbvh.append(tb); ttopo.append(t)
        out["frame%d_nodes_sha" % k] = np.frombuffer(hashlib.sha256(n.tobytes()).digest(), np.uint8)
        out["frame%d_order" % k] = tb["id"].astype(np.uint32)
        out["frame%d_n_nodes" % k] = np.int32(len(n))
    # closest triangle on the canonical mesh (BASELINE config 1 in miniature)
    idx, dist = R.closest_triangle(rec[:, :3].copy(), cn, ct)
    out["canon_closest_idx"] = idx
    out["canon_closest_dist"] = dist
    # warp, both interpolation modes
    unit = np.array([0, 0, 0, 1, 1, 1], np.float32)
    for name, di in (("interp", False), ("nointerp", True)):
        r = rec.copy()
        tri, bvh_idx = R.warp(r, 27, -1, nodes, tbvh, ttopo, canon, adj, unit, disable_interp=di)
        out["warp_%s_tri" % name] = tri
        out["warp_%s_rec" % name] = r[:, [0, 1, 2, 8, 9, 10, 11, 26]].copy()
    # render_deformed: fixed frame 1
    r = rec.copy()
    tri, _ = R.warp(r, 27, 1, nodes, tbvh, ttopo, canon, adj, unit)
    out["render_tri"] = tri
    out["render_rec"] = r[:, [0, 1, 2, 8, 9, 10]].copy()
    # xyz path with explicit frames
    xyz = rec[:, :3].copy()
    fids = (np.arange(N_POINTS) % len(FRAMES)).astype(np.int32)
    tri, _ = R.warp(xyz, 3, -1, nodes, tbvh, ttopo, canon, adj, unit, frame_ids=fids)
    out["points_tri"] = tri
    out["points_xyz"] = xyz.reshape(-1, 3)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stage1_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
