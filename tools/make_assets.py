"""Derive the compact synthetic-input base from the reference's mesh assets (data, not source):
   mesh_masks/canonical.obj (5023 v / 8090 f) and the red-vertex region masks mesh_masks/<region>.obj.
Run in the build container only (needs /root/reference); output: insta_b200/assets/flame_canonical.npz.
"""
import os
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "insta_b200", "assets", "flame_canonical.npz")

# FlameMasks enum values (include/rta/masking.h:32-50) of the regions mesh_to_mask() loads (masking.cpp:150-157)
REGIONS = {"boundary": 0, "eye_region": 1, "face": 2, "left_ear": 4, "lips": 7, "right_ear": 10}


def parse_obj(path):
    v, c, f = [], [], []
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == "v":
            v.append([float(x) for x in t[1:4]])
            if len(t) >= 7:
                c.append([float(x) for x in t[4:7]])
        elif t[0] == "f":
            f.append([int(x.split("/")[0]) - 1 for x in t[1:4]])
    return np.asarray(v, np.float32), np.asarray(c, np.float32), np.asarray(f, np.uint32)


verts, _, faces = parse_obj(os.path.join(REF, "mesh_masks", "canonical.obj"))
out = {"verts": verts, "faces": faces}
for name, rid in REGIONS.items():
    _, col, _ = parse_obj(os.path.join(REF, "mesh_masks", name + ".obj"))
    red = np.nonzero((col[:, 0] >= 1.0) & (col[:, 1] == 0.0) & (col[:, 2] == 0.0))[0].astype(np.uint32)  # masking.cpp:172-176
    out["region_%d" % rid] = red
    print(name, rid, len(red))
np.savez_compressed(OUT, **out)
print(verts.shape, faces.shape, os.path.getsize(OUT))
