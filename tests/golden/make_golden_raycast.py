"""Generates tests/golden/raycast_ref.npz by running the REFERENCE's own host-compiled ray cast (oracle/_ref/libref_host.so: ray_intersect of
instant-ngp/include/neural-graphics-primitives/bvh.cuh:98-144 with Triangle::ray_intersect / BoundingBox::ray_intersect) on the seeded rays of
tests/raycast_common.py.  Run in the build container only; the tests regenerate the rays from the same seeds, only the reference's OUTPUTS are stored."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from common import Scene  # noqa: E402
from oracle.pyoracle import Oracle, RefHost  # noqa: E402
from raycast_common import make_rays  # noqa: E402


def main():
    O, R = Oracle(), RefHost()
    scene = Scene(O, frames=(0, 7))
    out = {}
    for frame in (0, 1):
        ro, rd = make_rays(scene, frame, n_random=4000)
        nodes, tris = R.bvh_build(scene.ttopo[frame])  # the reference's own tree
        idx, t = R.ray_intersect(ro, rd, nodes, tris)
        out["tri_id_%d" % frame] = np.where(idx >= 0, tris["id"][np.maximum(idx, 0)].astype(np.int64), -1).astype(np.int32)
        out["t_%d" % frame] = t
    np.savez_compressed(os.path.join(HERE, "raycast_ref.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
