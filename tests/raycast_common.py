"""Rays for the mesh ray-cast tests (rta::Core::raycast_flame): random ones through the mesh, the tie-heavy ones (through vertices, edge midpoints and
centroids, where two or more triangles are hit at the same parameter), axis-parallel ones (zero direction components: infinities in the slab test),
rays starting inside the mesh, rays that miss."""
import numpy as np


def make_rays(scene, frame, n_random=20000, seed=5):
    rng = np.random.default_rng(seed)
    v = scene.frame_verts[frame]
    lo, hi = scene.aabb[:3], scene.aabb[3:]
    centre = ((lo + hi) * np.float32(0.5)).astype(np.float32)
    # random: from a shell around the box towards a random point of the box
    o = (centre + (rng.standard_normal((n_random, 3)) * 1.5)).astype(np.float32)
    tgt = (lo + (hi - lo) * rng.random((n_random, 3))).astype(np.float32)
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    # through vertices / edge midpoints / centroids from a common eye point (not normalised: the query takes directions as given)
    f = scene.faces[:1500]
    special = np.concatenate([v[:2000], (v[f[:, 0]] + v[f[:, 1]]) * np.float32(0.5), (v[f[:, 0]] + v[f[:, 1]] + v[f[:, 2]]) / np.float32(3)]).astype(np.float32)
    eye = (centre + np.array([0.1, 0.05, 1.7], np.float32)).astype(np.float32)
    o2 = np.broadcast_to(eye, special.shape).copy()
    d2 = (special - eye).astype(np.float32)
    # axis-parallel rays over a lattice in front of the mesh, and the same from inside the mesh (origin at the centre)
    g = np.linspace(0.05, 0.95, 40, dtype=np.float32)
    gx, gy = np.meshgrid(lo[0] + (hi[0] - lo[0]) * g, lo[1] + (hi[1] - lo[1]) * g)
    o3 = np.stack([gx.ravel(), gy.ravel(), np.full(gx.size, hi[2] + 0.5, np.float32)], 1).astype(np.float32)
    d3 = np.broadcast_to(np.array([0, 0, -1], np.float32), o3.shape).copy()
    o4 = np.broadcast_to(centre, (600, 3)).copy()
    d4 = rng.standard_normal((600, 3)).astype(np.float32)
    d4[:200, 0] = 0.0  # zero components from inside, too
    d4[200:400, 1] = 0.0
    # misses: pointing away from the box
    o5 = o[:500].copy()
    d5 = -d[:500]
    return np.concatenate([o, o2, o3, o4, o5]).astype(np.float32), np.concatenate([d, d2, d3, d4, d5]).astype(np.float32)
