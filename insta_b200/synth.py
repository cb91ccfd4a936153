"""Synthetic inputs for the INSTA hot path (the avatar datasets are gated): the canonical FLAME template shipped in
mesh_masks/canonical.obj (compacted to assets/flame_canonical.npz by tools/make_assets.py), randomly perturbed
per-frame vertex sets, sample records, cameras and grid weights.  Definitions follow SURVEY.md 8(d).

Everything here is deterministic numpy; nothing touches the GPU.
"""
import os

import numpy as np

_ASSET = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "flame_canonical.npz")

N_EXTRA_DIMS = 16               # exp_params, configs/nerf/main.json:17
RECORD_FLOATS = 11 + N_EXTRA_DIMS  # NerfCoordinate (nerf.h:86-115) + extra dims = 27 floats = 108 B


def load_canonical():
    """Returns (verts float32 [5023,3] already offset by +0.5 as TinyMesh::load does (tiny_mesh.cu:63-74),
    faces uint32 [8090,3], regions {FlameMasks id: red vertex ids})."""
    z = np.load(_ASSET)
    verts = (z["verts"].astype(np.float32) * np.float32(1.0) + np.float32(0.5)).astype(np.float32)
    regions = {int(k.split("_")[1]): z[k] for k in z.files if k.startswith("region_")}
    return verts, z["faces"].astype(np.uint32), regions


def mask_ids_from_regions(faces, regions):
    """Per-triangle FlameMasks id, masking.cpp:118-148: start at Torso (15); regions in ascending id order; a face takes
    the id of every region holding one of its vertices, later regions overwriting earlier ones."""
    out = np.full(len(faces), 15, np.uint32)
    for rid in sorted(regions):
        sel = np.zeros(int(faces.max()) + 1, bool)
        sel[regions[rid]] = True
        out[sel[faces].any(axis=1)] = rid
    return out


def _vertex_smoother(faces, n_verts):
    import scipy.sparse as sp

    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]]).astype(np.int64)
    e = np.concatenate([e, e[:, ::-1]])
    A = sp.coo_matrix((np.ones(len(e), np.float64), (e[:, 0], e[:, 1])), shape=(n_verts, n_verts)).tocsr()
    A.data[:] = 1.0
    A = A + sp.identity(n_verts)
    d = np.asarray(A.sum(axis=1)).ravel()
    return sp.diags(1.0 / d) @ A


class FrameSynth:
    """v_f = R_f (v - c) + c + smooth(0.003 * N(0,1)),  R_f a random-axis rotation of <= 5 degrees, seed = frame id."""

    def __init__(self, verts, faces, sigma=0.003, max_deg=5.0, smooth_iters=3):
        self.verts = verts.astype(np.float32)
        self.S = _vertex_smoother(faces, len(verts))
        self.sigma, self.max_deg, self.iters = sigma, max_deg, smooth_iters
        self.center = self.verts.mean(axis=0, dtype=np.float64)

    def frame(self, f):
        rng = np.random.default_rng(100003 + int(f))
        noise = rng.standard_normal(self.verts.shape) * self.sigma
        for _ in range(self.iters):
            noise = self.S @ noise
        axis = rng.standard_normal(3)
        axis /= np.linalg.norm(axis)
        ang = np.deg2rad(rng.uniform(-self.max_deg, self.max_deg))
        K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
        R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
        v = (self.verts.astype(np.float64) - self.center) @ R.T + self.center + noise
        return v.astype(np.float32)


def mesh_aabb(verts):
    """TinyMesh render AABB: vertex bounds inflated by 0.15 * |diag| (tiny_mesh.cu:157-162); fp32."""
    mn = verts.min(axis=0).astype(np.float32)
    mx = verts.max(axis=0).astype(np.float32)
    d = (mx - mn).astype(np.float32)
    amount = np.float32(np.sqrt(np.float32(d[0] * d[0]) + (np.float32(d[1] * d[1]) + np.float32(d[2] * d[2])))) * np.float32(0.15)
    return np.concatenate([mn - amount, mx + amount]).astype(np.float32)


def make_records(n, aabb6, frame_ids, seed=7, pad_fraction=0.0):
    """n sample records of 27 floats: pos ~ U(aabb), dt, holder = frame id, deformed_pos = pos, dir = (d+1)/2 with d uniform
    on the sphere, 16 expression floats ~ N(0,1).  pad_fraction of the records get holder = -1 (padding, skipped)."""
    rng = np.random.default_rng(seed)
    rec = np.empty((n, RECORD_FLOATS), np.float32)
    lo, hi = aabb6[:3].astype(np.float64), aabb6[3:].astype(np.float64)
    pos = (lo + (hi - lo) * rng.random((n, 3))).astype(np.float32)
    rec[:, 0:3] = pos
    rec[:, 3] = np.float32(0.0)
    rec[:, 4] = np.asarray(frame_ids, np.float32)
    rec[:, 5:8] = pos
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rec[:, 8:11] = ((d + 1.0) * 0.5).astype(np.float32)
    rec[:, 11:] = rng.standard_normal((n, N_EXTRA_DIMS)).astype(np.float32)
    if pad_fraction > 0:
        rec[rng.random(n) < pad_fraction, 4] = -1.0
    return rec


def train_regime_frame_ids(n, n_frames):
    """Runs of consecutive samples per frame (2^18 rays over n_frames images x ~16 samples/ray)."""
    run = max(1, -(-n // n_frames))
    return (np.arange(n) // run).astype(np.int32) % n_frames


def make_grid_table(n_entries, seed=1337, scale=1.0):
    """fp16 table [n_entries, 8] ~ U(-scale, scale).  tcnn initialises U(-1e-4, 1e-4) (grid.h:1275-1280); the default
    here is the U(+-1) variant so encodings are non-trivial."""
    rng = np.random.default_rng(seed)
    return ((rng.random((n_entries, 8), dtype=np.float32) * 2 - 1) * np.float32(scale)).astype(np.float16)


def make_camera_matrix(yaw_deg, pitch_deg, distance=1.0, target=(0.5, 0.5, 0.5)):
    """Camera-to-world 3x4 (columns right, down, forward, origin) of a camera at `distance` from `target` looking at it
    (yaw in U(+-30), pitch in U(+-10) degrees)."""
    yaw, pitch = np.deg2rad(yaw_deg), np.deg2rad(pitch_deg)
    fwd = -np.array([np.sin(yaw) * np.cos(pitch), np.sin(pitch), np.cos(yaw) * np.cos(pitch)])
    origin = np.asarray(target, np.float64) - fwd * distance
    up = np.array([0.0, 1.0, 0.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    return np.stack([right, down, fwd, origin], axis=1).astype(np.float32)


def random_camera(seed, distance=1.0):
    rng = np.random.default_rng(77000 + int(seed))
    return make_camera_matrix(rng.uniform(-30, 30), rng.uniform(-10, 10), distance)


def morton3d(x, y, z):
    def expand(v):
        v = v.astype(np.uint64)
        v = (v * 0x00010001) & 0xFF0000FF
        v = (v * 0x00000101) & 0x0F00F00F
        v = (v * 0x00000011) & 0xC30C30C3
        v = (v * 0x00000005) & 0x49249249
        return v

    return (expand(x) | (expand(y) << 1) | (expand(z) << 2)).astype(np.uint32)


def occupancy_bitfield(vertex_sets, radius=0.03, grid=128):
    """Occupancy bitfield of cascade 0 (128^3 bits, Morton order, ngp testbed_nerf.cu:313-337): a cell is occupied when its
    centre is within `radius` of a vertex of any of the given meshes (the tracked meshes are dense, ~5 mm edges, so this
    is the union-of-frames shell of SURVEY.md 8d config 3).  Returns uint8 [grid^3 / 8]."""
    from scipy.spatial import cKDTree

    c = (np.arange(grid, dtype=np.float32) + 0.5) / grid
    occ = np.zeros(grid ** 3, bool)
    X, Y, Z = np.meshgrid(np.arange(grid), np.arange(grid), np.arange(grid), indexing="ij")
    X, Y, Z = X.ravel(), Y.ravel(), Z.ravel()
    centres = np.stack([c[X], c[Y], c[Z]], axis=1)
    idx = morton3d(X, Y, Z)
    for v in vertex_sets:
        lo, hi = v.min(axis=0) - radius, v.max(axis=0) + radius
        sel = np.all((centres >= lo) & (centres <= hi), axis=1)
        d, _ = cKDTree(v).query(centres[sel], k=1, distance_upper_bound=radius)
        hit = np.zeros(len(centres), bool)
        hit[np.flatnonzero(sel)[np.isfinite(d)]] = True
        occ[idx[hit]] = True
    return np.packbits(occ, bitorder="little")


MLP_SHAPES = [(64, 144), (64, 64), (16, 64), (64, 32), (64, 64), (64, 64), (16, 64)]  # [out, in] per layer: density MLP, colour MLP
N_MLP_PARAMS = sum(o * i for o, i in MLP_SHAPES)  # 25,600


def make_network_params(n_params, seed=5, grid_scale=1.0):
    """fp16 parameter vector in the trainer's order (nerf_network.h:492-525): density MLP (64x144, 64x64, 16x64), colour MLP
    (64x32, 64x64, 64x64, 16x64), weights row-major [out][in] Xavier-uniform like tcnn (gpu_matrix.h:292-304), then the
    grid table ~ U(+-grid_scale) (tcnn initialises U(+-1e-4); the default makes encodings non-trivial)."""
    rng = np.random.default_rng(seed)
    parts = []
    for o, i in MLP_SHAPES:
        s = np.sqrt(6.0 / (i + o))
        parts.append(((rng.random((o, i)) * 2 - 1) * s).astype(np.float16).ravel())
    n_grid = n_params - N_MLP_PARAMS
    parts.append(((rng.random(n_grid, dtype=np.float32) * 2 - 1) * np.float32(grid_scale)).astype(np.float16))
    return np.concatenate(parts)


def make_training_images(n_images, res, focal, seed=3, with_rays=True, mask_fraction=0.05):
    """Synthetic training set: random smooth RGBA float images with alpha = 1 inside a disc (the projected
    head), a few masked-away pixels (negative red), a segmentation mask (region ids 0 / 120 / 200 in the red byte), a depth map,
    random cameras looking at the unit-cube centre and, optionally, the explicit per-pixel rays of that camera (origin, unnormalised
    direction) computed in fp32 numpy.  Returns a list of dicts of numpy arrays (capi.Context.training_set_images takes torch tensors)."""
    rng = np.random.default_rng(seed)
    W, H = res
    out = []
    yy, xx = np.meshgrid((np.arange(H) + 0.5) / H, (np.arange(W) + 0.5) / W, indexing="ij")
    for k in range(n_images):
        cam = random_camera(5000 + seed * 100 + k)
        f = rng.random(3) * 6 + 1
        rgb = np.stack([0.5 + 0.5 * np.sin(f[c] * xx + 3 * c + k) * np.cos(f[(c + 1) % 3] * yy) for c in range(3)], axis=-1)
        disc = ((xx - 0.5) ** 2 + (yy - 0.5) ** 2) < 0.16
        alpha = disc.astype(np.float32)
        pix = np.concatenate([rgb * alpha[..., None], alpha[..., None]], axis=-1).astype(np.float32)
        masked = rng.random((H, W)) < mask_fraction
        pix[masked] = -1.0
        seg = np.zeros((H, W, 4), np.uint8)
        seg[..., 0] = np.where(disc, np.where(xx < 0.5, 120, 200), 0)
        depth = (0.8 + 0.2 * np.sin(5 * xx) * np.cos(4 * yy)).astype(np.float32)
        im = {"pixels": pix, "seg": seg, "depth": depth, "res": (W, H), "focal": (float(focal), float(focal)), "pp": (0.5, 0.5), "xform": cam, "type": 3}
        if with_rays:
            lx = ((xx.astype(np.float32) - np.float32(0.5)) * np.float32(W) / np.float32(focal)).astype(np.float32)
            ly = ((yy.astype(np.float32) - np.float32(0.5)) * np.float32(H) / np.float32(focal)).astype(np.float32)
            d = (cam[:, 0][None, None] * lx[..., None] + cam[:, 1][None, None] * ly[..., None] + cam[:, 2][None, None]).astype(np.float32)
            o = np.broadcast_to(cam[:, 3].astype(np.float32), d.shape)
            im["rays"] = np.concatenate([o, d], axis=-1).astype(np.float32).copy()
        out.append(im)
    return out
