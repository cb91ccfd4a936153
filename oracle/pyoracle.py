"""TEST INFRASTRUCTURE ONLY: ctypes loaders for the CPU oracle (oracle/liboracle.so) and, when it was built,
the reference's own host code (oracle/_ref/libref_host.so).  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs only -- never by the insta_b200 package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_HOST_SO = os.path.join(HERE, "_ref", "libref_host.so")
REF_NERF_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_nerf.so")
REF_CUDA_SO = os.path.join(HERE, "_ref", "libref_cuda.so")

# ngp::Triangle / ngp::TriangleBvhNode byte layouts (triangle.cuh:234-238, triangle_bvh.cuh:29-33)
TRI_DTYPE = np.dtype([("a", "<f4", 3), ("b", "<f4", 3), ("c", "<f4", 3), ("scaling", "<f4"), ("id", "<u4"), ("mask_id", "<u4")])
NODE_DTYPE = np.dtype([("mn", "<f4", 3), ("mx", "<f4", 3), ("left_idx", "<i4"), ("right_idx", "<i4")])
assert TRI_DTYPE.itemsize == 48 and NODE_DTYPE.itemsize == 32


def build_oracle(force=False):
    src = [os.path.join(HERE, "oracle.cpp"), os.path.join(HERE, "oracle.h")]
    if force or not os.path.exists(ORACLE_SO) or any(os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in src):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _ptr_array(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


class _Lib:
    """Common wrapper: the oracle (prefix orc_) and the reference host harness (prefix ref_) share signatures."""

    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.path = path
        f = getattr(self.lib, prefix + "max_threads")
        f.restype = C.c_int
        self.max_threads = f()

    def _fn(self, name, restype=None):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        return f

    def bvh_build(self, tris, leaf_size=2):
        """Returns (nodes, tris_in_bvh_order); `tris` is not modified."""
        t = np.ascontiguousarray(tris.copy())
        max_nodes = 4 * len(t) + 16
        nodes = np.zeros(max_nodes, NODE_DTYPE)
        n = self._fn("bvh_build", C.c_int)(_p(t), C.c_int(len(t)), C.c_int(leaf_size), _p(nodes), C.c_int(max_nodes))
        assert n > 0, n
        return nodes[:n].copy(), t

    def closest_triangle(self, pts, nodes, tris, n_threads=None):
        pts = np.ascontiguousarray(pts, np.float32)
        n = len(pts)
        idx = np.empty(n, np.int32)
        dist = np.empty(n, np.float32)
        self._fn("closest_triangle")(_p(pts), C.c_int(n), _p(nodes), _p(tris), _p(idx), _p(dist), C.c_int(n_threads or self.max_threads))
        return idx, dist

    def ray_intersect(self, origins, dirs, nodes, tris):
        """TriangleBvh4::ray_intersect per ray: (BVH-order index of the first hit or -1, ray parameter; 10 = MAX_DIST when nothing is hit)."""
        ro = np.ascontiguousarray(origins, np.float32)
        rd = np.ascontiguousarray(dirs, np.float32)
        idx = np.empty(len(ro), np.int32)
        t = np.empty(len(ro), np.float32)
        self._fn("ray_intersect")(_p(ro), _p(rd), C.c_int(len(ro)), _p(nodes), _p(tris), _p(idx), _p(t))
        return idx, t

    def closest_triangle_bounded(self, pts, nodes, tris, init_best, n_threads=None):
        """closest_triangle with max_distance_sq = init_best[i] (bvh.cuh:48) instead of FLT_MAX; oracle only."""
        pts = np.ascontiguousarray(pts, np.float32)
        ib = np.ascontiguousarray(init_best, np.float32)
        idx = np.empty(len(pts), np.int32)
        self._fn("closest_triangle_bounded")(_p(pts), C.c_int(len(pts)), _p(nodes), _p(tris), _p(ib), _p(idx), C.c_int(n_threads or self.max_threads))
        return idx

    def cyl_build(self, nodes, tris):
        """The product's cylinder bounds (deform.cu build_cyl_kernel) restated: uint16[n_nodes, 8] fp16 bits."""
        out = np.zeros((len(nodes), 8), np.uint16)
        self._fn("cyl_build")(_p(nodes), C.c_int(len(nodes)), _p(tris), _p(out))
        return out

    def closest_triangle_cyl(self, pts, nodes, tris, cyl, init_best=None, check=False):
        """Reference traversal + the product's cylinder cull; returns (idx, stats4 = expansions, leaf visits, culled children, violations)."""
        pts = np.ascontiguousarray(pts, np.float32)
        ib = np.ascontiguousarray(init_best, np.float32) if init_best is not None else None
        idx = np.empty(len(pts), np.int32)
        st = np.zeros(4, np.float64)
        self._fn("closest_triangle_cyl")(_p(pts), C.c_int(len(pts)), _p(nodes), _p(tris), _p(cyl), _p(ib), _p(idx), _p(st), C.c_int(int(check)))
        return idx, st

    def warp(self, samples, stride, frame_id, nodes_list, tris_bvh_list, tris_topo_list, tris_canon, adjacency, aabb6,
             disable_interp=False, extra_dims=16, use_geo=False, frame_ids=None, n_threads=None):
        """In-place on `samples` (float32, n*stride). Returns (topology tri id, bvh-order idx) per sample."""
        assert samples.dtype == np.float32 and samples.flags.c_contiguous
        n = samples.size // stride
        out_tri = np.empty(n, np.int32)
        out_bvh = np.empty(n, np.int32)
        adjacency = np.ascontiguousarray(adjacency, np.int32)
        aabb6 = np.ascontiguousarray(aabb6, np.float32)
        fids = np.ascontiguousarray(frame_ids, np.int32) if frame_ids is not None else None
        self._fn("warp")(_p(samples), C.c_uint32(n), C.c_uint32(stride), C.c_int32(frame_id), C.c_uint32(len(nodes_list)),
                         _ptr_array(nodes_list), _ptr_array(tris_bvh_list), _ptr_array(tris_topo_list), _p(tris_canon), _p(adjacency),
                         C.c_int(int(disable_interp)), _p(aabb6), C.c_uint32(extra_dims), C.c_int(int(use_geo)), _p(fids),
                         _p(out_tri), _p(out_bvh), C.c_int(n_threads or self.max_threads))
        return out_tri, out_bvh


class Oracle(_Lib):
    def __init__(self):
        super().__init__(build_oracle(), "orc_")
        L = self.lib
        L.orc_triangle_distance_sq.restype = C.c_float
        L.orc_bbox_distance_sq.restype = C.c_float
        L.orc_f32_to_f16.restype = C.c_uint16
        L.orc_f32_to_f16.argtypes = [C.c_float]
        L.orc_f16_to_f32.restype = C.c_float
        L.orc_f16_to_f32.argtypes = [C.c_uint16]

    def set_arith(self, mode):
        """'host' = arithmetic of the reference's host build (default), 'device' = of its CUDA build (FMA-contracted)."""
        assert mode in ("host", "device")
        self.lib.orc_set_arith(C.c_int(1 if mode == "device" else 0))

    # the warp entry point is named orc_warp in the oracle, ref_warp_cpu in the reference harness
    def build_triangles(self, verts, faces, mask_ids, canonical=None):
        verts = np.ascontiguousarray(verts, np.float32)
        faces = np.ascontiguousarray(faces, np.uint32)
        mask_ids = np.ascontiguousarray(mask_ids, np.uint32)
        out = np.zeros(len(faces), TRI_DTYPE)
        self.lib.orc_build_triangles(_p(verts), _p(faces), C.c_uint32(len(faces)), _p(mask_ids), _p(canonical), _p(out))
        return out

    def mesh_aabb(self, verts):
        verts = np.ascontiguousarray(verts, np.float32)
        out = np.empty(6, np.float32)
        self.lib.orc_mesh_aabb(_p(verts), C.c_uint32(len(verts)), _p(out))
        return out

    def adjacency(self, faces, n_verts):
        faces = np.ascontiguousarray(faces, np.uint32)
        out = np.empty((len(faces), 3), np.int32)
        self.lib.orc_adjacency(_p(faces), C.c_uint32(len(faces)), C.c_uint32(n_verts), _p(out))
        return out

    def triangle_masks(self, faces, regions):
        """regions: dict {FlameMasks id: vertex index array}; applied in ascending id order (std::map iteration)."""
        faces = np.ascontiguousarray(faces, np.uint32)
        ids = sorted(regions)
        offs = np.zeros(len(ids) + 1, np.uint32)
        offs[1:] = np.cumsum([len(regions[k]) for k in ids])
        vs = np.ascontiguousarray(np.concatenate([np.asarray(regions[k], np.uint32) for k in ids]))
        rid = np.asarray(ids, np.int32)
        out = np.empty(len(faces), np.uint32)
        self.lib.orc_triangle_masks(_p(faces), C.c_uint32(len(faces)), _p(rid), _p(offs), _p(vs), C.c_uint32(len(ids)), _p(out))
        return out

    def triangle_distance_sq(self, tri, p):
        t = np.asarray(tri, TRI_DTYPE).reshape(1)
        p = np.ascontiguousarray(p, np.float32)
        return np.float32(self.lib.orc_triangle_distance_sq(_p(t), _p(p)))

    def bbox_distance_sq(self, mn, mx, p):
        mn, mx, p = (np.ascontiguousarray(v, np.float32) for v in (mn, mx, p))
        return np.float32(self.lib.orc_bbox_distance_sq(_p(mn), _p(mx), _p(p)))

    def tri2projection(self, tri):
        t = np.asarray(tri, TRI_DTYPE).reshape(1)
        out = np.empty(16, np.float32)
        self.lib.orc_tri2projection(_p(t), _p(out))
        return out.reshape(4, 4).T  # column-major -> [row, col]

    def random_frame_ids(self, n, n_meshes, seed=42):
        # pcg32(seed): state=0; inc=(1<<1)|1; next; state+=seed; next   (pcg32.h seed())
        mult = 0x5851F42D4C957F2D
        mask = (1 << 64) - 1
        inc = 3
        state = (0 * mult + inc) & mask
        state = (state + seed) & mask
        state = (state * mult + inc) & mask
        out = np.empty(n, np.int32)
        self.lib.orc_random_frame_ids(C.c_uint32(n), C.c_uint32(n_meshes), C.c_uint64(state), C.c_uint64(inc), _p(out))
        return out

    # ---- hash grid ----
    def grid_layout(self, n_levels=16, log2_hashmap_size=18, base_resolution=16, per_level_scale=None, aabb_scale=1):
        if per_level_scale is None:
            # testbed.cu:2246-2258: exp(log(desired_resolution * aabb_scale / base) / (n_levels - 1)), desired 2048, fp32
            per_level_scale = np.exp(np.log(np.float32(2048.0 * aabb_scale) / np.float32(base_resolution)) / np.float32(n_levels - 1)).astype(np.float32)
        offs = np.empty(n_levels + 1, np.uint32)
        scales = np.empty(n_levels, np.float32)
        res = np.empty(n_levels, np.uint32)
        self.lib.orc_grid_layout(C.c_uint32(n_levels), C.c_uint32(log2_hashmap_size), C.c_uint32(base_resolution),
                                 C.c_float(float(per_level_scale)), _p(offs), _p(scales), _p(res))
        return offs, scales, res

    def grid_slots(self, pos, pos_stride, layout, want_weights=False):
        offs, scales, res = layout
        pos = np.ascontiguousarray(pos, np.float32)
        n = pos.size // pos_stride
        L = len(scales)
        slots = np.empty((n, L, 8), np.uint32)
        w = np.empty((n, L, 8), np.float32) if want_weights else None
        self.lib.orc_grid_slots(C.c_uint32(n), _p(pos), C.c_uint32(pos_stride), C.c_uint32(L), _p(offs), _p(scales), _p(res), _p(slots), _p(w))
        return (slots, w) if want_weights else slots

    def grid_forward(self, pos, pos_stride, layout, table_f16, n_threads=None):
        offs, scales, res = layout
        pos = np.ascontiguousarray(pos, np.float32)
        n = pos.size // pos_stride
        L = len(scales)
        table = np.ascontiguousarray(table_f16).view(np.uint16)
        out = np.empty((L * 8, n), np.uint16)
        self.lib.orc_grid_forward(C.c_uint32(n), _p(pos), C.c_uint32(pos_stride), C.c_uint32(L), _p(offs), _p(scales), _p(res), _p(table),
                                  _p(out), C.c_int(n_threads or self.max_threads))
        return out.view(np.float16)

    def grid_backward(self, pos, pos_stride, layout, dLdy_soa_f16, want_hits=False):
        offs, scales, res = layout
        pos = np.ascontiguousarray(pos, np.float32)
        n = pos.size // pos_stride
        L = len(scales)
        dy = np.ascontiguousarray(dLdy_soa_f16).view(np.uint16)
        grad = np.zeros((int(offs[-1]), 8), np.float64)
        hits = np.zeros(int(offs[-1]), np.uint32) if want_hits else None
        self.lib.orc_grid_backward(C.c_uint32(n), _p(pos), C.c_uint32(pos_stride), C.c_uint32(L), _p(offs), _p(scales), _p(res), _p(dy), _p(grad), _p(hits))
        return (grad, hits) if want_hits else grad


class RefHost(_Lib):
    """The reference's own host code.  Present only where oracle/Makefile `ref` was run (needs /root/reference)."""

    def __init__(self):
        super().__init__(REF_HOST_SO, "ref_")
        assert self._fn("sizeof_triangle", C.c_int)() == 48 and self._fn("sizeof_node", C.c_int)() == 32

    def _fn(self, name, restype=None):
        if name == "warp":
            name = "warp_cpu"
        return super()._fn(name, restype)

    def triangle_distance_sq(self, tri, pts):
        t = np.asarray(tri, TRI_DTYPE).reshape(1)
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.empty(len(pts), np.float32)
        self.lib.ref_triangle_distance_sq(_p(pts), C.c_int(len(pts)), _p(t), _p(out))
        return out

    def bbox_distance_sq(self, mn, mx, pts):
        mn, mx = (np.ascontiguousarray(v, np.float32) for v in (mn, mx))
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.empty(len(pts), np.float32)
        self.lib.ref_bbox_distance_sq(_p(pts), C.c_int(len(pts)), _p(mn), _p(mx), _p(out))
        return out

    def tri2projection(self, tri):
        t = np.asarray(tri, TRI_DTYPE).reshape(1)
        out = np.empty(16, np.float32)
        self.lib.ref_tri2projection(_p(t), _p(out))
        return out.reshape(4, 4).T


def have_ref_host():
    return os.path.exists(REF_HOST_SO)


def have_ref_cuda():
    return os.path.exists(REF_CUDA_SO)


# ---- recorder read-back (numpy restatement of rta::Recorder save_rgba / save_depth, src/recorder.cu:36-60, 75-105) ----
def recorder_rgba8(rgba, srgb_to_linear=False):
    """rgba float32 [n, 4] -> uint8 [n, 4]: (uint8_t) tcnn::clamp(transform(v) * 255.f, 0.f, 255.f) per channel incl. alpha (recorder.cu:46-49);
    transform = identity, or srgb_to_linear (recorder.cu:116-122) in Normals mode (341-343)."""
    v = np.asarray(rgba, np.float32)
    if srgb_to_linear:
        v = np.where(v <= np.float32(0.04045), v / np.float32(12.92), np.power((v + np.float32(0.055)) / np.float32(1.055), np.float32(2.4), dtype=np.float32)).astype(np.float32)
    s = (v * np.float32(255.0)).astype(np.float32)
    s = np.where(s < 0, np.float32(0), np.where(s > 255, np.float32(255), s))
    return np.nan_to_num(s, nan=0.0).astype(np.uint8)  # truncation, like the C cast


def recorder_depth_mm(depth, w, h):
    """depth float32 [w*h] -> uint32 [(h-1)*w]: output pixel (x, y) reads index x + w + y*w and is skipped when that is past the buffer
    (recorder.cu:93-94); d outside [0, 5] -> 0; uint32(d * 1000.0f) (recorder.cu:96-99)."""
    d = np.asarray(depth, np.float32).reshape(-1)
    out = []
    for y in range(h):
        i = np.arange(w) + w + y * w
        i = i[i < d.size]
        v = d[i].copy()
        v[(v < 0) | (v > 5) | np.isnan(v)] = 0
        out.append((v * np.float32(1000.0)).astype(np.float32).astype(np.uint32))
    return np.concatenate(out) if out else np.zeros(0, np.uint32)
