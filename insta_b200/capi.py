"""ctypes binding of libinsta_b200.so (include/insta_b200.h) for the test-suite and bench.py.

PyTorch is used only as the owner of device buffers and streams (`tensor.data_ptr()`, `current_stream()`); every
computation goes through the C ABI.  There is no fallback: if the shared library is missing or no sm_100 GPU is
present, loading / context creation raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

LIB_PATH = _build.LIB


class InstaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("insta_b200 error %d: %s" % (code, msg))
        self.code = code


class Params(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("disable_interp", C.c_int32),
        ("max_cached_bvh", C.c_uint32),
        ("exp_params", C.c_uint32),
        ("geo_params", C.c_uint32),
        ("n_primitives_per_leaf", C.c_uint32),
        ("aabb_min", C.c_float * 3),
        ("aabb_max", C.c_float * 3),
        ("n_levels", C.c_uint32),
        ("n_features_per_level", C.c_uint32),
        ("log2_hashmap_size", C.c_uint32),
        ("base_resolution", C.c_uint32),
        ("per_level_scale", C.c_float),
        ("aabb_scale", C.c_uint32),
        ("rng_seed", C.c_uint64),
        ("arith", C.c_uint32),
        ("reserved_", C.c_uint32),
    ]


class Camera(C.Structure):
    _fields_ = [("camera_matrix", C.c_float * 12), ("focal_length", C.c_float * 2), ("screen_center", C.c_float * 2), ("resolution", C.c_uint32 * 2),
                ("sample_index", C.c_uint32), ("snap_to_pixel_centers", C.c_uint32)]


class TrainImage(C.Structure):
    _fields_ = [("d_pixels", C.c_void_p), ("d_seg_mask", C.c_void_p), ("d_depth", C.c_void_p), ("d_rays", C.c_void_p), ("resolution", C.c_uint32 * 2),
                ("focal_length", C.c_float * 2), ("principal_point", C.c_float * 2), ("xform", C.c_float * 12), ("image_data_type", C.c_uint32), ("reserved_", C.c_uint32)]


class TrainStep(C.Structure):
    _fields_ = [("n_rays", C.c_uint32), ("n_rays_total", C.c_uint32), ("max_samples", C.c_uint32), ("target_batch_size", C.c_uint32), ("stride_floats", C.c_uint32),
                ("snap_to_pixel_centers", C.c_uint32), ("random_bg_color", C.c_uint32), ("penalize_density_alpha", C.c_uint32), ("rng_seed", C.c_uint64),
                ("rng_advances", C.c_uint64), ("cone_angle_constant", C.c_float), ("loss_scale", C.c_float), ("background_color", C.c_float * 3),
                ("depth_weight", C.c_float), ("seg_weight", C.c_float), ("near_distance", C.c_float), ("beta_loss", C.c_float),
                ("d_density_grid_bitfield", C.c_void_p), ("d_extra_dims", C.c_void_p), ("d_coords", C.c_void_p), ("d_counters", C.c_void_p),
                ("d_ray_indices", C.c_void_p), ("d_rays", C.c_void_p), ("d_numsteps", C.c_void_p)]


class Optimizer(C.Structure):
    _fields_ = [("learning_rate", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("epsilon", C.c_float), ("l2_reg", C.c_float), ("ema_decay", C.c_float),
                ("loss_scale", C.c_float), ("current_step", C.c_uint32), ("d_params_fp32", C.c_void_p), ("d_params_f16", C.c_void_p), ("d_gradients_f16", C.c_void_p),
                ("d_first_moments", C.c_void_p), ("d_second_moments", C.c_void_p), ("d_param_steps", C.c_void_p), ("d_params_ema_f16", C.c_void_p)]


class TrainerConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("target_batch_size", C.c_uint32), ("rays_per_batch", C.c_uint32), ("snap_to_pixel_centers", C.c_uint32),
                ("random_bg_color", C.c_uint32), ("penalize_density_alpha", C.c_uint32), ("rng_seed", C.c_uint64), ("cone_angle_constant", C.c_float),
                ("loss_scale", C.c_float), ("background_color", C.c_float * 3), ("depth_weight", C.c_float), ("seg_weight", C.c_float),
                ("near_distance", C.c_float), ("beta_loss", C.c_float), ("fixed_rays_per_batch", C.c_uint32), ("freeze_parameters", C.c_uint32)]


class TrajectoryFrame(C.Structure):
    _fields_ = [("camera_matrix", C.c_float * 12), ("frame_index", C.c_uint32), ("suffix", C.c_char_p)]


class TrainStats(C.Structure):
    _fields_ = [("training_step", C.c_uint32), ("n_rays", C.c_uint32), ("n_rays_accepted", C.c_uint32), ("rays_per_batch", C.c_uint32), ("n_rays_total", C.c_uint32),
                ("measured_batch_size", C.c_uint32), ("measured_batch_size_before_compaction", C.c_uint32), ("loss", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class DensityGrid(C.Structure):
    _fields_ = [("d_density_grid", C.c_void_p), ("d_density_grid_bitfield", C.c_void_p), ("d_density_grid_mean", C.c_void_p), ("n_cascades", C.c_uint32),
                ("ema_step", C.c_uint32), ("rng_seed", C.c_uint64), ("rng_advances", C.c_uint64), ("decay", C.c_float), ("training_step", C.c_uint32),
                ("has_rays", C.c_uint32), ("images_changed", C.c_uint32)]


class Snapshot(C.Structure):
    _fields_ = [("n_params", C.c_uint64), ("h_params_f16", C.c_void_p), ("n_density_grid", C.c_uint64), ("h_density_grid_f16", C.c_void_p), ("training_step", C.c_uint32),
                ("loss", C.c_float), ("aabb", C.c_float * 6), ("render_aabb", C.c_float * 6), ("bounding_radius", C.c_float), ("aabb_scale", C.c_uint32),
                ("rays_per_batch", C.c_uint32), ("measured_batch_size", C.c_uint32), ("measured_batch_size_before_compaction", C.c_uint32), ("ema_decay", C.c_float),
                ("beta1", C.c_float), ("beta2", C.c_float), ("epsilon", C.c_float), ("l2_reg", C.c_float), ("opt_base_learning_rate", C.c_float),
                ("has_optimizer", C.c_uint32), ("opt_current_step", C.c_uint32), ("h_first_moments", C.c_void_p), ("h_second_moments", C.c_void_p),
                ("h_param_steps", C.c_void_p), ("h_weights_ema_f16", C.c_void_p)]


class RenderStats(C.Structure):
    _fields_ = [("iterations", C.c_uint32), ("reserved_", C.c_uint32), ("samples", C.c_uint64)]


class RenderOpts(C.Structure):
    _fields_ = [("d_density_grid_bitfield", C.c_void_p), ("cone_angle_constant", C.c_float), ("min_transmittance", C.c_float), ("use_render_aabb", C.c_uint32),
                ("render_aabb", C.c_float * 6), ("linear_colors", C.c_uint32), ("max_steps_per_iteration", C.c_uint32)]


def make_camera(cam3x4, focal, resolution, screen_center=(0.5, 0.5), sample_index=0, snap=True):
    """cam3x4: numpy [3, 4] camera-to-world (columns right, down, forward, origin) in ngp space."""
    c = Camera()
    m = np.asarray(cam3x4, np.float32)
    for col in range(4):
        for row in range(3):
            c.camera_matrix[col * 3 + row] = float(m[row, col])
    c.focal_length[0], c.focal_length[1] = (focal, focal) if np.isscalar(focal) else focal
    c.screen_center[0], c.screen_center[1] = screen_center
    c.resolution[0], c.resolution[1] = resolution
    c.sample_index = sample_index
    c.snap_to_pixel_centers = int(snap)
    return c


# every symbol include/insta_b200.h declares (tests/test_abi.py checks the header against this list and the .so)
EXPORTS = [
    "insta_params_default", "insta_last_error", "insta_version", "insta_ctx_create", "insta_ctx_destroy",
    "insta_set_canonical", "insta_load_canonical_obj", "insta_add_frame", "insta_add_frames", "insta_add_frame_obj",
    "insta_clear_frames", "insta_n_frames", "insta_get_render_aabb", "insta_set_aabb", "insta_get_topology",
    "insta_get_frame_bvh_order", "insta_get_frame_cylinders", "insta_deform_samples", "insta_deform_samples_n", "insta_network_inference_n", "insta_deform_points", "insta_grid_layout", "insta_grid_n_params",
    "insta_hash_encode_fwd", "insta_hash_encode_bwd", "insta_hash_slots", "insta_deform_encode_host", "insta_deform_infer_host", "insta_host_alloc",
    "insta_host_free", "insta_launch_count", "insta_host_bvh_build", "insta_host_build_triangles", "insta_host_adjacency",
    "insta_host_load_canonical_obj", "insta_host_mesh_aabb", "insta_recorder_readback", "insta_host_write_png", "insta_host_write_depth_bin",
    "insta_network_n_params", "insta_network_inference", "insta_network_density", "insta_network_forward_backward", "insta_training_set_images", "insta_generate_training_samples", "insta_compute_loss",
    "insta_optimizer_step",
    "insta_trainer_config_default", "insta_trainer_create", "insta_trainer_destroy", "insta_trainer_state", "insta_train_nerf",
    "insta_update_density_grid", "insta_training_prep_nerf", "insta_update_density_grid_mean_and_bitfield",
    "insta_replace_frames", "insta_refit_frames", "insta_host_write_trajectory_json", "insta_shard_frame_count", "insta_shard_owner", "insta_shard_local_ids",
    "insta_microbench_gather", "insta_microbench_scatter", "insta_raycast_mesh", "insta_mesh_ray_intersect", "insta_snapshot_save", "insta_snapshot_open", "insta_snapshot_close",
    "insta_comm_unique_id", "insta_comm_init_rank", "insta_comm_destroy", "insta_comm_exchange_mode", "insta_grad_allreduce",
    "insta_render_frame", "insta_tracer_init_rays", "insta_tracer_n_steps", "insta_tracer_generate_samples", "insta_tracer_composite",
]

_lib = None


def load_library(build_if_missing=True):
    """dlopen the in-tree library.  Raises if it is absent and cannot be built -- never falls back to anything."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise FileNotFoundError(LIB_PATH + " is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _build.build()
    lib = C.CDLL(LIB_PATH)
    lib.insta_last_error.restype = C.c_char_p
    lib.insta_version.restype = C.c_char_p
    lib.insta_n_frames.restype = C.c_uint32
    lib.insta_grid_n_params.restype = C.c_uint32
    lib.insta_launch_count.restype = C.c_uint64
    lib.insta_tracer_n_steps.restype = C.c_uint32
    lib.insta_network_n_params.restype = C.c_uint32
    lib.insta_ctx_destroy.restype = None
    lib.insta_params_default.restype = None
    lib.insta_trainer_config_default.restype = None
    lib.insta_snapshot_close.restype = None
    for name in EXPORTS:
        getattr(lib, name)  # AttributeError if the .so lost a symbol
    _lib = lib
    return lib


def default_params():
    p = Params()
    load_library().insta_params_default(C.byref(p))
    return p


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype)


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _dev(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(stream=None):
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


class Context:
    """Host-side mirror of the rta::Core hooks on top of the C ABI (same names as the reference where one exists)."""

    def __init__(self, device=None, **overrides):
        import torch

        self.lib = load_library()
        p = default_params()
        for k, v in overrides.items():
            if not hasattr(p, k):
                raise KeyError(k)
            setattr(p, k, v)
        self.params = p
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.device = int(device)
        h = C.c_void_p()
        self._check(self.lib.insta_ctx_create(C.byref(p), C.c_int(self.device), C.byref(h)))
        self.h = h
        self.n_verts = self.n_faces = 0

    def _check(self, rc):
        if rc != 0:
            raise InstaError(rc, (self.lib.insta_last_error() or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.insta_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- mesh ingest (rta::Core::load_meshes) ----
    def set_canonical(self, verts, faces, mask_ids=None):
        verts, faces = _np(verts, np.float32), _np(faces, np.uint32)
        m = _np(mask_ids, np.uint32) if mask_ids is not None else None
        self._check(self.lib.insta_set_canonical(self.h, _vp(verts), C.c_uint32(len(verts)), _vp(faces), C.c_uint32(len(faces)), _vp(m) if m is not None else None))
        self.n_verts, self.n_faces = len(verts), len(faces)

    def load_canonical_obj(self, canonical_obj, mesh_masks_dir):
        self._check(self.lib.insta_load_canonical_obj(self.h, canonical_obj.encode(), mesh_masks_dir.encode() if mesh_masks_dir else None))

    def add_frame(self, verts):
        verts = _np(verts, np.float32)
        idx = C.c_int32(-1)
        self._check(self.lib.insta_add_frame(self.h, _vp(verts), C.byref(idx)))
        return idx.value

    def add_frames(self, verts_all, n_threads=0):
        verts_all = _np(verts_all, np.float32)
        self._check(self.lib.insta_add_frames(self.h, _vp(verts_all), C.c_uint32(verts_all.shape[0]), C.c_uint32(n_threads)))

    def replace_frames(self, indices, verts_all, n_threads=0):
        """rta::Core::reload_training_data for the meshes: new vertex sets for resident frames, host builds (the reference's order), storage reused"""
        idx = _np(indices, np.int32)
        verts_all = _np(verts_all, np.float32)
        assert len(idx) == verts_all.shape[0]
        self._check(self.lib.insta_replace_frames(self.h, _vp(idx), _vp(verts_all), C.c_uint32(len(idx)), C.c_uint32(n_threads)))

    def refit_frames(self, indices, verts_all, stream=None):
        """non-parity turnover: the frames keep their trees, boxes refitted on the GPU"""
        idx = _np(indices, np.int32)
        verts_all = _np(verts_all, np.float32)
        assert len(idx) == verts_all.shape[0]
        self._check(self.lib.insta_refit_frames(self.h, _vp(idx), _vp(verts_all), C.c_uint32(len(idx)), _stream(stream)))

    def clear_frames(self):
        self._check(self.lib.insta_clear_frames(self.h))

    @property
    def n_frames(self):
        return self.lib.insta_n_frames(self.h)

    def render_aabb(self):
        out = np.empty(6, np.float32)
        self._check(self.lib.insta_get_render_aabb(self.h, _vp(out)))
        return out

    def set_aabb(self, aabb6):
        a = _np(aabb6, np.float32)
        self._check(self.lib.insta_set_aabb(self.h, _vp(a)))

    def topology(self):
        m = np.empty(self.n_faces, np.uint32)
        adj = np.empty((self.n_faces, 3), np.int32)
        self._check(self.lib.insta_get_topology(self.h, _vp(m), _vp(adj)))
        return m, adj

    def frame_bvh_order(self, frame):
        ids = np.empty(self.n_faces, np.uint32)
        nn = C.c_uint32(0)
        self._check(self.lib.insta_get_frame_bvh_order(self.h, C.c_int32(frame), _vp(ids), C.byref(nn)))
        return ids, nn.value

    def frame_cylinders(self, frame):
        """uint16 [n_quads, 4, 8]: fp16 bits of the traversal's cylinder bounds (empty when the topology has none)"""
        nq = C.c_uint32(0)
        self._check(self.lib.insta_get_frame_cylinders(self.h, C.c_int32(frame), None, C.byref(nq)))
        out = np.empty((nq.value, 4, 8), np.uint16)
        if nq.value:
            self._check(self.lib.insta_get_frame_cylinders(self.h, C.c_int32(frame), _vp(out), C.byref(nq)))
        return out

    # ---- stage 1 ----
    def surface_closest_point(self, coords, tri_ids=None, stream=None):
        """rta::Core::surface_closest_point: [n, 27] records (frame = holder) or [n, 3] points (reference RNG frame draw)."""
        n, m = coords.shape
        assert coords.is_cuda and coords.is_contiguous() and coords.dtype.is_floating_point and coords.element_size() == 4
        if m == 3:
            self._check(self.lib.insta_deform_points(self.h, _dev(coords), C.c_uint32(n), None, _dev(tri_ids), _stream(stream)))
        else:
            self._check(self.lib.insta_deform_samples(self.h, _dev(coords), C.c_uint32(n), C.c_uint32(m), C.c_int32(-1), _dev(tri_ids), _stream(stream)))
        return coords

    def surface_closest_point_n(self, coords, n_valid, tri_ids=None, frame_id=-1, stream=None):
        """like surface_closest_point, the record count comes from the device tensor n_valid (uint32 / int32, 1 element)"""
        n, stride = coords.shape
        self._check(self.lib.insta_deform_samples_n(self.h, _dev(coords), C.c_uint32(n), _dev(n_valid), C.c_uint32(stride), C.c_int32(frame_id), _dev(tri_ids), _stream(stream)))
        return coords

    def render_deformed(self, coords, frame_id, tri_ids=None, stream=None):
        """rta::Core::render_deformed: one fixed frame (m_target_deform_frame) for every record."""
        n, m = coords.shape
        self._check(self.lib.insta_deform_samples(self.h, _dev(coords), C.c_uint32(n), C.c_uint32(m), C.c_int32(frame_id), _dev(tri_ids), _stream(stream)))
        return coords

    def deform_points(self, xyz, frame_ids=None, tri_ids=None, stream=None):
        self._check(self.lib.insta_deform_points(self.h, _dev(xyz), C.c_uint32(xyz.shape[0]), _dev(frame_ids), _dev(tri_ids), _stream(stream)))
        return xyz

    # ---- stage 2 ----
    def grid_layout(self):
        L = self.params.n_levels
        offs = np.empty(L + 1, np.uint32)
        scales = np.empty(L, np.float32)
        res = np.empty(L, np.uint32)
        self._check(self.lib.insta_grid_layout(self.h, _vp(offs), _vp(scales), _vp(res)))
        return offs, scales, res

    @property
    def grid_n_params(self):
        return self.lib.insta_grid_n_params(self.h)

    def hash_encode_fwd(self, pos, table, out, stream=None):
        """pos: float32 [n, stride] (xyz = first three floats of each row); table: fp16 [entries, 8]; out: fp16 [L*8, n]."""
        n, stride = pos.shape
        self._check(self.lib.insta_hash_encode_fwd(self.h, _dev(pos), C.c_uint32(stride), C.c_uint32(n), _dev(table), _dev(out), _stream(stream)))
        return out

    def hash_encode_bwd(self, pos, dLdy, grad, zero_first=True, stream=None):
        n, stride = pos.shape
        self._check(self.lib.insta_hash_encode_bwd(self.h, _dev(pos), C.c_uint32(stride), C.c_uint32(n), _dev(dLdy), _dev(grad), C.c_int(int(zero_first)), _stream(stream)))
        return grad

    def hash_slots(self, pos, slots, stream=None):
        n, stride = pos.shape
        self._check(self.lib.insta_hash_slots(self.h, _dev(pos), C.c_uint32(stride), C.c_uint32(n), _dev(slots), _stream(stream)))
        return slots

    # ---- training ----
    def training_set_images(self, images):
        """images: list of dicts with torch tensors pixels (+ optional seg, depth, rays), res, focal, pp, xform [3,4], type."""
        arr = (TrainImage * len(images))()
        self._keep_images = images
        for k, im in enumerate(images):
            t = arr[k]
            t.d_pixels = im["pixels"].data_ptr()
            t.d_seg_mask = im["seg"].data_ptr() if im.get("seg") is not None else None
            t.d_depth = im["depth"].data_ptr() if im.get("depth") is not None else None
            t.d_rays = im["rays"].data_ptr() if im.get("rays") is not None else None
            t.resolution[0], t.resolution[1] = im["res"]
            t.focal_length[0], t.focal_length[1] = im["focal"]
            t.principal_point[0], t.principal_point[1] = im.get("pp", (0.5, 0.5))
            m = np.asarray(im["xform"], np.float32)
            for col in range(4):
                for row in range(3):
                    t.xform[col * 3 + row] = float(m[row, col])
            t.image_data_type = im.get("type", 3)
        self._check(self.lib.insta_training_set_images(self.h, arr, C.c_uint32(len(images))))

    def make_train_step(self, n_rays, n_rays_total, coords, counters, ray_indices, rays, numsteps, bitfield, extra_dims=None, target_batch_size=1 << 18,
                        rng_seed=1337, rng_advances=0, cone_angle=0.0, loss_scale=128.0, snap=True, random_bg=True, penalize_density=False,
                        background=(0.0, 0.0, 0.0), depth_weight=1.25, seg_weight=40.0, near_distance=0.2, beta_loss=0.0):
        t = TrainStep()
        t.n_rays, t.n_rays_total, t.max_samples, t.target_batch_size, t.stride_floats = n_rays, n_rays_total, coords.shape[0], target_batch_size, coords.shape[1]
        t.snap_to_pixel_centers, t.random_bg_color, t.penalize_density_alpha = int(snap), int(random_bg), int(penalize_density)
        t.rng_seed, t.rng_advances, t.cone_angle_constant, t.loss_scale = rng_seed, rng_advances, cone_angle, loss_scale
        for k in range(3):
            t.background_color[k] = background[k]
        t.depth_weight, t.seg_weight, t.near_distance, t.beta_loss = depth_weight, seg_weight, near_distance, beta_loss
        t.d_density_grid_bitfield = bitfield.data_ptr()
        t.d_extra_dims = extra_dims.data_ptr() if extra_dims is not None else None
        t.d_coords, t.d_counters, t.d_ray_indices, t.d_rays, t.d_numsteps = coords.data_ptr(), counters.data_ptr(), ray_indices.data_ptr(), rays.data_ptr(), numsteps.data_ptr()
        t._keep = (coords, counters, ray_indices, rays, numsteps, bitfield, extra_dims)
        return t

    def generate_training_samples(self, step, stream=None):
        self._check(self.lib.insta_generate_training_samples(self.h, C.byref(step), _stream(stream)))

    def compute_loss(self, step, net_out, out_width, coords_compacted, dloss, loss, mean_density, stream=None):
        self._check(self.lib.insta_compute_loss(self.h, C.byref(step), _dev(net_out), C.c_uint32(out_width), _dev(coords_compacted), _dev(dloss), _dev(loss),
                                                _dev(mean_density), _stream(stream)))

    def optimizer_step(self, current_step, params32, params16, grads16, m1, m2, steps, ema16, lr=0.00175, beta1=0.9, beta2=0.99, eps=1e-15, l2_reg=1e-6,
                       ema_decay=0.95, loss_scale=128.0, stream=None):
        o = Optimizer()
        o.learning_rate, o.beta1, o.beta2, o.epsilon, o.l2_reg, o.ema_decay, o.loss_scale, o.current_step = lr, beta1, beta2, eps, l2_reg, ema_decay, loss_scale, current_step
        o.d_params_fp32, o.d_params_f16, o.d_gradients_f16 = params32.data_ptr(), params16.data_ptr(), grads16.data_ptr()
        o.d_first_moments, o.d_second_moments, o.d_param_steps, o.d_params_ema_f16 = m1.data_ptr(), m2.data_ptr(), steps.data_ptr(), ema16.data_ptr()
        self._check(self.lib.insta_optimizer_step(self.h, C.byref(o), _stream(stream)))

    def make_optimizer(self, current_step, params32, params16, grads16, m1, m2, steps, ema16, lr=0.00175, beta1=0.9, beta2=0.99, eps=1e-15, l2_reg=1e-6,
                       ema_decay=0.95, loss_scale=128.0):
        o = Optimizer()
        o.learning_rate, o.beta1, o.beta2, o.epsilon, o.l2_reg, o.ema_decay, o.loss_scale, o.current_step = lr, beta1, beta2, eps, l2_reg, ema_decay, loss_scale, current_step
        o.d_params_fp32, o.d_params_f16, o.d_gradients_f16 = params32.data_ptr(), params16.data_ptr(), grads16.data_ptr()
        o.d_first_moments, o.d_second_moments, o.d_param_steps, o.d_params_ema_f16 = m1.data_ptr(), m2.data_ptr(), steps.data_ptr(), ema16.data_ptr()
        o._keep = (params32, params16, grads16, m1, m2, steps, ema16)
        return o

    def trainer_create(self, **overrides):
        """Testbed::train_nerf state + workspace; keyword overrides are insta_trainer_config fields."""
        c = TrainerConfig()
        self.lib.insta_trainer_config_default(C.byref(c))
        for k, v in overrides.items():
            if k == "background_color":
                for i in range(3):
                    c.background_color[i] = v[i]
            else:
                if not hasattr(c, k):
                    raise AttributeError(k)
                setattr(c, k, v)
        self._check(self.lib.insta_trainer_create(self.h, C.byref(c)))
        return c

    def trainer_destroy(self):
        self._check(self.lib.insta_trainer_destroy(self.h))

    def trainer_state(self):
        st = TrainStats()
        self._check(self.lib.insta_trainer_state(self.h, C.byref(st)))
        return st.as_dict()

    def train_nerf(self, opt, bitfield, extra_dims, mean_density, get_loss_scalar=False, stream=None):
        """One full training step (insta_train_nerf); opt from make_optimizer (its current_step is advanced here)."""
        st = TrainStats()
        self._check(self.lib.insta_train_nerf(self.h, C.byref(opt), _dev(bitfield), _dev(extra_dims), _dev(mean_density), C.c_int(int(get_loss_scalar)),
                                              C.byref(st), _stream(stream)))
        opt.current_step += 1
        return st.as_dict()

    # ---- occupancy grid ----
    def make_density_grid(self, grid, bitfield, mean, n_cascades=1, rng_seed=0, decay=0.95, has_rays=False):
        """grid float32 [128^3 * n_cascades], bitfield uint8 [2 MiB], mean float32 [1] (torch tensors, kept alive by the struct)."""
        g = DensityGrid()
        g.d_density_grid, g.d_density_grid_bitfield, g.d_density_grid_mean = grid.data_ptr(), bitfield.data_ptr(), mean.data_ptr()
        g.n_cascades, g.ema_step, g.rng_seed, g.rng_advances, g.decay, g.training_step, g.has_rays, g.images_changed = n_cascades, 0, rng_seed, 0, decay, 0, int(has_rays), 0
        g._keep = (grid, bitfield, mean)
        return g

    def update_density_grid(self, g, params, n_uniform, n_nonuniform, stream=None):
        self._check(self.lib.insta_update_density_grid(self.h, C.byref(g), _dev(params), C.c_uint32(n_uniform), C.c_uint32(n_nonuniform), _stream(stream)))

    def training_prep_nerf(self, g, params, stream=None):
        did = C.c_int(0)
        self._check(self.lib.insta_training_prep_nerf(self.h, C.byref(g), _dev(params), C.byref(did), _stream(stream)))
        return bool(did.value)

    def update_density_grid_mean_and_bitfield(self, g, stream=None):
        self._check(self.lib.insta_update_density_grid_mean_and_bitfield(self.h, C.byref(g), _stream(stream)))

    # ---- data-parallel gradient exchange ----
    def comm_unique_id(self):
        buf = (C.c_char * 128)()
        self._check(self.lib.insta_comm_unique_id(buf))
        return bytes(buf)

    def comm_init_rank(self, world_size, rank, unique_id):
        self._check(self.lib.insta_comm_init_rank(self.h, C.c_int(world_size), C.c_int(rank), C.c_char_p(unique_id) if unique_id is not None else None))

    def comm_destroy(self):
        self._check(self.lib.insta_comm_destroy(self.h))

    @property
    def comm_exchange_mode(self):
        """0 = single rank, 1 = ncclAllReduce, 2 = the library's own peer-memory kernel (after the first training step)"""
        return int(self.lib.insta_comm_exchange_mode(self.h))

    def grad_allreduce(self, grads16, stream=None):
        self._check(self.lib.insta_grad_allreduce(self.h, _dev(grads16), C.c_uint64(grads16.numel()), _stream(stream)))

    # ---- network ----
    @property
    def network_n_params(self):
        return self.lib.insta_network_n_params(self.h)

    def network_inference(self, records, params, out, out_stride=0, n=None, stream=None):
        """records float32 [n, stride]; params fp16 [network_n_params]; out fp16 [n, 4] (out_stride 0) or [rows, out_stride]."""
        n = records.shape[0] if n is None else n
        self._check(self.lib.insta_network_inference(self.h, _dev(records), C.c_uint32(records.shape[1]), C.c_uint32(n), _dev(params), _dev(out),
                                                     C.c_uint32(out_stride), _stream(stream)))
        return out

    def network_inference_n(self, records, n_valid, params, out, out_stride=0, stream=None):
        self._check(self.lib.insta_network_inference_n(self.h, _dev(records), C.c_uint32(records.shape[1]), C.c_uint32(records.shape[0]), _dev(n_valid), _dev(params),
                                                       _dev(out), C.c_uint32(out_stride), _stream(stream)))
        return out

    def network_density(self, xyz, params, logits, stream=None):
        """xyz float32 [n, stride>=3]; logits fp16 [n]."""
        self._check(self.lib.insta_network_density(self.h, _dev(xyz), C.c_uint32(xyz.shape[1]), C.c_uint32(xyz.shape[0]), _dev(params), _dev(logits), _stream(stream)))
        return logits

    def network_forward_backward(self, records, params, dloss, grads, out=None, n=None, stream=None):
        """records float32 [n, stride]; dloss fp16 [n, width]; grads fp16 [network_n_params] (overwritten); out fp16 [n, 4] or None."""
        n = records.shape[0] if n is None else n
        self._check(self.lib.insta_network_forward_backward(self.h, _dev(records), C.c_uint32(records.shape[1]), C.c_uint32(n), _dev(params), _dev(dloss),
                                                            C.c_uint32(dloss.shape[1]), _dev(grads), _dev(out), _stream(stream)))
        return grads

    # ---- stage 3: render tracer (NerfTracer) ----
    def tracer_init_rays(self, cam, frame_rgba, frame_depth, bitfield=None, render_aabb=None, cone_angle=0.0, min_transmittance=0.01, linear_colors=False,
                         dbg_origin_t=None, dbg_dir_alive=None, stream=None):
        o = RenderOpts()
        o.d_density_grid_bitfield = bitfield.data_ptr() if bitfield is not None else None
        o.cone_angle_constant = cone_angle
        o.min_transmittance = min_transmittance
        o.use_render_aabb = int(render_aabb is not None)
        for k in range(6):
            o.render_aabb[k] = float(render_aabb[k]) if render_aabb is not None else 0.0
        o.linear_colors = int(linear_colors)
        self._keep_opts = (o, bitfield)
        n = C.c_uint32(0)
        self._check(self.lib.insta_tracer_init_rays(self.h, C.byref(cam), C.byref(o), _dev(frame_rgba), _dev(frame_depth), _dev(dbg_origin_t), _dev(dbg_dir_alive),
                                                    C.byref(n), _stream(stream)))
        return n.value

    def _opts(self, bitfield, render_aabb, cone_angle, min_transmittance, linear_colors, max_steps=0):
        o = RenderOpts()
        o.max_steps_per_iteration = int(max_steps)
        o.d_density_grid_bitfield = bitfield.data_ptr() if bitfield is not None else None
        o.cone_angle_constant = cone_angle
        o.min_transmittance = min_transmittance
        o.use_render_aabb = int(render_aabb is not None)
        for k in range(6):
            o.render_aabb[k] = float(render_aabb[k]) if render_aabb is not None else 0.0
        o.linear_colors = int(linear_colors)
        return o

    def render_frame(self, cam, frame_id, params, frame_rgba, frame_depth, bitfield=None, extra_dims=None, render_aabb=None, cone_angle=0.0,
                     min_transmittance=0.01, linear_colors=False, stream=None, max_steps=0):
        """Testbed::render_nerf with this library's deformation + network; returns (iterations, samples).  max_steps: march steps per
        iteration (8 = the reference's schedule, 0 = library default; the frame is the same either way)."""
        o = self._opts(bitfield, render_aabb, cone_angle, min_transmittance, linear_colors, max_steps)
        st = RenderStats()
        self._check(self.lib.insta_render_frame(self.h, C.byref(cam), C.byref(o), C.c_int32(frame_id), _dev(params), _dev(extra_dims), _dev(frame_rgba),
                                                _dev(frame_depth), C.byref(st), _stream(stream)))
        return st.iterations, st.samples

    def tracer_n_steps(self):
        return self.lib.insta_tracer_n_steps(self.h)

    def tracer_generate_samples(self, records, n_steps, frame_id, extra_dims=None, stream=None):
        n = C.c_uint32(0)
        self._check(self.lib.insta_tracer_generate_samples(self.h, _dev(records), C.c_uint32(records.shape[1]), C.c_uint32(n_steps), C.c_int32(frame_id),
                                                           _dev(extra_dims), C.byref(n), _stream(stream)))
        return n.value

    def tracer_composite(self, records, net_out, out_stride, n_steps, frame_rgba, frame_depth, stream=None):
        n = C.c_uint32(0)
        self._check(self.lib.insta_tracer_composite(self.h, _dev(records), C.c_uint32(records.shape[1]), _dev(net_out), C.c_uint32(out_stride), C.c_uint32(n_steps),
                                                    _dev(frame_rgba), _dev(frame_depth), C.byref(n), _stream(stream)))
        return n.value

    # ---- recorder read-back (rta::Recorder::dump_frame_buffer) ----
    def recorder_readback(self, frame_rgba, frame_depth, width, height, h_rgba8=None, h_depth_mm=None, srgb_to_linear=False, stream=None):
        """frame_rgba [w*h, 4] float32 / frame_depth [w*h] float32 on the device -> h_rgba8 (uint8 [h, w, 4]) and h_depth_mm (int32 / uint32
        [h-1, w]) host tensors (pinned for asynchronous copies).  No synchronisation: sync the stream before reading them."""
        self._check(self.lib.insta_recorder_readback(self.h, _dev(frame_rgba), _dev(frame_depth), C.c_uint32(width), C.c_uint32(height), C.c_uint32(int(srgb_to_linear)),
                                                     C.c_void_p(h_rgba8.data_ptr()) if h_rgba8 is not None else None,
                                                     C.c_void_p(h_depth_mm.data_ptr()) if h_depth_mm is not None else None, _stream(stream)))

    # ---- host-buffer path ----
    def deform_encode_host(self, h_coords, frame_id, table, h_out):
        """h_coords: pinned float32 torch tensor [n, stride] (mutated in place); h_out: pinned fp16 [L*8, n]."""
        n, stride = h_coords.shape
        self._check(self.lib.insta_deform_encode_host(self.h, C.c_void_p(h_coords.data_ptr()), C.c_uint32(n), C.c_uint32(stride), C.c_int32(frame_id),
                                                      _dev(table), C.c_void_p(h_out.data_ptr())))

    def deform_infer_host(self, h_coords, frame_id, params, h_out):
        """h_coords: pinned float32 torch tensor [n, stride] (read only); h_out: pinned fp16 [n, 4] (r, g, b, sigma logits)."""
        n, stride = h_coords.shape
        self._check(self.lib.insta_deform_infer_host(self.h, C.c_void_p(h_coords.data_ptr()), C.c_uint32(n), C.c_uint32(stride), C.c_int32(frame_id),
                                                     _dev(params), C.c_void_p(h_out.data_ptr())))

    def microbench_gather(self, buffer_bytes, chunk_bytes=16, iters=5):
        """GB/s of fully divergent chunk_bytes gathers from a buffer of buffer_bytes (L2-resident when it fits)."""
        g = C.c_float(0)
        self._check(self.lib.insta_microbench_gather(self.h, C.c_uint64(buffer_bytes), C.c_uint32(chunk_bytes), C.c_uint32(iters), C.byref(g), _stream(None)))
        return float(g.value)

    def mesh_ray_intersect(self, origins, dirs, tri_ids, t=None, frame_id=-1, stream=None):
        """origins / dirs: float32 [n, 3] device tensors; tri_ids int32 [n] (topology id of the first hit, -1 = none); t float32 [n] or None."""
        n = origins.shape[0]
        self._check(self.lib.insta_mesh_ray_intersect(self.h, C.c_int32(frame_id), _dev(origins), _dev(dirs), C.c_uint32(n), _dev(tri_ids), _dev(t), _stream(stream)))

    def raycast_mesh(self, cam, frame_rgba, frame_depth, frame_id=-1, color_normal=False, render_aabb=None, stream=None):
        """rta::Core::raycast_flame: the shaded mesh of frame `frame_id` (-1 = canonical) into frame_rgba [h, w, 4] / frame_depth [h, w] (float32, device)."""
        o = RenderOpts()
        o.use_render_aabb = int(render_aabb is not None)
        for k in range(6):
            o.render_aabb[k] = float(render_aabb[k]) if render_aabb is not None else 0.0
        self._check(self.lib.insta_raycast_mesh(self.h, C.byref(cam), C.byref(o), C.c_int32(frame_id), C.c_int(int(color_normal)), _dev(frame_rgba), _dev(frame_depth), _stream(stream)))

    def microbench_scatter(self, buffer_bytes, chunk_bytes=16, iters=5):
        """GB/s of fully divergent 16-byte fp16 vector reductions into a buffer of buffer_bytes (chunk_bytes 32 / 64: 2 / 4 adjacent lanes on adjacent entries)."""
        g = C.c_float(0)
        self._check(self.lib.insta_microbench_scatter(self.h, C.c_uint64(buffer_bytes), C.c_uint32(chunk_bytes), C.c_uint32(iters), C.byref(g), _stream(None)))
        return float(g.value)

    @property
    def launch_count(self):
        return self.lib.insta_launch_count(self.h)


# ---- host-only utilities (no GPU) ----------------------------------------------------------------------------------
TRI_DTYPE = np.dtype([("a", "<f4", 3), ("b", "<f4", 3), ("c", "<f4", 3), ("scaling", "<f4"), ("id", "<u4"), ("mask_id", "<u4")])
NODE_DTYPE = np.dtype([("mn", "<f4", 3), ("mx", "<f4", 3), ("left_idx", "<i4"), ("right_idx", "<i4")])


def _hcheck(rc):
    if rc != 0:
        raise InstaError(rc, (load_library().insta_last_error() or b"").decode())


def host_build_triangles(verts, faces, mask_ids=None, canonical=None):
    verts, faces = _np(verts, np.float32), _np(faces, np.uint32)
    m = _np(mask_ids, np.uint32) if mask_ids is not None else None
    out = np.zeros(len(faces), TRI_DTYPE)
    _hcheck(load_library().insta_host_build_triangles(_vp(verts), _vp(faces), C.c_uint32(len(faces)), _vp(m) if m is not None else None,
                                                      _vp(canonical) if canonical is not None else None, _vp(out)))
    return out


def host_bvh_build(tris, leaf_size=2):
    t = np.ascontiguousarray(tris.copy())
    nodes = np.zeros(4 * len(t) + 16, NODE_DTYPE)
    nn = C.c_uint32(0)
    _hcheck(load_library().insta_host_bvh_build(_vp(t), C.c_uint32(len(t)), C.c_uint32(leaf_size), _vp(nodes), C.c_uint32(len(nodes)), C.byref(nn)))
    return nodes[: nn.value].copy(), t


def host_write_png(path, rgba8):
    """rgba8: uint8 [h, w, 4] -> PNG (rta::Recorder save_rgba's stbi_write_png, recorder.cu:59)."""
    a = _np(rgba8, np.uint8)
    h, w = a.shape[:2]
    _hcheck(load_library().insta_host_write_png(str(path).encode(), C.c_uint32(w), C.c_uint32(h), _vp(a)))


def host_write_depth_bin(path, depth_mm):
    a = _np(depth_mm, np.uint32)
    _hcheck(load_library().insta_host_write_depth_bin(str(path).encode(), _vp(a), C.c_uint64(a.size)))


def host_write_trajectory_json(path, width, height, fl, c, cameras, frame_indices, suffixes=None):
    """rta::Recorder::dump_cameras_json: cameras [n, 3, 4] (Core::m_camera per recorded frame), frame_indices [n]"""
    cams = _np(cameras, np.float32).reshape(-1, 12)
    n = len(cams)
    arr = (TrajectoryFrame * max(n, 1))()
    for i in range(n):
        for k in range(12):
            arr[i].camera_matrix[k] = float(cams[i, k])
        arr[i].frame_index = int(frame_indices[i])
        arr[i].suffix = suffixes[i].encode() if suffixes and suffixes[i] else None
    lib = load_library()
    lib.insta_host_write_trajectory_json.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_uint32]
    _hcheck(lib.insta_host_write_trajectory_json(str(path).encode(), width, height, fl[0], fl[1], c[0], c[1], C.cast(arr, C.c_void_p), n))


def host_adjacency(faces, n_verts):
    faces = _np(faces, np.uint32)
    out = np.empty((len(faces), 3), np.int32)
    _hcheck(load_library().insta_host_adjacency(_vp(faces), C.c_uint32(len(faces)), C.c_uint32(n_verts), _vp(out)))
    return out


def host_load_canonical_obj(canonical_obj, mesh_masks_dir):
    lib = load_library()
    nv, nf = C.c_uint32(0), C.c_uint32(0)
    md = mesh_masks_dir.encode() if mesh_masks_dir else None
    _hcheck(lib.insta_host_load_canonical_obj(canonical_obj.encode(), md, None, None, None, C.byref(nv), C.byref(nf)))
    verts = np.empty((nv.value, 3), np.float32)
    faces = np.empty((nf.value, 3), np.uint32)
    masks = np.empty(nf.value, np.uint32)
    _hcheck(lib.insta_host_load_canonical_obj(canonical_obj.encode(), md, _vp(verts), _vp(faces), _vp(masks), C.byref(nv), C.byref(nf)))
    return verts, faces, masks


def host_mesh_aabb(verts):
    verts = _np(verts, np.float32)
    out = np.empty(6, np.float32)
    _hcheck(load_library().insta_host_mesh_aabb(_vp(verts), C.c_uint32(len(verts)), _vp(out)))
    return out


# ---- snapshots (host-only) ----
def snapshot_save(path, params16, density_grid16=None, training_step=0, loss=0.0, aabb=(0, 0, 0, 1, 1, 1), render_aabb=(0, 0, 0, 1, 1, 1), bounding_radius=1.0,
                  aabb_scale=1, rays_per_batch=1 << 12, measured_batch_size=0, measured_batch_size_before_compaction=0, optimizer=None, lib_params=None,
                  learning_rate=1.75e-3, ema_decay=0.95, beta1=0.9, beta2=0.99, epsilon=1e-15, l2_reg=1e-6):
    """numpy arrays in, reference-layout MessagePack out.  optimizer: dict(current_step, m1, m2, steps, ema16) or None."""
    lib = load_library()
    s = Snapshot()
    keep = [np.ascontiguousarray(params16, np.float16)]
    s.n_params, s.h_params_f16 = keep[0].size, keep[0].ctypes.data
    if density_grid16 is not None:
        keep.append(np.ascontiguousarray(density_grid16, np.float16))
        s.n_density_grid, s.h_density_grid_f16 = keep[-1].size, keep[-1].ctypes.data
    s.training_step, s.loss, s.bounding_radius, s.aabb_scale = training_step, loss, bounding_radius, aabb_scale
    for k in range(6):
        s.aabb[k], s.render_aabb[k] = aabb[k], render_aabb[k]
    s.rays_per_batch, s.measured_batch_size, s.measured_batch_size_before_compaction = rays_per_batch, measured_batch_size, measured_batch_size_before_compaction
    s.ema_decay, s.beta1, s.beta2, s.epsilon, s.l2_reg, s.opt_base_learning_rate = ema_decay, beta1, beta2, epsilon, l2_reg, learning_rate
    if optimizer is not None:
        m1, m2 = np.ascontiguousarray(optimizer["m1"], np.float32), np.ascontiguousarray(optimizer["m2"], np.float32)
        st, ema = np.ascontiguousarray(optimizer["steps"], np.uint32), np.ascontiguousarray(optimizer["ema16"], np.float16)
        keep += [m1, m2, st, ema]
        s.has_optimizer, s.opt_current_step = 1, int(optimizer["current_step"])
        s.h_first_moments, s.h_second_moments, s.h_param_steps, s.h_weights_ema_f16 = m1.ctypes.data, m2.ctypes.data, st.ctypes.data, ema.ctypes.data
    _hcheck(lib.insta_snapshot_save(str(path).encode(), C.byref(lib_params) if lib_params is not None else None, C.byref(s)))


def snapshot_load(path):
    """dict of numpy arrays / scalars (copies; the file is closed before returning)"""
    lib = load_library()
    s, f = Snapshot(), C.c_void_p()
    _hcheck(lib.insta_snapshot_open(str(path).encode(), C.byref(f), C.byref(s)))
    try:
        def arr(ptr, n, dt):
            if not ptr or n == 0:
                return None
            return np.frombuffer(C.string_at(ptr, n * np.dtype(dt).itemsize), dtype=dt).copy()
        out = {k: getattr(s, k) for k in ("n_params", "training_step", "loss", "bounding_radius", "aabb_scale", "rays_per_batch", "measured_batch_size",
                                          "measured_batch_size_before_compaction", "ema_decay", "beta1", "beta2", "epsilon", "l2_reg", "opt_base_learning_rate",
                                          "has_optimizer", "opt_current_step")}
        out["aabb"], out["render_aabb"] = np.array(list(s.aabb), np.float32), np.array(list(s.render_aabb), np.float32)
        out["params16"] = arr(s.h_params_f16, s.n_params, np.float16)
        out["density_grid16"] = arr(s.h_density_grid_f16, s.n_density_grid, np.float16)
        if s.has_optimizer:
            out["m1"], out["m2"] = arr(s.h_first_moments, s.n_params, np.float32), arr(s.h_second_moments, s.n_params, np.float32)
            out["steps"] = arr(s.h_param_steps, s.n_params, np.uint32)
            out["ema16"] = arr(s.h_weights_ema_f16, s.n_params, np.float16)
        return out
    finally:
        lib.insta_snapshot_close(f)
