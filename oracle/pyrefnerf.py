"""TEST INFRASTRUCTURE ONLY: ctypes launcher for the reference's own stage-3 kernels (oracle/_ref/libref_nerf.so, built by
`make -C oracle refnerf` from ngp/src/testbed_nerf.cu).  Device buffers are torch tensors."""
import ctypes as C
import os

import numpy as np

from .pyoracle import REF_NERF_SO


def available():
    return os.path.exists(REF_NERF_SO)


class RefNerf:
    def __init__(self):
        self.lib = C.CDLL(REF_NERF_SO)

    @staticmethod
    def _s():
        import torch

        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def init_rays(self, cam3x4, focal, center, res, spp, snap, render_aabb, train_aabb, bitfield, cone, min_T, frame_rgba, frame_depth):
        m = np.asarray(cam3x4, np.float32)
        cam = (C.c_float * 12)(*[float(m[r, c]) for c in range(4) for r in range(3)])
        f = (C.c_float * 2)(focal, focal)
        ce = (C.c_float * 2)(*center)
        rs = (C.c_uint32 * 2)(*res)
        ra = (C.c_float * 6)(*[float(x) for x in render_aabb])
        ta = (C.c_float * 6)(*[float(x) for x in train_aabb])
        self._keep = bitfield
        rc = self.lib.refnerf_init_rays(cam, f, ce, rs, C.c_uint32(spp), C.c_int(int(snap)), ra, ta, C.c_void_p(bitfield.data_ptr()) if bitfield is not None else None,
                                        C.c_float(cone), C.c_float(min_T), C.c_void_p(frame_rgba.data_ptr()), C.c_void_p(frame_depth.data_ptr()), self._s())
        assert rc == 0, rc

    def export_init(self, o_t, d_alive):
        rc = self.lib.refnerf_export_init(C.c_void_p(o_t.data_ptr()), C.c_void_p(d_alive.data_ptr()), self._s())
        assert rc == 0, rc

    def compact_and_generate(self, records, n_extra, extra_dims, frame_id):
        ns = C.c_uint32(0)
        n = self.lib.refnerf_compact_and_generate(C.c_void_p(records.data_ptr()), C.c_uint32(n_extra), C.c_void_p(extra_dims.data_ptr()) if extra_dims is not None else None,
                                                  C.c_uint32(frame_id), C.byref(ns), self._s())
        return n, ns.value

    def composite(self, records, n_extra, net_out, out_stride, n_steps):
        rc = self.lib.refnerf_composite(C.c_void_p(records.data_ptr()), C.c_uint32(n_extra), C.c_void_p(net_out.data_ptr()), C.c_uint32(out_stride), C.c_uint32(n_steps),
                                        self._s())
        assert rc == 0, rc

    def shade(self, frame_rgba, frame_depth, linear_colors=False):
        return self.lib.refnerf_shade(C.c_void_p(frame_rgba.data_ptr()), C.c_void_p(frame_depth.data_ptr()), C.c_int(int(linear_colors)), self._s())


    # ---- training kernels ----
    def train_set_images(self, images):
        n = len(images)
        self._keep_images = images
        ptrs = (C.c_void_p * (4 * n))()
        res = (C.c_uint32 * (2 * n))()
        focal = (C.c_float * (2 * n))()
        pp = (C.c_float * (2 * n))()
        xf = (C.c_float * (12 * n))()
        typ = (C.c_uint32 * n)()
        for k, im in enumerate(images):
            for j, key in enumerate(("pixels", "seg", "depth", "rays")):
                ptrs[4 * k + j] = im[key].data_ptr() if im.get(key) is not None else None
            res[2 * k], res[2 * k + 1] = im["res"]
            focal[2 * k], focal[2 * k + 1] = im["focal"]
            pp[2 * k], pp[2 * k + 1] = im.get("pp", (0.5, 0.5))
            m = np.asarray(im["xform"], np.float32)
            for col in range(4):
                for row in range(3):
                    xf[12 * k + col * 3 + row] = float(m[row, col])
            typ[k] = im.get("type", 3)
        rc = self.lib.refnerf_train_set_images(C.c_uint32(n), ptrs, res, focal, pp, xf, typ)
        assert rc == 0, rc

    def train_generate(self, n_rays, n_rays_total, coords, counters, ray_indices, rays, numsteps, bitfield, aabb, extra_dims, rng_seed=1337, rng_advances=0,
                       cone=0.0, snap=True):
        a = (C.c_float * 6)(*[float(x) for x in aabb])
        n_extra = coords.shape[1] - 11
        rc = self.lib.refnerf_train_generate(C.c_uint32(n_rays), C.c_uint32(n_rays_total), C.c_uint32(coords.shape[0]), C.c_uint64(rng_seed), C.c_uint64(rng_advances), a,
                                             C.c_void_p(bitfield.data_ptr()), C.c_float(cone), C.c_int(int(snap)), C.c_void_p(extra_dims.data_ptr()), C.c_uint32(n_extra),
                                             C.c_void_p(coords.data_ptr()), C.c_void_p(counters.data_ptr()), C.c_void_p(ray_indices.data_ptr()), C.c_void_p(rays.data_ptr()),
                                             C.c_void_p(numsteps.data_ptr()), self._s())
        assert rc == 0, rc

    def train_loss(self, n_rays, n_rays_total, target_batch, coords, counters, ray_indices, rays, numsteps, net_out, out_width, coords_compacted, dloss, loss,
                   mean_density, aabb, rng_seed=1337, rng_advances=0, loss_scale=128.0, bg=(0, 0, 0), random_bg=True, snap=True, depth_lambda=1.25, seg_lambda=40.0,
                   near_distance=0.2, penalize_density=False, beta=0.0):
        a = (C.c_float * 6)(*[float(x) for x in aabb])
        b = (C.c_float * 3)(*bg)
        n_extra = coords.shape[1] - 11
        rc = self.lib.refnerf_train_loss(C.c_uint32(n_rays), C.c_uint32(n_rays_total), C.c_uint32(target_batch), C.c_uint64(rng_seed), C.c_uint64(rng_advances), a,
                                         C.c_float(loss_scale), C.c_uint32(out_width), b, C.c_int(int(random_bg)), C.c_void_p(net_out.data_ptr()), C.c_void_p(counters.data_ptr()),
                                         C.c_void_p(ray_indices.data_ptr()), C.c_void_p(rays.data_ptr()), C.c_void_p(numsteps.data_ptr()), C.c_void_p(coords.data_ptr()),
                                         C.c_void_p(coords_compacted.data_ptr()), C.c_uint32(n_extra), C.c_void_p(dloss.data_ptr()), C.c_void_p(loss.data_ptr()), C.c_int(int(snap)),
                                         C.c_void_p(mean_density.data_ptr()), C.c_float(depth_lambda), C.c_float(seg_lambda), C.c_float(near_distance),
                                         C.c_int(int(penalize_density)), C.c_float(beta), self._s())
        assert rc == 0, rc

    # ---- occupancy-grid maintenance kernels ----
    def grid_mark_untrained(self, grid, clear_visible):
        rc = self.lib.refnerf_grid_mark_untrained(C.c_uint32(grid.numel()), C.c_void_p(grid.data_ptr()), C.c_int(int(clear_visible)), self._s())
        assert rc == 0, rc

    def grid_generate(self, n, rng_seed, rng_advances, step, aabb, grid, positions, indices, n_cascades, thresh):
        a = (C.c_float * 6)(*[float(x) for x in aabb])
        rc = self.lib.refnerf_grid_generate(C.c_uint32(n), C.c_uint64(rng_seed), C.c_uint64(rng_advances), C.c_uint32(step), a, C.c_void_p(grid.data_ptr()),
                                            C.c_void_p(positions.data_ptr()), C.c_void_p(indices.data_ptr()), C.c_uint32(n_cascades), C.c_float(thresh), self._s())
        assert rc == 0, rc

    def grid_splat_ema(self, indices, logits, decay, ema_step, grid, grid_tmp):
        rc = self.lib.refnerf_grid_splat_ema(C.c_uint32(indices.numel()), C.c_void_p(indices.data_ptr()), C.c_void_p(logits.data_ptr()), C.c_uint32(grid.numel()),
                                             C.c_float(decay), C.c_uint32(ema_step), C.c_void_p(grid.data_ptr()), C.c_void_p(grid_tmp.data_ptr()), self._s())
        assert rc == 0, rc

    def grid_mean_and_bitfield(self, grid, bitfield, mean, n_cascades):
        rc = self.lib.refnerf_grid_mean_and_bitfield(C.c_void_p(grid.data_ptr()), C.c_void_p(bitfield.data_ptr()), C.c_void_p(mean.data_ptr()), C.c_uint32(n_cascades), self._s())
        assert rc == 0, rc
