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
