"""TEST INFRASTRUCTURE ONLY: ctypes wrapper of oracle/_ref/libref_json.so (oracle/ref_json.cpp: the reference's snapshot load / save restated over the
nlohmann::json header the reference vendors)."""
import ctypes as C
import os

import numpy as np

SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_json.so")


def available():
    return os.path.exists(SO)


class RefJson:
    def __init__(self):
        self.lib = C.CDLL(SO)
        self.lib.refjson_last_message.restype = C.c_char_p

    def load_like_reference(self, path):
        """Raises with the reference loader's own exception text when it would throw; returns what it read otherwise."""
        out = (C.c_uint64 * 8)()
        rc = self.lib.refjson_load_like_reference(str(path).encode(), out)
        if rc:
            raise RuntimeError(self.lib.refjson_last_message().decode())
        k = ("n_params", "n_density_grid", "training_step_resumed", "rays_per_batch", "has_optimizer", "current_step", "exp_params", "geo_params")
        return dict(zip(k, [int(v) for v in out]))

    def save_like_reference(self, path, main_json_text, params16, grid16, training_step, loss, aabb, render_aabb, bounding_radius, aabb_scale, rays_per_batch,
                            measured, measured_before, optimizer=None):
        p = np.ascontiguousarray(params16, np.float16)
        g = np.ascontiguousarray(grid16 if grid16 is not None else np.zeros(0, np.float16), np.float16)
        a = (C.c_float * 6)(*aabb)
        r = (C.c_float * 6)(*render_aabb)
        vp = lambda x: x.ctypes.data_as(C.c_void_p)
        if optimizer:
            m1, m2 = np.ascontiguousarray(optimizer["m1"], np.float32), np.ascontiguousarray(optimizer["m2"], np.float32)
            st, ema = np.ascontiguousarray(optimizer["steps"], np.uint32), np.ascontiguousarray(optimizer["ema16"], np.float16)
            args = (1, C.c_uint32(optimizer["current_step"]), C.c_float(optimizer.get("base_lr", 1.75e-3)), vp(m1), vp(m2), vp(st), vp(ema))
        else:
            args = (0, C.c_uint32(0), C.c_float(0), None, None, None, None)
        rc = self.lib.refjson_save_like_reference(str(path).encode(), main_json_text.encode(), C.c_uint64(len(p)), vp(p), C.c_uint64(len(g)), vp(g), C.c_uint32(training_step),
                                                  C.c_float(loss), a, r, C.c_float(bounding_radius), C.c_int(aabb_scale), C.c_uint32(rays_per_batch), C.c_uint32(measured),
                                                  C.c_uint32(measured_before), *args)
        if rc:
            raise RuntimeError(self.lib.refjson_last_message().decode())

    def dump_cameras(self, w, h, fl, c, cams, ids):
        """the text Recorder::dump_cameras_json writes for these cameras (n x 12 floats, 3x4 row-major) and frame indices"""
        cams = np.ascontiguousarray(cams, np.float32).reshape(-1, 12)
        ids = np.ascontiguousarray(ids, np.uint32)
        rc = self.lib.refjson_dump_cameras(C.c_int(w), C.c_int(h), C.c_float(fl[0]), C.c_float(fl[1]), C.c_float(c[0]), C.c_float(c[1]), cams.ctypes.data_as(C.c_void_p),
                                           ids.ctypes.data_as(C.c_void_p), C.c_int(len(ids)))
        msg = self.lib.refjson_last_message().decode()
        if rc:
            raise RuntimeError(msg)
        return msg
