"""TEST INFRASTRUCTURE ONLY: ctypes launcher for the reference's own network (oracle/_ref/libref_net.so = ngp::NerfNetwork<__half> over
tiny-cuda-nn, built by `make -C oracle refnet`).  Device buffers are torch tensors."""
import ctypes as C
import os

import numpy as np

from .pyoracle import HERE

REF_NET_SO = os.path.join(HERE, "_ref", "libref_net.so")


def available():
    return os.path.exists(REF_NET_SO)


class RefNet:
    def __init__(self, n_extra_dims=16, n_levels=16, n_features=8, log2_hashmap_size=18, base_resolution=16, per_level_scale=None):
        self.lib = C.CDLL(REF_NET_SO)
        self.lib.refnet_create.restype = C.c_long
        self.lib.refnet_last_error.restype = C.c_char_p
        if per_level_scale is None:  # testbed.cu:2246-2258 in fp32
            per_level_scale = float(np.exp(np.log(np.float32(2048.0) / np.float32(base_resolution)) / np.float32(n_levels - 1)).astype(np.float32))
        n = self.lib.refnet_create(C.c_uint32(n_extra_dims), C.c_uint32(n_levels), C.c_uint32(n_features), C.c_uint32(log2_hashmap_size), C.c_uint32(base_resolution),
                                   C.c_float(per_level_scale))
        if n <= 0:
            raise RuntimeError("refnet_create: " + self.lib.refnet_last_error().decode())
        self.n_params = int(n)
        self.out_width = int(self.lib.refnet_padded_output_width())
        self.stride = 11 + n_extra_dims

    @staticmethod
    def _s():
        import torch

        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _check(self, rc, who):
        if rc != 0:
            raise RuntimeError(who + ": " + self.lib.refnet_last_error().decode())

    def set_params(self, params16):
        assert params16.numel() == self.n_params
        self._check(self.lib.refnet_set_params(C.c_void_p(params16.data_ptr())), "set_params")

    def inference(self, records, out_rm):
        """records [n, stride] float32 (n % 128 == 0); out_rm fp16 [out_width, n]"""
        n = records.shape[0]
        assert n % 128 == 0 and records.shape[1] == self.stride and out_rm.shape == (self.out_width, n)
        self._check(self.lib.refnet_inference(C.c_void_p(records.data_ptr()), C.c_uint32(n), C.c_void_p(out_rm.data_ptr()), self._s()), "inference")

    def density(self, xyz, out_rm):
        n = xyz.shape[0]
        assert n % 128 == 0 and xyz.shape[1] == 3 and out_rm.shape[1] == n
        self._check(self.lib.refnet_density(C.c_void_p(xyz.data_ptr()), C.c_uint32(n), C.c_void_p(out_rm.data_ptr()), self._s()), "density")

    def forward_backward(self, records, dloss_cm, out_cm, grads):
        n = records.shape[0]
        assert n % 128 == 0 and dloss_cm.shape == (n, self.out_width) and out_cm.shape == (n, self.out_width) and grads.numel() == self.n_params
        self._check(self.lib.refnet_forward_backward(C.c_void_p(records.data_ptr()), C.c_uint32(n), C.c_void_p(dloss_cm.data_ptr()), C.c_void_p(out_cm.data_ptr()),
                                                     C.c_void_p(grads.data_ptr()), self._s()), "forward_backward")

    def inference_cm(self, records, out_cm):
        """out_cm fp16 [n, out_width] (one row of out_width halves per sample, as train_nerf_step reads it)"""
        n = records.shape[0]
        assert n % 128 == 0 and out_cm.shape[1] == self.out_width and out_cm.shape[0] >= n
        self._check(self.lib.refnet_inference_cm(C.c_void_p(records.data_ptr()), C.c_uint32(n), C.c_void_p(out_cm.data_ptr()), self._s()), "inference_cm")

    def optimizer_step(self, loss_scale=128.0):
        self._check(self.lib.refnet_optimizer_step(C.c_float(loss_scale), self._s()), "optimizer_step")

    def set_grads(self, grads16):
        assert grads16.numel() == self.n_params
        self._check(self.lib.refnet_set_grads(C.c_void_p(grads16.data_ptr()), self._s()), "set_grads")

    def optimizer_state(self, master32, params16, ema16):
        self._check(self.lib.refnet_get_optimizer_state(C.c_void_p(master32.data_ptr()), C.c_void_p(params16.data_ptr()), C.c_void_p(ema16.data_ptr()), self._s()), "optimizer_state")

    def reset_optimizer(self):
        self.lib.refnet_reset_optimizer()
