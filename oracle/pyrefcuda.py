"""TEST INFRASTRUCTURE ONLY: ctypes launcher for the reference's own CUDA kernels (oracle/_ref/libref_cuda.so).
Device buffers are torch tensors; used by tests/test_gpu_vs_ref_cuda.py and tests/bench_ref_cuda.py."""
import ctypes as C
import os

import numpy as np

from .pyoracle import REF_CUDA_SO


def available():
    return os.path.exists(REF_CUDA_SO)


class RefCuda:
    def __init__(self):
        self.lib = C.CDLL(REF_CUDA_SO)

    @staticmethod
    def _s():
        import torch

        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def upload_scene(self, nodes_list, tris_bvh_list, tris_topo_list, canon, adjacency):
        """numpy arrays in the reference layouts -> device buffers + pointer tables (rta::Core::load_meshes staging)."""
        import torch

        def dev(a):
            return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).cuda()

        self._keep = [[dev(a) for a in lst] for lst in (nodes_list, tris_bvh_list, tris_topo_list)]
        self.ptrs = [torch.tensor([t.data_ptr() for t in lst], dtype=torch.int64, device="cuda") for lst in self._keep]
        self.canon = dev(canon)
        self.adj = torch.from_numpy(np.ascontiguousarray(adjacency, np.int32)).cuda()
        self.n_meshes = len(nodes_list)

    def warp(self, d_samples, stride, frame_id=-1, disable_interp=False, aabb6=(0, 0, 0, 1, 1, 1), extra_dims=16, use_geo=False):
        n = d_samples.numel() // stride
        aabb = (C.c_float * 6)(*aabb6)
        rc = self.lib.refcuda_warp(C.c_void_p(d_samples.data_ptr()), C.c_uint32(n), C.c_uint32(stride), C.c_int32(frame_id), C.c_uint32(self.n_meshes),
                                   C.c_void_p(self.ptrs[0].data_ptr()), C.c_void_p(self.ptrs[1].data_ptr()), C.c_void_p(self.ptrs[2].data_ptr()),
                                   C.c_void_p(self.canon.data_ptr()), C.c_void_p(self.adj.data_ptr()), C.c_int(int(disable_interp)), aabb,
                                   C.c_uint32(extra_dims), C.c_int(int(use_geo)), self._s())
        assert rc == 0, rc

    def warp_random_samples(self, d_xyz, disable_interp=False):
        n = d_xyz.numel() // 3
        rc = self.lib.refcuda_warp_random_samples(C.c_void_p(d_xyz.data_ptr()), C.c_uint32(n), C.c_uint32(self.n_meshes), C.c_void_p(self.ptrs[0].data_ptr()),
                                                  C.c_void_p(self.ptrs[1].data_ptr()), C.c_void_p(self.ptrs[2].data_ptr()), C.c_void_p(self.canon.data_ptr()),
                                                  C.c_void_p(self.adj.data_ptr()), C.c_int(int(disable_interp)), self._s())
        assert rc == 0, rc

    def grid_forward(self, d_pos, stride, d_table, d_out, n_levels=16, log2_hashmap_size=18, base_resolution=16, per_level_scale=1.3819128274917603):
        n = d_pos.numel() // stride
        rc = self.lib.refcuda_grid_forward(C.c_uint32(n), C.c_void_p(d_pos.data_ptr()), C.c_uint32(stride), C.c_void_p(d_table.data_ptr()),
                                           C.c_void_p(d_out.data_ptr()), C.c_uint32(n_levels), C.c_uint32(log2_hashmap_size), C.c_uint32(base_resolution),
                                           C.c_float(per_level_scale), self._s())
        assert rc == 0, rc

    def grid_backward(self, d_pos, stride, d_dy, d_grad, n_levels=16, log2_hashmap_size=18, base_resolution=16, per_level_scale=1.3819128274917603):
        n = d_pos.numel() // stride
        rc = self.lib.refcuda_grid_backward(C.c_uint32(n), C.c_void_p(d_pos.data_ptr()), C.c_uint32(stride), C.c_void_p(d_dy.data_ptr()),
                                            C.c_void_p(d_grad.data_ptr()), C.c_uint32(n_levels), C.c_uint32(log2_hashmap_size), C.c_uint32(base_resolution),
                                            C.c_float(per_level_scale), self._s())
        assert rc == 0, rc

    def closest_triangle(self, d_pts, frame):
        """reference closest_triangle (bvh.cuh:48) on the device for one uploaded frame -> (bvh-order idx, distance)."""
        import torch

        n = d_pts.shape[0]
        idx = torch.empty(n, dtype=torch.int32, device="cuda")
        dist = torch.empty(n, dtype=torch.float32, device="cuda")
        rc = self.lib.refcuda_closest_triangle(C.c_uint32(n), C.c_void_p(d_pts.data_ptr()), C.c_void_p(self._keep[0][frame].data_ptr()),
                                               C.c_void_p(self._keep[1][frame].data_ptr()), C.c_void_p(idx.data_ptr()), C.c_void_p(dist.data_ptr()), self._s())
        assert rc == 0, rc
        return idx, dist

    def distances(self, d_pts, frame, tri_index, node_index):
        import torch

        n = d_pts.shape[0]
        a = torch.empty(n, dtype=torch.float32, device="cuda")
        b = torch.empty(n, dtype=torch.float32, device="cuda")
        rc = self.lib.refcuda_distances(C.c_uint32(n), C.c_void_p(d_pts.data_ptr()), C.c_void_p(self._keep[1][frame].data_ptr() + 48 * tri_index),
                                        C.c_void_p(self._keep[0][frame].data_ptr() + 32 * node_index), C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), self._s())
        assert rc == 0, rc
        return a, b
