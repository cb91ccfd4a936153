"""Shared scene construction for the tests: canonical mesh + synthetic frames, prepared with the oracle."""
import numpy as np

from insta_b200 import synth

UNIT_AABB = np.array([0, 0, 0, 1, 1, 1], np.float32)


class Scene:
    def __init__(self, oracle, frames=(0, 7)):
        self.verts, self.faces, regions = synth.load_canonical()
        self.masks = synth.mask_ids_from_regions(self.faces, regions)
        self.synth = synth.FrameSynth(self.verts, self.faces)
        self.frame_verts = [self.synth.frame(f) for f in frames]
        self.aabb = synth.mesh_aabb(self.verts)
        self.canon = oracle.build_triangles(self.verts, self.faces, self.masks)
        self.adj = oracle.adjacency(self.faces, len(self.verts))
        self.canon_nodes, self.canon_bvh = oracle.bvh_build(self.canon)
        self.nodes, self.tbvh, self.ttopo = [], [], []
        for v in self.frame_verts:
            t = oracle.build_triangles(v, self.faces, self.masks, self.canon)
            n, tb = oracle.bvh_build(t)
            self.nodes.append(n)
            self.tbvh.append(tb)
            self.ttopo.append(t)

    def golden_records(self, n=8192, seed=11):
        fids = np.repeat(np.arange(len(self.frame_verts)), n // len(self.frame_verts)).astype(np.int32)
        return synth.make_records(n, self.aabb, fids, seed=seed, pad_fraction=0.02)

    def oracle_warp(self, oracle, rec, stride=27, frame_id=-1, disable_interp=False, frame_ids=None, extra_dims=16, use_geo=False, aabb=UNIT_AABB):
        r = rec.copy()
        tri, bvh_idx = oracle.warp(r, stride, frame_id, self.nodes, self.tbvh, self.ttopo, self.canon, self.adj, aabb,
                                   disable_interp=disable_interp, extra_dims=extra_dims, use_geo=use_geo, frame_ids=frame_ids)
        return r, tri
