"""Host halves of the product's mesh ingest (C ABI, no GPU) against the oracle: triangles, BVH topology + leaf order,
adjacency, masks, AABB, OBJ parsing."""
import os

import numpy as np
import pytest

from insta_b200 import capi


def test_triangles_and_bvh_identical_to_oracle(oracle, scene):
    hc = capi.host_build_triangles(scene.verts, scene.faces, scene.masks)
    assert hc.tobytes() == scene.canon.tobytes()
    for k, v in enumerate(scene.frame_verts):
        t = capi.host_build_triangles(v, scene.faces, scene.masks, hc)
        assert t.tobytes() == scene.ttopo[k].tobytes()
        nodes, order = capi.host_bvh_build(t, 2)
        assert nodes.tobytes() == scene.nodes[k].tobytes()
        assert order.tobytes() == scene.tbvh[k].tobytes()


@pytest.mark.parametrize("leaf", [1, 2, 4])
def test_bvh_leaf_sizes_and_tiny_meshes(oracle, scene, leaf):
    for n in (1, 2, 3, 5, 17, 100):
        t = scene.canon[:n].copy()
        a, ta = capi.host_bvh_build(t, leaf)
        b, tb = oracle.bvh_build(t, leaf)
        assert a.tobytes() == b.tobytes() and ta.tobytes() == tb.tobytes()


def test_adjacency_and_aabb(oracle, scene):
    assert (capi.host_adjacency(scene.faces, len(scene.verts)) == scene.adj).all()
    assert (capi.host_mesh_aabb(scene.verts) == oracle.mesh_aabb(scene.verts)).all()
    # boundary triangles have fewer than 3 edge neighbours
    assert (scene.adj < 0).sum() == 30


def test_obj_loader(tmp_path, scene):
    p = tmp_path / "canonical.obj"
    with open(p, "w") as f:
        f.write("# test\n")
        for v in scene.verts[:50] - np.float32(0.5):
            f.write("v %.9g %.9g %.9g\n" % tuple(v))
        f.write("f 1 2 3\nf 2/1/1 3/2/2 4/3/3\nf -1 -2 -3\nf 5 6 7 8\n")
    v, fc, m = capi.host_load_canonical_obj(str(p), None)
    assert v.shape == (50, 3) and np.abs(v - scene.verts[:50]).max() < 1e-7
    assert fc.tolist() == [[0, 1, 2], [1, 2, 3], [49, 48, 47], [4, 5, 6], [4, 6, 7]]
    assert (m == 15).all()
    with pytest.raises(capi.InstaError) as e:
        capi.host_load_canonical_obj(str(tmp_path / "missing.obj"), None)
    assert e.value.code == -3


@pytest.mark.skipif(not os.path.exists("/root/reference/mesh_masks/canonical.obj"), reason="reference assets not present")
def test_reference_assets(scene):
    v, f, m = capi.host_load_canonical_obj("/root/reference/mesh_masks/canonical.obj", "/root/reference/mesh_masks")
    assert (v == scene.verts).all() and (f == scene.faces).all() and (m == scene.masks).all()
