"""Snapshot files (insta_snapshot_save / insta_snapshot_open) against the reference's layout -- ngp::Testbed::save_snapshot
(ngp/src/testbed.cu:2916-2955) -- checked with an independent MessagePack implementation (python-msgpack): what we write parses to the
reference's keys, and a file written the way nlohmann::json::to_msgpack writes the reference's snapshot loads here."""
import numpy as np
import pytest

msgpack = pytest.importorskip("msgpack")

from insta_b200 import capi

N = 25600 + 4096 * 8


def _state(seed=0):
    r = np.random.default_rng(seed)
    return dict(params16=r.standard_normal(N).astype(np.float16), grid16=(r.random(128 ** 3) * 0.05).astype(np.float16), m1=r.standard_normal(N).astype(np.float32),
                m2=r.random(N).astype(np.float32), steps=r.integers(0, 100, N).astype(np.uint32), ema16=r.standard_normal(N).astype(np.float16))


def test_written_file_has_the_reference_layout(tmp_path):
    st = _state()
    path = tmp_path / "snap.msgpack"
    capi.snapshot_save(path, st["params16"], st["grid16"], training_step=1234, loss=0.0125, aabb=(0, 0, 0, 1, 1, 1), render_aabb=(0.39, 0.3, 0.34, 0.61, 0.63, 0.58),
                       bounding_radius=1.0, rays_per_batch=90112, measured_batch_size=255989, measured_batch_size_before_compaction=3355697,
                       optimizer=dict(current_step=1234, m1=st["m1"], m2=st["m2"], steps=st["steps"], ema16=st["ema16"]))
    doc = msgpack.unpackb(path.read_bytes(), raw=False)
    assert doc["encoding"] == {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 8, "log2_hashmap_size": 18, "base_resolution": 16}
    assert doc["network"]["n_neurons"] == 64 and doc["rgb_network"]["n_hidden_layers"] == 3 and doc["optimizer"]["nested"]["otype"] == "Adam"
    assert abs(doc["optimizer"]["nested"]["learning_rate"] - 1.75e-3) < 1e-9 and doc["optimizer"]["decay"] == pytest.approx(0.95)
    s = doc["snapshot"]
    assert s["version"] == 1 and s["density_grid_size"] == 128 and s["n_params"] == N and s["training_step"] == 1234
    assert s["params_binary"] == st["params16"].tobytes()
    assert s["density_grid_binary"] == st["grid16"].tobytes()
    assert s["aabb"] == {"min": [0.0, 0.0, 0.0], "max": [1.0, 1.0, 1.0]}
    assert np.allclose(s["render_aabb"]["min"], [0.39, 0.3, 0.34]) and np.allclose(s["render_aabb"]["max"], [0.61, 0.63, 0.58])
    assert s["nerf"] == {"aabb_scale": 1, "rgb": {"rays_per_batch": 90112, "measured_batch_size": 255989, "measured_batch_size_before_compaction": 3355697}}
    o = s["optimizer"]
    assert o["weights_ema_binary"] == st["ema16"].tobytes()
    assert o["nested"]["first_moments_binary"] == st["m1"].tobytes() and o["nested"]["second_moments_binary"] == st["m2"].tobytes()
    assert o["nested"]["param_steps_binary"] == st["steps"].tobytes() and o["nested"]["current_step"] == 1234
    assert s["loss"] == pytest.approx(0.0125)
    # keys in nlohmann's order (std::map: sorted), so the bytes equal what the reference's writer produces for the same document
    assert list(s.keys()) == sorted(s.keys()) and list(doc.keys()) == sorted(doc.keys())


def test_round_trip(tmp_path):
    st = _state(1)
    path = tmp_path / "rt.msgpack"
    capi.snapshot_save(path, st["params16"], st["grid16"], training_step=7, loss=0.5, optimizer=dict(current_step=8, m1=st["m1"], m2=st["m2"], steps=st["steps"], ema16=st["ema16"]))
    got = capi.snapshot_load(path)
    assert got["n_params"] == N and got["training_step"] == 7 and got["has_optimizer"] == 1 and got["opt_current_step"] == 8
    for k in ("params16", "m1", "m2", "steps", "ema16"):
        assert np.array_equal(got[k].view(np.uint8), st[k].view(np.uint8)), k
    assert np.array_equal(got["density_grid16"].view(np.uint16), st["grid16"].view(np.uint16))
    path2 = tmp_path / "noopt.msgpack"
    capi.snapshot_save(path2, st["params16"])
    got2 = capi.snapshot_load(path2)
    assert got2["has_optimizer"] == 0 and got2["density_grid16"] is None and np.array_equal(got2["params16"], st["params16"])


def test_loads_a_reference_written_file(tmp_path):
    """the document Testbed::save_snapshot builds (incl. keys this library does not write: nerf.dataset, INSTA's own config keys, float64 values)"""
    st = _state(2)
    doc = {"loss": {"otype": "Huber"}, "max_cached_bvh": 4000, "exp_params": 16, "seg_weight": 40, "depth_weight": 1.25,
           "optimizer": {"otype": "Ema", "decay": 0.95, "nested": {"otype": "Adam", "learning_rate": 0.00175, "beta1": 0.9, "beta2": 0.99, "epsilon": 1e-15, "l2_reg": 1e-6}},
           "encoding": {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 8, "log2_hashmap_size": 18, "base_resolution": 16},
           "snapshot": {"n_params": N, "params_binary": st["params16"].tobytes(), "version": 1, "density_grid_size": 128, "density_grid_binary": st["grid16"].tobytes(),
                        "nerf": {"aabb_scale": 1, "rgb": {"rays_per_batch": 4096, "measured_batch_size": 100, "measured_batch_size_before_compaction": 200},
                                 "dataset": {"n_images": 3, "paths": ["a.png", "b.png", "c.png"], "xforms": [[[1.0, 0.0], [0.0, 1.0]]] * 3, "is_hdr": False, "scale": 0.33}},
                        "training_step": 33000, "loss": 0.001953125, "aabb": {"min": [0.0, 0.0, 0.0], "max": [1.0, 1.0, 1.0]},
                        "render_aabb": {"min": [0.25, 0.125, 0.5], "max": [0.75, 0.875, 1.0]}, "bounding_radius": 1.0,
                        "optimizer": {"weights_ema_binary": st["ema16"].tobytes(),
                                      "nested": {"current_step": 33000, "base_learning_rate": 0.00175, "first_moments_binary": st["m1"].tobytes(),
                                                 "second_moments_binary": st["m2"].tobytes(), "param_steps_binary": st["steps"].tobytes()}}}}
    path = tmp_path / "ref.msgpack"
    path.write_bytes(msgpack.packb(doc, use_bin_type=True))
    got = capi.snapshot_load(path)
    assert got["training_step"] == 33000 and got["rays_per_batch"] == 4096 and got["measured_batch_size_before_compaction"] == 200
    assert np.array_equal(got["render_aabb"], np.array([0.25, 0.125, 0.5, 0.75, 0.875, 1.0], np.float32))
    assert got["loss"] == pytest.approx(0.001953125) and got["opt_base_learning_rate"] == pytest.approx(0.00175) and got["epsilon"] == pytest.approx(1e-15)
    assert got["has_optimizer"] == 1 and got["opt_current_step"] == 33000
    for k in ("params16", "m1", "m2", "steps", "ema16"):
        assert np.array_equal(got[k].view(np.uint8), st[k].view(np.uint8)), k


def test_rejects_garbage(tmp_path):
    bad = tmp_path / "bad.msgpack"
    bad.write_bytes(b"\\x93\\x01\\x02")
    with pytest.raises(capi.InstaError):
        capi.snapshot_load(bad)
    nosnap = tmp_path / "nosnap.msgpack"
    nosnap.write_bytes(msgpack.packb({"encoding": {"otype": "HashGrid"}}))
    with pytest.raises(capi.InstaError):
        capi.snapshot_load(nosnap)
    old = tmp_path / "old.msgpack"
    old.write_bytes(msgpack.packb({"snapshot": {"version": 0, "params_binary": b"12"}}, use_bin_type=True))
    with pytest.raises(capi.InstaError):
        capi.snapshot_load(old)
    with pytest.raises(capi.InstaError):
        capi.snapshot_load(tmp_path / "missing.msgpack")
