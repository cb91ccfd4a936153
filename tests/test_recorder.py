"""recorder read-back (rta::Recorder::dump_frame_buffer, src/recorder.cu:299-359).  Host part (PNG / depth .bin writers) on CPU;
the GPU conversion against the numpy restatement of save_rgba / save_depth (oracle/pyoracle.py), bit-exact for the default (identity) transform."""
import struct
import zlib

import numpy as np
import pytest


def _decode_png(path):
    b = open(path, "rb").read()
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(b):
        n, typ = struct.unpack(">I4s", b[pos:pos + 8])
        data = b[pos + 8:pos + 8 + n]
        (crc,) = struct.unpack(">I", b[pos + 8 + n:pos + 12 + n])
        assert crc == zlib.crc32(typ + data) & 0xffffffff, typ
        if typ == b"IHDR":
            w, h, depth, ctype, comp, flt, il = struct.unpack(">IIBBBBB", data)
            assert (depth, ctype, comp, flt, il) == (8, 6, 0, 0, 0)
        elif typ == b"IDAT":
            idat += data
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 4 * w)
    assert (raw[:, 0] == 0).all()  # filter type 0 on every row
    return raw[:, 1:].reshape(h, w, 4)


@pytest.mark.parametrize("shape", [(1, 1), (7, 5), (300, 200)])  # the last one needs several stored blocks
def test_png_writer_round_trip(tmp_path, shape):
    from insta_b200 import capi

    h, w = shape
    img = np.random.default_rng(h * w).integers(0, 256, (h, w, 4), dtype=np.uint8)
    p = tmp_path / "f.png"
    capi.host_write_png(p, img)
    assert np.array_equal(_decode_png(p), img)


def test_depth_bin_writer(tmp_path):
    from insta_b200 import capi

    d = np.random.default_rng(1).integers(0, 5000, 1234, dtype=np.uint32)
    p = tmp_path / "d.bin"
    capi.host_write_depth_bin(p, d)
    assert np.array_equal(np.fromfile(p, "<u4"), d)


def test_oracle_recorder_conversions():
    from oracle import pyoracle

    rgba = np.array([[0.0, 0.5, 1.0, 2.0], [-1.0, 0.999, 1 / 255, np.nan]], np.float32)
    assert pyoracle.recorder_rgba8(rgba).tolist() == [[0, 127, 255, 255], [0, 254, 1, 0]]
    depth = np.array([9, 9, 9, 0.5, -1.0, 5.5, 1.2345, 5.0, 0.0], np.float32)  # 3 x 3: row 0 is never written
    assert pyoracle.recorder_depth_mm(depth, 3, 3).tolist() == [500, 0, 0, 1234, 5000, 0]


@pytest.mark.gpu
@pytest.mark.parametrize("res", [(64, 48), (517, 3), (5, 1)])
def test_recorder_readback_matches_oracle(gpu_ctx, res):
    import torch

    from oracle import pyoracle

    w, h = res
    g = torch.Generator(device="cuda").manual_seed(w * h)
    rgba = torch.rand((w * h, 4), device="cuda", generator=g) * 1.4 - 0.2  # values below 0 and above 1
    depth = torch.rand(w * h, device="cuda", generator=g) * 7.0 - 1.0     # values below 0 and above 5
    rgba[0, 0] = float("nan")
    h8 = torch.zeros((h, w, 4), dtype=torch.uint8).pin_memory()
    hd = torch.full((max(h - 1, 0), w), -1, dtype=torch.int32).pin_memory()
    gpu_ctx.recorder_readback(rgba, depth, w, h, h8, hd)
    torch.cuda.synchronize()
    assert np.array_equal(h8.numpy().reshape(-1, 4), pyoracle.recorder_rgba8(rgba.cpu().numpy()))
    assert np.array_equal(hd.numpy().reshape(-1).view(np.uint32), pyoracle.recorder_depth_mm(depth.cpu().numpy(), w, h))
    # Normals mode (srgb_to_linear): powf on the device against numpy's -> at most one 8-bit step
    gpu_ctx.recorder_readback(rgba, None, w, h, h8, None, srgb_to_linear=True)
    torch.cuda.synchronize()
    ref = pyoracle.recorder_rgba8(rgba.cpu().numpy(), srgb_to_linear=True).astype(np.int32)
    assert np.abs(h8.numpy().reshape(-1, 4).astype(np.int32) - ref).max() <= 1


@pytest.mark.gpu
def test_rendered_frame_to_png(gpu_ctx, scene, tmp_path):
    """render -> read back -> PNG: the decoded pixels are the oracle's conversion of the frame buffer."""
    import torch

    from insta_b200 import capi, synth
    from oracle import pyoracle
    from test_gpu_nerf import FOCAL, RES, _bitfield

    n = RES[0] * RES[1]
    pn = synth.make_network_params(gpu_ctx.network_n_params, seed=11)
    pn[13312:13312 + 64] = np.float16(0.6)
    fr, fd = torch.zeros((n, 4), device="cuda"), torch.zeros(n, device="cuda")
    gpu_ctx.render_frame(capi.make_camera(synth.random_camera(5), FOCAL, RES), 1, torch.from_numpy(pn).cuda(), fr, fd, bitfield=_bitfield(scene),
                         extra_dims=torch.zeros(16, device="cuda"))
    h8 = torch.zeros((RES[1], RES[0], 4), dtype=torch.uint8).pin_memory()
    hd = torch.zeros((RES[1] - 1, RES[0]), dtype=torch.int32).pin_memory()
    gpu_ctx.recorder_readback(fr, fd, RES[0], RES[1], h8, hd)
    torch.cuda.synchronize()
    capi.host_write_png(tmp_path / "00000.png", h8.numpy())
    capi.host_write_depth_bin(tmp_path / "00000.bin", hd.numpy().view(np.uint32))
    img = _decode_png(tmp_path / "00000.png")
    assert np.array_equal(img.reshape(-1, 4), pyoracle.recorder_rgba8(fr.cpu().numpy()))
    assert int((img[..., 3] > 50).sum()) > 500
    assert np.array_equal(np.fromfile(tmp_path / "00000.bin", "<u4"), pyoracle.recorder_depth_mm(fd.cpu().numpy(), RES[0], RES[1]))


def test_trajectory_json_equals_the_reference_writer(tmp_path):
    """insta_host_write_trajectory_json against Recorder::dump_cameras_json restated over the reference's vendored nlohmann::json (oracle/ref_json.cpp):
    the same text, byte for byte -- key order, 4-space indentation, shortest round-trip floats, frames sorted by mesh path."""
    import json

    from insta_b200 import capi
    from oracle import pyrefjson

    r = np.random.default_rng(4)
    n = 7
    cams = r.standard_normal((n, 3, 4)).astype(np.float32)
    cams[0] = np.array([[1, 0, 0, 0.5], [0, 1, 0, 0.75], [0, 0, 1, 2.0]], np.float32)  # exactly representable values print as x.0 / 0.25
    cams[1, 0, 0] = np.float32(1e-7)   # exponent notation
    cams[1, 0, 1] = np.float32(123456789.0)
    ids = np.array([12, 3, 100000, 7, 0, 99999, 41], np.uint32)  # sorted by mesh path on output, not by recording order
    path = tmp_path / "transforms_test.json"
    capi.host_write_trajectory_json(path, 512, 512, (1200.0, 1199.5), (256.0, 255.25), cams, ids)
    text = path.read_text()
    doc = json.loads(text)
    assert doc["w"] == 512 and doc["h"] == 512 and doc["synthetic"] is True and doc["fl_y"] == 1199.5 and abs(doc["integer_depth_scale"] - 0.001) < 1e-9
    assert [f["mesh_path"] for f in doc["frames"]] == sorted("meshes/%05d.obj" % i for i in ids)
    f0 = next(f for f in doc["frames"] if f["mesh_path"] == "meshes/00012.obj")
    assert f0["transform_matrix"] == [[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.25], [0.0, 0.0, 1.0, 1.5], [0.0, 0.0, 0.0, 1.0]]
    assert f0["exp_path"] == "flame/exp/00012.txt" and f0["eyes_path"] == "flame/eyes/00012.txt" and f0["depth_path"] == ""
    if not pyrefjson.available():
        pytest.skip("oracle/_ref/libref_json.so not built (make -C oracle refjson; needs /root/reference)")
    want = pyrefjson.RefJson().dump_cameras(512, 512, (1200.0, 1199.5), (256.0, 255.25), cams, ids)
    if text != want:
        k = next(i for i in range(min(len(text), len(want))) if text[i] != want[i])
        pytest.fail("trajectory text differs from the reference writer at offset %d: %r vs %r" % (k, text[max(0, k - 60):k + 30], want[max(0, k - 60):k + 30]))
    # no frames at all
    capi.host_write_trajectory_json(path, 64, 48, (10.0, 10.0), (32.0, 24.0), np.zeros((0, 3, 4), np.float32), [])
    assert path.read_text() == pyrefjson.RefJson().dump_cameras(64, 48, (10.0, 10.0), (32.0, 24.0), np.zeros((0, 12), np.float32), np.zeros(0, np.uint32))
