"""Hash-grid oracle self-checks on the CPU: ctor layout numbers quoted in SURVEY.md, fp16 helpers, slot arithmetic and
forward/backward consistency.  (The reference has no host implementation of kernel_grid; the device kernels are pinned
against the reference's CUDA kernels on the GPU box, tests/test_gpu_vs_ref_cuda.py.)"""
import numpy as np

from insta_b200 import synth


def test_layout_numbers(oracle):
    offs, scales, res = oracle.grid_layout()
    sizes = np.diff(offs.astype(np.int64))
    assert sizes[:5].tolist() == [4096, 12168, 29792, 79512, 205384]      # dense levels, SURVEY.md hard part 3
    assert (sizes[5:] == 262144).all()
    assert int(offs[-1]) == 3214536 and int(offs[-1]) * 8 == 25716288     # n_params of the shipped config
    assert res[0] == 16 and scales[0] == 15.0
    offs22, _, res22 = oracle.grid_layout(log2_hashmap_size=22)
    assert int(offs22[-1]) == 39474024                                      # stress config 5


def test_f16_roundtrip_and_rounding(oracle):
    L = oracle.lib
    allh = np.arange(65536, dtype=np.uint16)
    f = allh.view(np.float16).astype(np.float32)
    ok = ~np.isnan(f)
    mine = np.array([L.orc_f16_to_f32(int(h)) for h in allh[::7]], np.float32)
    ref = f[::7]
    assert (mine.view(np.uint32)[ok[::7]] == ref.view(np.uint32)[ok[::7]]).all()
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(20000).astype(np.float32) * s for s in (1e-8, 1e-5, 1e-3, 1.0, 300.0, 7e4)])
    x = np.concatenate([x, np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, 2.0 ** -24, 2.0 ** -25, 1.5 * 2.0 ** -25, 2.0 ** -14], np.float32)])
    mine = np.array([L.orc_f32_to_f16(float(v)) for v in x], np.uint16)
    assert (mine == x.astype(np.float16).view(np.uint16)).all()


def test_slots_weights_and_forward(oracle):
    layout = oracle.grid_layout()
    offs, scales, res = layout
    rng = np.random.default_rng(5)
    n = 3000
    pos = rng.random((n, 3), dtype=np.float32)
    pos[:8] = [[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0.999999, 0.5, 0.25], [1e-7, 1e-7, 1e-7]]
    slots, w = oracle.grid_slots(pos, 3, layout, want_weights=True)
    for l in range(16):
        assert slots[:, l].min() >= offs[l] and slots[:, l].max() < offs[l + 1]
    assert np.abs(w.sum(axis=2) - 1).max() < 1e-5
    # independent numpy evaluation of the index rule for one dense and one hashed level
    for l in (2, 9):
        p = np.float32(pos) * scales[l] + np.float32(0.5)   # numpy has no fma; compare only where fma == mul+add
        g = np.floor(p).astype(np.int64)
        hs = int(offs[l + 1] - offs[l])
        c0 = (g[:, 0] + g[:, 1] * int(res[l]) + g[:, 2] * int(res[l]) ** 2) % hs if l == 2 else \
            ((g[:, 0].astype(np.uint32)) ^ (g[:, 1].astype(np.uint32) * np.uint32(2654435761)) ^ (g[:, 2].astype(np.uint32) * np.uint32(805459861))) % np.uint32(hs)
        agree = (slots[:, l, 0] - offs[l]) == c0
        assert agree.mean() > 0.999
    table = synth.make_grid_table(int(offs[-1]), seed=1)
    out = oracle.grid_forward(pos, 3, layout, table)
    assert out.shape == (128, n)
    # fp32 trilinear interpolation agrees with the fp16-accumulated result to a few fp16 ulps
    ref = (w[:, :, :, None] * table[slots].astype(np.float32)).sum(axis=2)      # [n, L, 8]
    got = out.astype(np.float32).reshape(16, 8, n).transpose(2, 0, 1)
    assert np.abs(got - ref).max() < 8 * 2.0 ** -11


def test_backward_is_adjoint_of_forward(oracle):
    layout = oracle.grid_layout()
    offs = layout[0]
    rng = np.random.default_rng(6)
    n = 500
    pos = rng.random((n, 3), dtype=np.float32)
    slots, w = oracle.grid_slots(pos, 3, layout, want_weights=True)
    dy = (rng.standard_normal((128, n)) * 0.01).astype(np.float16)
    grad, hits = oracle.grid_backward(pos, 3, layout, dy, want_hits=True)
    assert hits.sum() == n * 16 * 8
    ref = np.zeros((int(offs[-1]), 8))
    d = dy.astype(np.float64).reshape(16, 8, n).transpose(2, 0, 1)            # [n, L, 8]
    np.add.at(ref, slots.reshape(-1), (w[:, :, :, None] * d[:, :, None, :]).reshape(-1, 8))
    touched = hits > 0
    assert np.abs(grad[touched] - ref[touched]).max() < 1e-4
    assert (grad[~touched] == 0).all()
