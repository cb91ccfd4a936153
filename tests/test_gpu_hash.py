"""Stage 2 on the GPU, through the C ABI, against the oracle: slot indices and per-level resolutions bit-exact,
fp16 encodings bit-exact (same fp32 multiply / fp16 accumulate order), gradients within an fp16 accumulation bound."""
import numpy as np
import pytest

from insta_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def layout(gpu_ctx, oracle):
    offs, scales, res = gpu_ctx.grid_layout()
    o2, s2, r2 = oracle.grid_layout()
    assert (offs == o2).all() and (res == r2).all()
    assert gpu_ctx.grid_n_params == 25716288
    # the device-evaluated scales are what both sides use from here on; they must agree with the host formula to 1 ulp
    assert np.abs(scales.view(np.int32) - s2.view(np.int32)).max() <= 1
    return offs, scales, res


def positions(n, seed, stride=3):
    rng = np.random.default_rng(seed)
    p = rng.random((n, stride), dtype=np.float32)
    k = min(n, 8)
    p[:k, :3] = np.array([[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0.999999, 0.5, 0.25], [1e-7, 1e-7, 1e-7]], np.float32)[:k]
    return p


@pytest.mark.parametrize("n,stride", [(20000, 3), (4097, 27), (1, 3), (2, 27)])
def test_slots_bit_exact(gpu_ctx, oracle, layout, n, stride):
    import torch

    pos = positions(n, 1, stride)
    d = torch.from_numpy(pos).cuda()
    slots = torch.empty((n, 16, 8), dtype=torch.int32, device="cuda")
    gpu_ctx.hash_slots(d, slots)
    want = oracle.grid_slots(pos, stride, layout)
    assert (slots.cpu().numpy().view(np.uint32) == want).all()


@pytest.mark.parametrize("n,stride", [(20000, 3), (4097, 27), (1, 3), (7, 27), (4096, 3)])
def test_forward_bit_exact(gpu_ctx, oracle, layout, n, stride):
    import torch

    pos = positions(n, 2, stride)
    table = synth.make_grid_table(int(layout[0][-1]), seed=1337)
    d_table = torch.from_numpy(table).cuda()
    out = torch.full((128, n), float("nan"), dtype=torch.float16, device="cuda")
    gpu_ctx.hash_encode_fwd(torch.from_numpy(pos).cuda(), d_table, out)
    want = oracle.grid_forward(pos, stride, layout, table)
    got = out.cpu().numpy()
    assert (got.view(np.uint16) == want.view(np.uint16)).all()


def test_forward_tcnn_init_range(gpu_ctx, oracle, layout):
    """tcnn's own initialisation U(-1e-4, 1e-4) (grid.h:1275-1280): products land in the fp16 subnormal range."""
    import torch

    n = 5000
    pos = positions(n, 3)
    table = synth.make_grid_table(int(layout[0][-1]), seed=7, scale=1e-4)
    out = torch.empty((128, n), dtype=torch.float16, device="cuda")
    gpu_ctx.hash_encode_fwd(torch.from_numpy(pos).cuda(), torch.from_numpy(table).cuda(), out)
    want = oracle.grid_forward(pos, 3, layout, table)
    assert (out.cpu().numpy().view(np.uint16) == want.view(np.uint16)).all()


def test_backward_against_exact_sum(gpu_ctx, oracle, layout):
    import torch

    n = 6000
    pos = positions(n, 4)
    rng = np.random.default_rng(8)
    dy = (rng.standard_normal((128, n)) * 0.02).astype(np.float16)
    dy[:, ::5] = 0                                             # zero-gradient (padding) samples contribute nothing
    grad = torch.full((int(layout[0][-1]), 8), 3.0, dtype=torch.float16, device="cuda")
    gpu_ctx.hash_encode_bwd(torch.from_numpy(pos).cuda(), torch.from_numpy(dy).cuda(), grad, zero_first=True)
    want, hits = oracle.grid_backward(pos, 3, layout, dy, want_hits=True)
    got = grad.cpu().numpy().astype(np.float64)
    assert (got[hits == 0] == 0).all()                          # zeroed first (EGradientMode::Overwrite)
    # fp16 accumulation in arbitrary order: |err| <= hits * ulp16(max partial sum) -- statistical bound with slack
    mag = np.maximum(np.abs(want), 1e-3)
    bound = (np.maximum(hits, 1)[:, None] ** 0.5) * mag * 2.0 ** -10 * 4 + 1e-6
    assert (np.abs(got - want) <= bound).mean() > 0.999
    assert np.abs(got - want).max() < 0.05 * np.abs(want).max()
    # accumulate mode: a second pass without zeroing doubles the table
    gpu_ctx.hash_encode_bwd(torch.from_numpy(pos).cuda(), torch.from_numpy(dy).cuda(), grad, zero_first=False)
    got2 = grad.cpu().numpy().astype(np.float64)
    big = np.abs(want) > 0.05
    assert np.abs(got2[big] / want[big] - 2).max() < 0.05


@pytest.mark.parametrize("n", [1, 3, 63, 65, 4097, 4099])
def test_backward_ragged_sizes(gpu_ctx, oracle, layout, n):
    """The scatter kernel's lanes take four consecutive samples each and read dL/dy eight bytes at a time when the row pitch allows it: sizes that are not a multiple
    of 4 (scalar loads), a single sample, one sample past a 64-sample tile.  Against the exact fp64 sum like the test above."""
    import torch

    pos = positions(n, 3)
    rng = np.random.default_rng(80 + n)
    dy = (rng.standard_normal((128, n)) * 0.02).astype(np.float16)
    grad = torch.full((int(layout[0][-1]), 8), -1.0, dtype=torch.float16, device="cuda")
    gpu_ctx.hash_encode_bwd(torch.from_numpy(pos).cuda(), torch.from_numpy(dy).cuda(), grad, zero_first=True)
    want, hits = oracle.grid_backward(pos, 3, layout, dy, want_hits=True)
    got = grad.cpu().numpy().astype(np.float64)
    assert (got[hits == 0] == 0).all() and (np.abs(got[hits > 0]).sum(axis=1) > 0).mean() > 0.9  # the touched entries, and only those (a tiny product may underflow fp16)
    mag = np.maximum(np.abs(want), 1e-3)
    bound = (np.maximum(hits, 1)[:, None] ** 0.5) * mag * 2.0 ** -10 * 4 + 1e-6
    assert (np.abs(got - want) <= bound).mean() > 0.999
    assert np.abs(got - want).max() < 0.05 * max(np.abs(want).max(), 1e-3)


def test_scatter_microbenchmark_runs(gpu_ctx):
    """insta_microbench_scatter (the reduction twin of the gather roofline): a positive figure, argument checks."""
    g16 = gpu_ctx.microbench_scatter(1 << 22, 16, iters=1)
    g32 = gpu_ctx.microbench_scatter(1 << 22, 32, iters=1)
    assert g16 > 10 and g32 > 10
    with pytest.raises(Exception):
        gpu_ctx.microbench_scatter(1 << 22, 24)


def test_full_size_forward_properties(gpu_ctx, oracle, layout):
    """4,194,304 samples: subsample against the oracle + two size-independent properties (a sample's encoding does not
    depend on its neighbours in the batch; a table of constant c encodes to ~c everywhere)."""
    import torch

    n = 1 << 22
    rng = np.random.default_rng(12)
    pos = rng.random((n, 3), dtype=np.float32)
    table = synth.make_grid_table(int(layout[0][-1]), seed=1337)
    d_pos = torch.from_numpy(pos).cuda()
    d_table = torch.from_numpy(table).cuda()
    out = torch.empty((128, n), dtype=torch.float16, device="cuda")
    gpu_ctx.hash_encode_fwd(d_pos, d_table, out)
    sub = np.arange(0, n, 419)
    want = oracle.grid_forward(pos[sub], 3, layout, table)
    assert (out[:, torch.from_numpy(sub).cuda()].cpu().numpy().view(np.uint16) == want.view(np.uint16)).all()
    # batch independence: encode a permuted batch, un-permute, compare bits
    perm = torch.randperm(n, device="cuda")
    out2 = torch.empty_like(out)
    gpu_ctx.hash_encode_fwd(d_pos[perm].contiguous(), d_table, out2)
    assert torch.equal(out2, out[:, perm])
    # partition of unity
    const = torch.full_like(d_table, 0.75)
    gpu_ctx.hash_encode_fwd(d_pos, const, out)
    assert float((out.float() - 0.75).abs().max()) < 4e-3


def test_host_buffer_path_matches_device_path(gpu_ctx, scene, layout):
    import torch

    n = (1 << 19) + 12345                                       # two chunks, ragged
    fids = synth.train_regime_frame_ids(n, 2)
    rec = synth.make_records(n, scene.aabb, fids, seed=21, pad_fraction=0.01)
    table = torch.from_numpy(synth.make_grid_table(int(layout[0][-1]), seed=1337)).cuda()
    h_rec = torch.from_numpy(rec.copy()).pin_memory()
    h_out = torch.empty((128, n), dtype=torch.float16).pin_memory()
    gpu_ctx.deform_encode_host(h_rec, -1, table, h_out)
    d = torch.from_numpy(rec.copy()).cuda()
    gpu_ctx.surface_closest_point(d)
    out = torch.empty((128, n), dtype=torch.float16, device="cuda")
    gpu_ctx.hash_encode_fwd(d, table, out)
    torch.cuda.synchronize()
    assert torch.equal(h_rec, d.cpu())
    assert torch.equal(h_out, out.cpu())


def test_host_buffer_inference_path_matches_device_path(gpu_ctx, scene):
    """insta_deform_infer_host (records in, r g b sigma out; nothing else crosses PCIe) = insta_deform_samples + insta_network_inference on the device"""
    import torch

    n = (1 << 19) + 4321                                        # two chunks, ragged
    fids = synth.train_regime_frame_ids(n, 2)
    rec = synth.make_records(n, scene.aabb, fids, seed=22, pad_fraction=0.01)
    params = torch.from_numpy(synth.make_network_params(gpu_ctx.network_n_params, seed=1337)).cuda()
    h_rec = torch.from_numpy(rec.copy()).pin_memory()
    h_out = torch.zeros((n, 4), dtype=torch.float16).pin_memory()
    gpu_ctx.deform_infer_host(h_rec, -1, params, h_out)
    assert torch.equal(h_rec, torch.from_numpy(rec))             # read only
    d = torch.from_numpy(rec.copy()).cuda()
    gpu_ctx.surface_closest_point(d)
    out = torch.empty((n, 4), dtype=torch.float16, device="cuda")
    gpu_ctx.network_inference(d, params, out)
    torch.cuda.synchronize()
    assert torch.equal(h_out.view(torch.int16), out.cpu().view(torch.int16))
