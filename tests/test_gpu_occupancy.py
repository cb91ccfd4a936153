"""GPU parity of the occupancy-grid maintenance (insta_update_density_grid / insta_training_prep_nerf) against the REFERENCE's own CUDA
kernels (oracle/_ref/libref_nerf.so: mark_untrained_density_grid, generate_grid_samples_nerf_nonuniform,
splat_grid_samples_nerf_max_nearest_neighbor, ema_grid_samples_nerf, grid_to_bitfield, bitfield_max_pool; oracle/_ref/libref_cuda.so:
warp_random_samples).  The density network has no reference build here (tcnn's FullyFusedMLP needs its whole library): both pipelines take
the density logits from insta_network_density, which tests/test_gpu_network.py checks against fp32 PyTorch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CELLS = 128 ** 3
N_IMAGES, RES, FOCAL = 6, (96, 80), 260.0


@pytest.fixture(scope="module")
def refnerf():
    from oracle import pyrefnerf

    if not pyrefnerf.available():
        pytest.skip("oracle/_ref/libref_nerf.so not built (needs /root/reference at build time)")
    return pyrefnerf.RefNerf()


@pytest.fixture(scope="module")
def images():
    import torch

    from insta_b200 import synth

    ims = synth.make_training_images(N_IMAGES, RES, FOCAL, seed=3, with_rays=False)
    dev = []
    for im in ims:
        d = dict(im)
        for k in ("pixels", "seg", "depth"):
            d[k] = torch.from_numpy(np.ascontiguousarray(im[k])).cuda()
        dev.append(d)
    return dev


@pytest.fixture(scope="module")
def bare_ctx(scene):
    """context with the canonical mesh but NO frames: the update skips the deformation hook"""
    from insta_b200 import capi

    capi.load_library(build_if_missing=False)
    ctx = capi.Context(device=0)
    ctx.set_canonical(scene.verts, scene.faces, scene.masks)
    yield ctx
    ctx.close()


def _params(ctx, seed=21):
    import torch

    from insta_b200 import synth

    p = synth.make_network_params(ctx.network_n_params, seed=seed, grid_scale=0.5)
    p[13312:13312 + 64] = np.float16(0.05)  # a mildly positive density row so that logits spread around the threshold
    return torch.from_numpy(p).cuda()


def _grid_buffers():
    import torch

    return (torch.zeros(CELLS, device="cuda"), torch.zeros(8 * CELLS // 8, dtype=torch.uint8, device="cuda"), torch.zeros(1, device="cuda"))


def test_mark_untrained_and_bitfield_vs_reference(bare_ctx, refnerf, images):
    import torch

    bare_ctx.training_set_images(images)
    refnerf.train_set_images(images)
    grid, bf, mean = _grid_buffers()
    grid.fill_(0.37)
    g = bare_ctx.make_density_grid(grid, bf, mean, rng_seed=5)
    bare_ctx.update_density_grid(g, _params(bare_ctx), 0, 0)  # step 0: frustum culling, then mean + bitfield of an empty update
    rgrid, rbf, rmean = _grid_buffers()
    rgrid.fill_(0.37)
    refnerf.grid_mark_untrained(rgrid, True)
    refnerf.grid_mean_and_bitfield(rgrid, rbf, rmean, 1)
    torch.cuda.synchronize()
    a, b = grid.cpu().numpy(), rgrid.cpu().numpy()
    assert set(np.unique(b)) <= {0.0, -1.0} and (b == -1).any() and (b == 0).any()
    # visibility is a float comparison at the frustum boundary: allow a handful of cells to flip with the contraction of the dot products
    assert (a != b).mean() < 1e-4
    assert g.ema_step == 1 and g.rng_advances == 1
    assert torch.equal(bf, rbf) or (bf != rbf).float().mean() < 1e-4
    assert abs(float(mean) - float(rmean)) < 1e-7


def _reference_update(ctx, refnerf, refcuda, grid, params, n_uni, n_non, seed, advances, step, decay, aabb=(0, 0, 0, 1, 1, 1)):
    import torch

    n = n_uni + n_non
    pos = torch.zeros((n, 3), device="cuda")
    idx = torch.zeros(n, dtype=torch.int32, device="cuda")
    if n_uni:
        refnerf.grid_generate(n_uni, seed, advances, step, aabb, grid, pos, idx, 1, -0.01)
    if n_non:
        refnerf.grid_generate(n_non, seed, advances, step, aabb, grid, pos[n_uni:], idx[n_uni:], 1, 0.01)
    if refcuda is not None:
        refcuda.warp_random_samples(pos)
    logits = torch.zeros(n, dtype=torch.float16, device="cuda")
    ctx.network_density(pos, params, logits)
    tmp = torch.zeros_like(grid)
    refnerf.grid_splat_ema(idx, logits, decay, step, grid, tmp)
    return pos, idx, logits


def _seed_grid(seed=3):
    import torch

    g = torch.Generator(device="cuda").manual_seed(seed)
    grid = torch.rand(CELLS, device="cuda", generator=g) * 0.03          # around NERF_MIN_OPTICAL_THICKNESS
    grid[torch.rand(CELLS, device="cuda", generator=g) < 0.2] = -1.0      # untrained cells stay untrained
    return grid


def test_update_without_deformation_bit_exact(bare_ctx, refnerf):
    """sample generation (cell choice, pcg32 stream, positions), splat, decay/max update, mean, bitfield and its mip chain"""
    import torch

    params = _params(bare_ctx)
    grid0 = _seed_grid()
    n_uni, n_non = 1 << 18, 3 << 16
    grid, bf, mean = _grid_buffers()
    grid.copy_(grid0)
    g = bare_ctx.make_density_grid(grid, bf, mean, rng_seed=777)
    g.training_step, g.ema_step, g.rng_advances = 40, 3, 2
    bare_ctx.update_density_grid(g, params, n_uni, n_non)
    rgrid, rbf, rmean = _grid_buffers()
    rgrid.copy_(grid0)
    _reference_update(bare_ctx, refnerf, None, rgrid, params, n_uni, n_non, 777, 2, 3, 0.95)
    refnerf.grid_mean_and_bitfield(rgrid, rbf, rmean, 1)
    torch.cuda.synchronize()
    assert torch.equal(grid.view(torch.int32), rgrid.view(torch.int32)), "density grid differs from the reference kernels"
    assert (grid != grid0).float().mean() > 0.05
    assert abs(float(mean) - float(rmean)) <= 1e-5 * abs(float(rmean))  # both sum with float atomics
    # the threshold is min(0.01, mean): cells within the mean's rounding of it may flip
    assert torch.equal(bf, rbf) or float((bf != rbf).float().mean()) < 1e-5, "bitfield / mip chain differs"
    assert int(bf[: CELLS // 8].count_nonzero()) > 1000 and int(bf[CELLS // 8: 2 * CELLS // 8].count_nonzero()) > 10
    assert (g.ema_step, g.rng_advances) == (4, 3)


def test_update_with_deformation_vs_reference(gpu_ctx, scene, refnerf):
    """the full update: generate -> warp_random_samples (random frame per point) -> density -> splat -> update -> bitfield"""
    import torch

    from oracle import pyrefcuda

    if not pyrefcuda.available():
        pytest.skip("oracle/_ref/libref_cuda.so not built")
    R = pyrefcuda.RefCuda()
    R.upload_scene(scene.nodes, scene.tbvh, scene.ttopo, scene.canon, scene.adj)
    params = _params(gpu_ctx)
    grid0 = _seed_grid(9)
    n_uni, n_non = 1 << 17, 1 << 17
    grid, bf, mean = _grid_buffers()
    grid.copy_(grid0)
    g = gpu_ctx.make_density_grid(grid, bf, mean, rng_seed=4242)
    g.training_step, g.ema_step = 300, 7
    gpu_ctx.update_density_grid(g, params, n_uni, n_non)
    rgrid, rbf, rmean = _grid_buffers()
    rgrid.copy_(grid0)
    _reference_update(gpu_ctx, refnerf, R, rgrid, params, n_uni, n_non, 4242, 0, 7, 0.95)
    refnerf.grid_mean_and_bitfield(rgrid, rbf, rmean, 1)
    torch.cuda.synchronize()
    a, b = grid.cpu().numpy(), rgrid.cpu().numpy()
    assert ((a < 0) == (b < 0)).all()
    # canonical positions agree to 1e-5 (non-blended) / 1e-4 (blended); the fp16 density logit moves by at most a few ulp with them
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-6)
    assert (rel > 2e-2).mean() < 2e-3, "density grid: %.4f%% of the cells differ by more than 2%%" % (100 * (rel > 2e-2).mean())
    assert abs(float(mean) - float(rmean)) <= 1e-3 * abs(float(rmean))
    x = (bf.cpu().numpy() ^ rbf.cpu().numpy())
    assert np.unpackbits(x).mean() < 1e-4


def test_training_prep_schedule(bare_ctx, images):
    """Testbed::train's prep schedule (testbed.cu:2480): every step up to 31, then every step // 16 steps, at most every 16"""
    bare_ctx.training_set_images(images)
    params = _params(bare_ctx)
    grid, bf, mean = _grid_buffers()
    g = bare_ctx.make_density_grid(grid, bf, mean, rng_seed=1)
    did = []
    for step in (0, 1, 31, 32, 33, 34, 48, 49, 255, 256, 257, 272, 1000, 1008):
        g.training_step = step
        did.append(bare_ctx.training_prep_nerf(g, params))
    assert did == [True, True, True, True, False, True, True, False, True, True, False, True, False, True]
    assert g.ema_step == sum(did)
