"""Worker of tests/test_gpu_nccl.py: one rank of a 2-rank insta_train_nerf step (frozen parameters, fixed ray count, its own RNG seed); writes the gradient vector
the step leaves on the rank (after the NCCL sum) to OUT_DIR/grads_rank<r>.npy.  With WORLD = 1 the same step without a communicator, seed from SEED."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

N_IMAGES, RES, FOCAL, TARGET, N_RAYS = 6, (96, 80), 260.0, 1 << 16, 512


def run(rank, world, seed, out_path, uid_path=None):
    import torch

    from common import Scene
    from insta_b200 import capi, synth
    from oracle.pyoracle import Oracle

    torch.cuda.set_device(rank if world > 1 else 0)
    capi.load_library(build_if_missing=False)
    scene = Scene(Oracle(), frames=())
    ctx = capi.Context(device=torch.cuda.current_device())
    ctx.set_canonical(scene.verts, scene.faces, scene.masks)
    verts = [scene.synth.frame(f) for f in range(N_IMAGES)]
    ctx.add_frames(np.stack(verts))
    ims = synth.make_training_images(N_IMAGES, RES, FOCAL, seed=3)
    dev = []
    for im in ims:
        d = dict(im)
        for k in ("pixels", "seg", "depth", "rays"):
            d[k] = torch.from_numpy(np.ascontiguousarray(im[k])).cuda()
        dev.append(d)
    ctx.training_set_images(dev)
    bf = torch.from_numpy(synth.occupancy_bitfield(verts, radius=0.03)).cuda()
    extra = torch.randn((N_IMAGES, 16), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    n = ctx.network_n_params
    p16 = torch.from_numpy(synth.make_network_params(n, seed=11, grid_scale=0.1)).cuda()
    p32, grads = p16.float(), torch.zeros(n, dtype=torch.float16, device="cuda")
    m1, m2, steps, ema = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda"), torch.zeros(n, dtype=torch.int32, device="cuda"), p16.clone()
    opt = ctx.make_optimizer(0, p32, p16, grads, m1, m2, steps, ema)
    if world > 1:
        if rank == 0:
            uid = ctx.comm_unique_id()
            with open(uid_path + ".tmp", "wb") as f:
                f.write(uid)
            os.replace(uid_path + ".tmp", uid_path)
        else:
            import time

            for _ in range(600):
                if os.path.exists(uid_path):
                    break
                time.sleep(0.1)
            uid = open(uid_path, "rb").read()
        ctx.comm_init_rank(world, rank, uid)
    ctx.trainer_create(target_batch_size=TARGET, rays_per_batch=N_RAYS, fixed_rays_per_batch=N_RAYS, freeze_parameters=1, depth_weight=1.25, seg_weight=40.0, rng_seed=seed)
    st = ctx.train_nerf(opt, bf, extra, torch.tensor([0.02], device="cuda"), get_loss_scalar=True)
    torch.cuda.synchronize()
    assert torch.equal(p16, ema) and torch.equal(p32, p16.float()), "frozen parameters moved"
    np.save(out_path, grads.cpu().numpy())
    mode = ctx.comm_exchange_mode
    if world > 1:
        want = os.environ.get("EXPECT_EXCHANGE_MODE")
        assert want is None or mode == int(want), "exchange mode %d, expected %s" % (mode, want)
        ctx.comm_destroy()
    ctx.close()
    return st


if __name__ == "__main__":
    rank, world, seed, out_dir = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    st = run(rank, world, seed, os.path.join(out_dir, "grads_rank%d_world%d_seed%d.npy" % (rank, world, seed)), os.path.join(out_dir, "nccl_uid"))
    print("rank", rank, st, flush=True)
