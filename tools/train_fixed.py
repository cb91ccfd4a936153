"""configs[3] controlled leg alone (torchrun): fixed 2^18 rays, frozen parameters, exchange modes.  usage: torchrun ... tools/train_fixed.py"""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from insta_b200 import capi
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
class A: pass
a = A(); a.train_images = 128; a.train_steps = 20; a.train_warmup = int(os.environ.get("WARMUP", 70))
verts, faces, masks, fv = bench.build_scene(128 * world, rank, world)
ctx = capi.Context(device=local); ctx.set_canonical(verts, faces, masks); ctx.add_frames(fv)
r = bench.train_leg(ctx, verts, fv, a, rank, world)
if rank == 0:
    print(json.dumps({k: r[k] for k in ("ms_per_step", "rays_per_step", "grad_allreduce_ms", "fixed_262144_rays")}, indent=1))
ctx.close()
if world > 1: dist.destroy_process_group()
