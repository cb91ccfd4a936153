"""The N > 1 path on CPU: two `gloo` processes shard frames and sample batches the way bench.py / a multi-GPU caller does (insta_b200/shard.py),
each runs ITS shard through the CPU oracle (the checker stands in for the GPU kernels here), and the union must equal the single-process
result; the gradient exchange (pre-scaled sum == mean over the global batch) is checked with a gloo all_reduce."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
N_FRAMES, N_SAMPLES = 6, 6000


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _scene():
    from insta_b200 import synth
    from oracle.pyoracle import Oracle

    O = Oracle()
    verts, faces, regions = synth.load_canonical()
    masks = synth.mask_ids_from_regions(faces, regions)
    fs = synth.FrameSynth(verts, faces)
    canon = O.build_triangles(verts, faces, masks)
    adj = O.adjacency(faces, len(verts))
    return O, synth, verts, faces, masks, fs, canon, adj


def _warp(O, fs, faces, masks, canon, adj, frame_list, rec, holder):
    nodes, tbvh, ttopo = [], [], []
    for f in frame_list:
        t = O.build_triangles(fs.frame(f), faces, masks, canon)
        n, tb = O.bvh_build(t)
        nodes.append(n); tbvh.append(tb); ttopo.append(t)
    r = rec.copy()
    r[:, 4] = holder
    tri, _ = O.warp(r, 27, -1, nodes, tbvh, ttopo, canon, adj, np.array([0, 0, 0, 1, 1, 1], np.float32), n_threads=2)
    return r, tri


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from insta_b200 import shard

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        O, synth, verts, faces, masks, fs, canon, adj = _scene()
        mine = shard.frames_of_rank(N_FRAMES, rank, world)
        gids = (np.arange(N_SAMPLES) * 7919) % N_FRAMES
        rec = synth.make_records(N_SAMPLES, synth.mesh_aabb(verts), gids.astype(np.int32), seed=5)
        holder = shard.local_frame_ids(gids, rank, world)
        out, tri = _warp(O, fs, faces, masks, canon, adj, mine, rec, holder.astype(np.float32))
        # every rank deformed only the samples of its frames; the union over ranks (MAX of ids, -1 = skipped) covers every sample once
        t = torch.from_numpy(tri.astype(np.int64))
        cnt = torch.from_numpy((tri >= 0).astype(np.int64))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        pos = torch.from_numpy(np.where((holder >= 0)[:, None], out[:, :3], 0.0).astype(np.float64))
        dist.all_reduce(pos, op=dist.ReduceOp.SUM)
        # gradient exchange: each rank's gradient is scaled by 1 / R before the sum
        g_local = torch.full((1000,), float(rank + 1)) * shard.loss_scale_of_rank(128.0, world) / 128.0
        g = g_local.clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        if rank == 0:
            q.put((t.numpy(), cnt.numpy(), pos.numpy(), g.numpy(), mine))
        else:
            q.put(("frames", mine))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_cover_the_batch_exactly_once():
    import torch.multiprocessing as mp

    sys.path.insert(0, ROOT)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    main = next(g for g in got if g[0] is not None and not isinstance(g[0], str))
    other = next(g for g in got if isinstance(g[0], str))
    tri, cnt, pos, grad, frames0 = main
    assert sorted(frames0 + other[1]) == list(range(N_FRAMES)) and not set(frames0) & set(other[1])
    # single-process truth with all frames
    O, synth, verts, faces, masks, fs, canon, adj = _scene()
    gids = (np.arange(N_SAMPLES) * 7919) % N_FRAMES
    rec = synth.make_records(N_SAMPLES, synth.mesh_aabb(verts), gids.astype(np.int32), seed=5)
    want, want_tri = _warp(O, fs, faces, masks, canon, adj, list(range(N_FRAMES)), rec, gids.astype(np.float32))
    assert (cnt == (want_tri >= 0)).all(), "a sample was deformed by zero or two ranks"
    assert (tri == want_tri).all(), "closest-triangle ids of the sharded run differ from the single-process run"
    assert np.array_equal(pos.astype(np.float32)[want_tri >= 0], want[want_tri >= 0, :3])
    assert np.allclose(grad, (1 + 2) / 2.0)  # mean of the per-rank gradients


def test_shard_helpers():
    from insta_b200 import shard

    assert shard.frames_of_rank(10, 1, 4) == [1, 5, 9]
    assert shard.frames_of_rank(3, 3, 4) == []
    assert [shard.owner_of_frame(f, 4) for f in (0, 5, 11)] == [(0, 0), (1, 1), (3, 2)]
    assert shard.local_frame_ids([0, 1, 2, 3, 4, -1], 1, 2).tolist() == [-1, 0, -1, 1, -1, -1]
    assert shard.render_schedule([4, 1, 2, 7], 2) == [[4, 2], [1, 7]]
    assert shard.loss_scale_of_rank(128.0, 8) == 16.0
    with pytest.raises(ValueError):
        shard.frames_of_rank(4, 2, 2)


def test_c_abi_sharding_rule_equals_the_python_helper():
    """insta_shard_* (for C / C++ callers) states the same rule as insta_b200/shard.py; host only, no device needed"""
    import ctypes as C

    import numpy as np

    from insta_b200 import capi, shard

    lib = capi.load_library()
    lib.insta_shard_frame_count.restype = C.c_uint32
    for n, world in ((10, 4), (1700, 8), (3, 8), (0, 2), (4000, 1)):
        for rank in range(world):
            assert lib.insta_shard_frame_count(n, rank, world) == len(shard.frames_of_rank(n, rank, world))
    r, li = C.c_uint32(), C.c_uint32()
    assert lib.insta_shard_owner(13, 4, C.byref(r), C.byref(li)) == 0 and (r.value, li.value) == shard.owner_of_frame(13, 4)
    g = np.array([0, 1, 2, 3, 9, -1, 17], np.int32)
    out = np.empty_like(g)
    assert lib.insta_shard_local_ids(g.ctypes.data_as(C.c_void_p), C.c_uint64(len(g)), 1, 4, out.ctypes.data_as(C.c_void_p)) == 0
    assert np.array_equal(out, shard.local_frame_ids(g, 1, 4))
    assert lib.insta_shard_local_ids(g.ctypes.data_as(C.c_void_p), C.c_uint64(len(g)), 4, 4, out.ctypes.data_as(C.c_void_p)) != 0
