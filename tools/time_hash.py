"""A/B timing of the stand-alone hash-grid forward kernel (INSTA_HASH_FWD=v1 selects the round-1 kernel at first use, so run once per setting)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from insta_b200 import capi, synth
ctx = capi.Context(device=0)
layout = ctx.grid_layout()
table = torch.from_numpy(synth.make_grid_table(int(layout[0][-1]))).cuda()
verts, faces, regions = synth.load_canonical()
rec = torch.from_numpy(synth.make_records(bench.N_SAMPLES, synth.mesh_aabb(verts), np.zeros(bench.N_SAMPLES, np.int32), seed=3)).cuda()
out = torch.empty((128, bench.N_SAMPLES), dtype=torch.float16, device="cuda")
ts = []
for k in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.hash_encode_fwd(rec, table, out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("INSTA_HASH_FWD=%s: %.3f ms (min %.3f) checksum %d" % (os.environ.get("INSTA_HASH_FWD", "pairs"), np.mean(ts[3:]), np.min(ts), int(out.view(torch.int16).to(torch.int64).sum())))
ctx.close()
