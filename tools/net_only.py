"""Times the fused inference network kernel alone on the 4M-record batch (variants by environment: INSTA_NET_FWD=v1, INSTA_NET_ENC_WARPS=8..16).  GPU box only."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from insta_b200 import capi, synth
ctx = capi.Context(device=0)
verts, faces, regions = synth.load_canonical()
rec = torch.from_numpy(synth.make_records(bench.N_SAMPLES, synth.mesh_aabb(verts), np.zeros(bench.N_SAMPLES, np.int32), seed=7)).cuda()
params = torch.from_numpy(synth.make_network_params(ctx.network_n_params, seed=1337)).cuda()
out = torch.empty((rec.shape[0], 4), dtype=torch.float16, device="cuda")
ts = []
for k in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.network_inference(rec, params, out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("NET_FWD=%s ENC_WARPS=%s: %.4f ms (min %.4f) checksum %d" % (os.environ.get("INSTA_NET_FWD", "pairs"), os.environ.get("INSTA_NET_ENC_WARPS", "12"), np.mean(ts[3:]), np.min(ts),
                                                                   int(out.view(torch.int16).to(torch.int64).sum())))
ctx.close()
