"""A/B timing of the backward side on the 2^18-sample training batch: scatter roofline (insta_microbench_scatter), stand-alone hash backward and the fused network
forward + backward.  The kernel variants are chosen by environment variables read at first use (INSTA_HASH_BWD=v1, INSTA_NET_BWD_SPLIT=1), so run once per setting."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from insta_b200 import capi, synth

ctx = capi.Context(device=0)
tag = "HASH_BWD=%s NET_BWD_SPLIT=%s" % (os.environ.get("INSTA_HASH_BWD", "pairs"), os.environ.get("INSTA_NET_BWD_SPLIT", "2"))
if "--roofline" in sys.argv:
    tb = 3214536 * 16 // 128 * 128
    for cb in (16, 32, 64):
        print("scatter into an L2-resident %d MB buffer, %2d-byte chunks: %.0f GB/s of reduced bytes" % (tb >> 20, cb, ctx.microbench_scatter(tb, cb)), flush=True)
layout = ctx.grid_layout()
n_entries = int(layout[0][-1])
verts, faces, regions = synth.load_canonical()
nb = 1 << 18
rec = torch.from_numpy(synth.make_records(nb, synth.mesh_aabb(verts), np.zeros(nb, np.int32), seed=3)).cuda()
dy = torch.full((128, nb), 0.01, dtype=torch.float16, device="cuda")
grad = torch.empty((n_entries, 8), dtype=torch.float16, device="cuda")


def timed(fn, reps=12, skip=3):
    ts = []
    for k in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts[skip:])), float(np.min(ts))


m, lo = timed(lambda: ctx.hash_encode_bwd(rec, dy, grad, zero_first=True))
print("%s: hash backward 2^18 samples %.4f ms (min %.4f), |grad| sum %.1f" % (tag, m, lo, float(grad.float().abs().sum())), flush=True)
params = torch.from_numpy(synth.make_network_params(ctx.network_n_params, seed=1337)).cuda()
dl = torch.full((nb, 4), 0.01, dtype=torch.float16, device="cuda")
grads = torch.empty(ctx.network_n_params, dtype=torch.float16, device="cuda")
m, lo = timed(lambda: ctx.network_forward_backward(rec, params, dl, grads))
print("%s: network forward + backward 2^18 samples %.4f ms (min %.4f), |grads| sum %.2f, MLP part %.4f" % (tag, m, lo, float(grads.float().abs().sum()), float(grads[:25600].float().abs().sum())), flush=True)
ctx.close()
