import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from insta_b200 import capi
ctx = capi.Context(device=0)
tb = 3214536 * 16 // 128 * 128
for cb in (16, 32, 64, 128):
    print("L2-resident %d MB buffer, %3d-byte chunks: %.0f GB/s" % (tb >> 20, cb, ctx.microbench_gather(tb, cb)), flush=True)
ctx.close()
