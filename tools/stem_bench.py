"""Stem kernels at the bench shape (B=8, 480x640): tensor-core in-SM im2col (stem_tc) vs TMA-gathered (stem_s2d)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dynmm_b200 import ops

model = bench.build_model().cuda()
eng = model.engine(torch.device("cuda"))
rgb, depth = (t.cuda() for t in bench.synthetic_batch(1000, bench.BATCH))
wr, sr, br = eng.stem["encoder_rgb"]
wd, sdp, bd = eng.stem["encoder_depth"]
packed = ops.stem_s2d_pack_weights(wr, wd)

def timeit(fn, name):
    for _ in range(3):
        out = fn()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            fn()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")
    return out

a = timeit(lambda: ops.stem(rgb, depth, wr, sr, br, wd, sdp, bd), "stem_tc ")
b = timeit(lambda: ops.stem_s2d(rgb, depth, packed, sr, br, sdp, bd), "stem_s2d (BN from shared memory)")
bn_host = ops.stem_s2d_bn_host(sr, br, sdp, bd)
timeit(lambda: ops.stem_s2d(rgb, depth, packed, sr, br, sdp, bd, bn_host=bn_host), "stem_s2d (BN as kernel parameter)")
timeit(lambda: ops.stem_s2d(rgb, depth, packed, sr, br, sdp, bd, bn_host=bn_host, want_bf16=False), "stem_s2d (parameter BN, fp32 outputs only)")
for x, y, n in zip(a, b, ("r32", "d32", "r16", "d16")):
    print(n, "max abs diff", (x.float() - y.float()).abs().max().item(), "max", x.float().abs().max().item())
