"""Clock stamps of the stem epilogue (library built with -DDYNMM_STEM_TRACE=1, experiments only): per-tile phases of one
epilogue warp per group and of the MMA issuer of CTA 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from dynmm_b200 import ops

model = bench.build_model().cuda()
eng = model.engine(torch.device("cuda"))
rgb, depth = (t.cuda() for t in bench.synthetic_batch(1000, bench.BATCH))
wr, sr, br = eng.stem["encoder_rgb"]
wd, sdp, bd = eng.stem["encoder_depth"]
packed = ops.stem_s2d_pack_weights(wr, wd)
for _ in range(3):
    out = ops.stem_s2d(rgb, depth, packed, sr, br, sdp, bd)
torch.cuda.synchronize()
d16 = out[3]
raw = d16.contiguous().view(torch.int64).reshape(-1)[: 2 * 20 * 1024].cpu().numpy().reshape(2, 20, 1024)
for cta in range(2):
    m = raw[cta, 16]
    t0 = m[0]
    print(f"CTA {cta}: MMA issuer, per tile (cycles since start): before acc_empty wait, after, after full wait")
    for i in range(0, 3 * 12, 3):
        print("   tile", i // 3, [int(x - t0) for x in m[i:i + 3]])
    for ew in (0, 3, 7, 8):
        e = raw[cta, ew]
        print(f" epilogue warp {ew}: tile start | acc_full wait | per pass: tmem_ld, math+shuffles, sts, barrier1, pooling+stores, barrier2(+loop)")
        for tl in range(8):
            s = [int(x) for x in e[tl * 22:(tl + 1) * 22 + 1]]
            if len(s) < 23 or s[21] == 0:
                break
            d = [s[k + 1] - s[k] for k in range(22)]
            print(f"   tile {tl}: start {s[0] - int(t0):7d} wait {d[0]:5d} | " + " | ".join(" ".join(f"{x:4d}" for x in d[1 + 5 * ps: 6 + 5 * ps]) for ps in range(4)) + f" | next {d[21]:5d}")
