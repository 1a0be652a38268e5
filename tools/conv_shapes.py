"""Per-shape microbenchmark of the tensor-core conv kernel (B=8 stage shapes of the
480x640 R34-NBt1D network).  Prints time, TFLOP/s and algorithmic GB/s per launch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynmm_b200 import ops

SHAPES = [
    # name, n, h, w, cin, cout, kh, kw, stride, residual
    ("s1 1x3 c64", 8, 120, 160, 64, 64, 1, 3, (1, 1), False),
    ("s1 1x3 c64 +res", 8, 120, 160, 64, 64, 1, 3, (1, 1), True),
    ("s1 3x1 c64", 8, 120, 160, 64, 64, 3, 1, (1, 1), False),
    ("s2 1x3 c128", 8, 60, 80, 128, 128, 1, 3, (1, 1), False),
    ("s2 3x1 c128 +res", 8, 60, 80, 128, 128, 3, 1, (1, 1), True),
    ("s2 3x1 s2 64->128", 8, 120, 160, 64, 128, 3, 1, (2, 1), False),
    ("s3 1x3 c256", 8, 30, 40, 256, 256, 1, 3, (1, 1), False),
    ("s3 3x1 c256 +res", 8, 30, 40, 256, 256, 3, 1, (1, 1), True),
    ("s3 1x3 c256 n6", 6, 30, 40, 256, 256, 1, 3, (1, 1), False),
    ("s4 1x3 c512", 8, 15, 20, 512, 512, 1, 3, (1, 1), False),
    ("s4 1x3 c512 n6", 6, 15, 20, 512, 512, 1, 3, (1, 1), False),
    ("s4 1x3 s2 512", 8, 15, 40, 512, 512, 1, 3, (1, 2), False),
    ("s3 1x3 s2 256", 8, 30, 80, 256, 256, 1, 3, (1, 2), False),
    ("s4 3x1 c512 +res", 8, 15, 20, 512, 512, 3, 1, (1, 1), True),
    ("dec 3x3 c128 15x20", 8, 15, 20, 128, 128, 3, 3, (1, 1), False),
    ("dec 1x3 c128 120x160", 8, 120, 160, 128, 128, 1, 3, (1, 1), False),
    ("dec 3x3 128->40 120x160", 8, 120, 160, 128, 40, 3, 3, (1, 1), False),
    ("skip 1x1 64->128 120x160", 8, 120, 160, 64, 128, 1, 1, (1, 1), False),
]

def main():
    only = os.environ.get("ONLY")
    eager = os.environ.get("EAGER") == "1"
    tile_ns = [0] + [int(x) for x in sys.argv[1:]]
    dev = "cuda"
    for name, n, h, w, cin, cout, kh, kw, stride, res in SHAPES:
        if only and only not in name:
            continue
        x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
        wt = ops.pack_conv_weight(torch.randn(cout, cin, kh, kw, device=dev) * 0.05)
        ho = (h + 2 * (kh // 2) - kh) // stride[0] + 1
        wo = (w + 2 * (kw // 2) - kw) // stride[1] + 1
        r = torch.randn(n, ho, wo, cout, device=dev).to(torch.bfloat16) if res else None
        sc = torch.rand(cout, device=dev) + 0.5
        sh = torch.randn(cout, device=dev)
        out = torch.empty(n, ho, wo, cout, dtype=torch.bfloat16, device=dev)
        for tn, dual in [(t, d) for t in tile_ns for d in ((None, False, True) if cin >= 256 else (None,))]:
            if tn > (cout + 15) // 16 * 16:
                continue
            kw_ = dict(c_out=cout, kh=kh, kw=kw, stride=stride, pad=(kh // 2, kw // 2), scale=sc, shift=sh,
                       residual=r, relu=True, out=out, tile_n=tn, dual=dual)
            for _ in range(3):
                ops.conv(x, wt, **kw_)
            torch.cuda.synchronize()
            if eager:
                continue
            R = 20
            graph = torch.cuda.CUDAGraph()      # kernel time only: no host launch cost between launches
            with torch.cuda.graph(graph):
                for _ in range(R):
                    ops.conv(x, wt, **kw_)
            graph.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / R * 1e3
            flops = 2.0 * n * ho * wo * cout * cin * kh * kw
            byts = 2.0 * (x.numel() + out.numel() + (r.numel() if res else 0) + wt.numel())
            print(f"{name:28s} tile_n={tn:3d} dual={str(dual):5s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s  {byts / us / 1e3:7.1f} GB/s")

if __name__ == "__main__":
    main()
