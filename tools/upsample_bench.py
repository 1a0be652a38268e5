"""Decoder up-sampling kernel (nearest x2 + depth-wise 3x3 [+ skip]) at the bench shapes, bf16 and [hi | lo] split."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynmm_b200 import ops

torch.manual_seed(0)
dev = torch.device("cuda")


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for split in (False, True):
    for (h, w, c, skip) in ((15, 20, 512, True), (30, 40, 256, True), (60, 80, 128, True), (120, 160, 40, False)):
        ld = 2 * c if split else c
        x = torch.randn(8, h, w, ld, device=dev).to(torch.bfloat16)
        wgt = torch.randn(9, c, device=dev)
        bias = torch.randn(c, device=dev)
        sk = torch.randn(8, 2 * h, 2 * w, ld, device=dev).to(torch.bfloat16) if skip else None
        out = torch.empty(8, 2 * h, 2 * w, ld, device=dev, dtype=torch.bfloat16)
        us = timeit(lambda: ops.upsample2x_dw3x3(x, wgt, bias, skip=sk, out=out, split=split))
        mb = (x.numel() + out.numel() * (2 if skip else 1)) * 2 / 1e6
        print(f"split={int(split)} {h}x{w} c={c} skip={int(skip)}: {us:7.1f} us  {mb / us * 1e-3 * 1e3:7.0f} GB/s")
