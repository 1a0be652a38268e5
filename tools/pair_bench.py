"""Fused NonBottleneck1D pair vs its two convolutions at the bench shape (B=8, 120x160, C=64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynmm_b200 import ops

g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(8, 120, 160, 64, device="cuda", generator=g).to(torch.bfloat16)
mk = lambda kh, kw: (ops.pack_conv_weight(torch.randn(64, 64, kh, kw, device="cuda", generator=g) * 0.07),
                     torch.randn(64, device="cuda", generator=g) * 0.1)
(w1, b1), (w2, b2) = mk(3, 1), mk(1, 3)
y = torch.empty_like(x); z = torch.empty_like(x); o = torch.empty_like(x)

def two():
    ops.conv(x, w1, c_out=64, kh=3, kw=1, pad=(1, 0), shift=b1, relu=True, out=y)
    ops.conv(y, w2, c_out=64, kh=1, kw=3, pad=(0, 1), shift=b2, relu=True, residual=x, out=z)

def fused():
    ops.conv_pair(x, w1, b1, w2, b2, residual=x, relu2=True, out=o)

for fn, name in ((two, "two convs "), (fused, "fused pair")):
    for _ in range(3):
        fn()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(10):
            fn()
    gr.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")
print("bit-identical:", torch.equal(z, o))
