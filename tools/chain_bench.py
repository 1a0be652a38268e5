"""Chain kernel (dynmm_conv_chain_fwd) against the per-layer launches on the shapes of a 480x640 batch-8 step."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynmm_b200 import ops

CASES = [  # name, c, h, w, [(n, blocks, drop_last, count)]
    ("stage3 rgb+depth(4)", 256, 30, 40, [(8, 5, True, None), (8, 5, False, 4)]),
    ("stage3 rgb+depth(8)", 256, 30, 40, [(8, 5, True, None), (8, 5, False, 8)]),
    ("stage2 rgb+depth(4)", 128, 60, 80, [(8, 3, True, None), (8, 3, False, 4)]),
    ("decoder 15x20", 128, 15, 20, [(8, 3, False, None)]),
    ("decoder 30x40", 128, 30, 40, [(8, 3, False, None)]),
    ("decoder 60x80", 128, 60, 80, [(8, 3, False, None)]),
]


def layers(c, n_blocks, dev):
    blocks = []
    for _ in range(n_blocks):
        blk = []
        for i in range(4):
            shape = (c, c, 3, 1) if i % 2 == 0 else (c, c, 1, 3)
            w = torch.randn(shape, device=dev) * (1.5 / (3 * c) ** 0.5)
            blk.append((ops.pack_conv_weight(w), torch.randn(c, device=dev) * 0.1, True))
        blocks.append(blk)
    return blocks


def reference(x, lay, count):
    c = x.shape[3]
    cur, out = x, None
    for (w, shift, taps_h, relu, residual, store) in lay:
        kh, kw = (3, 1) if taps_h else (1, 3)
        res = {0: None, 1: x, 2: out}[residual]
        cur = ops.conv(cur, w, c_out=c, kh=kh, kw=kw, pad=(kh // 2, kw // 2), shift=shift, relu=bool(relu), residual=res,
                       count=count, count_settled=count is not None)
        if store == 1:
            out = cur
    return cur


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    dev = torch.device("cuda")
    torch.manual_seed(0)
    for name, c, h, w, jobs in CASES:
        js, refs, flop = [], [], 0.0
        for n, nb, drop, cnt in jobs:
            x = torch.randn(n, h, w, c, device=dev).to(torch.bfloat16)
            lay = ops.nbt1d_chain_layers(layers(c, nb, dev), drop_last=drop)
            count = torch.tensor([cnt], dtype=torch.int32, device=dev) if cnt is not None else None
            js.append(dict(x=x, image=ops.ChainImage(lay, c, dev), count=count, count_settled=True))
            refs.append((x, lay, count))
            flop += 2.0 * h * w * c * c * 3 * len(lay) * (cnt if cnt is not None else n)
        total = sum(j["x"].shape[0] for j in js)
        plan = ops.chain_plan(h, w, c, total)
        if plan is None:
            print(f"{name}: unsupported")
            continue
        flags = torch.zeros(plan[0] + total, dtype=torch.int32, device=dev)
        t_chain = timed(lambda: ops.conv_chain(js, flags=flags))
        t_ref = timed(lambda: [reference(*r) for r in refs])
        nl = sum(len(r[1]) for r in refs)
        print(f"{name:22s} units {plan[0]:4d}  chain {t_chain:8.1f} us ({flop / t_chain * 1e-6:6.1f} TF/s, "
              f"{t_chain / max(len(r[1]) for r in refs):5.2f} us/layer)   per-layer launches {t_ref:8.1f} us ({nl} launches)")


if __name__ == "__main__":
    main()
