"""In-kernel cycle trace of the conv kernel (DYNMM trace slots, see conv_igemm.cu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynmm_b200 import ops
from tools.conv_shapes import SHAPES

NAMES = ["entry", "prologue done", "first TMA issued", "first full", "tile0 MMAs issued", "tile0 acc_full",
         "tile0 epilogue done", "all tiles done", "stores drained", "exit", "(tiles)", "producer: tile0 issued",
         "mma: k-iter0 issued", "epi: tmem loaded", "epi: computed", "epi: barrier passed"]

def main():
    dev = "cuda"
    only = os.environ.get("ONLY")
    for name, n, h, w, cin, cout, kh, kw, stride, res in SHAPES:
        if only and only not in name:
            continue
        split = os.environ.get("SPLIT") == "1"          # fp32-grade operands ([hi | lo] halves, K = 3 * c_in)
        if split:
            x = ops.split_from_f32(torch.randn(n, h, w, cin, device=dev))
            wt, _ = ops.fold_pack_conv(torch.randn(cout, cin, kh, kw, device=dev) * 0.05, None, None, split=True)
        else:
            x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
            wt = ops.pack_conv_weight(torch.randn(cout, cin, kh, kw, device=dev) * 0.05)
        ho = (h + 2 * (kh // 2) - kh) // stride[0] + 1
        wo = (w + 2 * (kw // 2) - kw) // stride[1] + 1
        r = torch.randn(n, ho, wo, cout, device=dev).to(torch.bfloat16) if res else None
        if split and res:
            r = ops.split_from_f32(torch.randn(n, ho, wo, cout, device=dev))
        sh = torch.randn(cout, device=dev)
        out = torch.empty(n, ho, wo, cout * (2 if split else 1), dtype=torch.bfloat16, device=dev)
        tiles_mode = os.environ.get("TILES") == "1"      # library built with -DDYNMM_TRACE_TILES=1: 64 slots per CTA
        per = 64 if tiles_mode else 16
        trace = torch.zeros(148 * per, dtype=torch.int64, device=dev)
        dual = {"0": False, "1": True}.get(os.environ.get("DUAL", ""), None)
        kw_ = dict(c_out=cout, kh=kh, kw=kw, stride=stride, pad=(kh // 2, kw // 2), shift=sh, residual=r, relu=True,
                   out=out, dual=dual, split=split, c_in=cin)
        for _ in range(3):
            ops.conv(x, wt, **kw_)
        ops.conv(x, wt, trace=trace, **kw_)
        torch.cuda.synchronize()
        t = trace.view(148, per).cpu()
        used = t[:, 0] > 0
        t = t[used]
        rel = (t[:, :16] - t[:, :1]).float()
        tiles = t[:, 10].float()
        print(f"== {name} dual={dual}: {int(used.sum())} CTAs, tiles/CTA avg {tiles.mean():.1f} max {tiles.max():.0f}")
        order = [0, 1, 2, 3, 12, 11, 4, 5, 13, 14, 15, 6, 7, 8, 9]
        for i in order:
            nm = NAMES[i]
            print(f"   {nm:22s} mean {rel[:, i].mean():9.0f}  max {rel[:, i].max():9.0f} cycles")
        if tiles_mode:
            for cta in (0, 1, 77):
                if cta >= t.shape[0]:
                    continue
                row = t[cta]
                nt = int(row[10])
                t0 = int(row[0])
                print(f"   CTA {cta}: per tile (cycles since entry): MMAs issued | accumulator seen full | epilogue done")
                for l in range(min(nt, 16)):
                    print(f"      tile {l:2d}: {int(row[16 + l]) - t0:8d} {int(row[32 + l]) - t0:8d} {int(row[48 + l]) - t0:8d}")

if __name__ == "__main__":
    main()
