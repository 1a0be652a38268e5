"""In-kernel cycle stamps of two consecutive layers of the chain kernel (DYNMM_CHAIN_TRACE_LAYER, default 1 and 2)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynmm_b200 import ops
from tools.chain_bench import CASES, layers

NAMES = ["mma: operand ready", "mma: all issued", "epi: acc_full", "epi: math done", "epi: published", "epi: polled",
         "epi: a_ready arrive"]


def main():
    dev = torch.device("cuda")
    only = os.environ.get("ONLY")
    for name, c, h, w, jobs in CASES:
        if only and only not in name:
            continue
        js = []
        for n, nb, drop, cnt in jobs:
            x = torch.randn(n, h, w, c, device=dev).to(torch.bfloat16)
            lay = ops.nbt1d_chain_layers(layers(c, nb, dev), drop_last=drop)
            count = torch.tensor([cnt], dtype=torch.int32, device=dev) if cnt is not None else None
            js.append(dict(x=x, image=ops.ChainImage(lay, c, dev), count=count, count_settled=True))
        total = sum(j["x"].shape[0] for j in js)
        units, _ = ops.chain_plan(h, w, c, total)
        for _ in range(2):
            ops.conv_chain(js)
        trace = torch.zeros(units * 16, dtype=torch.int64, device=dev)
        ops.conv_chain(js, trace=trace)
        torch.cuda.synchronize()
        t = trace.view(units, 16).cpu()
        t = t[t[:, 2] > 0]
        base = t[:, 2:3]
        print(f"== {name}: {t.shape[0]} CTAs; entry -> traced layer {(t[:, 2] - t[:, 0]).float().mean():.0f} cycles")
        for li in range(2):
            for k, nm in enumerate(NAMES):
                col = 2 + 7 * li + k
                v = (t[:, col] - base[:, 0]).float()
                v = v[t[:, col] > 0]
                if v.numel():
                    print(f"   layer +{li} {nm:22s} mean {v.mean():8.0f}  min {v.min():8.0f}  max {v.max():8.0f}")


if __name__ == "__main__":
    main()
