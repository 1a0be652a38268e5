"""Fused multi-scale cross-entropy (dynmm_ce2d_fwd / dynmm_ce2d_bwd) against its HBM bound and against the PyTorch
statement, at the training step's sizes (train.py:312-316: batch 8, 40 classes, 480x640 + the three decoder side
outputs at 1/8, 1/16, 1/32).  Algorithmic bytes per scale: forward reads the fp32 logits once (+ int32 targets, writes
the per-pixel logsumexp); backward reads logits + lse + targets and writes the gradient."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynmm_b200 import _lib
from dynmm_b200.fusion import CrossEntropyLoss2d

_lib.require_device()
peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
hbm = json.load(open(peaks)).get("hbm_gbs", 6650.0) if os.path.exists(peaks) else 6650.0
B, C = 8, 40
g = torch.Generator().manual_seed(0)
loss_fn = CrossEntropyLoss2d(torch.device("cuda"), (0.5 + torch.rand(C, generator=g)).numpy())
for h, w in ((480, 640), (60, 80), (30, 40), (15, 20)):
    x0 = torch.randn(B, C, h, w, generator=g).cuda()
    t = torch.randint(0, C + 1, (B, h, w), generator=g).cuda()
    px = B * h * w
    fwd_bytes = px * (4 * C + 4 + 4)
    bwd_bytes = px * (4 * C + 4 + 4 + 4 * C)
    line = {"shape": [B, C, h, w], "fwd_MB": fwd_bytes / 1e6, "bwd_MB": bwd_bytes / 1e6}
    for flag, name in (("1", "cuda"), ("0", "torch")):
        os.environ["DYNMM_CE_CUDA"] = flag

        def step():
            x = x0.detach().requires_grad_(True)
            (loss,) = loss_fn([x], [t])
            return x, loss
        for _ in range(3):
            x, loss = step()
            loss.backward()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        reps = 20
        tf = tb = 0.0
        for _ in range(reps):
            ev[0].record()
            x, loss = step()
            ev[1].record()
            loss.backward()
            ev[2].record()
            torch.cuda.synchronize()
            tf += ev[0].elapsed_time(ev[1]) * 1e-3
            tb += ev[1].elapsed_time(ev[2]) * 1e-3
        tf, tb = tf / reps, tb / reps
        line[name] = {"fwd_us": tf * 1e6, "bwd_us": tb * 1e6, "fwd_GBs": fwd_bytes / tf / 1e9,
                      "bwd_GBs": bwd_bytes / tb / 1e9, "fwd_frac_hbm": fwd_bytes / tf / 1e9 / hbm,
                      "bwd_frac_hbm": bwd_bytes / tb / 1e9 / hbm}
    print(json.dumps(line))
