"""Gate-gradient norm ratios (bf16 training graph / fp32 graph) of the batch-statistics sanity test, for both stem
layouts of the bf16 graph (DYNMM_TRAIN_STEM=nchw: cuDNN NCHW kernels as in the fp32 graph)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_backward import _train_step, _cos
from oracle import fusion_oracle as fo

cfg = fo.FusionConfig(height=160, width=224)
sd = fo.make_state_dict(cfg, 0, gate_scale=40.0)
m32, out32, loss32 = _train_step(sd, (160, 224), 4, "fp32", bn_batch_stats=True)
p32 = dict(m32.named_parameters())
for mode in ("nchw", "nhwc"):
    if mode == "nchw":
        os.environ["DYNMM_TRAIN_STEM"] = "nchw"
    else:
        os.environ.pop("DYNMM_TRAIN_STEM", None)
    m16, out16, loss16 = _train_step(sd, (160, 224), 4, "bf16", bn_batch_stats=True)
    p16 = dict(m16.named_parameters())
    ratios = {}
    for nme, p in p32.items():
        if p.grad is None or p.grad.numel() < 64 or p.grad.norm().item() <= 1e-3:
            continue
        ratios[nme] = ((p16[nme].grad.norm() / p.grad.norm()).item(), _cos(p16[nme].grad, p.grad))
    vals = sorted(ratios.items(), key=lambda kv: -abs(kv[1][0] - 1))
    print(mode, "loss", loss16, loss32, "worst ratios:", [(k, round(r, 3), round(c, 3)) for k, (r, c) in vals[:6]])
    print("   gate:", [(k, round(r, 3), round(c, 3)) for k, (r, c) in ratios.items() if k.startswith("gate")])
    print("   stem:", [(k, round(r, 3), round(c, 3)) for k, (r, c) in ratios.items() if "conv1.weight" in k or ".bn1." in k and "layer" not in k])
