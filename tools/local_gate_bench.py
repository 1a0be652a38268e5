"""Local-gate SkipESANet at the bench shape: eval forward on the CUDA engine (per-stage device-side
re-planning, real skipping) against the module's own PyTorch graph (fp32 eager and bf16 tcgen05 convolutions)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dynmm_b200 import _lib
from dynmm_b200.fusion import SkipESANet

_lib.require_device()
dev = torch.device("cuda")
torch.manual_seed(0)
model = SkipESANet(height=bench.H, width=bench.W, num_classes=40, encoder_rgb="resnet34", encoder_depth="resnet34",
                   encoder_block="NonBottleneck1D", nr_decoder_blocks=[3, 3, 3], fuse_depth_in_rgb_encoder="add",
                   upsampling="learned-3x3-zeropad").to(dev).eval()
g = torch.Generator().manual_seed(1)
with torch.no_grad():
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
model.block_rule = [2, 2, 2, 2]
model.hard_gate = True
rgb, depth = (t.to(dev) for t in bench.synthetic_batch(0, bench.BATCH))


def rate(reps):
    with torch.no_grad():
        for _ in range(3):
            model(rgb, depth, True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            model(rgb, depth, True)
        e1.record()
        torch.cuda.synchronize()
    return bench.BATCH * reps / (e0.elapsed_time(e1) * 1e-3)


out = {"workload": "local-gate SkipESANet R34-NBt1D 480x640 batch 8, block_rule 2222, hard gates (test=True), eager launches"}
model.use_engine = True
torch.manual_seed(5)
out["engine_images_per_s"] = rate(30)
out["depth_samples_per_stage"] = [int(c.item()) for c in model.last_counts]
model.use_cuda_graph = True
out["engine_graph_replay_images_per_s"] = rate(100)
out["depth_samples_per_stage_graph"] = [int(c.item()) for c in model.last_counts]
model.use_cuda_graph = False
model.use_engine = False
model.train_precision = "fp32"
torch.manual_seed(5)
out["module_graph_fp32_images_per_s"] = rate(5)
model.train_precision = "bf16"
torch.manual_seed(5)
out["module_graph_bf16_convs_images_per_s"] = rate(10)
print(json.dumps(out))
