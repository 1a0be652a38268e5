import sys, os
sys.path.insert(0, "/root/repo")
import torch
from oracle import fusion_oracle as fo
from oracle.make_golden import sample_inputs
from dynmm_b200 import ops
from tests.test_gpu_kernels import _stem_weights
for (h, w, b) in ((64, 64, 2), (96, 160, 3), (70, 90, 2)):
    cfg = fo.FusionConfig(height=h, width=w)
    sd = fo.make_state_dict(cfg, 0, 40.0)
    rgb, depth = sample_inputs(11, b, h, w)
    wr, sr, br = _stem_weights(sd, "encoder_rgb")
    wd, sdp, bd = _stem_weights(sd, "encoder_depth")
    out = ops.stem(rgb.cuda(), depth.cuda(), wr, sr, br, wd, sdp, bd)
    part, inv = ops.stem_squeeze(rgb.cuda(), depth.cuda(), wr, sr, br, wd, sdp, bd)
    torch.cuda.synchronize()
    print(h, w, "ok", float(out[0].abs().sum()), float(part.sum()))
