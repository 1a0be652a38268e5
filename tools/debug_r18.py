import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import fusion_oracle as fo
from oracle.make_golden import sample_inputs
from dynmm_b200 import ops
from dynmm_b200.fusion.engine import _Packer
torch.backends.cudnn.allow_tf32 = False
cfg = fo.FusionConfig(height=64, width=64, encoder="resnet18", encoder_block="BasicBlock")
sd = fo.make_state_dict(cfg, 5, 40.0)
p = _Packer(sd, "cuda")
c = fo._Ctx(sd, False, "relu")
x = torch.randn(2, 64, 16, 16)
for s in range(4):
    for b in range(2):
        key = f"encoder_rgb.layer{s+1}.{b}"
        stride = 2 if (b == 0 and s > 0) else 1
        ref = fo.basic_block(c, key, x, stride)
        blk = p.basic(key, stride)
        xg = x.permute(0, 2, 3, 1).contiguous().cuda().to(torch.bfloat16)
        y = blk.convs[0](xg)
        y_ref = c.a(c.bn(c.conv(x, key + ".conv1", stride, 1), key + ".bn1"))
        e1 = ((y.float().permute(0, 3, 1, 2).cpu() - y_ref).norm() / y_ref.norm()).item()
        idn = blk.downsample(xg) if blk.downsample is not None else xg
        out = blk.convs[1](y, residual=idn)
        torch.cuda.synchronize()
        e2 = ((out.float().permute(0, 3, 1, 2).cpu() - ref).norm() / ref.norm()).item()
        print(key, tuple(x.shape), "conv1 err", f"{e1:.3e}", "block err", f"{e2:.3e}", "finite", torch.isfinite(out.float()).all().item())
        x = ref
