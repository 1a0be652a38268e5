"""Compare the bf16 kernel training graph with the fp32 reference graph module by module
(forward outputs) and parameter by parameter (gradients).  Writes gpurun_out/debug_train.txt."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import fusion_oracle as fo
from oracle.make_golden import sample_inputs
from dynmm_b200.fusion import SkipGateESANet

warnings.simplefilter("ignore")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
hh, ww = 160, 224
cfg = fo.FusionConfig(height=hh, width=ww)
sd = fo.make_state_dict(cfg, 0, gate_scale=40.0)
rgb, depth = (t.cuda() for t in sample_inputs(1, 4, hh, ww))
target = torch.randint(0, 40, (4, hh, ww), device="cuda")


def run(precision):
    m = SkipGateESANet(height=hh, width=ww).cuda()
    m.load_state_dict(sd)
    m.train()
    m.temp, m.hard_gate, m.train_precision = 1.0, True, ("bf16" if precision == "bf16" else "fp32")
    acts = {}
    hooks = []
    for name, mod in m.named_modules():
        if len(list(mod.children())) == 0:
            hooks.append(mod.register_forward_hook(
                lambda mod, inp, out, name=name: acts.__setitem__(name, out.detach().float().clone())
                if torch.is_tensor(out) else None))
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(precision == "autocast")):
        (o, s8, s16, s32), lf = m(rgb, depth)
    loss = F.cross_entropy(o.float(), target) + 0.1 * lf.float()
    loss.backward()
    for h in hooks:
        h.remove()
    return m, acts, o.detach().float(), float(loss.detach())


m16, a16, o16, l16 = run("bf16")
m32, a32, o32, l32 = run("fp32")
mac, aac, oac, lac = run("autocast")
lines = [f"loss bf16 {l16:.5f} fp32 {l32:.5f}  out rel {((o16 - o32).norm() / o32.norm()).item():.4f}   "
         f"library autocast-bf16: loss {lac:.5f} out rel {((oac - o32).norm() / o32.norm()).item():.4f}"]
pac = dict(mac.named_parameters())
for k in ("decoder.conv_out.weight", "decoder.decoder_module_3.conv3x3.conv.weight",
          "decoder.decoder_module_2.conv3x3.conv.weight", "encoder_rgb.layer3.5.conv1x3_2.weight",
          "encoder_rgb.layer1.0.conv3x1_1.weight"):
    a, b = pac[k].grad.flatten().double(), dict(m32.named_parameters())[k].grad.flatten().double()
    lines.append(f"library autocast-bf16 vs fp32 grad cos {k}: {float((a @ b) / (a.norm() * b.norm() + 1e-30)):.4f}")
lines.append("---- forward activations (relative L2)")
for k in a32:
    if k in a16 and a16[k].shape == a32[k].shape:
        rel = ((a16[k] - a32[k]).norm() / (a32[k].norm() + 1e-20)).item()
        lines.append(f"{k:60s} {rel:.4f}  |ref| {a32[k].norm().item():.3g}")
lines.append("---- gradients (cosine, norm ratio)")
p16, p32 = dict(m16.named_parameters()), dict(m32.named_parameters())
for k, p in p32.items():
    if p.grad is None or p16[k].grad is None:
        lines.append(f"{k:60s} grad missing: fp32 {p.grad is None} bf16 {p16[k].grad is None}")
        continue
    a, b = p16[k].grad.flatten().double(), p.grad.flatten().double()
    cos = float((a @ b) / (a.norm() * b.norm() + 1e-30))
    lines.append(f"{k:60s} cos {cos:.4f} ratio {float(a.norm() / (b.norm() + 1e-30)):.3f} |ref| {float(b.norm()):.3g}")
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/debug_train.txt", "w").write("\n".join(lines) + "\n")
print(lines[0])
