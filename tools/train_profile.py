"""Where one bf16 training step (B=8, 480x640) spends its device time: torch.profiler (CUPTI) table of
the kernels of ONE eager step after warm-up, grouped by kernel name -- the input for deciding which element-wise /
BatchNorm / layout passes to fuse into the convolution epilogues next."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import bench

warnings.simplefilter("ignore")
dev = torch.device("cuda", 0)
per_gpu = int(os.environ.get("PER_GPU_BATCH", 8))
precision = os.environ.get("TRAIN_PRECISION", "bf16")
model = bench.build_model().to(dev)
model.train()
model.hard_gate = False
model.train_precision = "bf16" if precision == "bf16" else "fp32"
opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, nesterov=True, weight_decay=1e-4)
rgb, depth = (t.to(dev)[:per_gpu] for t in bench.synthetic_batch(7, max(per_gpu, 8)))
target = torch.randint(0, 40, (per_gpu, bench.H, bench.W), device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(precision == "autocast")):
        (out, o8, o16, o32), loss_flop = model(rgb, depth)
    loss = torch.nn.functional.cross_entropy(out.float(), target) + 1e-4 * loss_flop.float()
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
total = sum(e.device_time_total for e in rows)
print(f"precision={precision} per_gpu_batch={per_gpu}: {total / 1e3:.2f} ms of kernel time in one eager step, "
      f"{sum(e.count for e in rows)} launches")
print(f"{'kernel':<90} {'n':>5} {'ms':>8} {'share':>6}")
for e in rows[:40]:
    print(f"{e.key[:90]:<90} {e.count:>5} {e.device_time_total / 1e3:>8.3f} {100 * e.device_time_total / total:>5.1f}%")
