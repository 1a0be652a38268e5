"""One steady-state TRAINING step (forward + backward + SGD) of the bench workload bracketed by
cudaProfilerStart/Stop (for `ncu --profile-from-start off`), eager launches so every kernel is listed.
TRAIN_PRECISION = bf16 (tcgen05 conv fwd/dgrad/wgrad) | autocast (cuDNN bf16) | fp32 (cuDNN)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

precision = os.environ.get("TRAIN_PRECISION", "bf16")
warnings.simplefilter("ignore")
model = bench.build_model().cuda()
model.train()
model.hard_gate = False
model.train_precision = "bf16" if precision == "bf16" else "fp32"
opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, nesterov=True, weight_decay=1e-4)
rgb, depth = (t.cuda() for t in bench.synthetic_batch(7, bench.BATCH))
target = torch.randint(0, 40, (bench.BATCH, bench.H, bench.W), device="cuda")


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(precision == "autocast")):
        (out, o8, o16, o32), loss_flop = model(rgb, depth)
    loss = torch.nn.functional.cross_entropy(out.float(), target) + 1e-4 * loss_flop.float()
    loss.backward()
    opt.step()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("precision", precision, "loss", float(loss))
