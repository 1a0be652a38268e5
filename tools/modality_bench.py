"""BASELINE.json configs[0] and configs[3]: modality-level DynMM forwards, samples/s on one B200 next to the CPU
restatement (oracle/modality_oracle.py) on the host cores.

  configs[3]  MM-IMDB image+text late-fusion DynMM, hard gate, batch 128 (imdb_dyn.py:89-101)
  configs[0]  CMU-MOSEI DynMMNetV2, soft gate over 2 experts, batch 32, T = 50 (affect_dyn.py:152-165)

Seeded random-init experts (the pretrained MultiBench pickles are not reachable offline), synthetic features of the
named shapes.  Device time with CUDA events over K forwards after W warm-ups; one JSON line per configuration."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynmm_b200 import _lib
from dynmm_b200.modality import DynMMNet, DynMMNetV2
from oracle import modality_oracle as mo          # CPU baseline leg only

_lib.require_device()
K, W = 200, 10
threads = os.cpu_count() or 1
torch.set_num_threads(threads)


def gpu_rate(fn, batch):
    with torch.no_grad():
        for _ in range(W):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(K):
            fn()
        e1.record()
        torch.cuda.synchronize()
    return batch * K / (e0.elapsed_time(e1) * 1e-3)


def cpu_rate(fn, batch, reps=5):
    with torch.no_grad():
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
    return batch * reps / (time.perf_counter() - t0)


g = torch.Generator().manual_seed(1)
# ---------------------------------------------------------------- configs[3]: MM-IMDB, hard gate, B = 128
torch.manual_seed(0)
model = DynMMNet(pretrain=False, freeze=True).eval()
inputs = [torch.randn(128, 300, generator=g), torch.randn(128, 4096, generator=g)]
sd = {k: v.clone() for k, v in model.state_dict().items()}
model = model.cuda()
dev_inputs = [t.cuda() for t in inputs]
for hard in (True, False):
    model.hard_gate = hard
    rate = gpu_rate(lambda: model(dev_inputs), 128)
    base = cpu_rate(lambda: mo.imdb_forward(sd, inputs, 1.0, hard), 128)
    print(json.dumps({"workload": "MM-IMDB DynMMNet eval forward, batch 128 (configs[3])", "hard_gate": hard,
                      "value": rate, "unit": "samples/s", "route_counts": model.last_route_counts if hard else None,
                      "cpu_baseline": {"value": base, "unit": "samples/s", "cores": threads, "kind": "port"}}))

# ---------------------------------------------------------------- configs[0]: CMU-MOSEI, soft gate, B = 32, T = 50
torch.manual_seed(0)
model = DynMMNetV2(1.0, False, True, None).eval()
feats = [torch.randn(32, 50, d, generator=g) for d in (35, 74, 300)]
lens = [torch.full((32,), 50)] * 3
sd = {k: v.clone() for k, v in model.state_dict().items()}
model = model.cuda()
dev_inputs = [[t.cuda() for t in feats], lens]
for hard in (False, True):
    model.hard_gate = hard
    rate = gpu_rate(lambda: model(dev_inputs), 32)
    base = cpu_rate(lambda: mo.mosei_forward(sd, [feats, lens], 1.0, hard), 32, reps=3)
    print(json.dumps({"workload": "CMU-MOSEI DynMMNetV2 eval forward, batch 32, T=50 (configs[0])", "hard_gate": hard,
                      "value": rate, "unit": "samples/s",
                      "cpu_baseline": {"value": base, "unit": "samples/s", "cores": threads, "kind": "port"}}))
