"""One steady-state forward of the bench workload bracketed by cudaProfilerStart/Stop
(for `ncu --profile-from-start off`).  Eager launches (no CUDA graph) so every kernel
is listed under its own name."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

model = bench.build_model().cuda()
model.engine_precision = os.environ.get("PRECISION", "f32x3")
rgb, depth = (t.cuda() for t in bench.synthetic_batch(0, bench.BATCH))
with torch.no_grad():
    for _ in range(3):
        model(rgb, depth, True)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    out, w = model(rgb, depth, True, True)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("precision", model.engine_precision, "branches", w.argmax(1).tolist(), "launches", model.engine().launches)
