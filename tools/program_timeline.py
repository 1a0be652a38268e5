"""Per-phase timeline of the convolution programs of one bench-shaped forward (B=8, 480x640): every program is
re-launched with the %globaltimer trace and the release time of each phase is printed (min/max over CTAs),
with the jobs of the phase (shape, tiles)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

model = bench.build_model().cuda()
rgb, depth = (t.cuda() for t in bench.synthetic_batch(int(os.environ.get("SEED", 1000)), bench.BATCH))
eng = model.engine(rgb.device)
with torch.no_grad():
    for _ in range(2):
        _, wgt = eng.forward(rgb, depth, temp=1.0, hard_gate=True)
torch.cuda.synchronize()
print("branches", wgt.argmax(1).tolist())
for pi, prog in enumerate(eng.programs):
    grid, nph = int(prog.cfg[0]), int(prog.cfg[2])
    trace = torch.zeros(grid, 97, dtype=torch.int64, device="cuda")
    for _ in range(3):
        prog.launch(upload=False, trace=trace)
    torch.cuda.synchronize()
    t = trace[:, :nph + 1].cpu().double()
    t0 = t[:, 0].min()
    start = (t - t0) / 1e3                     # us
    total = start[:, nph].max().item()
    print(f"== program {pi}: grid {grid}, {nph} phases, {len(prog.jobs)} jobs, {total:.1f} us, {prog.flops() / 1e9:.1f} GFLOP "
          f"-> {prog.flops() / 1e6 / total:.0f} TFLOP/s")
    by_phase = {}
    for p, ph, tensors in zip(prog.jobs, prog.phases, prog.keep):
        cnt = tensors[-1]
        act = min(int(cnt.item()), p.n) if cnt is not None else p.n
        by_phase.setdefault(ph, []).append(f"{p.kh}x{p.kw} s{p.stride_h}{p.stride_w} {p.c_in}->{p.c_out} "
                                           f"{p.h_out}x{p.w_out} n={act}")
    for ph in range(nph):
        rel = start[:, ph]
        dur = start[:, ph + 1].max().item() - rel.max().item()
        print(f"  phase {ph:3d}: released {rel.min().item():8.1f}..{rel.max().item():8.1f} us  duration {dur:6.1f} us   "
              + " | ".join(by_phase.get(ph, [])))
