"""Data-parallel TRAINING step of the gated RGB-D model (BASELINE config C3 shape family):
torchrun, one rank per GPU, NCCL.  Forward/backward through the differentiable graph (cuDNN convs,
custom CUDA DiffSoftmax / gated-blend with custom backward), ONE gradient exchange per step through
dynmm_b200.dist.GradBuckets (flat reverse-order buckets, async all-reduce), SGD-nesterov step.
Checks that replicas stay bit-identical and prints step time.

TRAIN_PRECISION selects the arithmetic of the step:
  bf16      encoder/decoder convolutions forward + data gradient + weight gradient on the tcgen05 kernels
            (model.train_precision = "bf16"; dynmm_b200/fusion/train_ops.py)            [default]
  autocast  the library baseline: torch.autocast(bf16) over cuDNN convolutions
  fp32      the reference's arithmetic (cuDNN fp32)"""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from dynmm_b200 import dist as ddp

def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    per_gpu = int(os.environ.get("PER_GPU_BATCH", 8))
    precision = os.environ.get("TRAIN_PRECISION", "bf16")
    warnings.simplefilter("ignore")
    model = bench.build_model().to(dev)
    model.train()
    model.hard_gate = False
    model.train_precision = "bf16" if precision == "bf16" else "fp32"
    ddp.broadcast_parameters(model)
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, nesterov=True, weight_decay=1e-4)
    buckets = ddp.GradBuckets(model.parameters())
    if os.environ.get("OVERLAP", "1") != "0":
        buckets.attach()      # gradients are bucket views, all-reduce launched from hooks during backward
    rgb, depth = (t.to(dev)[:per_gpu] for t in bench.synthetic_batch(7 + rank, max(per_gpu, 8)))
    target = torch.randint(0, 40, (per_gpu, bench.H, bench.W), device=dev)
    times = []
    use_graph = os.environ.get("GRAPH", "1") != "0"
    if use_graph:
        # whole step (forward, backward, gradient exchange, optimizer) as ONE CUDA graph, timed with events
        from dynmm_b200.fusion.train_graph import GraphedTrainStep

        def loss_fn(out, tgt):
            (o, o8, o16, o32), loss_flop = out
            return torch.nn.functional.cross_entropy(o.float(), tgt) + 1e-4 * loss_flop.float()
        gstep = GraphedTrainStep(model, opt, loss_fn, rgb, depth, target, buckets=buckets if world > 1 else None,
                                 autocast=torch.bfloat16 if precision == "autocast" else None)
        for _ in range(3):
            loss = gstep(rgb, depth, target)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_rep = 10
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record()
        for _ in range(n_rep):
            loss = gstep(rgb, depth, target)
        e1.record()
        torch.cuda.synchronize()
        times = [e0.elapsed_time(e1) / n_rep * 1e-3] * 3
    else:
        for step in range(7):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if buckets._attached:
                buckets.zero()
            else:
                opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(precision == "autocast")):
                (out, o8, o16, o32), loss_flop = model(rgb, depth)
            loss = torch.nn.functional.cross_entropy(out.float(), target) + 1e-4 * loss_flop.float()
            loss.backward()
            buckets.allreduce(average=True)           # the single exchange step of the path
            opt.step()
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
    # replicas must agree after synchronised steps
    probe = torch.stack([p.detach().float().sum() for p in list(model.parameters())[:8]])
    ref = probe.clone()
    if world > 1:
        dist.broadcast(ref, 0)
    ok = torch.equal(probe, ref)
    t = torch.tensor([min(times[2:])], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"train step [{precision}{', graph' if use_graph else ', eager'}]: world={world} per_gpu_batch={per_gpu} loss={loss.item():.4f} "
              f"step={t.item() * 1e3:.1f} ms -> {per_gpu * world / t.item():.1f} img/s, replicas_identical={ok}, "
              f"grad buckets={len(buckets.buckets)}")
    assert ok
    if world > 1:
        # captured graphs hold NCCL work: release them before tearing the communicator down (otherwise the
        # destroy can hang until the launcher kills the ranks)
        if use_graph:
            del gstep
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()

if __name__ == "__main__":
    main()
