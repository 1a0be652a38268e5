"""Data-parallel TRAINING step of the gated RGB-D model (BASELINE config C3 shape family):
torchrun, one rank per GPU, NCCL.  Forward/backward through the differentiable graph (cuDNN convs,
custom CUDA DiffSoftmax / gated-blend with custom backward), ONE gradient exchange per step through
dynmm_b200.dist.GradBuckets (flat reverse-order buckets, async all-reduce), SGD-nesterov step.
Checks that replicas stay bit-identical and prints step time."""
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
    warnings.simplefilter("ignore")
    model = bench.build_model().to(dev)
    model.train()
    model.hard_gate = False
    ddp.broadcast_parameters(model)
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, nesterov=True, weight_decay=1e-4)
    buckets = ddp.GradBuckets(model.parameters())
    rgb, depth = (t.to(dev)[:per_gpu] for t in bench.synthetic_batch(7 + rank, max(per_gpu, 8)))
    target = torch.randint(0, 40, (per_gpu, bench.H, bench.W), device=dev)
    times = []
    for step in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
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
        print(f"train step: world={world} per_gpu_batch={per_gpu} loss={loss.item():.4f} "
              f"step={t.item() * 1e3:.1f} ms -> {per_gpu * world / t.item():.1f} img/s, replicas_identical={ok}, "
              f"grad buckets={len(buckets.buckets)}")
    assert ok
    if world > 1:
        dist.destroy_process_group()

if __name__ == "__main__":
    main()
