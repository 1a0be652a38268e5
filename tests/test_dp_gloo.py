"""The N>1 path on CPU: world_size-2 gloo.  Sharded batches + bucketed gradient
all-reduce reproduce the single-process gradient; the gate histogram sums."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dynmm_b200 import dist as ddp
        from dynmm_b200.modality import DynMMNet
        torch.manual_seed(0)
        model = DynMMNet(pretrain=False, freeze=True).train()      # gate-only training (imdb_dyn.py --freeze)
        model.hard_gate = False
        model.branch3.eval()                                       # no dropout noise in the frozen expert
        if rank == 1:                                              # replicas start different, then get synchronised
            with torch.no_grad():
                for p in model.parameters():
                    p.add_(1.0)
        ddp.broadcast_parameters(model, 0)
        g = torch.Generator().manual_seed(3)
        text, image = torch.randn(16, 300, generator=g), torch.randn(16, 4096, generator=g)
        target = torch.randn(16, 23, generator=g)
        buckets = ddp.GradBuckets(model.parameters(), bucket_bytes=1024)
        assert len(buckets.buckets) >= 2
        out, reg = model([ddp.shard(text), ddp.shard(image)])
        loss = ((out - ddp.shard(target)) ** 2).mean() + 0.1 * reg
        loss.backward()
        buckets.allreduce(average=True)
        hist = torch.tensor([rank + 1, 0, 0, 0, 2 * rank], dtype=torch.int64)
        ddp.allreduce_histogram(hist)
        if rank == 0:
            ref = DynMMNet(pretrain=False, freeze=True).train()
            ref.hard_gate = False
            ref.branch3.eval()
            ref.load_state_dict(model.state_dict())
            out, reg = ref([text, image])
            (((out - target) ** 2).mean() + 0.1 * reg).backward()
            err = max((a.grad - b.grad).abs().max().item() / (b.grad.abs().max().item() + 1e-12)
                      for a, b in zip(model.gate.parameters(), ref.gate.parameters()))
            ret["err"] = err
            ret["hist"] = hist.tolist()
    finally:
        dist.destroy_process_group()


def test_two_rank_gradients_match_single_process():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret["err"] < 1e-5, ret["err"]
    assert ret["hist"] == [3, 0, 0, 0, 2]


def _worker_overlap(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dynmm_b200 import dist as ddp
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(12, 32), torch.nn.Tanh(), torch.nn.Linear(32, 32), torch.nn.Tanh(),
                                    torch.nn.Linear(32, 4))
        extra = torch.nn.Parameter(torch.zeros(3))            # never receives a gradient: finish() must still reduce it
        ddp.broadcast_parameters(model, 0)
        g = torch.Generator().manual_seed(5)
        x, y = torch.randn(8, 12, generator=g), torch.randn(8, 4, generator=g)
        buckets = ddp.GradBuckets(list(model.parameters()) + [extra], bucket_bytes=256).attach()
        assert len(buckets.buckets) >= 3
        opt = torch.optim.SGD(list(model.parameters()) + [extra], lr=0.1)
        ref = torch.nn.Sequential(torch.nn.Linear(12, 32), torch.nn.Tanh(), torch.nn.Linear(32, 32), torch.nn.Tanh(),
                                  torch.nn.Linear(32, 4))
        ref.load_state_dict(model.state_dict())
        ref_opt = torch.optim.SGD(ref.parameters(), lr=0.1)
        launched = []
        for step in range(3):                                  # several steps: zero() must re-arm the hooks
            buckets.zero()
            loss = ((model(ddp.shard(x)) - ddp.shard(y)) ** 2).mean()
            loss.backward()
            launched.append(buckets.launched_in_backward)
            buckets.finish()
            assert all(p.grad.data_ptr() == v.data_ptr() for p, v in
                       zip(buckets.buckets[0], [buckets.buckets[0][0].grad])), "gradients must stay bucket views"
            opt.step()
            ref_opt.zero_grad()
            ((ref(x) - y) ** 2).mean().backward()
            ref_opt.step()
        err = max((a - b).abs().max().item() for a, b in zip(model.state_dict().values(), ref.state_dict().values()))
        if rank == 0:
            ret["err"] = err
            ret["launched"] = launched
            ret["n_buckets"] = len(buckets.buckets)
    finally:
        dist.destroy_process_group()


def test_two_rank_overlapped_buckets_match_single_process():
    """attach(): p.grad are views into the flat buckets, every bucket that received all its gradients is reduced from
    a hook DURING backward, finish() reduces the rest; three SGD steps on 2 ranks == the single-process steps."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_overlap, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret["err"] < 1e-6, ret["err"]
    # every bucket except the one holding the gradient-less parameter is launched from a hook, in every step
    assert all(n == ret["n_buckets"] - 1 for n in ret["launched"]), (ret["launched"], ret["n_buckets"])
