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
