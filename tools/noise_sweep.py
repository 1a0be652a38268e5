#!/usr/bin/env python
"""BASELINE.json configs[4]: robustness sweep of the gated FusionDynMM forward -- Gaussian noise on the depth (or RGB)
input at sigma in {0, 0.3, 0.6, 1.0} (eval.py --mode/--noise/--num-runs), gate-branch distribution + images/s per level.

  python tools/noise_sweep.py [--batches 12] [--batch 8] [--runs 1] [--mode 1] [--noises 0,0.3,0.6,1.0] [--labels]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/noise_sweep.py ...

Synthetic NYUv2-shape inputs (bench.py's generator: N(0,1) with a per-sample gain/offset) and bench.py's seeded
random-init model whose gate head is widened so the untrained gate spreads over the branches.  Batches are sharded
round-robin over the ranks (weak scaling: --batches is PER RANK); the only collective is the all-reduce of the 5-bin
histogram and the totals.  Prints one JSON line per noise level and a final summary line (rank 0)."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, default=12, help="batches per rank and run")
    ap.add_argument("--batch", type=int, default=bench.BATCH)
    ap.add_argument("--runs", type=int, default=1, help="eval.py --num-runs (seeds 0..runs-1)")
    ap.add_argument("--mode", type=int, default=1, help="eval.py --mode: 0 rgb, 1 depth, 2 both, -1 none")
    ap.add_argument("--noises", default="0,0.3,0.6,1.0")
    ap.add_argument("--labels", action="store_true", help="predict_labels (arg-max fused, no logits written)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--precision", default="f32x3", choices=["bf16", "f32x3"], help="engine arithmetic (see bench.py)")
    args = ap.parse_args()

    import torch.distributed as dist
    from dynmm_b200 import _lib
    from dynmm_b200.fusion import robustness as rb
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    _lib.require_device()
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    model = bench.build_model().to(dev)
    model.engine_precision = args.precision
    model.use_cuda_graph = not args.no_graph
    total = args.batches * world
    # the same resident batch list on every rank (a rank only forwards its own share)
    resident = [tuple(t.to(dev) for t in bench.synthetic_batch(7000 + i, args.batch)) for i in range(min(total, 6))]

    def batches(run):
        for i in range(total):
            yield resident[i % len(resident)]

    noises = [float(s) for s in args.noises.split(",") if s]
    with torch.no_grad():
        rb.run_point(model, lambda r: list(batches(r))[:world * 2], -1, 0.0, 1, rank, world, args.labels)   # warm-up
    points = []
    for s in noises:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        pt = rb.reduce_point(rb.run_point(model, batches, args.mode, s, args.runs, rank, world, args.labels), model, dev)
        points.append(pt)
        if rank == 0:
            d = pt.as_dict()
            d.update({"n_gpus": world, "per_gpu_batch": args.batch, "height": bench.H, "width": bench.W,
                      "dtype": args.precision,
                      "api": "predict_labels" if args.labels else "forward(test=True, return_weight=True)"})
            print(json.dumps(d))
    if rank == 0:
        print(json.dumps({"workload": "FusionDynMM robustness sweep (configs[4])", "mode": args.mode, "n_gpus": world,
                          "dtype": args.precision,
                          "noises": noises, "images_per_s": [p.images_per_s for p in points],
                          "gate_branch_histograms": [p.histogram for p in points],
                          "flop_saved_pct": [p.saved_pct for p in points]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
