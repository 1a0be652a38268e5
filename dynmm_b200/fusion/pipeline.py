"""Host-buffer inference pipeline: what ``eval.py``'s loop does per batch
(``image.to(device)`` / ``depth.to(device)`` -> ``model(image, depth, True)`` -> ``argmax`` ->
``.cpu()``; eval.py:89-90,109-120,129), with the copies of batch i+1 and the label read-back of
batch i-1 overlapped with the forward of batch i on separate CUDA streams.  The arg-max is fused
into the model's last kernel (``SkipGateESANet.predict_labels``)."""
from __future__ import annotations

from typing import Iterable, Iterator, Tuple

import torch


class EvalPipeline:
    """``in_flight`` forwards run concurrently, each on its own stream and its own captured-graph instance (with
    ``model.use_cuda_graph``): at batch 8 a forward is a chain of ~130 dependent kernels of a few microseconds, most of
    which leave SMs idle -- a second batch in flight fills them.  Results come back in order."""

    def __init__(self, model, batch: int, height: int, width: int, device=None, depth: int = None, in_flight: int = 2):
        self.model = model
        self.dev = device or next(model.parameters()).device
        self.copy_in = torch.cuda.Stream(device=self.dev)
        self.copy_out = torch.cuda.Stream(device=self.dev)
        self.in_flight = max(1, int(in_flight))
        self.compute = [torch.cuda.Stream(device=self.dev) for _ in range(self.in_flight)]
        depth = depth or (self.in_flight + 1)
        self.slots = []
        for _ in range(depth):
            self.slots.append({
                "rgb": torch.empty(batch, 3, height, width, device=self.dev),
                "depth": torch.empty(batch, 1, height, width, device=self.dev),
                "labels_dev": torch.empty(batch, height, width, dtype=torch.uint8, device=self.dev),
                "labels_host": torch.empty(batch, height, width, dtype=torch.uint8).pin_memory(),
                "in_ready": torch.cuda.Event(), "computed": torch.cuda.Event(), "out_ready": torch.cuda.Event(),
                "free": torch.cuda.Event(),
            })
        self.h2d_bytes = batch * 4 * height * width * 4
        self.d2h_bytes = batch * height * width

    def _upload(self, slot, rgb_host, depth_host):
        with torch.cuda.stream(self.copy_in):
            self.copy_in.wait_event(slot["free"])            # the previous forward that read this slot is done
            slot["rgb"].copy_(rgb_host, non_blocking=True)
            slot["depth"].copy_(depth_host, non_blocking=True)
            slot["in_ready"].record(self.copy_in)

    @torch.no_grad()
    def run(self, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]]) -> Iterator[torch.Tensor]:
        """batches: pinned host (rgb [B,3,H,W], depth [B,1,H,W]) fp32.  Yields uint8 label maps [B,H,W]
        (pinned host tensors, valid until two more batches have been yielded), in order."""
        main = torch.cuda.current_stream(self.dev)
        it = iter(batches)
        pending = []
        nxt = next(it, None)
        i = 0
        for s in self.slots:
            s["free"].record(main)
        for st in self.compute:
            st.wait_stream(main)
        if nxt is not None:
            self._upload(self.slots[0], *nxt)
        while nxt is not None:
            slot = self.slots[i % len(self.slots)]
            nxt = next(it, None)
            if nxt is not None:                               # overlap: upload batch i+1 now
                self._upload(self.slots[(i + 1) % len(self.slots)], *nxt)
            inst = i % self.in_flight
            comp = self.compute[inst]
            with torch.cuda.stream(comp):                     # forwards of different instances overlap on the GPU
                comp.wait_event(slot["in_ready"])
                self.model.predict_labels(slot["rgb"], slot["depth"], out=slot["labels_dev"], instance=inst)
                slot["computed"].record(comp)
                slot["free"].record(comp)
            with torch.cuda.stream(self.copy_out):
                self.copy_out.wait_event(slot["computed"])
                slot["labels_host"].copy_(slot["labels_dev"], non_blocking=True)
                slot["out_ready"].record(self.copy_out)
            pending.append(slot)
            if len(pending) == len(self.slots):               # keep at most `depth` batches in flight
                done = pending.pop(0)
                done["out_ready"].synchronize()
                yield done["labels_host"]
            i += 1
        for done in pending:
            done["out_ready"].synchronize()
            yield done["labels_host"]
        for st in self.compute:
            main.wait_stream(st)
