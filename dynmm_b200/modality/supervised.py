"""``MMDL`` -- the late-fusion container of MultiBench's training structure
(ModalityDynMM/training_structures/Supervised_Learning.py:16-51)."""
from __future__ import annotations

import torch
from torch import nn


class MMDL(nn.Module):
    def __init__(self, encoders, fusion, head, has_padding=False):
        super().__init__()
        self.encoders = nn.ModuleList(encoders)
        self.fuse = fusion
        self.head = head
        self.has_padding = has_padding
        self.fuseout = None
        self.reps = []

    def forward(self, inputs):
        if self.has_padding:
            outs = [enc([x, ln]) for enc, x, ln in zip(self.encoders, inputs[0], inputs[1])]
        else:
            outs = [enc(x) for enc, x in zip(self.encoders, inputs)]
        self.reps = outs
        if self.has_padding and not isinstance(outs[0], torch.Tensor):
            out = self.fuse([o[0] for o in outs])
        else:
            out = self.fuse(outs)
        self.fuseout = out
        if type(out) is tuple:
            out = out[0]
        if self.has_padding and not isinstance(outs[0], torch.Tensor):
            out = self.head([out, inputs[1][0]])
        else:
            out = self.head(out)
        if type(out) is list:
            out = out[0]
        return out
