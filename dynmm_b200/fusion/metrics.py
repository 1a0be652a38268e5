"""Confusion matrix / mIoU on the device -- the role of ``ConfusionMatrixPytorch`` and
``miou_pytorch`` (FusionDynMM/src/confusion_matrix.py:85-178) together with the arg-max /
void-masking step of eval.py:120-131, without the per-batch GPU->CPU copy."""
from __future__ import annotations

import torch

from .. import ops


class ConfusionMatrix:
    def __init__(self, num_classes: int, device="cuda"):
        self.num_classes = num_classes
        self.confusion_matrix = torch.zeros(num_classes, num_classes, dtype=torch.int64, device=device)
        self._num_examples = 0

    def reset(self):
        self.confusion_matrix.zero_()
        self._num_examples = 0

    def update_from_logits(self, logits: torch.Tensor, label_orig: torch.Tensor, want_pred: bool = False):
        """logits [n,c,h,w] fp32 at label resolution (eval.py:117-119 resizes to it; for NYUv2 the sizes
        already match); label_orig [n,h,w] with 0 = void (eval.py:123-130)."""
        if logits.shape[2:] != label_orig.shape[1:]:
            logits = torch.nn.functional.interpolate(logits, label_orig.shape[1:], mode="bilinear",
                                                     align_corners=False)
        self._num_examples += logits.shape[0]
        return ops.argmax_confusion(logits.contiguous(), label_orig.to(logits.device, torch.uint8).contiguous(),
                                    self.confusion_matrix, want_pred)

    def compute_miou(self):
        """-> (mIoU, per-class IoU) as Python float / CPU tensor (one sync, at the end of an eval run)."""
        m, iou = ops.miou(self.confusion_matrix)
        return float(m.item()), iou.cpu()
