"""CPU oracle for the eval post-processing.  TEST INFRASTRUCTURE ONLY.
Restates FusionDynMM/eval.py:120-141 and src/confusion_matrix.py:122-133,139-178 with numpy
(the reference's own module needs tensorflow + ignite, which are not installed)."""
import numpy as np


def confusion_from_logits(logits: np.ndarray, label_orig: np.ndarray, num_classes: int) -> np.ndarray:
    pred = logits.argmax(1)                       # eval.py:120
    mask = label_orig > 0                         # eval.py:123
    label = label_orig[mask].astype(np.int64) - 1  # eval.py:124,130
    pred = pred[mask].astype(np.int64)
    idx = num_classes * label + pred              # confusion_matrix.py:130
    return np.bincount(idx, minlength=num_classes ** 2).reshape(num_classes, num_classes)


def iou(cm: np.ndarray) -> np.ndarray:
    cm = cm.astype(np.float64)
    d = np.diag(cm)
    return d / (cm.sum(1) + cm.sum(0) - d + 1e-15)   # confusion_matrix.py:153


def miou(cm: np.ndarray) -> float:
    return float(iou(cm).mean())                  # confusion_matrix.py:177-178
