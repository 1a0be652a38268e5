"""CPU oracle for the multi-scale weighted cross-entropy of the training step.
TEST INFRASTRUCTURE ONLY.  Restates FusionDynMM/src/utils.py:18-50 (`CrossEntropyLoss2d.forward`) with numpy in
float64; pinned by vectors generated from the reference class itself (oracle/make_golden_loss.py ->
tests/golden/loss_ce2d.npz)."""
import numpy as np


def ce2d_scale(logits: np.ndarray, targets: np.ndarray, weight: np.ndarray) -> float:
    """One scale: logits [n, c, h, w], targets [n, h, w] with 0 = void, 1..c = class + 1, weight [c]."""
    n, c, h, w = logits.shape
    x = logits.astype(np.float64)
    t = targets.astype(np.int64) - 1                      # utils.py:39-40: void -> -1 (ignore_index)
    m = x.max(1, keepdims=True)
    logp = x - m - np.log(np.exp(x - m).sum(1, keepdims=True))
    valid = t >= 0
    tc = np.where(valid, t, 0)
    picked = np.take_along_axis(logp, tc[:, None], 1)[:, 0]
    loss_all = np.where(valid, -weight.astype(np.float64)[tc] * picked, 0.0)   # reduction='none', weighted (utils.py:41)
    per_class = np.bincount(targets.reshape(-1).astype(np.int64), minlength=c + 1)   # utils.py:43-45
    divisor = float((per_class[1:] * weight.astype(np.float64)).sum())         # utils.py:46-47 (without void)
    return float(loss_all.sum() / divisor)                                     # utils.py:48


def ce2d(logits_scales, targets_scales, weight: np.ndarray):
    """utils.py:34-50: one loss per scale (train.py sums them)."""
    return [ce2d_scale(x, t, weight) for x, t in zip(logits_scales, targets_scales)]
