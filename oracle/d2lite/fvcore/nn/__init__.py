"""d2lite restatement of the fvcore.nn pieces cubercnn imports (rpn.py:11, fast_rcnn.py:7, cube_head.py:8)."""
import torch

from . import weight_init  # noqa: F401


def smooth_l1_loss(input, target, beta: float, reduction: str = "none"):
    if beta < 1e-5:
        loss = torch.abs(input - target)
    else:
        n = torch.abs(input - target)
        loss = torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta)
    if reduction == "mean":
        loss = loss.mean() if loss.numel() > 0 else 0.0 * loss.sum()
    elif reduction == "sum":
        loss = loss.sum()
    return loss


def giou_loss(*a, **k):
    raise NotImplementedError("d2lite: giou_loss is off the default path (fast_rcnn.py:229)")
