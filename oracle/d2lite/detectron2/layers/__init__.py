from collections import namedtuple
from typing import List

import torch
import torch.nn.functional as F
import torchvision
from torch import nn


class ShapeSpec(namedtuple("_ShapeSpec", ["channels", "height", "width", "stride"])):
    def __new__(cls, channels=None, height=None, width=None, stride=None):
        return super().__new__(cls, channels, height, width, stride)


def cat(tensors: List[torch.Tensor], dim: int = 0):
    assert isinstance(tensors, (list, tuple))
    if len(tensors) == 1:
        return tensors[0]
    return torch.cat(tensors, dim)


def nonzero_tuple(x):
    if x.dim() == 0:
        return x.unsqueeze(0).nonzero().unbind(1)
    return x.nonzero().unbind(1)


def cross_entropy(input, target, *, reduction="mean", **kwargs):
    if target.numel() == 0 and reduction == "mean":
        return input.sum() * 0.0
    return F.cross_entropy(input, target, reduction=reduction, **kwargs)


def batched_nms(boxes, scores, idxs, iou_threshold):
    """Per-category greedy NMS (suppress IoU > thr), kept indices sorted by score (torchvision)."""
    assert boxes.shape[-1] == 4
    return torchvision.ops.boxes.batched_nms(boxes.float(), scores, idxs, iou_threshold)


def get_norm(norm, out_channels):
    if norm is None or norm == "":
        return None
    if norm == "BN":
        return nn.BatchNorm2d(out_channels)
    raise ValueError(f"d2lite: norm '{norm}' not restated")


class Conv2d(nn.Conv2d):
    """nn.Conv2d with optional `norm` and `activation` attributes (detectron2.layers.Conv2d)."""

    def __init__(self, *args, **kwargs):
        norm = kwargs.pop("norm", None)
        activation = kwargs.pop("activation", None)
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x):
        x = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x
