"""DefaultAnchorGenerator: per level one size x ratios (0.5,1,2); order (y, x, a)."""
import math

import torch
from torch import nn

from detectron2.structures import Boxes
from detectron2.utils.registry import Registry

ANCHOR_GENERATOR_REGISTRY = Registry("ANCHOR_GENERATOR")


def _broadcast(params, num_features, name):
    if not isinstance(params[0], (list, tuple)):
        return [params] * num_features
    if len(params) == 1:
        return list(params) * num_features
    assert len(params) == num_features, name
    return params


@ANCHOR_GENERATOR_REGISTRY.register()
class DefaultAnchorGenerator(nn.Module):
    box_dim = 4

    def __init__(self, cfg, input_shape):
        super().__init__()
        sizes = cfg.MODEL.ANCHOR_GENERATOR.SIZES
        ratios = cfg.MODEL.ANCHOR_GENERATOR.ASPECT_RATIOS
        self.strides = [x.stride for x in input_shape]
        self.offset = cfg.MODEL.ANCHOR_GENERATOR.OFFSET
        n = len(self.strides)
        sizes, ratios = _broadcast(sizes, n, "sizes"), _broadcast(ratios, n, "aspect_ratios")
        self.cell_anchors = [self.generate_cell_anchors(s, a).float() for s, a in zip(sizes, ratios)]

    @property
    def num_anchors(self):
        return [len(c) for c in self.cell_anchors]

    @staticmethod
    def generate_cell_anchors(sizes, aspect_ratios):
        anchors = []
        for size in sizes:
            area = size ** 2.0
            for r in aspect_ratios:
                w = math.sqrt(area / r)
                h = r * w
                anchors.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
        return torch.tensor(anchors)

    def forward(self, features):
        out = []
        for feat, stride, base in zip(features, self.strides, self.cell_anchors):
            gh, gw = feat.shape[-2:]
            base = base.to(feat.device)
            sx = torch.arange(self.offset * stride, gw * stride, step=stride, dtype=torch.float32, device=feat.device)
            sy = torch.arange(self.offset * stride, gh * stride, step=stride, dtype=torch.float32, device=feat.device)
            yy, xx = torch.meshgrid(sy, sx, indexing="ij")
            xx, yy = xx.reshape(-1), yy.reshape(-1)
            shifts = torch.stack((xx, yy, xx, yy), dim=1)
            out.append(Boxes((shifts.view(-1, 1, 4) + base.view(1, -1, 4)).reshape(-1, 4)))
        return out


def build_anchor_generator(cfg, input_shape):
    return ANCHOR_GENERATOR_REGISTRY.get(cfg.MODEL.ANCHOR_GENERATOR.NAME)(cfg, input_shape)
