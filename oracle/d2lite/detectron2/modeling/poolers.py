"""ROIPooler over FPN levels with ROIAlignV2 (aligned=True) — SURVEY A.4."""
import math

import torch
from torch import nn
from torchvision.ops import roi_align

from detectron2.layers import cat, nonzero_tuple


def assign_boxes_to_levels(box_lists, min_level, max_level, canonical_box_size, canonical_level):
    box_sizes = torch.sqrt(cat([boxes.area() for boxes in box_lists]))
    level_assignments = torch.floor(canonical_level + torch.log2(box_sizes / canonical_box_size + 1e-8))
    level_assignments = torch.clamp(level_assignments, min=min_level, max=max_level)
    return level_assignments.to(torch.int64) - min_level


def convert_boxes_to_pooler_format(box_lists):
    boxes = torch.cat([x.tensor for x in box_lists], dim=0)
    sizes = torch.tensor([len(x) for x in box_lists], device=boxes.device)
    indices = torch.repeat_interleave(torch.arange(len(box_lists), dtype=boxes.dtype, device=boxes.device), sizes)
    return cat([indices[:, None], boxes], dim=1)


class ROIPooler(nn.Module):
    def __init__(self, output_size, scales, sampling_ratio, pooler_type, canonical_box_size=224, canonical_level=4):
        super().__init__()
        if isinstance(output_size, int):
            output_size = (output_size, output_size)
        assert pooler_type == "ROIAlignV2", "d2lite restates ROIAlignV2 only (Base.yaml / config.py:44)"
        self.output_size, self.scales, self.sampling_ratio = output_size, list(scales), sampling_ratio
        min_level = -(math.log2(scales[0]))
        max_level = -(math.log2(scales[-1]))
        assert math.isclose(min_level, int(min_level)) and math.isclose(max_level, int(max_level))
        self.min_level, self.max_level = int(min_level), int(max_level)
        assert len(scales) == self.max_level - self.min_level + 1
        self.canonical_level, self.canonical_box_size = canonical_level, canonical_box_size

    def _pool(self, feat, rois, scale):
        return roi_align(feat, rois.to(feat.dtype), self.output_size, scale, self.sampling_ratio, aligned=True)

    def forward(self, x, box_lists):
        num_level_assignments = len(self.scales)
        assert len(x) == num_level_assignments and len(box_lists) == x[0].size(0)
        if len(box_lists) == 0:
            return torch.zeros((0, x[0].shape[1]) + self.output_size, device=x[0].device, dtype=x[0].dtype)
        pooler_fmt_boxes = convert_boxes_to_pooler_format(box_lists)
        if num_level_assignments == 1:
            return self._pool(x[0], pooler_fmt_boxes, self.scales[0])
        level_assignments = assign_boxes_to_levels(box_lists, self.min_level, self.max_level,
                                                   self.canonical_box_size, self.canonical_level)
        num_boxes = pooler_fmt_boxes.size(0)
        output = torch.zeros((num_boxes, x[0].shape[1], self.output_size[0], self.output_size[1]),
                             dtype=x[0].dtype, device=x[0].device)
        for level, scale in enumerate(self.scales):
            inds = nonzero_tuple(level_assignments == level)[0]
            output.index_put_((inds,), self._pool(x[level], pooler_fmt_boxes[inds], scale))
        return output
