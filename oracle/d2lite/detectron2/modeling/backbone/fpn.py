"""FPN top-down pathway: lateral 1x1 + output 3x3 convs with bias, nearest x2 + sum."""
import math

import torch.nn.functional as F
from fvcore.nn import weight_init
from torch import nn

from detectron2.layers import Conv2d, get_norm

from .backbone import Backbone


class LastLevelMaxPool(nn.Module):
    def __init__(self):
        super().__init__()
        self.num_levels = 1
        self.in_feature = "p5"

    def forward(self, x):
        return [F.max_pool2d(x, kernel_size=1, stride=2, padding=0)]


class FPN(Backbone):
    def __init__(self, bottom_up, in_features, out_channels, norm="", top_block=None, fuse_type="sum",
                 square_pad=0):
        super().__init__()
        assert isinstance(bottom_up, Backbone) and in_features
        input_shapes = bottom_up.output_shape()
        strides = [input_shapes[f].stride for f in in_features]
        in_channels_per_feature = [input_shapes[f].channels for f in in_features]
        for i in range(1, len(strides)):
            assert strides[i] == 2 * strides[i - 1]
        lateral_convs, output_convs = [], []
        use_bias = norm == ""
        for idx, in_channels in enumerate(in_channels_per_feature):
            lateral_conv = Conv2d(in_channels, out_channels, kernel_size=1, bias=use_bias,
                                  norm=get_norm(norm, out_channels))
            output_conv = Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=use_bias,
                                 norm=get_norm(norm, out_channels))
            weight_init.c2_xavier_fill(lateral_conv)
            weight_init.c2_xavier_fill(output_conv)
            stage = int(math.log2(strides[idx]))
            self.add_module("fpn_lateral{}".format(stage), lateral_conv)
            self.add_module("fpn_output{}".format(stage), output_conv)
            lateral_convs.append(lateral_conv)
            output_convs.append(output_conv)
        # top-down order: coarsest first
        self.lateral_convs = lateral_convs[::-1]
        self.output_convs = output_convs[::-1]
        self.top_block = top_block
        self.in_features = tuple(in_features)
        self.bottom_up = bottom_up
        self._out_feature_strides = {"p{}".format(int(math.log2(s))): s for s in strides}
        if self.top_block is not None:
            stage = int(math.log2(strides[-1]))
            for s in range(stage, stage + self.top_block.num_levels):
                self._out_feature_strides["p{}".format(s + 1)] = 2 ** (s + 1)
        self._out_features = list(self._out_feature_strides.keys())
        self._out_feature_channels = {k: out_channels for k in self._out_features}
        self._size_divisibility = strides[-1]
        self._square_pad = square_pad
        assert fuse_type in {"avg", "sum"}
        self._fuse_type = fuse_type

    @property
    def size_divisibility(self):
        return self._size_divisibility

    def forward(self, x):
        bottom_up_features = self.bottom_up(x)
        results = []
        prev = self.lateral_convs[0](bottom_up_features[self.in_features[-1]])
        results.append(self.output_convs[0](prev))
        for idx, (lateral_conv, output_conv) in enumerate(zip(self.lateral_convs, self.output_convs)):
            if idx > 0:
                features = bottom_up_features[self.in_features[-idx - 1]]
                top_down = F.interpolate(prev, scale_factor=2.0, mode="nearest")
                prev = lateral_conv(features) + top_down
                if self._fuse_type == "avg":
                    prev /= 2
                results.insert(0, output_conv(prev))
        if self.top_block is not None:
            if self.top_block.in_feature in bottom_up_features:
                top_in = bottom_up_features[self.top_block.in_feature]
            else:
                top_in = results[self._out_features.index(self.top_block.in_feature)]
            results.extend(self.top_block(top_in))
        assert len(self._out_features) == len(results)
        return dict(zip(self._out_features, results))
