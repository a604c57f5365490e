from torch import nn

from detectron2.layers import ShapeSpec


class Backbone(nn.Module):
    @property
    def size_divisibility(self):
        return 0

    @property
    def padding_constraints(self):
        return {}

    def output_shape(self):
        return {name: ShapeSpec(channels=self._out_feature_channels[name], stride=self._out_feature_strides[name])
                for name in self._out_features}
