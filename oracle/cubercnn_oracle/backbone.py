"""DLA-34 / torchvision ResNet bottom-ups + FPN (restates cubercnn/modeling/backbone/dla.py:40-68,
156-321,417-507 and resnet.py:12-96 for the BASELINE backbones)."""
import math

import torch.nn.functional as F
from detectron2.modeling.backbone import FPN, Backbone, LastLevelMaxPool
from torch import nn


def _cbr(cin, cout, k, stride=1):
    return [nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2, bias=False), nn.BatchNorm2d(cout),
            nn.ReLU(inplace=True)]


class BasicBlock(nn.Module):
    def __init__(self, cin, cout, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)

    def forward(self, x, residual=None):
        residual = x if residual is None else residual
        y = self.relu(self.bn1(self.conv1(x)))
        return self.relu(self.bn2(self.conv2(y)) + residual)


class Root(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, *xs):
        import torch
        return self.relu(self.bn(self.conv(torch.cat(xs, 1))))     # residual_root False for dla34


class Tree(nn.Module):
    def __init__(self, levels, cin, cout, stride=1, level_root=False, root_dim=0):
        super().__init__()
        if root_dim == 0:
            root_dim = 2 * cout
        if level_root:
            root_dim += cin
        if levels == 1:
            self.tree1 = BasicBlock(cin, cout, stride)
            self.tree2 = BasicBlock(cout, cout, 1)
            self.root = Root(root_dim, cout)
        else:
            self.tree1 = Tree(levels - 1, cin, cout, stride, root_dim=0)
            self.tree2 = Tree(levels - 1, cout, cout, root_dim=root_dim + cout)
        self.level_root, self.levels = level_root, levels
        self.downsample = nn.MaxPool2d(stride, stride=stride) if stride > 1 else None
        self.project = None
        if cin != cout:
            # NB for levels == 2 the projected residual is handed to an inner Tree that ignores it
            # (dla.py:217-230): the parameters exist (state_dict parity) but receive no gradient.
            self.project = nn.Sequential(nn.Conv2d(cin, cout, 1, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x, residual=None, children=None):
        children = [] if children is None else children
        bottom = self.downsample(x) if self.downsample else x
        residual = self.project(bottom) if self.project else bottom
        if self.level_root:
            children.append(bottom)
        x1 = self.tree1(x, residual)
        if self.levels == 1:
            return self.root(self.tree2(x1), x1, *children)
        children.append(x1)
        return self.tree2(x1, children=children)


class DLA34Backbone(Backbone):
    CH = [16, 32, 64, 128, 256, 512]

    def __init__(self):
        super().__init__()
        c = self.CH
        self.base_layer = nn.Sequential(*_cbr(3, c[0], 7))
        self.level0 = nn.Sequential(*_cbr(c[0], c[0], 3))
        self.level1 = nn.Sequential(*_cbr(c[0], c[1], 3, stride=2))
        self.level2 = Tree(1, c[1], c[2], 2, level_root=False)
        self.level3 = Tree(2, c[2], c[3], 2, level_root=True)
        self.level4 = Tree(2, c[3], c[4], 2, level_root=True)
        self.level5 = Tree(1, c[4], c[5], 2, level_root=True)
        for m in self.modules():                      # dla.py:262-268
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / n))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
        self._out_feature_channels = {"p2": 64, "p3": 128, "p4": 256, "p5": 512, "p6": 512}
        self._out_feature_strides = {"p2": 4, "p3": 8, "p4": 16, "p5": 32, "p6": 64}
        self._out_features = ["p2", "p3", "p4", "p5", "p6"]

    def forward(self, x):
        x = self.level1(self.level0(self.base_layer(x)))
        out = {}
        for i, name in zip(range(2, 6), ["p2", "p3", "p4", "p5"]):
            x = getattr(self, "level%d" % i)(x)
            out[name] = x
        out["p6"] = F.max_pool2d(x, kernel_size=1, stride=2, padding=0)
        return out


class TVResNetBackbone(Backbone):
    def __init__(self, depth):
        super().__init__()
        from torchvision import models
        assert depth in (18, 34), "oracle restates the BasicBlock torchvision ResNets (BASELINE: ResNet34)"
        base = getattr(models, "resnet%d" % depth)(weights=None)
        for k in ("conv1", "bn1", "relu", "maxpool", "layer1", "layer2", "layer3", "layer4"):
            setattr(self, k, getattr(base, k))
        self._out_feature_channels = {"p2": 64, "p3": 128, "p4": 256, "p5": 512, "p6": 512}
        self._out_feature_strides = {"p2": 4, "p3": 8, "p4": 16, "p5": 32, "p6": 64}
        self._out_features = ["p2", "p3", "p4", "p5", "p6"]

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        out = {}
        for name, layer in zip(["p2", "p3", "p4", "p5"], [self.layer1, self.layer2, self.layer3, self.layer4]):
            x = layer(x)
            out[name] = x
        out["p6"] = F.max_pool2d(x, kernel_size=1, stride=2, padding=0)
        return out


def build_backbone(cfg):
    name = cfg.MODEL.BACKBONE.NAME
    assert cfg.MODEL.WEIGHTS_PRETRAIN + cfg.MODEL.WEIGHTS != "", \
        "no network: set MODEL.WEIGHTS_PRETRAIN non-empty for random init (dla.py:494, resnet.py:76)"
    kw = dict(in_features=cfg.MODEL.FPN.IN_FEATURES, out_channels=cfg.MODEL.FPN.OUT_CHANNELS,
              norm=cfg.MODEL.FPN.NORM, fuse_type=cfg.MODEL.FPN.FUSE_TYPE)
    if name == "build_dla_from_vision_fpn_backbone":
        assert cfg.MODEL.DLA.TYPE == "dla34", "oracle restates dla34 only (BASELINE configs)"
        return FPN(bottom_up=DLA34Backbone(), **kw)
    if name == "build_resnet_from_vision_fpn_backbone":
        assert cfg.MODEL.RESNETS.TORCHVISION
        return FPN(bottom_up=TVResNetBackbone(cfg.MODEL.RESNETS.DEPTH), top_block=LastLevelMaxPool(), **kw)
    raise NotImplementedError(f"oracle: backbone '{name}' is out of scope (SURVEY.md section 2.1 #2)")
