"""ROIHeads / StandardROIHeads plumbing; subclassed at cubercnn/.../roi_heads.py:40."""
from typing import Dict

import torch
from torch import nn

from detectron2.config import configurable
from detectron2.layers import ShapeSpec
from detectron2.utils.registry import Registry

from ..matcher import Matcher
from ..poolers import ROIPooler
from .box_head import build_box_head
from .fast_rcnn import FastRCNNOutputLayers

ROI_HEADS_REGISTRY = Registry("ROI_HEADS")


def build_roi_heads(cfg, input_shape):
    return ROI_HEADS_REGISTRY.get(cfg.MODEL.ROI_HEADS.NAME)(cfg, input_shape)


def select_foreground_proposals(proposals, bg_label):
    assert isinstance(proposals, (list, tuple))
    fg_proposals, fg_selection_masks = [], []
    for p in proposals:
        gt_classes = p.gt_classes
        fg_selection_mask = (gt_classes != -1) & (gt_classes != bg_label)
        fg_idxs = fg_selection_mask.nonzero().squeeze(1)
        fg_proposals.append(p[fg_idxs])
        fg_selection_masks.append(fg_selection_mask)
    return fg_proposals, fg_selection_masks


class ROIHeads(nn.Module):
    @configurable
    def __init__(self, *, num_classes, batch_size_per_image, positive_fraction, proposal_matcher,
                 proposal_append_gt=True):
        super().__init__()
        self.batch_size_per_image = batch_size_per_image
        self.positive_fraction = positive_fraction
        self.num_classes = num_classes
        self.proposal_matcher = proposal_matcher
        self.proposal_append_gt = proposal_append_gt

    @classmethod
    def from_config(cls, cfg):
        return {"batch_size_per_image": cfg.MODEL.ROI_HEADS.BATCH_SIZE_PER_IMAGE,
                "positive_fraction": cfg.MODEL.ROI_HEADS.POSITIVE_FRACTION,
                "num_classes": cfg.MODEL.ROI_HEADS.NUM_CLASSES,
                "proposal_append_gt": cfg.MODEL.ROI_HEADS.PROPOSAL_APPEND_GT,
                "proposal_matcher": Matcher(cfg.MODEL.ROI_HEADS.IOU_THRESHOLDS, cfg.MODEL.ROI_HEADS.IOU_LABELS,
                                            allow_low_quality_matches=False)}


@ROI_HEADS_REGISTRY.register()
class StandardROIHeads(ROIHeads):
    @configurable
    def __init__(self, *, box_in_features, box_pooler, box_head, box_predictor, mask_in_features=None,
                 mask_pooler=None, mask_head=None, keypoint_in_features=None, keypoint_pooler=None,
                 keypoint_head=None, train_on_pred_boxes=False, **kwargs):
        super().__init__(**kwargs)
        self.in_features = self.box_in_features = box_in_features
        self.box_pooler = box_pooler
        self.box_head = box_head
        self.box_predictor = box_predictor
        self.mask_on = self.keypoint_on = False
        self.train_on_pred_boxes = train_on_pred_boxes

    @classmethod
    def from_config(cls, cfg, input_shape):
        ret = super().from_config(cfg)
        ret["train_on_pred_boxes"] = cfg.MODEL.ROI_BOX_HEAD.TRAIN_ON_PRED_BOXES
        assert not cfg.MODEL.MASK_ON and not cfg.MODEL.KEYPOINT_ON, "d2lite: box head only (Base.yaml:35)"
        ret.update(cls._init_box_head(cfg, input_shape))
        return ret

    @classmethod
    def _init_box_head(cls, cfg, input_shape: Dict[str, ShapeSpec]):
        in_features = cfg.MODEL.ROI_HEADS.IN_FEATURES
        pooler_resolution = cfg.MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION
        pooler_scales = tuple(1.0 / input_shape[k].stride for k in in_features)
        in_channels = [input_shape[f].channels for f in in_features]
        assert len(set(in_channels)) == 1, in_channels
        box_pooler = ROIPooler(output_size=pooler_resolution, scales=pooler_scales,
                               sampling_ratio=cfg.MODEL.ROI_BOX_HEAD.POOLER_SAMPLING_RATIO,
                               pooler_type=cfg.MODEL.ROI_BOX_HEAD.POOLER_TYPE)
        box_head = build_box_head(cfg, ShapeSpec(channels=in_channels[0], height=pooler_resolution,
                                                 width=pooler_resolution))
        box_predictor = FastRCNNOutputLayers(cfg, box_head.output_shape)
        return {"box_in_features": in_features, "box_pooler": box_pooler, "box_head": box_head,
                "box_predictor": box_predictor}
