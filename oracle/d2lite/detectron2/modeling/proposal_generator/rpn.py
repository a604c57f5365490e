"""StandardRPNHead + RPN.  The reference subclasses RPN at cubercnn/.../rpn.py:19."""
from typing import Dict, List

import torch
import torch.nn.functional as F
from torch import nn

from detectron2.config import configurable
from detectron2.layers import Conv2d, ShapeSpec, cat
from detectron2.structures import Boxes, ImageList, Instances, pairwise_iou
from detectron2.utils.events import get_event_storage
from detectron2.utils.registry import Registry

from ..anchor_generator import build_anchor_generator
from ..box_regression import Box2BoxTransform, _dense_box_regression_loss
from ..matcher import Matcher
from ..sampling import subsample_labels
from .build import PROPOSAL_GENERATOR_REGISTRY
from .proposal_utils import find_top_rpn_proposals

RPN_HEAD_REGISTRY = Registry("RPN_HEAD")


def build_rpn_head(cfg, input_shape):
    return RPN_HEAD_REGISTRY.get(cfg.MODEL.RPN.HEAD_NAME)(cfg, input_shape)


@RPN_HEAD_REGISTRY.register()
class StandardRPNHead(nn.Module):
    @configurable
    def __init__(self, *, in_channels, num_anchors, box_dim=4, conv_dims=(-1,)):
        super().__init__()
        cur = in_channels
        assert len(conv_dims) == 1, "d2lite: single-conv RPN head only (detectron2 default CONV_DIMS [-1])"
        out = cur if conv_dims[0] == -1 else conv_dims[0]
        self.conv = Conv2d(cur, out, kernel_size=3, stride=1, padding=1, activation=nn.ReLU())
        cur = out
        self.objectness_logits = nn.Conv2d(cur, num_anchors, kernel_size=1, stride=1)
        self.anchor_deltas = nn.Conv2d(cur, num_anchors * box_dim, kernel_size=1, stride=1)
        for layer in [self.conv, self.objectness_logits, self.anchor_deltas]:
            nn.init.normal_(layer.weight, std=0.01)
            nn.init.constant_(layer.bias, 0)

    @classmethod
    def from_config(cls, cfg, input_shape):
        in_channels = [s.channels for s in input_shape]
        assert len(set(in_channels)) == 1
        ag = build_anchor_generator(cfg, input_shape)
        assert len(set(ag.num_anchors)) == 1
        return {"in_channels": in_channels[0], "num_anchors": ag.num_anchors[0], "box_dim": ag.box_dim,
                "conv_dims": cfg.MODEL.RPN.CONV_DIMS}

    def forward(self, features):
        logits, deltas = [], []
        for x in features:
            t = self.conv(x)
            logits.append(self.objectness_logits(t))
            deltas.append(self.anchor_deltas(t))
        return logits, deltas


@PROPOSAL_GENERATOR_REGISTRY.register()
class RPN(nn.Module):
    @configurable
    def __init__(self, *, in_features, head, anchor_generator, anchor_matcher, box2box_transform,
                 batch_size_per_image, positive_fraction, pre_nms_topk, post_nms_topk, nms_thresh=0.7,
                 min_box_size=0.0, anchor_boundary_thresh=-1.0, loss_weight=1.0, box_reg_loss_type="smooth_l1",
                 smooth_l1_beta=0.0):
        super().__init__()
        self.in_features = in_features
        self.rpn_head = head
        self.anchor_generator = anchor_generator
        self.anchor_matcher = anchor_matcher
        self.box2box_transform = box2box_transform
        self.batch_size_per_image = batch_size_per_image
        self.positive_fraction = positive_fraction
        self.pre_nms_topk = {True: pre_nms_topk[0], False: pre_nms_topk[1]}
        self.post_nms_topk = {True: post_nms_topk[0], False: post_nms_topk[1]}
        self.nms_thresh = nms_thresh
        self.min_box_size = float(min_box_size)
        self.anchor_boundary_thresh = anchor_boundary_thresh
        if isinstance(loss_weight, float):
            loss_weight = {"loss_rpn_cls": loss_weight, "loss_rpn_loc": loss_weight}
        self.loss_weight = loss_weight
        self.box_reg_loss_type = box_reg_loss_type
        self.smooth_l1_beta = smooth_l1_beta

    @classmethod
    def from_config(cls, cfg, input_shape: Dict[str, ShapeSpec]):
        in_features = cfg.MODEL.RPN.IN_FEATURES
        ret = {
            "in_features": in_features,
            "min_box_size": cfg.MODEL.PROPOSAL_GENERATOR.MIN_SIZE,
            "nms_thresh": cfg.MODEL.RPN.NMS_THRESH,
            "batch_size_per_image": cfg.MODEL.RPN.BATCH_SIZE_PER_IMAGE,
            "positive_fraction": cfg.MODEL.RPN.POSITIVE_FRACTION,
            "loss_weight": {"loss_rpn_cls": cfg.MODEL.RPN.LOSS_WEIGHT,
                            "loss_rpn_loc": cfg.MODEL.RPN.BBOX_REG_LOSS_WEIGHT * cfg.MODEL.RPN.LOSS_WEIGHT},
            "anchor_boundary_thresh": cfg.MODEL.RPN.BOUNDARY_THRESH,
            "box2box_transform": Box2BoxTransform(weights=cfg.MODEL.RPN.BBOX_REG_WEIGHTS),
            "box_reg_loss_type": cfg.MODEL.RPN.BBOX_REG_LOSS_TYPE,
            "smooth_l1_beta": cfg.MODEL.RPN.SMOOTH_L1_BETA,
        }
        ret["pre_nms_topk"] = (cfg.MODEL.RPN.PRE_NMS_TOPK_TRAIN, cfg.MODEL.RPN.PRE_NMS_TOPK_TEST)
        ret["post_nms_topk"] = (cfg.MODEL.RPN.POST_NMS_TOPK_TRAIN, cfg.MODEL.RPN.POST_NMS_TOPK_TEST)
        shapes = [input_shape[f] for f in in_features]
        ret["anchor_generator"] = build_anchor_generator(cfg, shapes)
        ret["anchor_matcher"] = Matcher(cfg.MODEL.RPN.IOU_THRESHOLDS, cfg.MODEL.RPN.IOU_LABELS,
                                        allow_low_quality_matches=True)
        ret["head"] = build_rpn_head(cfg, shapes)
        return ret

    def _subsample_labels(self, label):
        pos_idx, neg_idx = subsample_labels(label, self.batch_size_per_image, self.positive_fraction, 0)
        label.fill_(-1)
        label.scatter_(0, pos_idx, 1)
        label.scatter_(0, neg_idx, 0)
        return label

    @torch.no_grad()
    def label_and_sample_anchors(self, anchors, gt_instances):
        anchors = Boxes.cat(anchors)
        gt_labels, matched_gt_boxes = [], []
        for gt_boxes_i in [x.gt_boxes for x in gt_instances]:
            mqm = pairwise_iou(gt_boxes_i, anchors)
            matched_idxs, gt_labels_i = self.anchor_matcher(mqm)
            gt_labels_i = self._subsample_labels(gt_labels_i)
            matched = torch.zeros_like(anchors.tensor) if len(gt_boxes_i) == 0 else gt_boxes_i[matched_idxs].tensor
            gt_labels.append(gt_labels_i)
            matched_gt_boxes.append(matched)
        return gt_labels, matched_gt_boxes

    def losses(self, anchors, pred_objectness_logits, gt_labels, pred_anchor_deltas, gt_boxes):
        num_images = len(gt_labels)
        gt_labels = torch.stack(gt_labels)
        pos_mask = gt_labels == 1
        storage = get_event_storage()
        storage.put_scalar("rpn/num_pos_anchors", pos_mask.sum().item() / num_images)
        storage.put_scalar("rpn/num_neg_anchors", (gt_labels == 0).sum().item() / num_images)
        loc = _dense_box_regression_loss(anchors, self.box2box_transform, pred_anchor_deltas, gt_boxes, pos_mask,
                                         box_reg_loss_type=self.box_reg_loss_type, smooth_l1_beta=self.smooth_l1_beta)
        valid_mask = gt_labels >= 0
        obj = F.binary_cross_entropy_with_logits(cat(pred_objectness_logits, dim=1)[valid_mask],
                                                 gt_labels[valid_mask].to(torch.float32), reduction="sum")
        normalizer = self.batch_size_per_image * num_images
        losses = {"loss_rpn_cls": obj / normalizer, "loss_rpn_loc": loc / normalizer}
        return {k: v * self.loss_weight.get(k, 1.0) for k, v in losses.items()}

    def forward(self, images: ImageList, features: Dict[str, torch.Tensor], gt_instances=None):
        features = [features[f] for f in self.in_features]
        anchors = self.anchor_generator(features)
        pred_objectness_logits, pred_anchor_deltas = self.rpn_head(features)
        pred_objectness_logits = [score.permute(0, 2, 3, 1).flatten(1) for score in pred_objectness_logits]
        pred_anchor_deltas = [
            x.view(x.shape[0], -1, self.anchor_generator.box_dim, x.shape[-2], x.shape[-1])
            .permute(0, 3, 4, 1, 2).flatten(1, -2) for x in pred_anchor_deltas]
        if self.training:
            assert gt_instances is not None, "RPN requires gt_instances in training!"
            gt_labels, gt_boxes = self.label_and_sample_anchors(anchors, gt_instances)
            losses = self.losses(anchors, pred_objectness_logits, gt_labels, pred_anchor_deltas, gt_boxes)
        else:
            losses = {}
        proposals = self.predict_proposals(anchors, pred_objectness_logits, pred_anchor_deltas, images.image_sizes)
        return proposals, losses

    def predict_proposals(self, anchors, pred_objectness_logits, pred_anchor_deltas, image_sizes):
        with torch.no_grad():
            pred_proposals = self._decode_proposals(anchors, pred_anchor_deltas)
            return find_top_rpn_proposals(pred_proposals, pred_objectness_logits, image_sizes, self.nms_thresh,
                                          self.pre_nms_topk[self.training], self.post_nms_topk[self.training],
                                          self.min_box_size, self.training)

    def _decode_proposals(self, anchors, pred_anchor_deltas):
        N = pred_anchor_deltas[0].shape[0]
        proposals = []
        for anchors_i, deltas_i in zip(anchors, pred_anchor_deltas):
            B = anchors_i.tensor.size(1)
            deltas_i = deltas_i.reshape(-1, B)
            anchors_e = anchors_i.tensor.unsqueeze(0).expand(N, -1, -1).reshape(-1, B)
            proposals.append(self.box2box_transform.apply_deltas(deltas_i, anchors_e).view(N, -1, B))
        return proposals
