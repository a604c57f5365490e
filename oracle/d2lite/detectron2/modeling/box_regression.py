import math

import torch
from fvcore.nn import smooth_l1_loss

from detectron2.layers import cat
from detectron2.structures import Boxes

_DEFAULT_SCALE_CLAMP = math.log(1000.0 / 16)


class Box2BoxTransform:
    def __init__(self, weights, scale_clamp=_DEFAULT_SCALE_CLAMP):
        self.weights = weights
        self.scale_clamp = scale_clamp

    def get_deltas(self, src_boxes, target_boxes):
        src_w = src_boxes[:, 2] - src_boxes[:, 0]
        src_h = src_boxes[:, 3] - src_boxes[:, 1]
        src_cx = src_boxes[:, 0] + 0.5 * src_w
        src_cy = src_boxes[:, 1] + 0.5 * src_h
        tw = target_boxes[:, 2] - target_boxes[:, 0]
        th = target_boxes[:, 3] - target_boxes[:, 1]
        tcx = target_boxes[:, 0] + 0.5 * tw
        tcy = target_boxes[:, 1] + 0.5 * th
        wx, wy, ww, wh = self.weights
        dx = wx * (tcx - src_cx) / src_w
        dy = wy * (tcy - src_cy) / src_h
        dw = ww * torch.log(tw / src_w)
        dh = wh * torch.log(th / src_h)
        deltas = torch.stack((dx, dy, dw, dh), dim=1)
        assert (src_w > 0).all().item(), "Input boxes to Box2BoxTransform are not valid!"
        return deltas

    def apply_deltas(self, deltas, boxes):
        deltas = deltas.float()
        boxes = boxes.to(deltas.dtype)
        w = boxes[:, 2] - boxes[:, 0]
        h = boxes[:, 3] - boxes[:, 1]
        cx = boxes[:, 0] + 0.5 * w
        cy = boxes[:, 1] + 0.5 * h
        wx, wy, ww, wh = self.weights
        dx = deltas[:, 0::4] / wx
        dy = deltas[:, 1::4] / wy
        dw = torch.clamp(deltas[:, 2::4] / ww, max=self.scale_clamp)
        dh = torch.clamp(deltas[:, 3::4] / wh, max=self.scale_clamp)
        pcx = dx * w[:, None] + cx[:, None]
        pcy = dy * h[:, None] + cy[:, None]
        pw = torch.exp(dw) * w[:, None]
        ph = torch.exp(dh) * h[:, None]
        x1, y1, x2, y2 = pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph
        return torch.stack((x1, y1, x2, y2), dim=-1).reshape(deltas.shape)


def _dense_box_regression_loss(anchors, box2box_transform, pred_anchor_deltas, gt_boxes, fg_mask,
                               box_reg_loss_type="smooth_l1", smooth_l1_beta=0.0):
    anchors = type(anchors[0]).cat(anchors).tensor if isinstance(anchors[0], Boxes) else cat(anchors)
    assert box_reg_loss_type == "smooth_l1"
    gt_anchor_deltas = torch.stack([box2box_transform.get_deltas(anchors, k) for k in gt_boxes])
    return smooth_l1_loss(cat(pred_anchor_deltas, dim=1)[fg_mask], gt_anchor_deltas[fg_mask],
                          beta=smooth_l1_beta, reduction="sum")
