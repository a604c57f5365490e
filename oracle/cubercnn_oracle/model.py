"""ORACLE restatement of the Cube R-CNN model (default configuration path only):
  RCNN3D            cubercnn/modeling/meta_arch/rcnn3d.py:25-112, 247-272
  RPNWithIgnore     cubercnn/modeling/proposal_generator/rpn.py:19-354
  FastRCNNOutputs   cubercnn/modeling/roi_heads/fast_rcnn.py:16-260
  CubeHead          cubercnn/modeling/roi_heads/cube_head.py:19-202
  ROIHeads3D        cubercnn/modeling/roi_heads/roi_heads.py:39-941
Own code, written against the behaviour of those files (not copied); state_dict keys, module
construction order (=> same-seed init), RNG consumption order and outputs match the reference run.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F
from detectron2.layers import ShapeSpec, batched_nms, cat, cross_entropy, nonzero_tuple
from detectron2.modeling.box_regression import Box2BoxTransform
from detectron2.modeling.meta_arch.rcnn import GeneralizedRCNN
from detectron2.modeling.poolers import ROIPooler
from detectron2.modeling.proposal_generator.proposal_utils import add_ground_truth_to_proposals
from detectron2.modeling.proposal_generator.rpn import RPN
from detectron2.modeling.roi_heads import StandardROIHeads, select_foreground_proposals
from detectron2.modeling.roi_heads.fast_rcnn import FastRCNNOutputLayers, _log_classification_stats
from detectron2.structures import Boxes, Instances, pairwise_ioa, pairwise_iou
from detectron2.utils.events import get_event_storage
from fvcore.nn import weight_init
from pytorch3d.transforms import rotation_6d_to_matrix
from torch import nn

from .backbone import build_backbone
from .geometry import R_from_allocentric, cuboid_corners, virtual_scale

SQRT2 = 1.41421356


# ------------------------------------------------------------------------------------------------
def iou_weighted_subsample(labels, num_samples, positive_fraction, bg_label, matched_ious, eps=1e-4):
    """rpn.py:275-328: multinomial sampling weighted by matched IoU (+eps), positives first."""
    positive = nonzero_tuple((labels != -1) & (labels != bg_label))[0]
    negative = nonzero_tuple(labels == bg_label)[0]
    num_pos = min(positive.numel(), int(num_samples * positive_fraction))
    num_neg = min(negative.numel(), num_samples - num_pos)
    if num_pos > 0:
        perm1 = torch.multinomial(matched_ious[positive] + eps, num_pos)
    else:
        perm1 = torch.randperm(positive.numel(), device=positive.device)[:num_pos]
    if num_neg > 0:
        perm2 = torch.multinomial(matched_ious[negative] + eps, num_neg)
    else:
        perm2 = torch.randperm(negative.numel(), device=negative.device)[:num_neg]
    return positive[perm1], negative[perm2]


def paired_iou(b1, b2):
    """rpn.py:330-354 (no zero guard on the union)."""
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    wh = (torch.min(b1[:, 2:], b2[:, 2:]) - torch.max(b1[:, :2], b2[:, :2])).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    return inter / (a1 + a2 - inter)


class RPNWithIgnore(RPN):
    def __init__(self, cfg, input_shape):
        assert cfg.MODEL.RPN.OBJECTNESS_UNCERTAINTY == "IoUness", "oracle restates the IoUness RPN only"
        super().__init__(cfg, input_shape)
        self.ignore_thresh = cfg.MODEL.RPN.IGNORE_THRESHOLD

    @torch.no_grad()
    def label_and_sample_anchors(self, anchors, gt_instances):
        anchors = Boxes.cat(anchors)
        labels_out, boxes_out = [], []
        for x in gt_instances:
            gt_i, ign_i = x.gt_boxes[x.gt_classes >= 0], x.gt_boxes[x.gt_classes < 0]
            mqm = pairwise_iou(gt_i, anchors)
            matched_idxs, lab = self.anchor_matcher(mqm)
            matched_ious = mqm[matched_idxs, torch.arange(mqm.shape[1], device=mqm.device)]
            # the best anchor of every GT is forced positive if the matcher labelled it positive (:71-84)
            best = mqm.max(dim=1)[1] if mqm.shape[0] else torch.zeros(0, dtype=torch.long)
            best = torch.tensor(sorted(set(best.tolist()) & set((lab == 1).nonzero().squeeze(1).tolist())),
                                dtype=torch.long, device=lab.device)
            pos, neg = iou_weighted_subsample(lab, self.batch_size_per_image, self.positive_fraction, 0, matched_ious)
            lab.fill_(-1)
            lab.scatter_(0, pos, 1)
            lab.scatter_(0, neg, 0)
            if best.numel() > 0:
                lab[best] = 1
            matched = torch.zeros_like(anchors.tensor) if len(gt_i) == 0 else gt_i[matched_idxs].tensor
            if len(ign_i) > 0:
                bg = (lab == 0).nonzero().squeeze()
                if bg.numel() > 1:
                    ioa = pairwise_ioa(ign_i, anchors[bg])
                    lab[bg[ioa.max(0)[0] >= self.ignore_thresh]] = -1
            labels_out.append(lab)
            boxes_out.append(matched)
        return labels_out, boxes_out

    def losses(self, anchors, pred_objectness_logits, gt_labels, pred_anchor_deltas, gt_boxes):
        num_images = len(gt_labels)
        gt_labels = torch.stack(gt_labels)
        pos = gt_labels == 1
        storage = get_event_storage()
        storage.put_scalar("rpn/num_pos_anchors", pos.sum().item() / num_images)
        storage.put_scalar("rpn/num_neg_anchors", (gt_labels == 0).sum().item() / num_images)
        A = Boxes.cat(anchors).tensor
        n = len(gt_boxes)
        a_fg = A.unsqueeze(0).repeat([n, 1, 1])[pos]
        g_fg = torch.stack(gt_boxes)[pos].detach()
        target = paired_iou(a_fg, g_fg).detach()                     # "IoUness" target (:233)
        logits = torch.cat(pred_objectness_logits, dim=1)
        conf = (F.binary_cross_entropy_with_logits(logits[pos], target, reduction="none") * target).sum()
        storage.put_scalar("rpn/conf_pos_anchors", torch.sigmoid(logits[pos]).mean().item())
        storage.put_scalar("rpn/conf_neg_anchors", torch.sigmoid(logits[~pos]).mean().item())
        gt_deltas = torch.stack([self.box2box_transform.get_deltas(A, k) for k in gt_boxes])
        l1 = (cat(pred_anchor_deltas, dim=1)[pos] - gt_deltas[pos]).abs()   # smooth-L1 with beta 0
        loc = (l1.sum(dim=1) * target).sum()
        norm = self.batch_size_per_image * num_images
        return {"rpn/cls": conf / norm, "rpn/loc": loc / norm}


# ------------------------------------------------------------------------------------------------
class FastRCNNOutputs(FastRCNNOutputLayers):
    def losses(self, predictions, proposals):
        scores, deltas = predictions
        gt_classes = cat([p.gt_classes for p in proposals], dim=0)
        boxes = cat([p.proposal_boxes.tensor for p in proposals], dim=0)
        gt_boxes = cat([(p.gt_boxes if p.has("gt_boxes") else p.proposal_boxes).tensor for p in proposals], dim=0)
        _log_classification_stats(scores, gt_classes)
        loss_cls = cross_entropy(scores, gt_classes, reduction="mean")
        fg = nonzero_tuple((gt_classes >= 0) & (gt_classes < self.num_classes))[0]
        fg_deltas = deltas.view(-1, self.num_classes, 4)[fg, gt_classes[fg]]
        target = self.box2box_transform.get_deltas(boxes[fg], gt_boxes[fg])
        loss_box = (fg_deltas - target).abs().sum() / max(gt_classes.numel(), 1.0)
        return {"BoxHead/loss_cls": loss_cls, "BoxHead/loss_box_reg": loss_box}

    def inference(self, predictions, proposals):
        boxes = self.predict_boxes(predictions, proposals)
        scores = self.predict_probs(predictions, proposals)
        out = [self._inference_one(b, s, p.image_size) for b, s, p in zip(boxes, scores, proposals)]
        return [o[0] for o in out], [o[1] for o in out]

    def _inference_one(self, boxes, scores, image_shape):
        valid = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores).all(dim=1)
        if not valid.all():
            boxes, scores = boxes[valid], scores[valid]
        scores = scores[:, :-1]
        K = boxes.shape[1] // 4
        b = Boxes(boxes.reshape(-1, 4))
        b.clip(image_shape)
        boxes = b.tensor.view(-1, K, 4)
        mask = scores > self.test_score_thresh
        inds = mask.nonzero()
        boxes = boxes[mask]
        scores_full = scores[inds[:, 0]]
        scores = scores[mask]
        keep = batched_nms(boxes, scores, inds[:, 1], self.test_nms_thresh)
        if self.test_topk_per_image >= 0:
            keep = keep[: self.test_topk_per_image]
        res = Instances(image_shape)
        res.pred_boxes = Boxes(boxes[keep])
        res.scores = scores[keep]
        res.scores_full = scores_full[keep]
        res.pred_classes = inds[keep][:, 1]
        return res, inds[keep][:, 0]


class CubeHead(nn.Module):
    def __init__(self, cfg, input_shape):
        super().__init__()
        H = cfg.MODEL.ROI_CUBE_HEAD
        assert H.SHARED_FC and H.POSE_TYPE == "6d" and H.CLUSTER_BINS == 1 and H.NUM_CONV == 0 and H.USE_CONFIDENCE
        K = self.num_classes = cfg.MODEL.ROI_HEADS.NUM_CLASSES
        self.feature_generator = nn.Sequential()
        d = input_shape.channels * input_shape.height * input_shape.width
        for k in range(H.NUM_FC):
            fc = nn.Linear(d, H.FC_DIM)
            weight_init.c2_xavier_fill(fc)
            self.feature_generator.add_module("fc%d" % (k + 1), fc)
            self.feature_generator.add_module("fc_relu%d" % (k + 1), nn.ReLU())
            d = H.FC_DIM
        # construction order == cube_head.py:108-144 (dims, center deltas, pose, depth, uncertainty)
        for name, mult, bias in (("bbox_3D_dims", 3, 0), ("bbox_3D_center_deltas", 2, 0), ("bbox_3D_pose", 6, 0),
                                 ("bbox_3D_center_depth", 1, 0), ("bbox_3D_uncertainty", 1, 5)):
            lin = nn.Linear(d, K * mult)
            nn.init.normal_(lin.weight, std=0.001)
            nn.init.constant_(lin.bias, bias)
            setattr(self, name, lin)

    def forward(self, x):
        n, K = x.shape[0], self.num_classes
        f = self.feature_generator(x)
        deltas = self.bbox_3D_center_deltas(f).view(n, K, 2)
        dims = self.bbox_3D_dims(f).view(n, K, 3)
        pose = rotation_6d_to_matrix(self.bbox_3D_pose(f).view(-1, 6)).view(n, K, 3, 3)
        z = self.bbox_3D_center_depth(f).view(n, K, -1)
        uncert = self.bbox_3D_uncertainty(f).clip(0.01)
        return deltas, z, dims, pose, uncert


def finite_mean(loss):
    """roi_heads.py:932-941 safely_reduce_losses."""
    ok = (~loss.isinf()) & (~loss.isnan())
    return loss[ok].mean() if ok.any() else loss.mean() * 0.0


def chamfer8(a, b):
    d = (a.view(-1, 8, 1, 3) - b.view(-1, 1, 8, 3)).abs().sum(-1)
    return d.min(1).values.mean(-1) + d.min(2).values.mean(-1)


class ROIHeads3D(StandardROIHeads):
    def __init__(self, cfg, input_shape, priors=None):
        H = cfg.MODEL.ROI_CUBE_HEAD
        assert H.DISENTANGLED_LOSS and H.CHAMFER_POSE and H.ALLOCENTRIC_POSE and H.VIRTUAL_DEPTH
        assert H.Z_TYPE == "direct" and H.DIMS_PRIORS_ENABLED and H.DIMS_PRIORS_FUNC == "exp"
        assert not H.INVERSE_Z_WEIGHT and H.SCALE_ROI_BOXES == 0.0 and H.LOSS_W_3D > 0 and H.LOSS_W_JOINT > 0
        kw = StandardROIHeads.from_config(cfg, input_shape)
        kw["box_predictor"] = FastRCNNOutputs(cfg, kw["box_head"].output_shape)   # 2nd construction as at :151
        in_features = cfg.MODEL.ROI_HEADS.IN_FEATURES
        res = H.POOLER_RESOLUTION
        cube_pooler = ROIPooler(output_size=res, scales=tuple(1.0 / input_shape[k].stride for k in in_features),
                                sampling_ratio=H.POOLER_SAMPLING_RATIO, pooler_type=H.POOLER_TYPE)
        cube_head = CubeHead(cfg, ShapeSpec(channels=input_shape[in_features[0]].channels, width=res, height=res))
        super().__init__(**kw)
        self.cube_head, self.cube_pooler = cube_head, cube_pooler
        self.ignore_thresh = cfg.MODEL.RPN.IGNORE_THRESHOLD
        self.virtual_focal = H.VIRTUAL_FOCAL
        self.w = dict(w3d=H.LOSS_W_3D, xy=H.LOSS_W_XY, z=H.LOSS_W_Z, dims=H.LOSS_W_DIMS, pose=H.LOSS_W_POSE,
                      joint=H.LOSS_W_JOINT, conf=H.USE_CONFIDENCE)
        if priors is not None:
            self.priors_dims_per_cat = nn.Parameter(torch.FloatTensor(priors["priors_dims_per_cat"]).unsqueeze(0))
        else:
            self.priors_dims_per_cat = nn.Parameter(torch.ones(1, self.num_classes, 2, 3))
        self.priors_z_scales = nn.Parameter(torch.ones(self.num_classes, H.CLUSTER_BINS))

    # -- sampling ---------------------------------------------------------------------------------
    @torch.no_grad()
    def label_and_sample_proposals(self, proposals, targets):
        ign = [t[t.gt_classes < 0] for t in targets]
        targets = [t[t.gt_classes >= 0] for t in targets]
        if self.proposal_append_gt:
            proposals = add_ground_truth_to_proposals(targets, proposals)
        out, nfg, nbg = [], [], []
        for prop, tgt, tign in zip(proposals, targets, ign):
            has_gt = len(tgt) > 0
            mqm = pairwise_iou(tgt.gt_boxes, prop.proposal_boxes)
            midx, mlab = self.proposal_matcher(mqm)
            if len(tign) > 0:
                bg = (mlab == 0).nonzero().squeeze()
                if bg.numel() > 1:
                    ioa = pairwise_ioa(tign.gt_boxes, prop.proposal_boxes[bg])
                    mlab[bg[ioa.max(0)[0] >= self.ignore_thresh]] = -1
            mious = mqm[midx, torch.arange(mqm.shape[1], device=mqm.device)]
            if has_gt:
                cls = tgt.gt_classes[midx]
                cls[mlab == 0] = self.num_classes
                cls[mlab == -1] = -1
            else:
                cls = torch.zeros_like(midx) + self.num_classes
            fg_i, bg_i = iou_weighted_subsample(cls, self.batch_size_per_image, self.positive_fraction,
                                                self.num_classes, mious)
            sel = torch.cat([fg_i, bg_i], dim=0)
            prop = prop[sel]
            prop.gt_classes = cls[sel]
            if has_gt:
                for name, val in tgt.get_fields().items():
                    if name.startswith("gt_") and not prop.has(name):
                        prop.set(name, val[midx[sel]])
            nbg.append((prop.gt_classes == self.num_classes).sum().item())
            nfg.append(prop.gt_classes.numel() - nbg[-1])
            out.append(prop)
        storage = get_event_storage()
        storage.put_scalar("roi_head/num_fg_samples", np.mean(nfg))
        storage.put_scalar("roi_head/num_bg_samples", np.mean(nbg))
        return out

    # -- forward ----------------------------------------------------------------------------------
    def forward(self, images, features, proposals, Ks, im_scales_ratio, targets=None):
        im_dims = [image.shape[1:] for image in images]
        if self.training:
            proposals = self.label_and_sample_proposals(proposals, targets)
            losses = self._forward_box(features, proposals)
            inst, lc = self._forward_cube(features, proposals, Ks, im_dims, im_scales_ratio)
            losses.update(lc)
            return inst, losses
        assert all(isinstance(p, Instances) for p in proposals), "oracle: oracle2D bypass not restated"
        pred = self._forward_box(features, proposals)
        return self._forward_cube(features, pred, Ks, im_dims, im_scales_ratio), {}

    def _forward_box(self, features, proposals):
        feats = [features[f] for f in self.box_in_features]
        x = self.box_head(self.box_pooler(feats, [p.proposal_boxes for p in proposals]))
        pred = self.box_predictor(x)
        if not self.training:
            return self.box_predictor.inference(pred, proposals)[0]
        losses = self.box_predictor.losses(pred, proposals)
        for p, b in zip(proposals, self.box_predictor.predict_boxes_for_gt_classes(pred, proposals)):
            p.pred_boxes = Boxes(b)
        return losses

    def _forward_cube(self, features, instances, Ks, im_dims, im_scales_ratio):
        feats = [features[f] for f in self.in_features]
        if self.training:
            proposals, _ = select_foreground_proposals(instances, self.num_classes)
            boxes = [p.proposal_boxes for p in proposals]
            classes = torch.cat([p.gt_classes for p in proposals], dim=0)
            gt3 = torch.cat([p.gt_boxes3D for p in proposals], dim=0)
            gtR = torch.cat([p.gt_poses for p in proposals], dim=0)
        else:
            proposals = instances
            boxes = [p.pred_boxes for p in instances]
            classes = torch.cat([p.pred_classes for p in instances])
        x = self.cube_pooler(feats, boxes).flatten(1)
        n = x.shape[0]
        if n == 0:
            return instances if not self.training else (instances, {})
        counts = [len(p) for p in proposals]
        rep = lambda vals: torch.cat([torch.as_tensor(v, dtype=torch.float32).reshape(1, -1).repeat(c, 1)
                                      for v, c in zip(vals, counts)]).to(x.device)      # (the reference .cuda()s these)
        Kb = rep([(Ks[i] / im_scales_ratio[i]).reshape(-1) for i in range(len(Ks))]).view(n, 3, 3)
        Kb[:, -1, -1] = 1
        focal = rep([Ks[i][1, 1] for i in range(len(Ks))]).squeeze(1)
        ratio = rep(im_scales_ratio).squeeze(1)
        im_h = rep([d[0] for d in im_dims]).squeeze(1)
        v2r = virtual_scale(focal, im_h * ratio, self.virtual_focal, im_h)
        src = torch.cat([b.tensor for b in boxes], dim=0)
        sw, sh = src[:, 2] - src[:, 0], src[:, 3] - src[:, 1]
        scx, scy = src[:, 0] + 0.5 * sw, src[:, 1] + 0.5 * sh
        d2, z, dims, pose, uncert = self.cube_head(x)
        ar = torch.arange(n, device=x.device)
        z, dims, pose, uncert, d2 = z[ar, classes, :], dims[ar, classes, :], pose[ar, classes], uncert[ar, classes], \
            d2[ar, classes, :]
        cx, cy = scx + sw * d2[:, 0], scy + sh * d2[:, 1]
        cxy = torch.stack((cx, cy), dim=1)
        prior_mean = self.priors_dims_per_cat.detach().repeat([n, 1, 1, 1])[ar, classes][:, 0, :]
        dims = torch.exp(dims.clip(max=5)) * prior_mean
        pose = R_from_allocentric(Kb, pose, u=cx.detach(), v=cy.detach())
        z = z.squeeze() * v2r
        fx, fy, px, py = Kb[:, 0, 0], Kb[:, 1, 1], Kb[:, 0, 2], Kb[:, 1, 2]
        losses = {}
        if self.training:
            storage = get_event_storage()
            g2, gz, gdims = gt3[:, :2], gt3[:, 2], gt3[:, 3:6]
            lift = lambda zz, uu, vv: torch.stack((zz * (uu - px) / fx, zz * (vv - py) / fy, zz)).T
            g3 = lift(gz, g2[:, 0], g2[:, 1])
            gt_c = cuboid_corners(torch.cat((g3, gdims), dim=1), gtR)
            c_z = cuboid_corners(torch.cat((lift(z, g2[:, 0], g2[:, 1]), gdims), dim=1), gtR)
            c_xy = cuboid_corners(torch.cat((lift(gz, cx, cy), gdims), dim=1), gtR)
            c_pose = cuboid_corners(torch.cat((g3, gdims), dim=1), pose)
            c_dims = cuboid_corners(torch.cat((g3, dims), dim=1), gtR)
            l1 = lambda a: (a - gt_c).abs().contiguous().view(n, -1).mean(dim=1)
            loss_xy, loss_dims, loss_z = l1(c_xy), l1(c_dims), l1(c_z)
            loss_pose = chamfer8(c_pose, gt_c)
            w = self.w
            total = loss_dims * w["dims"] + loss_pose * w["pose"] + loss_xy * w["xy"] + loss_z * w["z"]
            total = total.detach()
            c_joint = cuboid_corners(torch.cat((lift(z, cx, cy), dims), dim=1), pose)
            loss_joint = chamfer8(c_joint, gt_c)
            valid_joint = loss_joint < np.inf
            total = total + (loss_joint * w["joint"]).detach()
            z_err = (z - gz).detach().abs()
            put = lambda k, v: storage.put_scalar("Cube/" + k, v, smoothing_hint=False)
            put("z_error", z_err.mean().item())
            put("dims_error", (dims - gdims).detach().abs().mean().item())
            put("xy_error", (cxy - g2).detach().abs().mean().item())
            put("z_close", (z_err < 0.20).float().mean().item())
            put("total_3D_loss", w["w3d"] * finite_mean(total))
            sf = SQRT2 * torch.exp(-uncert)
            loss_dims, loss_xy, loss_z, loss_pose, loss_joint = (l * sf for l in
                                                                 (loss_dims, loss_xy, loss_z, loss_pose, loss_joint))
            losses["Cube/uncert"] = w["conf"] * finite_mean(uncert.clone())
            put("conf", torch.exp(-uncert).mean().item())
            self.batch_losses = [b.mean().item() for b in total.split(counts)]
            losses["Cube/loss_dims"] = finite_mean(loss_dims) * w["dims"] * w["w3d"]
            losses["Cube/loss_xy"] = finite_mean(loss_xy) * w["xy"] * w["w3d"]
            losses["Cube/loss_z"] = finite_mean(loss_z) * w["z"] * w["w3d"]
            losses["Cube/loss_pose"] = finite_mean(loss_pose) * w["pose"] * w["w3d"]
            if valid_joint.any():
                losses["Cube/loss_joint"] = finite_mean(loss_joint[valid_joint]) * w["joint"] * w["w3d"]
        if z.dim() == 0:
            z = z.unsqueeze(0)
        out3 = torch.cat((torch.stack((z * (cx - px) / fx, z * (cy - py) / fy, z)).T, dims, cxy * ratio.unsqueeze(1),
                          torch.exp(-uncert).unsqueeze(1)), dim=1)
        preds = instances if not self.training else [Instances(s) for s in im_dims]
        inst_boxes = [p.pred_boxes for p in proposals]
        for c3, R, inst, cls_i, box_i in zip(out3.split(counts), pose.split(counts), preds, classes.split(counts),
                                             inst_boxes):
            inst.scores = (inst.scores * c3[:, -1]) ** 0.5 if inst.has("scores") else c3[:, -1]
            if not inst.has("pred_classes"):
                inst.pred_classes = cls_i
            if not inst.has("pred_boxes"):
                inst.pred_boxes = box_i
            inst.pred_bbox3D = cuboid_corners(c3[:, :6], R)
            inst.pred_center_cam = c3[:, :3]
            inst.pred_center_2D = c3[:, 6:8]
            inst.pred_dimensions = c3[:, 3:6]
            inst.pred_pose = R
        return (preds, losses) if self.training else preds


# ------------------------------------------------------------------------------------------------
class RCNN3D(GeneralizedRCNN):
    def __init__(self, cfg, priors=None):
        backbone = build_backbone(cfg)
        shape = backbone.output_shape()
        assert cfg.MODEL.PROPOSAL_GENERATOR.NAME == "RPNWithIgnore" and cfg.MODEL.ROI_HEADS.NAME == "ROIHeads3D"
        rpn = RPNWithIgnore(cfg, shape)
        heads = ROIHeads3D(cfg, shape, priors=priors)
        super().__init__(backbone=backbone, proposal_generator=rpn, roi_heads=heads,
                         pixel_mean=cfg.MODEL.PIXEL_MEAN, pixel_std=cfg.MODEL.PIXEL_STD,
                         input_format=cfg.INPUT.FORMAT, vis_period=0)

    def forward(self, batched_inputs):
        images = self.preprocess_image(batched_inputs)
        ratios = [info["height"] / im.shape[1] for info, im in zip(batched_inputs, images)]
        Ks = [torch.FloatTensor(info["K"]) for info in batched_inputs]
        features = self.backbone(images.tensor)
        if self.training:
            gt = [x["instances"].to(self.device) for x in batched_inputs]
            proposals, l_rpn = self.proposal_generator(images, features, gt)
            _, losses = self.roi_heads(images, features, proposals, Ks, ratios, gt)
            losses.update(l_rpn)
            return losses
        proposals, _ = self.proposal_generator(images, features, None)
        results, _ = self.roi_heads(images, features, proposals, Ks, ratios, None)
        return GeneralizedRCNN._postprocess(results, batched_inputs, images.image_sizes)


def build_model(cfg, priors=None):
    assert cfg.MODEL.META_ARCHITECTURE == "RCNN3D"
    model = RCNN3D(cfg, priors=priors)
    model.to(torch.device(cfg.MODEL.DEVICE))
    return model
