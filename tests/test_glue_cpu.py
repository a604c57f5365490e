"""CPU tests of the accelerated path's device-agnostic host logic (batched RPN/ROI glue) against the oracle's
per-image detectron2-style code, on identical fp32 inputs: anchors, matcher labels, box coding, level
assignment, sampler validity.  (Kernels themselves are covered by the -m gpu tests.)"""
import torch

from omni3d_b200 import cubercnn as pc
from omni3d_b200 import synth
from omni3d_b200.cubercnn import rpn as prpn
from omni3d_b200.cubercnn.model import collate_gt
from omni3d_b200.cubercnn.roi_heads import assign_levels
from oracle import cubercnn_oracle as co
from oracle import model_io


def _cfgs():
    return (pc.load_cfg("cubercnn_DLA34_FPN.yaml", ["MODEL.DEVICE", "cpu", "MODEL.WEIGHTS_PRETRAIN", "none"]),
            co.load_cfg("cubercnn_DLA34_FPN.yaml"))


def test_anchors_and_matcher_bit_exact():
    from detectron2.modeling.anchor_generator import DefaultAnchorGenerator
    from detectron2.layers import ShapeSpec
    from detectron2.modeling.matcher import Matcher
    from detectron2.structures import Boxes, pairwise_iou
    pcfg, ocfg = _cfgs()
    shapes = [(40, 48), (20, 24), (10, 12), (5, 6), (3, 3)]
    strides = [4, 8, 16, 32, 64]
    ag = prpn.AnchorGenerator(pcfg.MODEL.ANCHOR_GENERATOR.SIZES, pcfg.MODEL.ANCHOR_GENERATOR.ASPECT_RATIOS, strides)
    mine = ag(shapes, torch.device("cpu"))
    oag = DefaultAnchorGenerator(ocfg, [ShapeSpec(stride=s) for s in strides])
    theirs = oag([torch.zeros(1, 1, h, w) for h, w in shapes])
    for a, b in zip(mine, theirs):
        assert torch.equal(a, b.tensor)
    anchors = torch.cat(mine)
    items = synth.make_batch(3, 160, 192, num_gt=5, seed=4)
    gt = collate_gt(items, torch.device("cpu"))
    head = prpn.RPNWithIgnore.__new__(prpn.RPNWithIgnore)
    head.iou_thresholds = [0.05, 0.05]
    valid = gt["present"] & (gt["classes"] >= 0)
    idx, miou, lab, best, _ = prpn.RPNWithIgnore.match_anchors(head, anchors, gt["boxes"], valid)
    m = Matcher([0.05, 0.05], [0, -1, 1], allow_low_quality_matches=True)
    for i, it in enumerate(model_io.to_d2_inputs(items)):
        inst = it["instances"]
        g = inst.gt_boxes[inst.gt_classes >= 0]
        mq = pairwise_iou(g, Boxes(anchors))
        oi, ol = m(mq)
        assert torch.equal(lab[i], ol)
        # same matched GT box (indices differ only by the removed ignore rows)
        assert torch.equal(gt["boxes"][i][idx[i]], g.tensor[oi])
        assert torch.equal(miou[i], mq[oi, torch.arange(mq.shape[1])])
        ob = set(mq.max(dim=1)[1].tolist()) & set((ol == 1).nonzero().squeeze(1).tolist())
        assert set(best[i].nonzero().squeeze(1).tolist()) == ob


def test_box_coding_matches_oracle():
    from detectron2.modeling.box_regression import Box2BoxTransform
    g = torch.Generator().manual_seed(0)
    src = torch.rand(50, 4, generator=g) * 100
    src[:, 2:] += src[:, :2] + 1
    tgt = torch.rand(50, 4, generator=g) * 100
    tgt[:, 2:] += tgt[:, :2] + 1
    w = (10.0, 10.0, 5.0, 5.0)
    t = Box2BoxTransform(w)
    assert torch.equal(prpn.get_deltas(src, tgt, w), t.get_deltas(src, tgt))
    d = torch.randn(50, 12, generator=g)
    assert torch.equal(prpn.apply_deltas(d, src, w), t.apply_deltas(d, src))


def test_level_assignment_matches_oracle():
    from detectron2.modeling.poolers import assign_boxes_to_levels
    from detectron2.structures import Boxes
    g = torch.Generator().manual_seed(1)
    b = torch.rand(500, 4, generator=g) * 300
    b[:, 2:] = b[:, :2] + torch.rand(500, 2, generator=g) * 600
    assert torch.equal(assign_levels(b).long(), assign_boxes_to_levels([Boxes(b)], 2, 6, 224, 4))


def test_gumbel_sampler_is_valid_weighted_sample():
    g = torch.Generator().manual_seed(0)
    w = torch.zeros(4, 1000)
    w[:, :300] = torch.rand(4, 300, generator=g) + 1e-4
    idx, ok = prpn.gumbel_topk_sample(w, 256, g)
    assert ok.all() and (idx < 300).all()
    assert all(len(set(r.tolist())) == 256 for r in idx)          # without replacement
    w[1, 10:] = 0                                                # fewer candidates than k
    idx, ok = prpn.gumbel_topk_sample(w, 256, g)
    assert ok[1].sum() == 10 and set(idx[1][ok[1]].tolist()) == set(range(10))
    # heavier weights are picked more often
    w2 = torch.ones(2000, 10); w2[:, 0] = 20
    first = prpn.gumbel_topk_sample(w2, 1, g)[0][:, 0]
    assert (first == 0).float().mean() > 0.55


def test_product_refuses_cpu_forward():
    import pytest
    from omni3d_b200 import _lib
    pcfg, _ = _cfgs()
    m = pc.build_model(pcfg)
    with pytest.raises(_lib.C3DError):
        m(synth.make_batch(1, 64, 64))


def test_pixel_stride_of_channel_slices():
    """gradients of torch.cat inputs arrive as channel slices of a dense NHWC buffer: the BatchNorm / max-pool backward
    kernels read them in place when 16-byte aligned, and fall back to a copy otherwise."""
    from omni3d_b200.kernels import pixel_stride
    big = torch.zeros(2, 5, 7, 448, dtype=torch.bfloat16)
    assert pixel_stride(big) == 448
    assert pixel_stride(big[..., 128:256]) == 448
    assert pixel_stride(big[..., 4:132]) is None                 # 8-byte offset: not vector aligned
    assert pixel_stride(big[:, :, 1:4, :128]) is None            # not a plain channel slice
    assert pixel_stride(big.permute(0, 3, 1, 2)) is None
    assert pixel_stride(torch.zeros(1, 1, 1, 64, dtype=torch.bfloat16)) == 64


def test_stage_inputs_is_the_only_host_interface():
    """RCNN3D.stage_inputs: images keep their dtype (uint8 as the mapper emits them, or float), one (B,12) row of
    per-image scalars [h, w, height/h, K row-major], padded GT with the collate conventions."""
    from omni3d_b200 import cubercnn as pc
    cfg = pc.load_cfg("cubercnn_DLA34_FPN.yaml", ["MODEL.WEIGHTS_PRETRAIN", "none", "MODEL.DEVICE", "cpu"])
    model = pc.build_model(cfg).train()
    items = synth.make_batch(2, 96, 128, num_gt=3, seed=3, image_dtype=torch.uint8)
    items[1]["gt"] = {k: v[:2] for k, v in items[1]["gt"].items()}            # ragged GT counts
    items[0]["height"] = 192                                                   # original image was 2x larger
    st = model.stage_inputs(items)
    assert [im.dtype for im in st["images"]] == [torch.uint8, torch.uint8] and st["sizes"] == [(96, 128), (96, 128)]
    meta = st["meta"]
    assert meta.shape == (2, 12) and meta[0, :3].tolist() == [96.0, 128.0, 2.0] and meta[1, 2].item() == 1.0
    assert torch.equal(meta[0, 3:].reshape(3, 3), torch.tensor(items[0]["K"], dtype=torch.float32))
    gt = st["gt"]
    assert gt["boxes"].shape == (2, 3, 4) and gt["present"].tolist() == [[True, True, True], [True, True, False]]
    assert gt["classes"][1, 2].item() == -2 and torch.equal(gt["poses"][1, 2], torch.eye(3))
    f = synth.make_batch(1, 64, 64, num_gt=2, seed=4)
    assert model.stage_inputs(f)["images"][0].dtype == torch.float32


# ---- the product's OWN proposal labelling against the oracle's pre-sampling outputs ------------------
def synthetic_proposals(items, P, seed):
    """(B,P,4) proposals that exercise every label: jittered copies of the valid GT (foreground at varying IoU), boxes
    inside the ignore regions, random background; per-image count < P (padding slots)."""
    g = torch.Generator().manual_seed(seed)
    B = len(items)
    boxes = torch.zeros(B, P, 4)
    counts = torch.zeros(B, dtype=torch.int32)
    for i, it in enumerate(items):
        H, W = it["image"].shape[1:]
        gb = it["gt"]["boxes"]
        n = P - 7 * i
        rows = []
        for j in range(n):
            r = j % 3
            if r == 0:                                   # jittered GT (valid or ignore alike)
                b = gb[torch.randint(0, len(gb), (1,), generator=g)[0]].clone()
                wh = (b[2:] - b[:2])
                b += (torch.rand(4, generator=g) - 0.5) * 0.5 * torch.cat([wh, wh])
            elif r == 1:                                 # small box inside some GT (IoA with an ignore region = 1)
                k = gb[torch.randint(0, len(gb), (1,), generator=g)[0]]
                c = k[:2] + torch.rand(2, generator=g) * (k[2:] - k[:2]) * 0.5
                b = torch.cat([c, c + (k[2:] - k[:2]) * 0.3])
            else:
                c = torch.rand(2, generator=g) * torch.tensor([W * 0.8, H * 0.8])
                b = torch.cat([c, c + torch.rand(2, generator=g) * torch.tensor([W * 0.2, H * 0.2]) + 2])
            b[0::2] = b[0::2].clamp(0, W)
            b[1::2] = b[1::2].clamp(0, H)
            if b[2] - b[0] < 1 or b[3] - b[1] < 1:
                b = torch.tensor([0.0, 0.0, 8.0, 8.0])
            rows.append(b)
        boxes[i, :n] = torch.stack(rows)
        counts[i] = n
    return boxes, counts


def oracle_prelabels(orc, items, boxes, counts):
    """the oracle's (= reference's roi_heads.py:862-929) matched GT boxes / IoUs / class labels of every proposal
    (appended GT included), captured right before its multinomial sampling."""
    from detectron2.structures import Boxes, Instances
    from detectron2.utils.events import EventStorage
    import oracle.cubercnn_oracle.model as om
    rh = orc.roi_heads
    cap = {"labels": [], "ious": [], "midx": []}
    o_sub, o_match = om.iou_weighted_subsample, rh.proposal_matcher

    def sub(labels, num, frac, bg, ious, eps=1e-4):
        cap["labels"].append(labels.clone()); cap["ious"].append(ious.clone())
        return o_sub(labels, num, frac, bg, ious, eps)

    class M:
        def __call__(self, q):
            mi, ml = o_match(q)
            cap["midx"].append(mi.clone())
            return mi, ml
    d2 = model_io.to_d2_inputs(items)
    props = []
    for i, it in enumerate(d2):
        p = Instances(it["instances"].image_size)
        n = int(counts[i])
        p.proposal_boxes = Boxes(boxes[i, :n].clone())
        p.objectness_logits = torch.zeros(n)
        props.append(p)
    om.iou_weighted_subsample, rh.proposal_matcher = sub, M()
    try:
        torch.manual_seed(0)
        with EventStorage(0):
            sampled = rh.label_and_sample_proposals(props, [it["instances"] for it in d2])
    finally:
        om.iou_weighted_subsample, rh.proposal_matcher = o_sub, o_match
    return cap, sampled, d2


def check_prelabels(midx, miou, cls, gt, cap, d2, counts, P):
    """product (B,P+G) outputs == oracle per-image outputs, bit for bit (same fp32 IoU formula)."""
    for i, it in enumerate(d2):
        inst = it["instances"]
        vmask = inst.gt_classes >= 0
        valid_pos = vmask.nonzero().squeeze(1)                 # padded-GT index of the oracle's k-th valid target
        n, nv = int(counts[i]), int(vmask.sum())
        sel = torch.cat([torch.arange(n), P + valid_pos])       # proposals, then the appended valid GT
        assert len(sel) == len(cap["labels"][i])
        assert torch.equal(cls[i][sel].cpu(), cap["labels"][i]), i
        assert torch.equal(miou[i][sel].cpu(), cap["ious"][i]), i
        vb = inst.gt_boxes.tensor[vmask]
        assert torch.equal(gt["boxes"][i].cpu()[midx[i][sel].cpu()], vb[cap["midx"][i]]), i
        # padding slots and appended non-valid GT are never candidates
        rest = torch.ones(cls.shape[1], dtype=torch.bool); rest[sel] = False
        assert (cls[i].cpu()[rest] == -1).all()


def _label_case(seed, num_gt=6, extra_ignore=True):
    items = synth.make_batch(3, 160, 192, num_gt=num_gt, seed=seed)
    if extra_ignore:                                            # a second ignore region in image 0, none in image 2
        items[0]["gt"]["classes"][0] = -1
        items[2]["gt"]["classes"][-1] = 3
    return items


def test_match_proposals_equals_oracle_prelabels():
    pcfg, ocfg = _cfgs()
    torch.manual_seed(0)
    orc = co.build_model(ocfg)
    from omni3d_b200.cubercnn.roi_heads import ROIHeads3D
    rh = ROIHeads3D.__new__(ROIHeads3D)
    RH = pcfg.MODEL.ROI_HEADS
    rh.num_classes, rh.iou_thresh, rh.ignore_thresh = RH.NUM_CLASSES, RH.IOU_THRESHOLDS[0], pcfg.MODEL.RPN.IGNORE_THRESHOLD
    P = 96
    for seed in (3, 4):
        items = _label_case(seed)
        boxes, counts = synthetic_proposals(items, P, seed)
        cap, _, d2 = oracle_prelabels(orc, items, boxes, counts)
        gt = collate_gt(items, torch.device("cpu"))
        pvalid = torch.arange(P)[None] < counts[:, None]
        allb = torch.cat([boxes, gt["boxes"]], 1)
        allv = torch.cat([pvalid, gt["present"] & (gt["classes"] >= 0)], 1)
        midx, miou, cls = ROIHeads3D.match_proposals(rh, allb, allv, gt)
        check_prelabels(midx, miou, cls, gt, cap, d2, counts, P)
        # all three label kinds occur, so the comparison is not vacuous
        flat = torch.cat(cap["labels"])
        assert (flat == -1).any() and (flat == RH.NUM_CLASSES).any() and ((flat >= 0) & (flat < RH.NUM_CLASSES)).any()
