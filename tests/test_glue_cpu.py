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
