"""The MODEL-side statements of the reference's two entry scripts, executed verbatim against the product model:
tools/train_net.py:188-236 (`loss_dict = model(data)` on DatasetMapper3D-style dicts with a detectron2 `Instances`,
`losses = sum(loss_dict.values())`, `optimizer.zero_grad(); losses.backward(); optimizer.step()` with the torch SGD that
cubercnn/solver/build.py:47-56 builds) and demo/demo.py:83-103 (`model(batched)[0]['instances']` on an image dict with a
numpy K, then the zip over the 3D prediction fields).  The scripts' control plane (dataset registry, loaders, evaluator,
visualiser) is out of scope; what they ask of the model is what is checked here."""
import numpy as np
import pytest
import torch

from omni3d_b200 import synth

pytestmark = pytest.mark.gpu

TRAIN_LOSS_KEYS = {"BoxHead/loss_cls", "BoxHead/loss_box_reg", "Cube/uncert", "Cube/loss_dims", "Cube/loss_xy", "Cube/loss_z",
                   "Cube/loss_pose", "Cube/loss_joint", "rpn/cls", "rpn/loc"}


def _build(train):
    from omni3d_b200 import cubercnn as pc
    cfg = pc.load_cfg("cubercnn_DLA34_FPN.yaml", ["MODEL.WEIGHTS_PRETRAIN", "none"])
    torch.manual_seed(0)
    model = pc.build_model(cfg)
    return cfg, (model.train() if train else model.eval())


def _to_instances_schema(items):
    """the dict DatasetMapper3D emits (dataset_mapper.py:133-155): image uint8 CHW, height, width, K, `instances` with
    gt_boxes (Boxes), gt_classes, gt_boxes3D (G,9), gt_poses (G,3,3)"""
    from omni3d_b200.cubercnn.structures import Boxes, Instances
    out = []
    for it in items:
        g = it["gt"]
        inst = Instances((it["height"], it["width"]))
        inst.gt_boxes = Boxes(g["boxes"])
        inst.gt_classes = g["classes"]
        inst.gt_boxes3D = g["boxes3D"]
        inst.gt_poses = g["poses"]
        out.append({"image": it["image"], "height": it["height"], "width": it["width"], "K": it["K"], "instances": inst})
    return out


def test_train_net_call_pattern():
    cfg, model = _build(train=True)
    data = _to_instances_schema(synth.make_batch(2, 128, 192, num_gt=4, seed=11, image_dtype=torch.uint8))
    params = [p for p in model.parameters() if p.requires_grad]
    optimizer = torch.optim.SGD(params, lr=0.0025, momentum=0.9, weight_decay=1e-4)       # solver/build.py:47-56
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    # tools/train_net.py:201-236
    loss_dict = model(data)
    losses = sum(loss_dict.values())
    assert set(loss_dict) == TRAIN_LOSS_KEYS
    assert all(torch.isfinite(v).all() for v in loss_dict.values()) and torch.isfinite(losses)
    optimizer.zero_grad()
    losses.backward()
    for n, p in model.named_parameters():       # train_net.py:226-232: per-parameter finite check
        if p.grad is not None:
            assert torch.isfinite(p.grad).all(), n
    optimizer.step()
    moved = sum(int(not torch.equal(before[n], p.detach())) for n, p in model.named_parameters())
    assert moved > 150, moved
    # the same batch through the plain-tensor `gt` schema gives the same losses (same sampling stream)
    cfg2, model2 = _build(train=True)
    plain = synth.make_batch(2, 128, 192, num_gt=4, seed=11, image_dtype=torch.uint8)
    cfg3, model3 = _build(train=True)
    from omni3d_b200 import kernels as Kx
    dev = torch.device("cuda", torch.cuda.current_device())
    Kx.rng_state(dev, seed=5)                   # the sampling kernels' Philox stream {seed, step counter} lives on the device
    l_inst = {k: float(v) for k, v in model2(_to_instances_schema(plain)).items()}
    Kx.rng_state(dev, seed=5)
    l_gt = {k: float(v) for k, v in model3(plain).items()}
    for k in TRAIN_LOSS_KEYS:
        assert abs(l_inst[k] - l_gt[k]) <= 1e-3 * abs(l_gt[k]) + 1e-5, (k, l_inst[k], l_gt[k])


def test_demo_call_pattern():
    cfg, model = _build(train=False)
    h, w = 128, 192
    im = np.random.RandomState(0).randint(0, 256, (h, w, 3), dtype=np.uint8)                   # util.imread: HWC BGR uint8
    focal_length = 4.0 * h / 2                                                                # demo.py:64-66
    K = np.array([[focal_length, 0.0, w / 2], [0.0, focal_length, h / 2], [0.0, 0.0, 1.0]])
    batched = [{"image": torch.as_tensor(np.ascontiguousarray(im.transpose(2, 0, 1))).cuda(),   # demo.py:83-86
                "height": h, "width": w, "K": K}]
    with torch.no_grad():
        dets = model(batched)[0]["instances"]
    n_det = len(dets)
    assert n_det > 0
    seen = 0
    for idx, (corners3D, center_cam, center_2D, dimensions, pose, score, cat_idx) in enumerate(zip(      # demo.py:95-98
            dets.pred_bbox3D, dets.pred_center_cam, dets.pred_center_2D, dets.pred_dimensions,
            dets.pred_pose, dets.scores, dets.pred_classes)):
        bbox3D = center_cam.tolist() + dimensions.tolist()                                             # demo.py:106
        assert len(bbox3D) == 6 and tuple(corners3D.shape) == (8, 3) and tuple(pose.shape) == (3, 3)
        assert len(center_2D.tolist()) == 2 and 0.0 <= float(score) <= 1.0
        assert 0 <= int(cat_idx) < cfg.MODEL.ROI_HEADS.NUM_CLASSES
        R = pose.float()
        assert torch.allclose(R @ R.T, torch.eye(3, device=R.device), atol=1e-3)                      # a rotation
        seen += 1
    assert seen == n_det
    assert dets.pred_boxes.tensor.shape == (n_det, 4) and dets.image_size == (h, w)
