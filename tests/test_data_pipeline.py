"""SURVEY 8f-3 — input pipeline.  CPU: the oracle restatement of Pillow's 8-bit bilinear resize is pinned bit-exactly against
Pillow itself; the product's vectorised coefficient tables equal the oracle's; the batched annotation transform equals the
per-object restatement of dataset_mapper.py:74-155.  GPU (-m gpu): c3d_resize_bilinear_u8 == Pillow bit for bit (+ flip,
+ CHW), and the mapped batch trains."""
import numpy as np
import pytest
import torch
from PIL import Image

from oracle import mapper_oracle, pil_resize

SHAPES = [(37, 53, 22, 31), (480, 640, 512, 683), (64, 64, 64, 64), (100, 80, 100, 40), (90, 120, 45, 120),
          (333, 500, 640, 961), (720, 1280, 384, 683), (50, 60, 137, 91), (5, 7, 3, 2), (1, 9, 4, 30)]


def _img(h, w, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("h,w,nh,nw", SHAPES)
def test_oracle_resize_is_pillow_bit_exact(h, w, nh, nw):
    a = _img(h, w)
    ref = np.asarray(Image.fromarray(a).resize((nw, nh), Image.BILINEAR))
    assert np.array_equal(pil_resize.resize_bilinear_u8(a, nh, nw), ref)


def test_product_coefficients_equal_oracle():
    from omni3d_b200 import data
    for i, o in [(53, 31), (640, 683), (64, 64), (80, 40), (500, 961), (1280, 683), (60, 91), (7, 2), (9, 30), (1, 4), (4096, 640)]:
        b1, k1 = pil_resize.precompute_coeffs(i, o)
        b2, k2 = data.pil_bilinear_coeffs(i, o)
        assert np.array_equal(b1, b2) and np.array_equal(k1, k2), (i, o)
    assert data.shortest_edge_shape(480, 640, 512, 4096) == (512, 683)
    assert data.shortest_edge_shape(720, 1280, 384, 4096) == (384, 683)
    assert data.shortest_edge_shape(1000, 3000, 640, 1333) == (444, 1333)


def _annos(n, h, w, seed):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        x1, y1 = rng.uniform(0, w - 40), rng.uniform(0, h - 40)
        q, r = np.linalg.qr(rng.standard_normal((3, 3)))
        out.append({"bbox": [x1, y1, x1 + rng.uniform(5, 39), y1 + rng.uniform(5, 39)], "category_id": int(rng.integers(0, 50)),
                    "center_cam": [float(rng.uniform(-3, 3)), float(rng.uniform(-2, 2)), float(rng.uniform(2, 30))],
                    "dimensions": rng.uniform(0.3, 3, 3).tolist(), "pose": q.tolist(), "iscrowd": int(i == 3)})
    out[1]["center_cam"][2] = 0.0                       # the reference skips projection / mirroring for z == 0 (:86)
    out[2]["bbox"][2] = out[2]["bbox"][0]               # empty box: filtered
    return out


@pytest.mark.parametrize("flip", [False, True])
def test_annotation_transform_equals_per_object_oracle(flip):
    from omni3d_b200 import data
    h, w, nh, nw = 480, 640, 512, 683
    K = [[500.0, 0, 320.0], [0, 500.0, 240.0], [0, 0, 1]]
    annos = _annos(9, h, w, 4)
    annos[1]["center_cam_proj"] = [0.0, 0.0, 0.0]       # the dataset's stored value is kept by the reference when z == 0
    cls, box, b3, pose = mapper_oracle.map_annotations(annos, K, h, w, nh, nw, flip)
    got = data.transform_annotations(annos, K, h, w, nh, nw, flip)
    assert np.array_equal(got["classes"].numpy(), cls)
    assert np.array_equal(got["boxes"].numpy(), box)
    assert np.array_equal(got["poses"].numpy(), pose)
    assert np.array_equal(got["boxes3D"].numpy(), b3)


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,nh,nw", SHAPES)
@pytest.mark.parametrize("flip", [False, True])
def test_device_resize_is_pillow_bit_exact(h, w, nh, nw, flip):
    from omni3d_b200 import data
    a = _img(h, w, 3)
    ref = np.asarray(Image.fromarray(a).resize((nw, nh), Image.BILINEAR))
    if flip:
        ref = ref[:, ::-1]
    got = data.resize_flip_u8(torch.from_numpy(a).cuda(), nh, nw, flip)
    assert got.dtype == torch.uint8 and tuple(got.shape) == (3, nh, nw)
    assert np.array_equal(got.cpu().numpy(), np.ascontiguousarray(ref.transpose(2, 0, 1)))


@pytest.mark.gpu
def test_device_mapper_feeds_the_model():
    from omni3d_b200 import cubercnn as pc
    from omni3d_b200 import data
    cfg = pc.load_cfg("cubercnn_DLA34_FPN.yaml", ["MODEL.WEIGHTS_PRETRAIN", "none"])
    mapper = data.DeviceMapper3D(cfg, is_train=True, seed=0)
    K = [[500.0, 0, 320.0], [0, 500.0, 240.0], [0, 0, 1]]
    recs = [{"image_hwc": _img(240, 320, i), "K": K, "annotations": [a for a in _annos(6, 240, 320, i) if a["center_cam"][2] != 0]}
            for i in range(2)]
    items = [mapper(r, size=256, flip=bool(i)) for i, r in enumerate(recs)]
    assert tuple(items[0]["image"].shape) == (3, 256, 341) and items[0]["image"].is_cuda
    torch.manual_seed(0)
    model = pc.build_model(cfg).train()
    losses = model(items)
    assert len(losses) == 10 and all(torch.isfinite(v) for v in losses.values())
