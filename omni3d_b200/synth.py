"""Seeded synthetic Cube R-CNN batches (SURVEY.md section 8d "Synthetic inputs — model").

Plain tensors only (no detectron2 types): each item mirrors the batched-input schema the reference's
DatasetMapper3D emits (cubercnn/data/dataset_mapper.py:133-155) —
  image (3,H,W) BGR in [0,255] (float32, or uint8 as the mapper's torch.as_tensor(image) yields), height, width, K 3x3, and `gt` = dict(classes int64 (G,),
  boxes (G,4) XYXY, boxes3D (G,9) = [u,v,z,W,H,L,X,Y,Z], poses (G,3,3)).
One of the G boxes per image is an ignore region (class -1) to exercise the ignore path.
"""
import torch


def _rand_rot(n, g):
    q, r = torch.linalg.qr(torch.randn(n, 3, 3, generator=g))
    q = q * torch.sign(torch.diagonal(r, dim1=1, dim2=2)).unsqueeze(1)
    det = torch.linalg.det(q)
    q[:, :, 0] *= det.unsqueeze(1)
    return q


def make_batch(batch, height=640, width=640, num_gt=8, num_classes=50, seed=0, with_gt=True, with_ignore=True,
               image_dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    items = []
    for _ in range(batch):
        img = torch.randint(0, 256, (3, height, width), generator=g).to(image_dtype)   # uint8 = what DatasetMapper3D emits
        f = float(torch.empty(1).uniform_(400, 800, generator=g))
        K = [[f, 0.0, width / 2.0], [0.0, f, height / 2.0], [0.0, 0.0, 1.0]]
        item = {"image": img, "height": height, "width": width, "K": K}
        if with_gt:
            lo, hi = min(32.0, width / 8), min(256.0, width / 2.5)
            wh = torch.empty(num_gt, 2).uniform_(lo, hi, generator=g)
            x1 = torch.rand(num_gt, generator=g) * (width - wh[:, 0])
            y1 = torch.rand(num_gt, generator=g) * (height - wh[:, 1])
            boxes = torch.stack([x1, y1, x1 + wh[:, 0], y1 + wh[:, 1]], 1)
            classes = torch.randint(0, num_classes, (num_gt,), generator=g)
            if with_ignore and num_gt > 1:
                classes[-1] = -1
            z = torch.empty(num_gt).uniform_(2, 40, generator=g)
            dims = torch.empty(num_gt, 3).uniform_(0.3, 3.0, generator=g)
            u, v = (boxes[:, 0] + boxes[:, 2]) / 2, (boxes[:, 1] + boxes[:, 3]) / 2
            X = z * (u - K[0][2]) / f
            Y = z * (v - K[1][2]) / f
            boxes3D = torch.cat([u[:, None], v[:, None], z[:, None], dims, X[:, None], Y[:, None], z[:, None]], 1)
            item["gt"] = {"classes": classes, "boxes": boxes, "boxes3D": boxes3D, "poses": _rand_rot(num_gt, g)}
        items.append(item)
    return items
