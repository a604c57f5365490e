"""ORACLE (test infrastructure only): per-object numpy restatement of the annotation half of DatasetMapper3D —
cubercnn/data/dataset_mapper.py:74-155 (transform_instance_annotations, annotations_to_instances) over detectron2's
ResizeTransform / HFlipTransform / TransformList semantics (recalled upstream, SURVEY App. A: apply_coords scales x by
new_w / w and y by new_h / h; HFlip maps x -> width - x; apply_box transforms the 4 corners and takes min / max)."""
import copy

import numpy as np


class Resize:
    def __init__(self, h, w, new_h, new_w):
        self.h, self.w, self.new_h, self.new_w = h, w, new_h, new_w

    def apply_coords(self, c):
        c = np.asarray(c, np.float64).copy()
        c[:, 0] = c[:, 0] * (self.new_w * 1.0 / self.w)
        c[:, 1] = c[:, 1] * (self.new_h * 1.0 / self.h)
        return c


class HFlip:
    def __init__(self, width):
        self.width = width

    def apply_coords(self, c):
        c = np.asarray(c, np.float64).copy()
        c[:, 0] = self.width - c[:, 0]
        return c


def apply_coords(transforms, c):
    for t in transforms:
        c = t.apply_coords(c)
    return c


def apply_box(transforms, box):
    box = np.asarray(box, np.float64).reshape(-1, 4)
    idxs = np.array([(0, 1), (2, 1), (0, 3), (2, 3)]).flatten()
    coords = box[:, idxs].reshape(-1, 2)
    coords = apply_coords(transforms, coords).reshape((-1, 4, 2))
    minxy, maxxy = coords.min(axis=1), coords.max(axis=1)
    return np.concatenate((minxy, maxxy), axis=1)


_M1 = np.array([[1, 0, 0], [0, -1, 0], [0, 0, -1]])
_M2 = np.array([[-1., 0., 0.], [0., -1., 0.], [0., 0., 1.]])


def transform_instance_annotations(annotation, transforms, K):
    """dataset_mapper.py:74-131 (bbox already XYXY_ABS)."""
    annotation = copy.deepcopy(annotation)
    annotation["bbox"] = apply_box(transforms, np.array([annotation["bbox"]]))[0]
    if annotation["center_cam"][2] != 0:
        point2D = K @ np.array(annotation["center_cam"])
        point2D[:2] = point2D[:2] / point2D[-1]
        annotation["center_cam_proj"] = point2D.tolist()
        annotation["center_cam_proj"][0:2] = apply_coords(transforms, point2D[np.newaxis][:, :2])[0].tolist()
        for t in transforms:
            if isinstance(t, HFlip):
                pose = _M1 @ np.array(annotation["pose"]) @ _M2
                annotation["pose"] = pose.tolist()
    return annotation


def map_annotations(annos, K, h, w, new_h, new_w, flip):
    """-> classes (n,), boxes (n,4), boxes3D (n,9), poses (n,3,3) as float32 arrays (annotations_to_instances :134-143 +
    filter_empty_instances)."""
    transforms = [Resize(h, w, new_h, new_w)] + ([HFlip(new_w)] if flip else [])
    K = np.array(K)
    out = [transform_instance_annotations(a, transforms, K) for a in annos if a.get("iscrowd", 0) == 0]
    cls = np.array([int(a["category_id"]) for a in out], np.int64)
    box = np.array([a["bbox"] for a in out], np.float32).reshape(-1, 4)
    b3 = np.array([a["center_cam_proj"] + a["dimensions"] + a["center_cam"] for a in out], np.float32).reshape(-1, 9)
    pose = np.array([a["pose"] for a in out], np.float32).reshape(-1, 3, 3)
    keep = ((box[:, 2] - box[:, 0]) > 1e-5) & ((box[:, 3] - box[:, 1]) > 1e-5)
    return cls[keep], box[keep], b3[keep], pose[keep]
