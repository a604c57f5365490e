"""ORACLE helper: wrap omni3d_b200.synth plain-tensor batches into d2lite Instances (the
batched-input schema of cubercnn/data/dataset_mapper.py:133-155)."""
import os
import sys

_D2 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "d2lite")
if _D2 not in sys.path:
    sys.path.insert(0, _D2)


def to_d2_inputs(items):
    from detectron2.structures import Boxes, Instances
    out = []
    for it in items:
        d = {k: it[k] for k in ("image", "height", "width", "K")}
        if "gt" in it:
            g = it["gt"]
            inst = Instances((it["image"].shape[1], it["image"].shape[2]))
            inst.gt_classes = g["classes"].clone()
            inst.gt_boxes = Boxes(g["boxes"].clone())
            inst.gt_boxes3D = g["boxes3D"].clone()
            inst.gt_poses = g["poses"].clone()
            d["instances"] = inst
        out.append(d)
    return out
