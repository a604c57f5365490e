"""Generates tests/golden/iou3d_wrapper_golden.npz by EXECUTING THE REFERENCE'S OWN PYTHON
(cubercnn/evaluation/omni3d_evaluation.py:65-166: _check_coplanar, _check_nonzero, box3d_overlap)
from /root/reference in this container.

The reference module cannot be imported as-is (detectron2 / pytorch3d / pycocotools are absent),
so every missing third-party module is replaced by an inert stub; the ONLY third-party arithmetic
the three functions reach is `pytorch3d._C.iou_box3d` (not vendored anywhere), which is bound to the
C restatement in oracle/iou3d_oracle.c, and the two index tables `_box_planes/_box_triangles`
(PyTorch3D constants restated in SURVEY.md A.7).  The fixture therefore pins the *wrapper*
semantics (row masks, eps handling, the summed-offset coplanarity quirk) to the reference's code.

Run here only (needs /root/reference):   python tests/golden/make_iou_golden.py
"""
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"

from oracle import iou3d as oracle_iou  # noqa: E402
import boxgen  # noqa: E402


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (), {"__init__": lambda self, *a, **k: None})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    PREFIXES = ("detectron2", "pycocotools", "pytorch3d", "cubercnn", "fvcore", "iopath")

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.PREFIXES:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def load_reference_eval():
    sys.meta_path.append(_StubFinder())
    path = os.path.join(REF, "cubercnn/evaluation/omni3d_evaluation.py")
    spec = importlib.util.spec_from_file_location("ref_omni3d_evaluation", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod._box_planes = [[0, 1, 2, 3], [3, 2, 6, 7], [0, 1, 5, 4], [0, 3, 7, 4], [1, 2, 6, 5], [4, 5, 6, 7]]
    mod._box_triangles = [[0, 1, 2], [0, 3, 2], [4, 5, 6], [4, 6, 7], [1, 5, 6], [1, 6, 2],
                          [0, 4, 7], [0, 7, 3], [3, 2, 6], [3, 6, 7], [0, 1, 5], [0, 4, 5]]

    def iou_box3d(b1, b2):
        vol, iou = oracle_iou.iou_box3d(b1.numpy(), b2.numpy())
        return torch.from_numpy(vol), torch.from_numpy(iou)

    mod._C = types.SimpleNamespace(iou_box3d=iou_box3d)
    return mod


def main():
    ref = load_reference_eval()
    out = {}
    cases = [("dense", 48, 40, 1.0, 0), ("sparse", 64, 24, 10.0, 7), ("mid", 33, 17, 3.0, 11)]
    for name, n, m, L, seed in cases:
        dt = boxgen.random_boxes(n, L, seed)
        gt = boxgen.random_boxes(m, L, seed + 100)
        dt, bad_idx = boxgen.inject_degenerate(dt, frac=0.1, seed=seed + 1)
        # borderline coplanarity: perturbations around the 1e-4 threshold
        rng = np.random.default_rng(seed + 2)
        for t, i in enumerate(rng.choice(n, 6, replace=False)):
            dt[i, 2] += np.float32((t + 1) * 4e-5) * rng.standard_normal(3).astype(np.float32)
        tdt, tgt = torch.from_numpy(dt), torch.from_numpy(gt)
        cop = ref._check_coplanar(tdt, eps=1e-4).numpy()
        nz = ref._check_nonzero(tdt, eps=1e-8).numpy()
        iou = ref.box3d_overlap(tdt, tgt).numpy()
        out[f"{name}_dt"], out[f"{name}_gt"] = dt, gt
        out[f"{name}_coplanar_ok"], out[f"{name}_nonzero_ok"] = cop, nz
        out[f"{name}_iou"] = iou
        print(name, "bad coplanar", int((~cop).sum()), "bad nonzero", int((~nz).sum()), "iou>0", float((iou > 0).mean()))
    # closed-form known answers: unit cube vs shifted / rotated / nested copies
    unit = boxgen.corners([[0, 0, 0]], [[1, 1, 1]])
    th = np.pi / 4
    R = np.array([[[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]]])
    ka_boxes = np.concatenate([
        unit,
        boxgen.corners([[0, 0, 0.5]], [[1, 1, 1]]),        # z-shift 0.5 -> 1/3
        boxgen.corners([[0.25, 0.5, 0.125]], [[1, 1, 1]]),  # general axis-aligned offset
        boxgen.corners([[0, 0, 0]], [[0.5, 0.5, 0.5]]),    # nested scale 0.5 -> 1/8
        boxgen.corners([[0, 0, 0]], [[1, 1, 1]], R),       # 45 deg about y -> vol 2(sqrt2-1)
        boxgen.corners([[3, 0, 0]], [[1, 1, 1]]),          # disjoint -> 0
        boxgen.corners([[1, 0, 0]], [[1, 1, 1]]),          # face-touching -> 0
    ])
    d = [0.25, 0.5, 0.125]
    ov = np.prod([1 - x for x in d])
    ka_expect = np.array([1.0, 1 / 3, ov / (2 - ov), 0.125, (2 * (2 ** 0.5 - 1)) / (2 - 2 * (2 ** 0.5 - 1)), 0.0, 0.0])
    out["ka_boxes"], out["ka_expect"] = ka_boxes, ka_expect
    np.savez_compressed(os.path.join(ROOT, "tests/golden/iou3d_wrapper_golden.npz"), **out)
    print("wrote tests/golden/iou3d_wrapper_golden.npz")


if __name__ == "__main__":
    main()
