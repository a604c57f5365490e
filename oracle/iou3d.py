"""ORACLE (test infrastructure only): ctypes front-end of oracle/iou3d_oracle.c.

Restates pytorch3d._C.iou_box3d + the reference wrapper
cubercnn/evaluation/omni3d_evaluation.py:65-166.  Also holds an independent fp64
half-space oracle (scipy) used to pin the C restatement.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libiou3d_oracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libiou3d_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a, t=ctypes.c_float):
    return a.ctypes.data_as(ctypes.POINTER(t)) if a is not None else None


def iou_box3d(boxes1, boxes2, threads=1, with_counts=False):
    """(N,8,3),(M,8,3) -> vol (N,M), iou (N,M) [, nfaces (N,M) int32, nside (N,M,2)]."""
    b1, b2 = _f32(boxes1).reshape(-1, 8, 3), _f32(boxes2).reshape(-1, 8, 3)
    N, M = len(b1), len(b2)
    vol = np.zeros((N, M), np.float32)
    iou = np.zeros((N, M), np.float32)
    nf = np.zeros((N, M), np.int32) if with_counts else None
    ns = np.zeros((N, M, 2), np.int32) if with_counts else None
    lib().oracle_iou_box3d(_p(b1), N, _p(b2), M, _p(vol), _p(iou), _p(nf, ctypes.c_int32),
                           _p(ns, ctypes.c_int32), int(threads))
    return (vol, iou, nf, ns) if with_counts else (vol, iou)


def iou_box3d_paired(boxes1, boxes2, threads=1):
    b1, b2 = _f32(boxes1).reshape(-1, 8, 3), _f32(boxes2).reshape(-1, 8, 3)
    P = len(b1)
    assert len(b2) == P
    vol = np.zeros(P, np.float32)
    iou = np.zeros(P, np.float32)
    nf = np.zeros(P, np.int32)
    lib().oracle_iou_box3d_paired(_p(b1), _p(b2), P, _p(vol), _p(iou), _p(nf, ctypes.c_int32),
                                  int(threads))
    return vol, iou, nf


def check_boxes(boxes, eps_coplanar=1e-4, eps_nonzero=1e-8):
    """-> (coplanar_ok (N,) bool, nonzero_ok (N,) bool); omni3d_evaluation.py:65-104."""
    b = _f32(boxes).reshape(-1, 8, 3)
    N = len(b)
    c = np.zeros(max(N, 1), np.uint8)
    z = np.zeros(max(N, 1), np.uint8)
    lib().oracle_check_boxes(_p(b), N, ctypes.c_float(eps_coplanar), ctypes.c_float(eps_nonzero),
                             _p(c, ctypes.c_uint8), _p(z, ctypes.c_uint8))
    return c[:N].astype(bool), z[:N].astype(bool)


def box3d_overlap(boxes_dt, boxes_gt, eps_coplanar=1e-4, eps_nonzero=1e-8, threads=1):
    """The reference entry point omni3d_evaluation.py:106-166 -> iou (N,M) fp32."""
    b1, b2 = _f32(boxes_dt).reshape(-1, 8, 3), _f32(boxes_gt).reshape(-1, 8, 3)
    N, M = len(b1), len(b2)
    iou = np.zeros((N, M), np.float32)
    vol = np.zeros((N, M), np.float32)
    nbad = np.zeros(2, np.int32)
    lib().oracle_box3d_overlap(_p(b1), N, _p(b2), M, ctypes.c_float(eps_coplanar),
                               ctypes.c_float(eps_nonzero), _p(iou), _p(vol),
                               _p(nbad, ctypes.c_int32), int(threads))
    return iou, nbad


# ---------------------------------------------------------------------------
# independent fp64 oracle: convex-polytope intersection by half-spaces (scipy)
# ---------------------------------------------------------------------------
_PLANES = [[0, 1, 2, 3], [3, 2, 6, 7], [0, 1, 5, 4], [0, 3, 7, 4], [1, 2, 6, 5], [4, 5, 6, 7]]


def _halfspaces(box):
    box = np.asarray(box, np.float64)
    c = box.mean(0)
    hs = []
    for p in _PLANES:
        v = box[p]
        n = np.cross(v[1] - v[0], v[3] - v[0])
        n /= np.linalg.norm(n)
        pc = v.mean(0)
        if np.dot(c - pc, n) < 0:
            n = -n
        # inside: n.(x - pc) >= 0  <=>  -n.x + n.pc <= 0
        hs.append(np.concatenate([-n, [np.dot(n, pc)]]))
    return np.array(hs)


def iou_halfspace_fp64(box1, box2):
    """fp64 reference IoU/volume of two convex boxes (corner order of DATA.md:109-131)."""
    from scipy.optimize import linprog
    from scipy.spatial import ConvexHull, HalfspaceIntersection, QhullError

    b1, b2 = np.asarray(box1, np.float64), np.asarray(box2, np.float64)
    v1, v2 = ConvexHull(b1).volume, ConvexHull(b2).volume
    hs = np.vstack([_halfspaces(b1), _halfspaces(b2)])
    # Chebyshev centre -> strictly interior point (or none)
    nrm = np.linalg.norm(hs[:, :3], axis=1)
    res = linprog([0, 0, 0, -1], A_ub=np.hstack([hs[:, :3], nrm[:, None]]), b_ub=-hs[:, 3],
                  bounds=[(None, None)] * 3 + [(0, None)])
    if (not res.success) or res.x[3] < 1e-9:
        return 0.0, 0.0
    try:
        hi = HalfspaceIntersection(hs, res.x[:3])
        vol = ConvexHull(hi.intersections).volume
    except QhullError:
        return 0.0, 0.0
    return vol, vol / (v1 + v2 - vol)
