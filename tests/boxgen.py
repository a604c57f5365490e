"""Seeded synthetic oriented boxes for the IoU3D tests/bench.

Corner convention = cubercnn/util/math_util.py:116-219 get_cuboid_verts_faces
(== DATA.md:109-131 == PyTorch3D): dims (W,H,L) map to (z,y,x).
"""
import numpy as np

_SX = np.array([-1, 1, 1, -1, -1, 1, 1, -1], np.float64)
_SY = np.array([-1, -1, 1, 1, -1, -1, 1, 1], np.float64)
_SZ = np.array([-1, -1, -1, -1, 1, 1, 1, 1], np.float64)


def random_rotations(n, rng):
    q, r = np.linalg.qr(rng.standard_normal((n, 3, 3)))
    q = q * np.sign(np.diagonal(r, axis1=1, axis2=2))[:, None, :]
    det = np.linalg.det(q)
    q[:, :, 0] *= det[:, None]
    return q


def corners(center, dims_whl, R=None):
    """center (n,3), dims (n,3)=(W,H,L), R (n,3,3) -> (n,8,3) float32."""
    center = np.asarray(center, np.float64).reshape(-1, 3)
    d = np.asarray(dims_whl, np.float64).reshape(-1, 3)
    v = np.stack([_SX[None] * d[:, 2:3] / 2, _SY[None] * d[:, 1:2] / 2, _SZ[None] * d[:, 0:1] / 2], 1)
    if R is not None:
        v = np.asarray(R, np.float64) @ v
    v = v + center[:, :, None]
    return np.ascontiguousarray(v.transpose(0, 2, 1)).astype(np.float32)


def random_boxes(n, L=1.0, seed=0, rotate=True):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-L / 2, L / 2, (n, 3))
    d = rng.uniform(0.5, 2.0, (n, 3))
    R = random_rotations(n, rng) if rotate else None
    return corners(c, d, R)


def inject_degenerate(boxes, frac=0.01, seed=1):
    """1% degenerate dt boxes: half non-coplanar (one vertex moved 1e-2), half zero-thickness."""
    b = boxes.copy()
    rng = np.random.default_rng(seed)
    n = len(b)
    k = max(2, int(n * frac))
    idx = rng.choice(n, k, replace=False)
    for t, i in enumerate(idx):
        if t % 2 == 0:
            b[i, 6] += np.float32(1e-2) * rng.standard_normal(3).astype(np.float32)
        else:
            b[i, 4:8] = b[i, 0:4]
    return b, idx
