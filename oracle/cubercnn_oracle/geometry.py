"""Restates cubercnn/util/math_util.py:116-219 (corners), :581-592 (virtual scale), :651-679
(allocentric -> egocentric) in vectorised form."""
import torch
from pytorch3d.transforms import axis_angle_to_matrix

# corner signs of the unit cuboid in the (x<-L, y<-H, z<-W) convention of math_util.py:171-181
_SX = torch.tensor([-1., 1, 1, -1, -1, 1, 1, -1])
_SY = torch.tensor([-1., -1, 1, 1, -1, -1, 1, 1])
_SZ = torch.tensor([-1., -1, -1, -1, 1, 1, 1, 1])


def cuboid_corners(box3d, R=None):
    """box3d (n,6) = [X,Y,Z,W,H,L]; R (n,3,3) -> corners (n,8,3)."""
    box3d = box3d.float()
    sx, sy, sz = _SX.to(box3d.device), _SY.to(box3d.device), _SZ.to(box3d.device)
    l, h, w = box3d[:, 5:6], box3d[:, 4:5], box3d[:, 3:4]
    v = torch.stack([sx[None] * l / 2, sy[None] * h / 2, sz[None] * w / 2], dim=1)   # (n,3,8)
    if R is not None:
        v = R.float() @ v
    v = v + box3d[:, :3].unsqueeze(2)
    return v.transpose(1, 2)


def virtual_scale(f, H, f0, H0):
    return (H0 * f) / (f0 * H)


def R_from_allocentric(K, R_view, u, v):
    fx, fy, sx, sy = K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2]
    oray = torch.stack(((u - sx) / fx, (v - sy) / fy, torch.ones_like(u))).T
    oray = oray / torch.linalg.norm(oray, dim=1).unsqueeze(1)
    angle = torch.acos(oray[:, -1])
    axis = torch.zeros_like(oray)
    axis[:, 0] = axis[:, 0] - oray[:, 1]
    axis[:, 1] = axis[:, 1] + oray[:, 0]
    norms = torch.linalg.norm(axis, dim=1)
    valid = angle > 0
    M = axis_angle_to_matrix(angle.unsqueeze(1) * axis / norms.unsqueeze(1))
    R = R_view.clone()
    R[valid] = torch.bmm(M[valid], R_view[valid]).to(R.dtype)      # (.to: no-op in fp32; keeps autocast runs legal)
    return R
