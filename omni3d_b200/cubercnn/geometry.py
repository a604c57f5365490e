"""3D box geometry of the CubeHead decode / disentangled loss in batched torch fp32 (device resident):
cuboid corners (cubercnn/util/math_util.py:116-219), allocentric -> egocentric rotation (:651-679),
6D -> rotation matrix (pytorch3d rotation_6d_to_matrix, cube_head.py:176), chamfer / finite-mean
reductions (roi_heads.py:298-304, 932-941).  NaN-safe for masked rows (no 0*inf in backward)."""
import torch
import torch.nn.functional as F

_SIGNS = torch.tensor([[-1., 1, 1, -1, -1, 1, 1, -1],      # x <- +-L/2
                       [-1., -1, 1, 1, -1, -1, 1, 1],      # y <- +-H/2
                       [-1., -1, -1, -1, 1, 1, 1, 1]])     # z <- +-W/2


def cuboid_corners(center, dims_whl, R):
    """center (n,3), dims (n,3) = (W,H,L), R (n,3,3) -> (n,8,3) in the corner order of DATA.md:109-131."""
    s = _SIGNS.to(center.device)
    half = torch.stack((dims_whl[:, 2], dims_whl[:, 1], dims_whl[:, 0]), 1) / 2          # (n,3) = (L,H,W)/2
    v = s[None] * half[:, :, None]                                                       # (n,3,8)
    v = R @ v + center[:, :, None]
    return v.transpose(1, 2)


def rotation_6d_to_matrix(d6):
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = F.normalize(a1, dim=-1)
    b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=-1)
    return torch.stack((b1, b2, torch.cross(b1, b2, dim=-1)), dim=-2)


def _axis_angle_to_matrix(aa):
    angle = torch.norm(aa, p=2, dim=-1, keepdim=True)
    half = 0.5 * angle
    small = angle.abs() < 1e-6
    safe = torch.where(small, torch.ones_like(angle), angle)
    s = torch.where(small, 0.5 - angle * angle / 48, torch.sin(half) / safe)
    q = torch.cat([torch.cos(half), aa * s], dim=-1)
    r, i, j, k = q.unbind(-1)
    two_s = 2.0 / (q * q).sum(-1)
    m = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return m.reshape(aa.shape[:-1] + (3, 3))


def R_from_allocentric(K, R_view, u, v):
    fx, fy, sx, sy = K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2]
    oray = torch.stack(((u - sx) / fx, (v - sy) / fy, torch.ones_like(u)), 1)
    oray = oray / torch.linalg.norm(oray, dim=1, keepdim=True)
    angle = torch.acos(oray[:, 2])
    axis = torch.stack((-oray[:, 1], oray[:, 0], torch.zeros_like(angle)), 1)
    norms = torch.linalg.norm(axis, dim=1)
    valid = angle > 0
    axis_n = axis / torch.where(valid, norms, torch.ones_like(norms))[:, None]
    M = _axis_angle_to_matrix(angle[:, None] * axis_n)
    eye = torch.eye(3, device=K.device, dtype=K.dtype).expand_as(M)
    M = torch.where(valid[:, None, None], M, eye)
    return M @ R_view


def chamfer8(a, b):
    d = (a[:, :, None, :] - b[:, None, :, :]).abs().sum(-1)
    return d.min(1).values.mean(-1) + d.min(2).values.mean(-1)


def finite_mean(loss, mask=None):
    ok = torch.isfinite(loss)
    if mask is not None:
        ok = ok & mask
    cnt = ok.sum()
    total = torch.where(ok, loss, torch.zeros_like(loss)).sum()
    return total / cnt.clamp(min=1).to(total.dtype)
