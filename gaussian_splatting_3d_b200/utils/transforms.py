"""Quaternion helpers (reference utils/transforms.py).  The reference delegates to kornia 0.6.x
(`quaternion_to_rotation_matrix(q, QuaternionCoeffOrder.WXYZ)`, un-vendored, unpinned); its
arithmetic is restated here: L2-normalise (eps 1e-12), then the standard matrix from 2*q products."""
import numpy as np
import torch
import torch.nn.functional as F


def qvec2rotmat(qvec):
    w, x, y, z = qvec[0], qvec[1], qvec[2], qvec[3]
    return np.array([
        [1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * w * z, 2 * z * x + 2 * w * y],
        [2 * x * y + 2 * w * z, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * w * x],
        [2 * z * x - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x * x - 2 * y * y],
    ])


def qvec2rotmat_batched(qvec):
    q = F.normalize(qvec, p=2.0, dim=-1, eps=1e-12)
    w, x, y, z = q.unbind(-1)
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    one = torch.ones_like(w)
    m = torch.stack([
        one - (tyy + tzz), txy - twz, txz + twy,
        txy + twz, one - (txx + tzz), tyz - twx,
        txz - twy, tyz + twx, one - (txx + tyy),
    ], dim=-1)
    return m.view(*qvec.shape[:-1], 3, 3)


def qsvec2rotmat_batched(qvec, svec):
    """R(q) diag(s): scale copied along columns (reference :31-45)."""
    return svec.unsqueeze(-2) * qvec2rotmat_batched(qvec)


def rotmat2wxyz(rotmat):
    """Rotation matrix -> (w,x,y,z), Shepperd's method (reference :48-51 via kornia)."""
    m = rotmat
    t = m[..., 0, 0] + m[..., 1, 1] + m[..., 2, 2]
    w = torch.sqrt(torch.clamp(1 + t, min=1e-12)) / 2
    x = torch.sqrt(torch.clamp(1 + m[..., 0, 0] - m[..., 1, 1] - m[..., 2, 2], min=1e-12)) / 2
    y = torch.sqrt(torch.clamp(1 - m[..., 0, 0] + m[..., 1, 1] - m[..., 2, 2], min=1e-12)) / 2
    z = torch.sqrt(torch.clamp(1 - m[..., 0, 0] - m[..., 1, 1] + m[..., 2, 2], min=1e-12)) / 2
    x = torch.copysign(x, m[..., 2, 1] - m[..., 1, 2])
    y = torch.copysign(y, m[..., 0, 2] - m[..., 2, 0])
    z = torch.copysign(z, m[..., 1, 0] - m[..., 0, 1])
    return torch.stack([w, x, y, z], dim=-1)
