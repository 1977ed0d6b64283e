"""TEST INFRASTRUCTURE - restatement of the three `roma` entry points the alignment path uses.

`roma` is an un-vendored, un-pinned pip dependency of the reference (requirements.txt:3 says just
``roma``; the devcontainer pins only torch) and is absent from this image.  Call sites on the path:

  * ``roma.RigidUnitQuat(Q, T).normalize().to_homogeneous()``        reference base_opt.py:206
  * ``roma.rotmat_to_unitquat(R)``                                    reference base_opt.py:221
  * ``roma.rigid_points_registration(x, y, weights, compute_scaling)`` reference init_im_poses.py:447,541

Published semantics (roma >= 1.2): quaternions are XYZW; ``normalize`` divides the quaternion by
its L2 norm; ``to_homogeneous`` builds [R|t; 0 0 0 1]; ``rigid_points_registration`` is the weighted
Kabsch/Umeyama solve returning (R, t, scale) with y ~ scale * R x + t.

PARITY UNPINNED at this boundary: nothing in the reference's tree pins roma's numerics, so these
functions define the third-party behaviour for both the real reference run (oracle/ref_harness.py
installs this module as ``roma``) and for oracle/ga_oracle.py.
"""
import numpy as np
import torch
from scipy.spatial.transform import Rotation


def unitquat_to_rotmat(q):
    x, y, z, w = q.unbind(-1)
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    one = torch.ones_like(x)
    rows = [one - (tyy + tzz), txy - twz, txz + twy,
            txy + twz, one - (txx + tzz), tyz - twx,
            txz - twy, tyz + twx, one - (txx + tyy)]
    return torch.stack(rows, dim=-1).reshape(q.shape[:-1] + (3, 3))


class RigidUnitQuat:
    def __init__(self, linear, translation):
        self.linear = linear
        self.translation = translation

    def normalize(self):
        return RigidUnitQuat(self.linear / torch.norm(self.linear, dim=-1, keepdim=True), self.translation)

    def to_homogeneous(self):
        R = unitquat_to_rotmat(self.linear)
        top = torch.cat((R, self.translation.unsqueeze(-1)), dim=-1)
        bottom = torch.zeros_like(top[..., :1, :])
        bottom[..., 0, 3] = 1
        return torch.cat((top, bottom), dim=-2)


def rotmat_to_unitquat(R):
    q = Rotation.from_matrix(R.detach().cpu().double().numpy()).as_quat()   # xyzw
    return torch.as_tensor(q, dtype=R.dtype, device=R.device)


def special_procrustes(M):
    """Nearest rotation to M (in Frobenius norm) and the signed singular values."""
    U, D, Vt = torch.linalg.svd(M)
    d = torch.det(U) * torch.det(Vt)
    S = torch.ones_like(D)
    S[..., -1] = d
    return (U * S[..., None, :]) @ Vt, D * S


def rigid_points_registration(x, y, weights=None, compute_scaling=False):
    if weights is None:
        weights = torch.ones_like(x[..., 0])
    n = weights.sum(-1, keepdim=True)[..., None]
    xm = (weights[..., None] * x).sum(-2, keepdim=True) / n
    ym = (weights[..., None] * y).sum(-2, keepdim=True) / n
    xh, yh = x - xm, y - ym
    R, DS = special_procrustes(yh.transpose(-1, -2) @ (weights[..., None] * xh))
    if compute_scaling:
        s = DS.sum(-1) / (weights[..., None] * xh.square()).sum((-2, -1))
        t = ym - s[..., None, None] * (xm @ R.transpose(-1, -2))
        return R, t.squeeze(-2), s
    t = ym - xm @ R.transpose(-1, -2)
    return R, t.squeeze(-2)
