import torch


def hat(v):
    x, y, z = v.unbind(1)
    o = torch.zeros_like(x)
    return torch.stack([o, -z, y, z, o, -x, -y, x, o], dim=1).view(-1, 3, 3)


def so3_exp_map(log_rot, eps=0.0001):
    """Rodrigues with theta = sqrt(clamp(|w|^2, eps)) (models/model.py:229, modules' callers)."""
    if log_rot.dim() != 2 or log_rot.shape[1] != 3:
        raise ValueError("Input tensor shape has to be Nx3.")
    nrms = (log_rot * log_rot).sum(1)
    theta = torch.clamp(nrms, eps).sqrt()
    inv = 1.0 / theta
    fac1 = inv * theta.sin()
    fac2 = inv * inv * (1.0 - theta.cos())
    K = hat(log_rot)
    return fac1[:, None, None] * K + fac2[:, None, None] * torch.bmm(K, K) + torch.eye(3, dtype=log_rot.dtype, device=log_rot.device)[None]


def so3_log_map(R, eps=0.0001, cos_bound=1e-4):
    """Inverse of ``so3_exp_map`` (utils/pc_util.py:8 imports it for ``init_cov_from_pointcloud``, which GoMAvatar's
    mesh-based initialisation never calls)."""
    tr = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]
    phi = torch.acos(torch.clamp((tr - 1.0) * 0.5, -1.0 + cos_bound, 1.0 - cos_bound))
    s = phi.sin()
    fac = torch.where(s.abs() > 0.5 * eps, phi / (2.0 * s), 0.5 + phi * phi / 12.0)
    A = fac[:, None, None] * (R - R.permute(0, 2, 1))
    return torch.stack([A[:, 2, 1], A[:, 0, 2], A[:, 1, 0]], dim=1)
