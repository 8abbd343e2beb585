"""Rotation conversions used by the forward tail (ref:cs_vit/utils/geometry.py).

fp32 PyTorch on the GPU, negligible cost.  The branch structure of the reference is kept
because it is observable: ``matrix_to_axis_angle`` goes through the best-conditioned quaternion candidate with
a non-negative real part, so axis-angle outputs jump near pi exactly where the reference's do.
"""
import math

import torch
import torch.nn.functional as F


def rotation_6d_to_matrix(d6: torch.Tensor) -> torch.Tensor:
    """Zhou et al. 6D -> rotation matrix via Gram-Schmidt; rows are (b1, b2, b1 x b2)   (ref :111-132)."""
    b1 = F.normalize(d6[..., :3], dim=-1)
    a2 = d6[..., 3:]
    b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=-1)
    return torch.stack((b1, b2, torch.linalg.cross(b1, b2, dim=-1)), dim=-2)


def _sqrt_positive_part(x: torch.Tensor) -> torch.Tensor:
    """``sqrt(max(0, x))`` with a ZERO subgradient where ``x <= 0`` (ref :150-161).  ``sqrt(clamp(x, 0))`` is the same value but
    its backward is ``inf * 0 = NaN`` at a clamped entry, which poisons every gradient of a training step."""
    positive = x > 0
    safe = torch.where(positive, x, torch.ones_like(x))
    return torch.where(positive, torch.sqrt(safe), torch.zeros_like(x))


def matrix_to_quaternion(matrix: torch.Tensor) -> torch.Tensor:
    """Rotation matrices ``(...,3,3)`` -> quaternions ``(...,4)``, real part first and >= 0   (ref :164-223, :135-147)."""
    if matrix.size(-1) != 3 or matrix.size(-2) != 3:
        raise ValueError(f"Invalid rotation matrix shape {matrix.shape}.")
    m = matrix
    d0, d1, d2 = m[..., 0, 0], m[..., 1, 1], m[..., 2, 2]
    sq = torch.stack([1 + d0 + d1 + d2, 1 + d0 - d1 - d2, 1 - d0 + d1 - d2, 1 - d0 - d1 + d2], dim=-1)
    q_abs = _sqrt_positive_part(sq)
    a, b, c = m[..., 2, 1] - m[..., 1, 2], m[..., 0, 2] - m[..., 2, 0], m[..., 1, 0] - m[..., 0, 1]
    e, f, g = m[..., 1, 0] + m[..., 0, 1], m[..., 0, 2] + m[..., 2, 0], m[..., 1, 2] + m[..., 2, 1]
    q2 = q_abs ** 2
    cands = torch.stack([
        torch.stack([q2[..., 0], a, b, c], dim=-1),
        torch.stack([a, q2[..., 1], e, f], dim=-1),
        torch.stack([b, e, q2[..., 2], g], dim=-1),
        torch.stack([c, f, g, q2[..., 3]], dim=-1)], dim=-2)
    cands = cands / (2.0 * q_abs[..., None].clamp_min(0.1))
    best = q_abs.argmax(dim=-1)
    quat = torch.gather(cands, -2, best[..., None, None].expand(*best.shape, 1, 4)).squeeze(-2)
    return torch.where(quat[..., :1] < 0, -quat, quat)


def quaternion_to_axis_angle(quaternions: torch.Tensor) -> torch.Tensor:
    """ref :258-277."""
    norms = quaternions[..., 1:].norm(dim=-1, keepdim=True)
    half = torch.atan2(norms, quaternions[..., :1])
    return quaternions[..., 1:] / (0.5 * torch.sinc(half / math.pi))


def matrix_to_axis_angle(matrix: torch.Tensor) -> torch.Tensor:
    """ref :280-298 (the default, quaternion route)."""
    return quaternion_to_axis_angle(matrix_to_quaternion(matrix))


def axis_angle_to_quaternion(axis_angle: torch.Tensor) -> torch.Tensor:
    angles = axis_angle.norm(dim=-1, keepdim=True)
    k = 0.5 * torch.sinc(0.5 * angles / math.pi)
    return torch.cat([torch.cos(0.5 * angles), axis_angle * k], dim=-1)


def quaternion_to_matrix(q: torch.Tensor) -> torch.Tensor:
    r, i, j, k = q.unbind(-1)
    s = 2.0 / (q * q).sum(-1)
    rows = (1 - s * (j * j + k * k), s * (i * j - k * r), s * (i * k + j * r),
            s * (i * j + k * r), 1 - s * (i * i + k * k), s * (j * k - i * r),
            s * (i * k - j * r), s * (j * k + i * r), 1 - s * (i * i + j * j))
    return torch.stack(rows, dim=-1).reshape(q.shape[:-1] + (3, 3))


def axis_angle_to_matrix(axis_angle: torch.Tensor) -> torch.Tensor:
    return quaternion_to_matrix(axis_angle_to_quaternion(axis_angle))


def rotation_matrix_z(rad: torch.Tensor) -> torch.Tensor:
    """``[b]`` angles -> ``[b,3,3]`` rotations about +z, counter-clockwise for ``R @ v``   (ref:cs_vit/utils/geometry.py:9-40;
    the data sets' in-plane augmentation uses it, ref:cs_vit/dataset/DexYCB.py:169-172)."""
    c, s_ = torch.cos(rad), torch.sin(rad)
    z, o = torch.zeros_like(c), torch.ones_like(c)
    return torch.stack([torch.stack([c, -s_, z], -1), torch.stack([s_, c, z], -1), torch.stack([z, z, o], -1)], -2)
