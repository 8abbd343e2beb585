"""Deterministic synthetic stand-in for the MANO right-hand layer.

The reference builds its hand mesh with ``smplx.create(path, "mano", is_rhand=True, use_pca=False)``
(ref:cs_vit/net/ti_poser.py:268-270) and calls it as
``layer(betas=[n,10], global_orient=[n,3], hand_pose=[n,45], transl=[n,3]).vertices -> [n,778,3]`` metres
(ref:cs_vit/net/ti_poser.py:573-578).  The real MANO pickles are licence-gated and ``smplx`` is not in this
image, so neither the reference nor this repo can run the true layer here.  ``SyntheticMANO`` keeps MANO's
I/O contract and its *structure* (shape blend → joint regression → Rodrigues per joint → 16-joint kinematic
chain → linear-blend skinning) with seeded random template / blend weights, so rotation and shape errors
propagate to joints and vertices the way they do through the real layer.  The parity harness injects the
very same class into the reference through a ``smplx`` stub, so both sides see identical FK.

This is an asset substitute, not an approximation of MANO geometry: numbers produced through it are only
meaningful for parity and throughput work, never for accuracy claims.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn as nn

# MANO kinematic tree: wrist, then index / middle / pinky / ring / thumb chains of three joints each.
MANO_PARENTS = (-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14)
NUM_VERTS = 778
NUM_JOINTS = 16


def rodrigues(aa: torch.Tensor) -> torch.Tensor:
    """Axis-angle ``[..., 3]`` → rotation matrices ``[..., 3, 3]`` (safe at zero angle)."""
    theta = torch.sqrt((aa * aa).sum(-1, keepdim=True) + 1e-16)
    k = aa / theta
    kx, ky, kz = k.unbind(-1)
    zero = torch.zeros_like(kx)
    K = torch.stack([zero, -kz, ky, kz, zero, -kx, -ky, kx, zero], dim=-1).reshape(aa.shape[:-1] + (3, 3))
    s = torch.sin(theta)[..., None]
    c = torch.cos(theta)[..., None]
    eye = torch.eye(3, dtype=aa.dtype, device=aa.device).expand(K.shape)
    return eye + s * K + (1.0 - c) * (K @ K)


class SyntheticMANO(nn.Module):
    """Seeded LBS hand with MANO's call signature.  All tensors are buffers (no trainable state)."""

    def __init__(self, seed: int = 1234):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        # A flat, hand-sized point cloud (metres): palm blob plus five finger rays.
        tmpl = torch.randn(NUM_VERTS, 3, generator=g) * torch.tensor([0.035, 0.045, 0.010])
        finger = torch.randint(0, 5, (NUM_VERTS,), generator=g)
        reach = torch.rand(NUM_VERTS, generator=g) * 0.09
        ang = (finger.float() - 2.0) * 0.28
        tmpl[:, 0] += torch.sin(ang) * reach
        tmpl[:, 1] += torch.cos(ang) * reach
        self.register_buffer("v_template", tmpl)
        self.register_buffer("shapedirs", torch.randn(NUM_VERTS, 3, 10, generator=g) * 2.5e-3)
        jreg = torch.softmax(torch.randn(NUM_JOINTS, NUM_VERTS, generator=g) * 3.0, dim=-1)
        self.register_buffer("J_regressor", jreg)
        self.register_buffer("lbs_weights", torch.softmax(torch.randn(NUM_VERTS, NUM_JOINTS, generator=g) * 4.0, dim=-1))
        self.parents = MANO_PARENTS

    def forward(self, betas, global_orient, hand_pose, transl=None, **_unused):
        n = betas.shape[0]
        v_shaped = self.v_template[None] + torch.einsum("vdk,nk->nvd", self.shapedirs, betas)
        joints = torch.einsum("jv,nvd->njd", self.J_regressor, v_shaped)
        pose = torch.cat([global_orient.reshape(n, 1, 3), hand_pose.reshape(n, NUM_JOINTS - 1, 3)], dim=1)
        rot = rodrigues(pose)  # [n,16,3,3]

        # Forward kinematics along the tree, as 3x4 world transforms.
        world_r = [rot[:, 0]]
        world_t = [joints[:, 0]]
        for j in range(1, NUM_JOINTS):
            p = self.parents[j]
            rel = joints[:, j] - joints[:, p]
            world_r.append(world_r[p] @ rot[:, j])
            world_t.append(world_t[p] + (world_r[p] @ rel[..., None])[..., 0])
        world_r = torch.stack(world_r, dim=1)  # [n,16,3,3]
        world_t = torch.stack(world_t, dim=1)  # [n,16,3]
        # Remove the rest pose so that identity rotations reproduce v_shaped.
        rest_t = world_t - (world_r @ joints[..., None])[..., 0]

        blended_r = torch.einsum("vj,njab->nvab", self.lbs_weights, world_r)
        blended_t = torch.einsum("vj,nja->nva", self.lbs_weights, rest_t)
        verts = (blended_r @ v_shaped[..., None])[..., 0] + blended_t
        if transl is not None:
            verts = verts + transl[:, None, :]
            world_t = world_t + transl[:, None, :]
        return SimpleNamespace(vertices=verts, joints=world_t)


def synthetic_joint_regressor(seed: int = 4321) -> torch.Tensor:
    """``[21, 778]`` row-stochastic matrix standing in for ``sh_joint_regressor.npy``.

    The reference ships the real regressor as a data file next to ``ti_poser.py``
    (ref:cs_vit/net/ti_poser.py:274-277) and also stores it in every checkpoint as the persistent buffer
    ``J_regressor_mano``; a loaded checkpoint therefore overrides this stand-in.
    """
    g = torch.Generator().manual_seed(seed)
    return torch.softmax(torch.randn(21, NUM_VERTS, generator=g) * 3.0, dim=-1)
