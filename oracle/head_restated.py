"""CPU fp32 restatement of the CS-ViT head and of ``Poser.predict_batch``.  TEST INFRASTRUCTURE.

Pure functions of a ``state_dict`` in the reference's key schema plus the constructor options.
"ref:" = /root/reference.  Quirk numbers (Q1..Q9) refer to SURVEY.md §0.5 - they are reproduced on purpose.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F

from . import swin_restated as swin

Tensor = torch.Tensor
SD = Dict[str, Tensor]

IMAGENET_MEAN = (0.485, 0.456, 0.406)   # ref:cs_vit/net/ti_poser.py:239-243
IMAGENET_STD = (0.229, 0.224, 0.225)
BN_EPS = 1e-5

# ref:cs_vit/constants.py:96-121 TARGET_JOINTS_CONNECTION: wrist to the five finger bases, then each finger chain.
SKELETON_EDGES = [(0, b) for b in (1, 5, 9, 13, 17)] + [(b + k, b + k + 1) for b in (1, 5, 9, 13, 17) for k in range(3)]


@dataclass
class HeadOptions:
    """Constructor options of ``Poser`` that change the forward   (ref:cs_vit/net/ti_poser.py:192-210)."""
    num_heads: int
    depths: tuple
    swin_heads: tuple
    num_spatial_layer: int = 6
    spatial_layer_type: str = "decoder"
    num_temporal_layer: int = 2
    temporal_supervision: str = "full"
    trope_scalar: float = 20.0
    persp_embed_method: str = "dense"
    persp_decorate: str = "query"
    phase: str = "inference"           # "spatial" | "temporal" | "inference"
    window_size: int = 7


# ------------------------------------------------------------------------------------------------ building blocks
def linear(x: Tensor, sd: SD, p: str) -> Tensor:
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def batchnorm_tokens(x: Tensor, sd: SD, p: str, training: bool = False) -> Tensor:
    """``norm(x.transpose(-1,-2)).transpose(-1,-2)`` with ``BatchNorm1d(D)`` on x [n, L, D] (Q3).

    ref:cs_vit/net/transformer_module.py:312,316 (and :341,345,349, :370,374).  Eval: running statistics.
    Train: statistics over the (n, L) axes of this batch, biased variance.
    """
    if training:
        mean = x.mean(dim=(0, 1))
        var = x.var(dim=(0, 1), unbiased=False)
    else:
        mean, var = sd[p + ".running_mean"], sd[p + ".running_var"]
    return (x - mean) / torch.sqrt(var + BN_EPS) * sd[p + ".weight"] + sd[p + ".bias"]


def batchnorm_rows(x: Tensor, sd: SD, p: str, training: bool = False) -> Tensor:
    """``BatchNorm1d`` applied to x [n, D] (PerspectiveEncoder, ref:cs_vit/net/ti_poser.py:171-176)."""
    return batchnorm_tokens(x[:, None, :], sd, p, training)[:, 0, :]


def mha(x: Tensor, ctx: Tensor, sd: SD, p: str, heads: int) -> Tensor:
    """ref:cs_vit/net/transformer_module.py:250-282.  Logits are MULTIPLIED by sqrt(head_dim) (Q1, line 273)."""
    n, L, D = x.shape
    S = ctx.shape[1]
    d = D // heads
    q = linear(x, sd, p + ".query").reshape(n, L, heads, d).transpose(1, 2)
    k = linear(ctx, sd, p + ".key").reshape(n, S, heads, d).transpose(1, 2)
    v = linear(ctx, sd, p + ".value").reshape(n, S, heads, d).transpose(1, 2)
    inv_sqrt = 1.0 / (d ** 0.5)
    scores = (q @ k.transpose(-1, -2)) / inv_sqrt
    out = (scores.softmax(-1) @ v).transpose(1, 2).reshape(n, L, D)
    return linear(out, sd, p + ".output")


def ffn(x: Tensor, sd: SD, p: str) -> Tensor:
    """ref:cs_vit/net/transformer_module.py:285-297: Linear(D,4D) -> exact GELU -> Linear(4D,D)."""
    return linear(F.gelu(linear(x, sd, p + ".net.0")), sd, p + ".net.2")


def encoder_block(x: Tensor, sd: SD, p: str, heads: int, training=False) -> Tensor:
    """ref:cs_vit/net/transformer_module.py:309-319."""
    y = batchnorm_tokens(x, sd, p + ".norm1", training)
    x = x + mha(y, y, sd, p + ".attn", heads)
    y = batchnorm_tokens(x, sd, p + ".norm2", training)
    return x + ffn(y, sd, p + ".ffn")


def decoder_block(x: Tensor, ref: Tensor, sd: SD, p: str, heads: int, training=False) -> Tensor:
    """ref:cs_vit/net/transformer_module.py:333-353 (``ref`` is NOT normalised)."""
    y = batchnorm_tokens(x, sd, p + ".norm1", training)
    x = x + mha(y, y, sd, p + ".self_atten", heads)
    y = batchnorm_tokens(x, sd, p + ".norm2", training)
    x = x + mha(y, ref, sd, p + ".cross_atten", heads)
    y = batchnorm_tokens(x, sd, p + ".norm3", training)
    return x + ffn(y, sd, p + ".ffn")


def cross_attn_decoder(x: Tensor, ref: Tensor, sd: SD, p: str, heads: int, training=False) -> Tensor:
    """ref:cs_vit/net/transformer_module.py:364-378."""
    y = batchnorm_tokens(x, sd, p + ".norm1", training)
    x = x + mha(y, ref, sd, p + ".cross_atten", heads)
    y = batchnorm_tokens(x, sd, p + ".norm2", training)
    return x + ffn(y, sd, p + ".ffn")


def absolute_pe(x: Tensor, sd: SD, p: str) -> Tensor:
    """Learned absolute PE: rows 0..L-1 of ``Embedding(512, D)``   (ref:cs_vit/net/transformer_module.py:46-49)."""
    return x + sd[p + ".pe.weight"][: x.shape[1]][None]


def trope_pe(x: Tensor, t: Tensor, sd: SD, p: str) -> Tensor:
    """"trope": rotate input pairs (2i, 2i+1) by (t_last - t) * inv_freq_i (Q6).

    ref:cs_vit/net/transformer_module.py:54-81.  ``t`` is already divided by trope_scalar by the caller.
    """
    delta = (t[:, -1:] - t).float()                               # [n, T]
    ang = delta[..., None] * sd[p + ".inv_freq"][None, None]      # [n, T, D/2]
    c, s = torch.cos(ang), torch.sin(ang)
    pairs = x.reshape(*x.shape[:-1], -1, 2)
    a, b = pairs[..., 0], pairs[..., 1]
    return torch.stack([a * c - b * s, a * s + b * c], dim=-1).flatten(-2)


def spatial_encoder(q: Tensor, patches: Tensor, sd: SD, opt: HeadOptions, training=False, execute_all=True) -> Tensor:
    """ref:cs_vit/net/ti_poser.py:80-97.

    "encoder" type: every layer is applied to the SAME embedded input and only the last layer's result is
    returned (Q2).  ``execute_all`` keeps the five discarded layers in the computation, as the reference
    does (it matters only for the timed CPU baseline; the result is identical).
    """
    p = "spatial_encoder"
    if opt.spatial_layer_type == "decoder":
        x = absolute_pe(q, sd, p + ".pe_spatial")
        for l in range(opt.num_spatial_layer):
            x = decoder_block(x, patches, sd, f"{p}.layers.{l}", opt.num_heads, training)
        return x
    z = absolute_pe(torch.cat([q, patches], dim=1), sd, p + ".pe_spatial")
    out = None
    first = 0 if execute_all else opt.num_spatial_layer - 1
    for l in range(first, opt.num_spatial_layer):
        out = encoder_block(z, sd, f"{p}.layers.{l}", opt.num_heads, training)
    return out[:, : q.shape[1]]


def temporal_encoder(x: Tensor, timestamp: Optional[Tensor], sd: SD, p: str, opt: HeadOptions, training=False) -> Tensor:
    """ref:cs_vit/net/ti_poser.py:140-158.  Realtime: last frame attends to all frames (Q7: returns T=1)."""
    if opt.temporal_supervision == "realtime":
        e = trope_pe(x, timestamp / opt.trope_scalar, sd, p + ".pe_temporal")
        u = e[:, -1:]
        for l in range(opt.num_temporal_layer):
            u = cross_attn_decoder(u, e, sd, f"{p}.layers.{l}", opt.num_heads, training)
        return F.linear(u, sd[p + ".zero_conv.weight"])
    e = absolute_pe(x, sd, p + ".pe_temporal")
    for l in range(opt.num_temporal_layer):
        e = encoder_block(e, sd, f"{p}.layers.{l}", opt.num_heads, training)
    return F.linear(e, sd[p + ".zero_conv.weight"])


def perspective_encoder(v: Tensor, sd: SD, training=False) -> Tensor:
    """ref:cs_vit/net/ti_poser.py:161-182: proj, 3 x (BN, Linear, ReLU), Linear.  Sequential indices 0..9."""
    p = "perspective_mlp"
    y = linear(v, sd, p + ".proj")
    for k in range(3):
        y = batchnorm_rows(y, sd, f"{p}.layer.{3 * k}", training)
        y = F.relu(linear(y, sd, f"{p}.layer.{3 * k + 1}"))
    return linear(y, sd, p + ".layer.9")


def perspective_directions_dense(bbox: Tensor, focal: Tensor, princpt: Tensor, num: int = 16) -> Tensor:
    """Unit-ray (x, y) on a num x num grid over the box, [B,T,num,num,2]   (ref:cs_vit/net/ti_poser.py:609-639)."""
    g = torch.linspace(0.5 / num, 1 - 0.5 / num, num)
    xs = bbox[..., 0:1] + (bbox[..., 2:3] - bbox[..., 0:1]) * g          # [B,T,p]
    ys = bbox[..., 1:2] + (bbox[..., 3:4] - bbox[..., 1:2]) * g
    grid = torch.stack([xs[..., :, None].expand(-1, -1, -1, num), ys[..., None, :].expand(-1, -1, num, -1)], dim=-1)
    d = (grid - princpt[:, :, None, None]) / focal[:, :, None, None]
    d = torch.cat([d, torch.ones_like(d[..., :1])], dim=-1)
    return (d / d.norm(dim=-1, keepdim=True))[..., :2]


def perspective_directions_sparse(bbox: Tensor, focal: Tensor, princpt: Tensor) -> Tensor:
    """Normalised coordinates of the four box corners, [B,T,2,2,2]   (ref:cs_vit/net/ti_poser.py:670-683)."""
    u0 = (bbox[..., 0] - princpt[..., 0]) / focal[..., 0]
    u1 = (bbox[..., 2] - princpt[..., 0]) / focal[..., 0]
    v0 = (bbox[..., 1] - princpt[..., 1]) / focal[..., 1]
    v1 = (bbox[..., 3] - princpt[..., 1]) / focal[..., 1]
    top = torch.stack([torch.stack([u0, v0], -1), torch.stack([u1, v0], -1)], dim=2)
    bot = torch.stack([torch.stack([u0, v1], -1), torch.stack([u1, v1], -1)], dim=2)
    return torch.stack([top, bot], dim=2)


# ------------------------------------------------------------------------------------------------ rotations
def rotation_6d_to_matrix(d6: Tensor) -> Tensor:
    """Gram-Schmidt on the two 3-vectors; rows (b1, b2, b1 x b2)   (ref:cs_vit/utils/geometry.py:111-132)."""
    b1 = F.normalize(d6[..., :3], dim=-1)
    a2 = d6[..., 3:]
    b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=-1)
    return torch.stack([b1, b2, torch.linalg.cross(b1, b2, dim=-1)], dim=-2)


def matrix_to_quaternion(m: Tensor) -> Tensor:
    """Best-conditioned of the four candidate quaternions, real part made non-negative.

    ref:cs_vit/utils/geometry.py:164-223 (+ standardize_quaternion :135-147).
    """
    m00, m01, m02 = m[..., 0, 0], m[..., 0, 1], m[..., 0, 2]
    m10, m11, m12 = m[..., 1, 0], m[..., 1, 1], m[..., 1, 2]
    m20, m21, m22 = m[..., 2, 0], m[..., 2, 1], m[..., 2, 2]
    sq = torch.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22], dim=-1)
    q_abs = torch.where(sq > 0, torch.sqrt(sq.clamp_min(0)), torch.zeros_like(sq))
    cand = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1),
    ], dim=-2)
    cand = cand / (2.0 * q_abs[..., None].clamp_min(0.1))
    pick = q_abs.argmax(dim=-1)
    quat = torch.gather(cand, -2, pick[..., None, None].expand(*pick.shape, 1, 4))[..., 0, :]
    return torch.where(quat[..., :1] < 0, -quat, quat)


def quaternion_to_axis_angle(q: Tensor) -> Tensor:
    """ref:cs_vit/utils/geometry.py:258-277."""
    n = q[..., 1:].norm(dim=-1, keepdim=True)
    half = torch.atan2(n, q[..., :1])
    return q[..., 1:] / (0.5 * torch.sinc(half / math.pi))


def matrix_to_axis_angle(m: Tensor) -> Tensor:
    """ref:cs_vit/utils/geometry.py:280-298 (default, quaternion route)."""
    return quaternion_to_axis_angle(matrix_to_quaternion(m))


def axis_angle_to_matrix(aa: Tensor) -> Tensor:
    """Rodrigues via quaternion, for comparing rotations instead of raw axis-angle."""
    ang = aa.norm(dim=-1, keepdim=True)
    half = 0.5 * ang
    k = 0.5 * torch.sinc(half / math.pi)
    q = torch.cat([torch.cos(half), aa * k], dim=-1)
    r, i, j, kk = q.unbind(-1)
    s = 2.0 / (q * q).sum(-1)
    return torch.stack([
        1 - s * (j * j + kk * kk), s * (i * j - kk * r), s * (i * kk + j * r),
        s * (i * j + kk * r), 1 - s * (i * i + kk * kk), s * (j * kk - i * r),
        s * (i * kk - j * r), s * (j * kk + i * r), 1 - s * (i * i + j * j)], dim=-1).reshape(aa.shape[:-1] + (3, 3))


# ------------------------------------------------------------------------------------------------ top level
def mean_bone_length(joints: Tensor) -> Tensor:
    """ref:cs_vit/utils/joint.py:49-70 over ref:cs_vit/constants.py:96-121."""
    a = torch.tensor([e[0] for e in SKELETON_EDGES])
    b = torch.tensor([e[1] for e in SKELETON_EDGES])
    return (joints[..., a, :] - joints[..., b, :]).norm(dim=-1).mean(dim=-1)


def pose_fk(pose_aa: Tensor, shape: Tensor, root_norm: Tensor, mano, jreg: Tensor):
    """ref:cs_vit/net/ti_poser.py:561-607.  Returns joint_cam [B,T,21,3], verts_cam [B,T,778,3], root_transl (mm)."""
    B, T = pose_aa.shape[:2]
    flat_pose = pose_aa.reshape(B * T, -1)
    out = mano(betas=shape.reshape(B * T, -1), global_orient=flat_pose[:, :3], hand_pose=flat_pose[:, 3:],
               transl=torch.zeros(B * T, 3))
    verts = out.vertices
    joints = torch.einsum("nvd,jv->njd", verts, jreg)
    scale = 1e3 * mean_bone_length(joints).reshape(B, T, 1)
    root = root_norm * scale
    verts_cam = ((verts - joints[:, :1]) * 1e3).reshape(B, T, -1, 3) + root[:, :, None]
    joint_cam = ((joints - joints[:, :1]) * 1e3).reshape(B, T, -1, 3) + root[:, :, None]
    return joint_cam, verts_cam, root


def backbone_features(imgs: Tensor, sd: SD, opt: HeadOptions, prefix: str = "backbone.") -> Tensor:
    """Normalize + Swin (ref:cs_vit/net/ti_poser.py:424-426).  imgs [n,3,S,S] in [0,1]."""
    mean = torch.tensor(IMAGENET_MEAN)[None, :, None, None]
    std = torch.tensor(IMAGENET_STD)[None, :, None, None]
    bsd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    return swin.swin_forward((imgs - mean) / std, bsd, opt.depths, opt.swin_heads, opt.window_size)


def decode_pose(imgs: Tensor, timestamp: Tensor, persp_vec: Tensor, sd: SD, opt: HeadOptions,
                features_fn: Optional[Callable[[Tensor], Tensor]] = None, training=False, execute_all=True):
    """ref:cs_vit/net/ti_poser.py:404-559 without the training-only latent branch (num_latent_layer=None)."""
    B, T = imgs.shape[:2]
    flat = imgs.reshape(B * T, *imgs.shape[2:])
    patches = features_fn(flat) if features_fn is not None else backbone_features(flat, sd, opt)
    persp_bias = perspective_encoder(persp_vec.reshape(B * T, -1), sd, training)
    queries = sd["query_token"][None].repeat(B * T, 1, 1)
    if opt.persp_decorate == "query":
        queries = queries + persp_bias[:, None]
    else:
        patches = patches + persp_bias[:, None]
    tokens = spatial_encoder(queries, patches, sd, opt, training, execute_all)       # [BT, 3, D]
    tokens = tokens.reshape(B, T, 3, -1)
    if opt.phase in ("inference", "temporal"):
        streams = []
        for qi, name in enumerate(("pose", "shape", "root")):
            x = tokens[:, :, qi]                                                       # [B, T, D]
            enc = temporal_encoder(x, timestamp, sd, f"{name}_temporal_encoder", opt, training and opt.phase == "temporal")
            streams.append((x[:, -1:] if opt.temporal_supervision == "realtime" else x) + enc)
        pose_tok, shape_tok, root_tok = streams
    else:
        pose_tok, shape_tok, root_tok = tokens[:, :, 0], tokens[:, :, 1], tokens[:, :, 2]
    pose6d = linear(pose_tok, sd, "pose_decoder.0")
    pose6d = pose6d.reshape(*pose6d.shape[:2], -1, 6)
    pose_aa = matrix_to_axis_angle(rotation_6d_to_matrix(pose6d))
    return pose_aa, linear(shape_tok, sd, "shape_decoder.0"), linear(root_tok, sd, "root_decoder.0")


def predict_batch(inputs: Dict[str, Tensor], sd: SD, opt: HeadOptions, mano, features_fn=None, training=False,
                  execute_all=True) -> Dict[str, Tensor]:
    """ref:cs_vit/net/ti_poser.py:641-722 (global_positioning == "direct")."""
    bbox, focal, princpt = inputs["square_bboxes"], inputs["focal"], inputs["princpt"]
    if opt.persp_embed_method == "dense":
        dirs = perspective_directions_dense(bbox, focal, princpt, 16)
    else:
        dirs = perspective_directions_sparse(bbox, focal, princpt)
    pose_aa, shape, root_norm = decode_pose(inputs["patches"], inputs["timestamp"], dirs, sd, opt, features_fn, training,
                                            execute_all)
    joint_cam, verts_cam, root = pose_fk(pose_aa, shape, root_norm, mano, sd["J_regressor_mano"])
    return {"joint_cam": joint_cam, "verts_cam": verts_cam, "pose_aa": pose_aa, "shape": shape,
            "root_transl_norm": root_norm, "root_transl": root}


def criterion(pred: Dict[str, Tensor], batch: Dict[str, Tensor], opt: HeadOptions) -> Tensor:
    """ref:cs_vit/net/ti_poser.py:724-778."""
    T = pred["joint_cam"].shape[1]
    idx = [-1] if opt.temporal_supervision == "realtime" else list(range(T))
    pj, gj, valid = pred["joint_cam"][:, idx], batch["joint_cam"][:, idx], batch["joint_valid"][:, idx]
    loss = ((pj - gj).norm(dim=-1) * valid).mean()
    loss = loss + (((pj - pj[:, :, :1]) - (gj - gj[:, :, :1])).norm(dim=-1) * valid).mean()
    loss = loss + (pred["shape"][:, idx] - batch["mano_shape"][:, idx]).abs().mean()
    if opt.phase == "temporal" and opt.temporal_supervision == "full":
        def diff(x):
            return (x[:, 2:] - x[:, :-2]) / 2.0
        vp, vg = diff(pred["joint_cam"]), diff(batch["joint_cam"])
        ap, ag = diff(vp), diff(vg)
        loss = loss + 1e-2 * ((vp - vg).norm(dim=-1).mean() + (ap - ag).norm(dim=-1).mean())
    return loss
