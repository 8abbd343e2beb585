"""The two kernels of the fp32 tail (csrc/tail.cu) against the torch code they replace: cs_vit.utils.geometry (restating
ref:cs_vit/utils/geometry.py, pinned through the reference goldens' pose_aa) and Poser._pose_fk on the MANO stand-in."""
import math

import pytest
import torch

from helpers import build_product, rel

pytestmark = pytest.mark.gpu


def test_rot6d_to_axis_angle_matches_torch_form():
    from cs_vit import ops
    from cs_vit.utils.geometry import axis_angle_to_matrix, matrix_to_axis_angle, rotation_6d_to_matrix
    g = torch.Generator(device="cuda").manual_seed(0)
    d6 = torch.randn(4096, 16, 6, device="cuda", generator=g)
    # rotations close to pi about the axes exercise the three non-real quaternion candidates; tiny rotations the sinc limit
    for k, axis in enumerate(torch.eye(3, device="cuda")):
        R = axis_angle_to_matrix(axis[None] * (math.pi - 1e-3 * torch.rand(64, 1, device="cuda", generator=g)))
        d6[k * 64:(k + 1) * 64, 0] = R[:, :2].reshape(64, 6)
    d6[300:364, 1] = torch.tensor([1.0, 0, 0, 0, 1.0, 0], device="cuda") + 1e-6 * torch.randn(64, 6, device="cuda", generator=g)
    d6[400, 2] = torch.tensor([1.0, 0, 0, 0, 1.0, 0], device="cuda")                 # exact identity: half angle 0
    got = ops.rot6d_to_axis_angle(d6)
    want = matrix_to_axis_angle(rotation_6d_to_matrix(d6))
    assert got.shape == want.shape and torch.isfinite(got).all()
    # compare as rotations (axis-angle is discontinuous at pi) and element-wise away from pi
    assert rel(axis_angle_to_matrix(got), axis_angle_to_matrix(want)) < 1e-5
    away = want.norm(dim=-1) < 3.0
    assert (got - want)[away].abs().max().item() < 1e-4


@pytest.mark.parametrize("n", [1, 5, 256])
def test_mano_fk_matches_pose_fk(n):
    model, _, _, _ = build_product("swint_encoder_patch_spatial", "fp32")
    model = model.cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(n)
    pose = torch.randn(n, 1, 16, 3, device="cuda", generator=g) * 0.6
    shape = torch.randn(n, 1, 10, device="cuda", generator=g)
    root = torch.randn(n, 1, 3, device="cuda", generator=g)
    with torch.no_grad():
        got = model._pose_fk(pose, shape, root)                                   # fused kernel (no grad)
    with torch.enable_grad():
        want = model._pose_fk(pose.clone().requires_grad_(True), shape, root)     # torch path (differentiable)
    for a, b, name in zip(got, want, ("joint_cam", "verts_cam", "root_transl")):
        assert a.shape == b.shape, name
        assert rel(a, b.detach()) < 2e-5, (name, rel(a, b.detach()))
