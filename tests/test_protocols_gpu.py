"""GPU tests of the two warp-specialised barrier protocols added in round 2, beyond value parity: repeated launches must be bit-identical
(a race between the softmax sets' MUFU turns, the TMEM write-backs, or the fused MLP's separate residual-epilogue warps and its second
GEMM2 accumulator would show up as nondeterminism), on ragged / tiny / more-than-one-wave sizes."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from cs_vit import ops as o
    return o


@pytest.mark.parametrize("C", [128, 256])
@pytest.mark.parametrize("M", [1, 129, 5000, 148 * 128 + 1, 2 * 148 * 128 + 77])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_mlp_fused_repeatable_on_ragged_sizes(ops, C, M, dtype):
    g = torch.Generator(device="cuda").manual_seed(M + C)
    xn = torch.randn(M, C, device="cuda", generator=g).to(dtype)
    w1 = (torch.randn(4 * C, C, device="cuda", generator=g) * C ** -0.5).to(dtype)
    b1 = 0.1 * torch.randn(4 * C, device="cuda", generator=g)
    w2 = (torch.randn(C, 4 * C, device="cuda", generator=g) * (4 * C) ** -0.5).to(dtype)
    b2 = 0.1 * torch.randn(C, device="cuda", generator=g)
    x0 = torch.randn(M, C, device="cuda", generator=g)
    outs = []
    for _ in range(3):
        x = x0.clone()
        ops.mlp_fused(xn, w1, b1, w2, b2, x)
        outs.append(x)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    hidden = torch.nn.functional.gelu(xn.float() @ w1.float().T + b1).to(dtype).float()        # the kernel rounds the hidden chunk to 16 bit
    ref = x0 + hidden @ w2.float().T + b2
    err = ((outs[0] - ref).norm() / ref.norm()).item()
    assert err < (6e-3 if dtype == torch.bfloat16 else 1e-3), err


@pytest.mark.parametrize("H,heads,shift,B", [(16, 16, 0, 37), (32, 8, 8, 9), (64, 4, 8, 3), (32, 3, 0, 5), (16, 1, 0, 300)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_swinv2_attn_tc_repeatable(ops, H, heads, shift, B, dtype):
    g = torch.Generator(device="cuda").manual_seed(H + heads + B)
    C = heads * 32
    rows = B * H * H
    qkv = torch.randn(rows, 3 * C, device="cuda", generator=g).view(rows, 3, heads, 32)
    qkv[:, 0] = torch.nn.functional.normalize(qkv[:, 0], dim=-1) * 20.0
    qkv[:, 1] = torch.nn.functional.normalize(qkv[:, 1], dim=-1)
    qn = qkv.view(rows, 3 * C).to(dtype)
    bl = ops.swinv2_bias_log2((16 * torch.sigmoid(2 * torch.randn(heads, 961, device="cuda", generator=g))).contiguous())
    outs = [ops.swinv2_attn_tc(qn, bl, B, H, H, heads, shift, token_order=True) for _ in range(3)]
    assert torch.isfinite(outs[0].float()).all()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
