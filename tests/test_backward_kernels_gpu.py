"""Per-kernel GPU checks of the training-step (backward) kernels through the C ABI, each against fp32/fp64 torch math
or torch autograd of the same op on the same device tensors."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def ops():
    from cs_vit import ops as o
    return o


# ------------------------------------------------------------------------------------------------ gemm_ex
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 384, 192), (1000, 384, 136), (6272, 512, 2048), (77, 96, 1024),
                                   (128, 128, 6272), (1536, 512, 25088), (96, 48, 6272)])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("dtype", ["bf16", "fp16", "tf32", "fp32"])
def test_gemm_ex_layouts(ops, M, N, K, a_mn, b_mn, dtype):
    g = torch.Generator(device="cuda").manual_seed(M + 3 * N + 7 * K + int(a_mn) + 2 * int(b_mn))
    td = {"bf16": torch.bfloat16, "fp16": torch.float16}.get(dtype, torch.float32)

    def operand(rows, cols, scale):   # logical rows x cols inside a parent whose pitch is a 16-byte multiple (TMA)
        parent = (torch.randn(rows, -(-cols // 8) * 8, device="cuda", generator=g) * scale).to(td)
        return parent[:, :cols]

    A_ = operand(K, M, 1.0) if a_mn else operand(M, K, 1.0)
    B_ = operand(K, N, 0.05) if b_mn else operand(N, K, 0.05)
    if dtype in ("bf16", "fp16"):
        impl, tol = ops.GEMM_TC, 1e-5
    else:
        impl, tol = (ops.GEMM_TC, 2e-3) if dtype == "tf32" else (ops.GEMM_SIMT, 1e-5)
    Am = (A_.T if a_mn else A_).double()
    Bm = (B_.T if b_mn else B_).double()
    ref = Am @ Bm.T
    out = ops.gemm_ex(A_, a_mn, B_, b_mn, out_dtype=torch.float32, impl=impl)
    torch.cuda.synchronize()
    assert out.shape == (M, N)
    assert rel(out, ref) < tol, (dtype, M, N, K, a_mn, b_mn, rel(out, ref))


def test_gemm_ex_accumulate_and_split(ops):
    g = torch.Generator(device="cuda").manual_seed(5)
    T, O, I = 12544, 384, 128
    dy = torch.randn(T, O, device="cuda", generator=g).to(torch.bfloat16)
    x = torch.randn(T, I, device="cuda", generator=g).to(torch.bfloat16)
    ref = dy.double().T @ x.double()
    for split in (0, 1, 7):
        base = torch.randn(O, I, device="cuda", generator=g)
        out = base.clone()
        ops.gemm_ex(dy, True, x, True, out=out, accumulate=True, split_k=split)
        torch.cuda.synchronize()
        assert rel(out, ref + base.double()) < 3e-5, split
        out2 = ops.gemm_ex(dy, True, x, True, out_dtype=torch.float32, split_k=split)
        assert rel(out2, ref) < 3e-5, split
    # 16-bit output (dgrad feeding the next kernel)
    w = (torch.randn(O, I, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    dx = ops.gemm_ex(dy, False, w, True, out_dtype=torch.bfloat16)
    assert rel(dx, dy.double() @ w.double()) < 5e-3


# ------------------------------------------------------------------------------------------------ row kernels
@pytest.mark.parametrize("src", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("rows,C", [(6272, 96), (1000, 128), (392, 1024), (63, 3072)])
def test_col_reduce_modes(ops, src, rows, C):
    g = torch.Generator(device="cuda").manual_seed(rows + C)
    a = (torch.randn(rows, C, device="cuda", generator=g) + 0.3).to(src)
    b = torch.randn(rows, C, device="cuda", generator=g)
    s1, _, _ = ops.col_reduce(a)
    assert rel(s1, a.double().sum(0)) < 1e-5
    mean = (s1 / rows).contiguous()
    d1, d2, _ = ops.col_reduce(a, ops.CR_CENTERED, center=mean)
    assert rel(d2, ((a.double() - mean.double()) ** 2).sum(0)) < 1e-5
    assert d1.abs().max().item() < 1e-2 * rows ** 0.5
    e1, e2, cp = ops.col_reduce(a, ops.CR_DOT, b=b, copy_dtype=torch.bfloat16)
    assert rel(e1, a.double().sum(0)) < 1e-5 and rel(e2, (a.double() * b.double()).sum(0)) < 1e-5
    assert torch.equal(cp, a.float().to(torch.bfloat16))


@pytest.mark.parametrize("H,shift", [(56, 0), (28, 3), (14, 3)])
def test_col_reduce_window_gather(ops, H, shift):
    C, B = 128, 3
    g = torch.Generator(device="cuda").manual_seed(H + shift)
    a = torch.randn(B * H * H, C, device="cuda", generator=g)
    idx = ops.window_index_map(H, H, 7, shift).long()
    s1, _, cp = ops.col_reduce(a, window=(H, H, 7, shift), copy_dtype=torch.float16)
    want = a.view(B, H * H, C)[:, idx].reshape(-1, C)
    assert torch.equal(cp, want.to(torch.float16))
    assert rel(s1, a.double().sum(0)) < 1e-5
    _, _, cp32 = ops.col_reduce(a, window=(H, H, 7, shift), copy_dtype=torch.float32, sums=False)
    assert torch.equal(cp32, want)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
def test_eltwise(ops, dt):
    g = torch.Generator(device="cuda").manual_seed(3)
    x = (torch.randn(777, 512, device="cuda", generator=g) * 2).to(dt)
    dy = torch.randn(777, 512, device="cuda", generator=g).to(dt)
    tol = 1e-6 if dt == torch.float32 else 6e-3
    xr = x.double().requires_grad_(True)
    yr = torch.nn.functional.gelu(xr)
    assert rel(ops.eltwise(ops.EW_GELU_FWD, x), yr) < tol
    (gr,) = torch.autograd.grad(yr, xr, dy.double())
    assert rel(ops.eltwise(ops.EW_GELU_BWD, dy, x), gr) < tol
    y = torch.relu(x)
    assert torch.equal(ops.eltwise(ops.EW_RELU_BWD, dy, y), torch.where(y > 0, dy, torch.zeros_like(dy)))


def test_affine2_rows(ops):
    g = torch.Generator(device="cuda").manual_seed(4)
    dy, x, r = (torch.randn(333, 768, device="cuda", generator=g) for _ in range(3))
    a, b, c = (torch.randn(768, device="cuda", generator=g) for _ in range(3))
    assert rel(ops.affine2_rows(dy, x, a, b, c), a * dy + b * x + c) < 1e-6
    assert rel(ops.affine2_rows(dy, x, a, b, c, resid=r), a * dy + b * x + c + r) < 1e-6


@pytest.mark.parametrize("C", [96, 128, 384, 512, 1024])
@pytest.mark.parametrize("dyt", [torch.float32, torch.bfloat16])
def test_layernorm_bwd_identity(ops, C, dyt):
    g = torch.Generator(device="cuda").manual_seed(C)
    rows = 1237
    x = torch.randn(rows, C, device="cuda", generator=g) * 1.7 + 0.4
    gamma = torch.randn(C, device="cuda", generator=g)
    beta = torch.randn(C, device="cuda", generator=g)
    dy = torch.randn(rows, C, device="cuda", generator=g).to(dyt)
    dres = torch.randn(rows, C, device="cuda", generator=g)
    xr, gr, br = x.double().requires_grad_(True), gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    y = torch.nn.functional.layer_norm(xr, (C,), gr, br, 1e-5)
    ex, eg, eb = torch.autograd.grad(y, (xr, gr, br), dy.double())
    dx, dg, db = ops.layernorm_bwd(x, dy, gamma, 1e-5, dres=dres)
    assert rel(dx, ex + dres.double()) < 1e-5 and rel(dg, eg) < 1e-5 and rel(db, eb) < 1e-5
    dx2, _, _ = ops.layernorm_bwd(x, dy, gamma, 1e-5)
    assert rel(dx2, ex) < 1e-5


@pytest.mark.parametrize("H,shift,C", [(56, 3, 96), (28, 0, 256), (14, 3, 512)])
def test_layernorm_bwd_window(ops, H, shift, C):
    B = 2
    g = torch.Generator(device="cuda").manual_seed(H * C)
    x = torch.randn(B * H * H, C, device="cuda", generator=g)
    gamma = torch.randn(C, device="cuda", generator=g)
    dy = torch.randn(B * H * H, C, device="cuda", generator=g)   # window-ordered rows
    dres = torch.randn(B * H * H, C, device="cuda", generator=g)
    idx = ops.window_index_map(H, H, 7, shift).long()
    xr = x.double().requires_grad_(True)
    y = torch.nn.functional.layer_norm(xr, (C,), gamma.double(), None, 1e-5).view(B, H * H, C)[:, idx].reshape(-1, C)
    (ex,) = torch.autograd.grad(y, xr, dy.double())
    dx, dg, db = ops.layernorm_bwd(x, dy, gamma, 1e-5, mode=ops.LN_WINDOW, grid=(H, H), ws=7, shift=shift, dres=dres)
    assert rel(dx, ex + dres.double()) < 1e-5
    assert rel(db, dy.double().sum(0)) < 1e-5


@pytest.mark.parametrize("H,C", [(56, 96), (14, 512)])
def test_layernorm_bwd_merge(ops, H, C):
    B = 2
    g = torch.Generator(device="cuda").manual_seed(H + C)
    x = torch.randn(B * H * H, C, device="cuda", generator=g)
    gamma = torch.randn(4 * C, device="cuda", generator=g)
    dy = torch.randn(B * (H // 2) ** 2, 4 * C, device="cuda", generator=g)
    midx = ops.merge_index_map(H, H).long()     # [(H/2)^2, 4]
    xr = x.double().requires_grad_(True)
    cat = xr.view(B, H * H, C)[:, midx].reshape(B * (H // 2) ** 2, 4 * C)
    y = torch.nn.functional.layer_norm(cat, (4 * C,), gamma.double(), None, 1e-5)
    (ex,) = torch.autograd.grad(y, xr, dy.double())
    dx, dg, db = ops.layernorm_bwd(x, dy, gamma, 1e-5, mode=ops.LN_MERGE2X2, grid=(H, H))
    assert rel(dx, ex) < 1e-5 and rel(db, dy.double().sum(0)) < 1e-5


# ------------------------------------------------------------------------------------------------ attention backward
def _attn_ref(q, k, v, n, L, S, h, scale, bias=None, mask=None):
    qh = q.view(n, L, h, 32).transpose(1, 2)
    kh = k.view(n, S, h, 32).transpose(1, 2)
    vh = v.view(n, S, h, 32).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) * scale
    if bias is not None:
        s = s + bias[None]
    if mask is not None:
        s = s + mask
    return (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(n * L, h * 32)


@pytest.mark.parametrize("L,S,heads,n", [(52, 52, 24, 5), (3, 49, 32, 7), (1, 8, 24, 33), (3, 3, 24, 4), (67, 67, 8, 3), (3, 128, 4, 2), (100, 65, 2, 2)])
def test_attention_bwd_dense(ops, L, S, heads, n):
    g = torch.Generator(device="cuda").manual_seed(L * S)
    D = heads * 32
    q = torch.randn(n * L, D, device="cuda", generator=g) * 0.3
    k = torch.randn(n * S, D, device="cuda", generator=g) * 0.3
    v = torch.randn(n * S, D, device="cuda", generator=g)
    do = torch.randn(n * L, D, device="cuda", generator=g)
    scale = math.sqrt(32.0)     # quirk Q1: logits multiplied by sqrt(d)
    qr, kr, vr = (t.double().requires_grad_(True) for t in (q, k, v))
    out = _attn_ref(qr, kr, vr, n, L, S, heads, scale)
    eq, ek, ev = torch.autograd.grad(out, (qr, kr, vr), do.double())
    assert rel(ops.attention(q, k, v, n, L, S, heads, scale), out) < 1e-5
    dq, dk, dv, db = ops.attention_bwd(q, k, v, do, n, L, S, heads, scale)
    assert db is None
    assert rel(dq, eq) < 2e-5 and rel(dk, ek) < 2e-5 and rel(dv, ev) < 2e-5, (rel(dq, eq), rel(dk, ek), rel(dv, ev))


@pytest.mark.parametrize("H,shift,heads", [(14, 3, 12), (28, 0, 6), (7, 0, 24), (28, 3, 8)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
def test_window_attention_bwd(ops, H, shift, heads, dt):
    B, ws, L = 2, 7, 49
    C = heads * 32
    nW = (H // ws) ** 2
    g = torch.Generator(device="cuda").manual_seed(H * 10 + shift + heads)
    qkv = (torch.randn(B * H * H, 3 * C, device="cuda", generator=g) * 0.7).to(dt)
    do = torch.randn(B * H * H, C, device="cuda", generator=g).to(dt)
    table = torch.randn(169, heads, device="cuda", generator=g) * 0.5
    bias = ops.expand_rel_bias(table, ws)
    mask = ops.shift_mask(H, H, ws, shift).view(1, nW, 1, L, L).expand(B, -1, -1, -1, -1).reshape(B * nW, 1, L, L) if shift else None
    qr = qkv.double().requires_grad_(True)
    br = bias.double().requires_grad_(True)
    out = _attn_ref(qr[:, :C], qr[:, C:2 * C], qr[:, 2 * C:], B * nW, L, L, heads, 1 / math.sqrt(32.0), br, mask.double() if shift else None)
    eqkv, ebias = torch.autograd.grad(out, (qr, br), do.double())
    dqkv = torch.empty_like(qkv)
    _, _, _, dbias = ops.attention_bwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], do, B * nW, L, L, heads, 1 / math.sqrt(32.0),
                                      bias=bias, mask=(H, H, ws, shift), dq=dqkv[:, :C], dk=dqkv[:, C:2 * C], dv=dqkv[:, 2 * C:])
    # 16-bit: tensor-core kernel (attention_bwd_mma.cu): P and dS are rounded to the operand format for the second MMA
    tol = {torch.float32: 2e-5, torch.bfloat16: 8e-3, torch.float16: 2e-3}[dt]
    assert rel(dqkv, eqkv) < tol, rel(dqkv, eqkv)
    assert rel(dbias, ebias) < (1e-4 if dt == torch.float32 else tol), rel(dbias, ebias)
