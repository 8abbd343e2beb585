"""Per-kernel GPU checks through the C ABI: each CUDA kernel against plain fp32 torch math on the same
device tensors (tolerances: bf16 operands -> 1e-2 relative, tf32 -> 2e-3, fp32 SIMT -> 1e-5)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def ops():
    from cs_vit import ops as o
    return o


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 128), (1000, 384, 128), (12544, 1024, 1024),
                                   (6272, 512, 2048), (200, 96, 48), (77, 10, 1024), (3136 * 2, 1536, 512)])
@pytest.mark.parametrize("dtype", ["bf16", "fp16", "tf32", "fp32"])
def test_linear_plain(ops, M, N, K, dtype):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) * 0.05
    b = torch.randn(N, device="cuda", generator=g)
    if dtype in ("bf16", "fp16"):
        td = torch.bfloat16 if dtype == "bf16" else torch.float16
        a_, w_ = a.to(td), w.to(td)
        ref = a_.float() @ w_.float().T + b
        out = ops.linear(a_, w_, b, out_dtype=torch.float32)
        tol = 1e-5
    elif dtype == "tf32":
        ref = a.double() @ w.double().T + b.double()
        out = ops.linear(a, w, b)
        tol = 2e-3
    else:
        ref = a.double() @ w.double().T + b.double()
        out = ops.linear(a, w, b, impl=ops.GEMM_SIMT)
        tol = 1e-5
    torch.cuda.synchronize()
    assert out.shape == (M, N)
    assert rel(out, ref) < tol, (dtype, M, N, K, rel(out, ref))


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_linear_pair_kernel_matches_single(ops, dt):
    """cta_group::2 kernel (auto-selected for large problems) against the single-CTA kernel and fp32 math."""
    g = torch.Generator(device="cuda").manual_seed(11)
    M, N, K = 148 * 2 * 256 + 77, 512, 320
    a = torch.randn(M, K, device="cuda", generator=g).to(dt)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(dt)
    b = torch.randn(N, device="cuda", generator=g)
    x = torch.randn(M, N, device="cuda", generator=g)
    try:
        outs = {}
        for pair in (0, 1):
            ops.set_gemm_tuning(0, -1, 0, pair)
            outs[pair] = (ops.linear(a, w, b, act=ops.ACT_GELU, out_dtype=dt), ops.linear(a, w, b, resid=x, out_dtype=torch.float32))
    finally:
        ops.set_gemm_tuning()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    ref = a[-3000:].float() @ w.float().T + b + x[-3000:]
    assert rel(outs[1][1][-3000:], ref) < 1e-5


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K", [(50176, 512, 512), (50176, 512, 2048), (50176 - 131, 2048, 512), (12544, 1024, 1024)])
def test_pair_kernel_inplace_residual_tail_split_gelu16(ops, dt, M, N, K):
    """Swin-B stage-2/3 sizes at batch 256 through the CTA-pair kernel's round-2 paths: the IN-PLACE residual epilogue (x += a W^T + b by TMA
    reduce-add), the tail split (N = 512: 392 tiles on 74 pairs -> 22 left-over tiles as 44 half tiles) and the sixteen-warp GELU epilogue
    (K <= 512) - bit-identical to the single-CTA kernel (same tcgen05 accumulation order), equal to fp32 math, and reproducible."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(dt)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(dt)
    b = torch.randn(N, device="cuda", generator=g)
    x0 = torch.randn(M, N, device="cuda", generator=g)
    try:
        outs = {}
        for pair in (0, 1, 1):
            ops.set_gemm_tuning(0, -1, 0, pair)
            x = x0.clone()
            ops.linear(a, w, b, resid=x, out=x)                       # in place: the reduce-add epilogue
            h = ops.linear(a, w, b, act=ops.ACT_GELU, out_dtype=dt)     # 16-bit TMA-store epilogue (16 warps when K <= 512 on the pair kernel)
            outs.setdefault(pair, []).append((x, h))
    finally:
        ops.set_gemm_tuning()
    (xs, hs), (xp, hp), (xp2, hp2) = outs[0][0], outs[1][0], outs[1][1]
    assert torch.equal(xp, xp2) and torch.equal(hp, hp2), "pair kernel not reproducible"
    assert torch.equal(hs, hp), "GELU output differs between the single-CTA and the CTA-pair kernel"
    assert torch.equal(xs, xp), "in-place residual differs between the single-CTA and the CTA-pair kernel"
    for lo in (0, M - 4000):
        acc = a[lo:lo + 4000].float() @ w.float().T + b
        assert rel(xp[lo:lo + 4000], x0[lo:lo + 4000] + acc) < 1e-5
        assert rel(hp[lo:lo + 4000], torch.nn.functional.gelu(acc)) < (1e-2 if dt == torch.bfloat16 else 2e-3)


@pytest.mark.parametrize("C,M", [(128, 128 * 148 * 2 + 77), (256, 128 * 150 + 5), (128, 64)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_mlp_fused_matches_unfused(ops, C, M, dt):
    """Fused fc1+GELU+fc2+residual against the two-GEMM path and against fp32 torch math."""
    g = torch.Generator(device="cuda").manual_seed(C + M)
    xn = torch.randn(M, C, device="cuda", generator=g).to(dt)
    w1 = (torch.randn(4 * C, C, device="cuda", generator=g) * 0.06).to(dt)
    b1 = torch.randn(4 * C, device="cuda", generator=g) * 0.1
    w2 = (torch.randn(C, 4 * C, device="cuda", generator=g) * 0.04).to(dt)
    b2 = torch.randn(C, device="cuda", generator=g) * 0.1
    x = torch.randn(M, C, device="cuda", generator=g)
    fused = ops.mlp_fused(xn, w1, b1, w2, b2, x.clone())
    hid = ops.linear(xn, w1, b1, act=ops.ACT_GELU, out_dtype=dt)
    unfused = ops.linear(hid, w2, b2, resid=x.clone(), out_dtype=torch.float32)
    href = torch.nn.functional.gelu(xn.float() @ w1.float().T + b1)
    ref = x + href @ w2.float().T + b2
    tol = 6e-3 if dt == torch.bfloat16 else 1e-3
    assert rel(fused - x, ref - x) < tol and rel(unfused - x, ref - x) < tol
    assert rel(fused, unfused) < 1e-4


@pytest.mark.parametrize("M,N", [(1000, 256), (1000, 512), (777, 768), (128 * 300 + 40, 512), (128 * 150 + 5, 1024), (64, 256)])
def test_linear_fp32_residual_tma_epilogue(ops, M, N):
    """fp32 output on plain rows with N >= 256 takes the TMA-staged epilogue (residual chunks in, results out by TMA): ragged M,
    in place / out of place / no residual, single-CTA and CTA-pair kernels, against fp32 torch math."""
    g = torch.Generator(device="cuda").manual_seed(M + N)
    K = 256
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    x = torch.randn(M, N, device="cuda", generator=g)
    y = a.float() @ w.float().T + b
    out = ops.linear(a, w, b, out_dtype=torch.float32)                       # no residual
    assert rel(out, y) < 1e-5
    out2 = ops.linear(a, w, b, resid=x, out_dtype=torch.float32)             # out of place
    assert rel(out2, x + y) < 1e-5
    xi = x.clone()
    ops.linear(a, w, b, resid=xi, out=xi)                                    # in place
    assert torch.equal(xi, out2)
    big = torch.full((M + 2, N + 32), 7.0, device="cuda")                    # pitched output / residual views, untouched borders
    view = big[1:M + 1, :N]
    view.copy_(x)
    ops.linear(a, w, b, resid=view, out=view)
    assert torch.equal(view, out2)
    assert (big[0] == 7).all() and (big[M + 1] == 7).all() and (big[:, N:] == 7).all()


def test_linear_epilogues(ops):
    g = torch.Generator(device="cuda").manual_seed(5)
    B, H, W, C = 3, 14, 14, 256
    M = B * H * W
    a = torch.randn(M, C, device="cuda", generator=g).bfloat16()
    w = (torch.randn(4 * C, C, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(4 * C, device="cuda", generator=g)
    # GELU -> bf16
    out = ops.linear(a, w, b, act=ops.ACT_GELU, out_dtype=torch.bfloat16)
    ref = torch.nn.functional.gelu(a.float() @ w.float().T + b)
    assert rel(out, ref) < 5e-3
    # residual, in place, with window scatter
    w2 = (torch.randn(C, C, device="cuda", generator=g) * 0.05).bfloat16()
    b2 = torch.randn(C, device="cuda", generator=g)
    x = torch.randn(M, C, device="cuda", generator=g)
    for shift in (0, 3):
        idx = ops.window_index_map(H, W, 7, shift).long()
        full = (torch.arange(B, device="cuda")[:, None] * (H * W) + idx[None]).reshape(-1)
        y = a.float() @ w2.float().T + b2
        ref = x.clone()
        ref[full] += y
        xi = x.clone()
        ops.linear(a, w2, b2, resid=xi, out=xi, scatter=(H, W, 7, shift))
        assert rel(xi, ref) < 1e-5


@pytest.mark.parametrize("C", [96, 128, 192, 256, 384, 512])
@pytest.mark.parametrize("rows", [1, 7, 1001, 4099])
def test_layernorm_ragged_row_counts(ops, C, rows):
    """Identity mode with row counts that are not multiples of the rows a warp handles per pass (narrow-row kernel: 4-16 rows per
    warp; wide rows: 1-2), fp32 and 16-bit outputs, and the window mode of the same widths on whole images."""
    g = torch.Generator(device="cuda").manual_seed(C + rows)
    x = torch.randn(rows, C, device="cuda", generator=g) * 3 - 1
    gamma = torch.randn(C, device="cuda", generator=g)
    beta = torch.randn(C, device="cuda", generator=g)
    ref = torch.nn.functional.layer_norm(x, (C,), gamma, beta, 1e-5)
    assert rel(ops.layernorm(x, gamma, beta, 1e-5), ref) < 2e-6
    assert rel(ops.layernorm(x, gamma, beta, 1e-5, out_dtype=torch.bfloat16), ref) < 4e-3
    if rows == 1001:
        B, H = 3, 14
        xi = torch.randn(B * H * H, C, device="cuda", generator=g)
        idx = ops.window_index_map(H, H, 7, 3).long()
        want = torch.nn.functional.layer_norm(xi.view(B, H * H, C)[:, idx].reshape(-1, C), (C,), gamma, beta, 1e-5)
        assert rel(ops.layernorm(xi, gamma, beta, 1e-5, mode=ops.LN_WINDOW, grid=(H, H), ws=7, shift=3), want) < 2e-6


@pytest.mark.parametrize("C,mode", [(96, 0), (128, 1), (512, 1), (1024, 0), (256, 2), (512, 2)])
def test_layernorm(ops, C, mode):
    g = torch.Generator(device="cuda").manual_seed(C + mode)
    B, H, W = 2, 14, 14
    x = torch.randn(B * H * W, C, device="cuda", generator=g) * 2 + 0.5
    width = 4 * C if mode == 2 else C
    gamma = torch.randn(width, device="cuda", generator=g)
    beta = torch.randn(width, device="cuda", generator=g)
    for shift in ((0, 3) if mode == 1 else (0,)):
        if mode == 0:
            src = x
        elif mode == 1:
            idx = ops.window_index_map(H, W, 7, shift).long()
            src = x.view(B, H * W, C)[:, idx].reshape(-1, C)
        else:
            idx = ops.merge_index_map(H, W).long()  # [No,4]
            src = x.view(B, H * W, C)[:, idx].reshape(B * idx.shape[0], 4 * C)
        ref = torch.nn.functional.layer_norm(src, (width,), gamma, beta, 1e-5)
        out = ops.layernorm(x, gamma, beta, 1e-5, mode=mode, grid=(H, W), ws=7, shift=shift)
        assert rel(out, ref) < 1e-5
        outb = ops.layernorm(x, gamma, beta, 1e-5, mode=mode, grid=(H, W), ws=7, shift=shift, out_dtype=torch.bfloat16)
        assert rel(outb, ref) < 4e-3


def test_patch_im2col(ops):
    g = torch.Generator(device="cuda").manual_seed(3)
    img = torch.rand(2, 3, 224, 224, device="cuda", generator=g)
    out = ops.patch_im2col(img, out_dtype=torch.float32)
    mean = torch.tensor([0.485, 0.456, 0.406], device="cuda")[None, :, None, None]
    std = torch.tensor([0.229, 0.224, 0.225], device="cuda")[None, :, None, None]
    n = (img - mean) / std
    ref = torch.nn.functional.unfold(n, kernel_size=4, stride=4).transpose(1, 2).reshape(-1, 48)
    assert rel(out, ref) < 1e-6


@pytest.mark.parametrize("H,heads,shift", [(14, 16, 0), (14, 16, 3), (28, 8, 3), (56, 4, 3), (7, 32, 0), (7, 3, 0)])
@pytest.mark.parametrize("dtype", [torch.float32])
def test_window_attention(ops, H, heads, shift, dtype):
    """Exact fp32 window attention of the validation mode (16-bit operands: test_swin_attn_core / test_swin_attn_fused)."""
    g = torch.Generator(device="cuda").manual_seed(H * heads + shift)
    B, W, C, ws, L = 2, H, heads * 32, 7, 49
    nW = (H // ws) * (W // ws)
    qkv = torch.randn(B * H * W, 3 * C, device="cuda", generator=g).to(dtype)
    table = torch.randn(169, heads, device="cuda", generator=g)
    bias = ops.expand_rel_bias(table, ws)
    out = ops.window_attention(qkv, bias, B, H, W, heads, ws, shift)
    q, k, v = qkv.float().view(B * nW, L, 3, heads, 32).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / math.sqrt(32) + bias[None]
    if shift:
        s = s.view(B, nW, heads, L, L) + ops.shift_mask(H, W, ws, shift)[None, :, None]
        s = s.view(B * nW, heads, L, L)
    ref = (s.softmax(-1) @ v).transpose(1, 2).reshape(B * H * W, C)
    assert rel(out, ref) < 1e-5


@pytest.mark.parametrize("B,H,heads,shift", [(2, 14, 16, 0), (2, 14, 16, 3), (3, 28, 8, 3), (2, 56, 4, 3), (5, 7, 32, 0), (3, 7, 3, 0),
                                            (2, 14, 6, 3), (3, 14, 12, 3), (1, 7, 24, 0), (160, 14, 16, 3)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_swin_attn_core(ops, B, H, heads, shift, dtype):
    """csvit_swin_attn_core (tcgen05 attention core on TMA-loaded q/k/v tiles) vs fp32 torch math on the same 16-bit qkv: every
    Swin-B and Swin-T width (even and odd head counts), masked and unmasked windows, odd window counts (half-empty last tile),
    more tiles than SMs; window- and token-ordered output; logits scaled in the kernel or pre-scaled q."""
    g = torch.Generator(device="cuda").manual_seed(H * heads + shift + B)
    W, C, ws, L = H, heads * 32, 7, 49
    nW = (H // ws) * (W // ws)
    qkv = torch.randn(B * H * W, 3 * C, device="cuda", generator=g).to(dtype)
    table = torch.randn(169, heads, device="cuda", generator=g)
    bias = ops.expand_rel_bias(table, ws)
    bias_l2 = ops.pack_rel_bias_log2(table, ops.rel_pos_index(ws).long())
    q, k, v = qkv.float().view(B * nW, L, 3, heads, 32).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / math.sqrt(32) + bias[None]
    if shift:
        s = s.view(B, nW, heads, L, L) + ops.shift_mask(H, W, ws, shift)[None, :, None]
        s = s.view(B * nW, heads, L, L)
    ref = (s.softmax(-1) @ v).transpose(1, 2).reshape(B * H * W, C)
    tol = 6e-3 if dtype == torch.bfloat16 else 1e-3       # the 16-bit rounding of P and of the output
    out = ops.swin_attn_core(qkv, bias_l2, B, H, W, heads, ws, shift)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert rel(out, ref) < tol, rel(out, ref)
    assert torch.equal(out, ops.swin_attn_core(qkv, bias_l2, B, H, W, heads, ws, shift)), "must be deterministic"
    # token order = the window-ordered rows scattered by the window index map (window_reverse + un-shift), bit for bit
    out_tok = ops.swin_attn_core(qkv, bias_l2, B, H, W, heads, ws, shift, token_order=True)
    idx = ops.window_index_map(H, W, ws, shift).long()
    want_tok = torch.empty_like(out).view(B, H * W, C)
    want_tok[:, idx] = out.view(B, H * W, C)
    assert torch.equal(out_tok.view(B, H * W, C), want_tok)
    # q pre-scaled by log2(e)/sqrt(32) (what the inference path folds into the Q/K/V GEMM): same result up to the rounding of q
    qs = qkv.clone()
    qs[:, :C] = (qkv[:, :C].float() * (math.log2(math.e) / math.sqrt(32))).to(dtype)
    out_ps = ops.swin_attn_core(qs, bias_l2, B, H, W, heads, ws, shift, q_prescaled=True)
    assert rel(out_ps, ref) < 2 * tol


def fused_attention_case(ops, B, H, heads, shift, dtype, seed, zero_bias=False):
    """Inputs + fp32 torch reference of the attention half up to the token-ordered context (HF:404-459, 604-636)."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    W, C, ws, L = H, heads * 32, 7, 49
    N = H * W
    nW = (H // ws) * (W // ws)
    x = torch.randn(B * N, C, device="cuda", generator=g) * 1.5 + 0.3
    gamma = 1.0 + 0.2 * torch.randn(C, device="cuda", generator=g)
    beta = 0.1 * torch.randn(C, device="cuda", generator=g)
    wq, wk, wv = (torch.randn(C, C, device="cuda", generator=g) * C ** -0.5 for _ in range(3))
    bq, bk, bv = (0.2 * torch.randn(C, device="cuda", generator=g) for _ in range(3))
    table = torch.randn(169, heads, device="cuda", generator=g) * (0.0 if zero_bias else 1.0)
    rel_index = ops.rel_pos_index(ws).long()
    xn = torch.nn.functional.layer_norm(x, (C,), gamma, beta, 1e-5)
    idx = ops.window_index_map(H, W, ws, shift).long()
    xw = xn.view(B, N, C)[:, idx].reshape(B * nW, L, C)      # exact fp32 math: the tolerance below is the 16-bit operand rounding
    q = (xw @ wq.T + bq).view(B * nW, L, heads, 32).transpose(1, 2)
    k = (xw @ wk.T + bk).view(B * nW, L, heads, 32).transpose(1, 2)
    v = (xw @ wv.T + bv).view(B * nW, L, heads, 32).transpose(1, 2)
    s = q @ k.transpose(-1, -2) / math.sqrt(32) + ops.expand_rel_bias(table, ws)[None]
    if shift:
        s = (s.view(B, nW, heads, L, L) + ops.shift_mask(H, W, ws, shift)[None, :, None]).view(B * nW, heads, L, L)
    ref_win = (s.softmax(-1) @ v).transpose(1, 2).reshape(B, N, C)
    ref = torch.empty_like(ref_win)
    ref[:, idx] = ref_win
    packed = ops.pack_attn_fused(wq, wk, wv, bq, bk, bv, table, rel_index, dtype, gamma, beta)
    return x, gamma, beta, packed, ref.reshape(B * N, C)


@pytest.mark.parametrize("B,H,heads,shift", [(2, 56, 4, 0), (2, 56, 4, 3), (3, 28, 8, 0), (3, 28, 8, 3), (3, 14, 4, 3), (3, 7, 4, 0),
                                            (5, 7, 8, 0), (40, 28, 4, 3)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_swin_attn_fused(ops, B, H, heads, shift, dtype):
    """csvit_swin_attn_fused (LN + shift/partition + QKV + window attention + reverse on tcgen05) vs fp32 torch math; odd window
    counts (a half-empty last tile), masked and unmasked windows, both operand formats, more tiles than SMs."""
    x, gamma, beta, (wqkv_h, bqkv_h, bias_op), ref = fused_attention_case(ops, B, H, heads, shift, dtype, seed=B * 131 + H * 7 + heads + shift)
    out = ops.swin_attn_fused(x, 1e-5, wqkv_h, bqkv_h, bias_op, B, H, H, heads, 7, shift)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    tol = 1.2e-2 if dtype == torch.bfloat16 else 2.5e-3
    assert rel(out, ref) < tol
    out2 = ops.swin_attn_fused(x, 1e-5, wqkv_h, bqkv_h, bias_op, B, H, H, heads, 7, shift)
    assert torch.equal(out, out2), "fused attention must be deterministic"


@pytest.mark.parametrize("Lq,S", [(52, 52), (3, 49), (3, 3), (1, 8)])
def test_dense_attention(ops, Lq, S):
    g = torch.Generator(device="cuda").manual_seed(Lq + S)
    n, heads = 5, 24
    D = heads * 32
    q = torch.randn(n * Lq, D, device="cuda", generator=g)
    kv = torch.randn(n * S, 2 * D, device="cuda", generator=g)
    k, v = kv[:, :D], kv[:, D:]
    scale = math.sqrt(32.0) * 0.1
    out = ops.attention(q, k, v, n, Lq, S, heads, scale)
    qh = q.view(n, Lq, heads, 32).transpose(1, 2)
    kh = k.reshape(n, S, heads, 32).transpose(1, 2)
    vh = v.reshape(n, S, heads, 32).transpose(1, 2)
    ref = ((qh @ kh.transpose(-1, -2) * scale).softmax(-1) @ vh).transpose(1, 2).reshape(n * Lq, D)
    assert rel(out, ref) < 1e-5
