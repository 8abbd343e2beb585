"""GPU parity for the SwinV2 row, through the C ABI: the two new kernels against plain fp32 torch math, the
backbone against the goldens written by the live HF ``Swinv2Model``, and ``Poser`` on a swinv2 backbone against the composed
oracle (oracle/swinv2_restated.py pinned to HF + oracle/head_restated.py pinned to the reference)."""
import math
import os

import pytest
import torch

from test_swinv2_oracle import V2_CASES, v2_case

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def ops():
    from cs_vit import ops as o
    return o


def _bias_table(ws, heads, g):
    """[heads, (2ws-1)^2] in (0, 16), like 16 sigmoid(cpb_mlp)."""
    return (16 * torch.sigmoid(2 * torch.randn(heads, (2 * ws - 1) ** 2, device="cuda", generator=g))).contiguous()


@pytest.mark.parametrize("H,ws,heads,shift", [(32, 16, 4, 0), (32, 16, 4, 8), (64, 16, 1, 8), (16, 16, 16, 0), (8, 8, 32, 0),
                                              (16, 8, 3, 4), (32, 8, 2, 4)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_swinv2_window_attention(ops, H, ws, heads, shift, dtype):
    g = torch.Generator(device="cuda").manual_seed(H * heads + shift + ws)
    B, W, C, L = 2, H, heads * 32, ws * ws
    nW = (H // ws) * (W // ws)
    qkv = torch.randn(B * H * W, 3 * C, device="cuda", generator=g).to(dtype)
    tab = _bias_table(ws, heads, g)
    scale = (math.log(10.0) + 0.5 * torch.randn(heads, device="cuda", generator=g)).clamp(max=math.log(100.0)).exp().contiguous()
    out = ops.swinv2_window_attention(qkv, tab, scale, B, H, W, heads, ws, shift)
    q, k, v = qkv.float().view(B * nW, L, 3, heads, 32).permute(2, 0, 3, 1, 4)
    s = torch.nn.functional.normalize(q, dim=-1) @ torch.nn.functional.normalize(k, dim=-1).transpose(-1, -2)
    idx = ops.rel_pos_index(ws).long().reshape(-1)
    s = s * scale.view(1, heads, 1, 1) + tab[:, idx].view(1, heads, L, L)
    if shift:
        m = ops.shift_mask(H, W, ws, shift)
        s = s.view(B, nW, heads, L, L) + 2 * m[None, :, None]          # HF adds the mask twice (V2:466-474)
        s = s.view(B * nW, heads, L, L)
    ref = (s.softmax(-1) @ v).transpose(1, 2).reshape(B * H * W, C)
    tol = {torch.bfloat16: 1.5e-2, torch.float16: 2e-3, torch.float32: 1e-5}[dtype]
    assert rel(out, ref) < tol
    # token-ordered output = the window-ordered rows scattered by the window index map, bit for bit
    out_tok = ops.swinv2_window_attention(qkv, tab, scale, B, H, W, heads, ws, shift, token_order=True)
    idx = ops.window_index_map(H, W, ws, shift).long()
    want_tok = torch.empty_like(out).view(B, H * W, C)
    want_tok[:, idx] = out.view(B, H * W, C)
    assert torch.equal(out_tok.view(B, H * W, C), want_tok)
    if shift:   # mask_repeat is honoured (a single add changes the result only where -100 does not already saturate)
        out1 = ops.swinv2_window_attention(qkv, tab, scale, B, H, W, heads, ws, shift, mask_repeat=0)
        assert rel(out1, ref) > tol


@pytest.mark.parametrize("H,heads,shift,B", [(32, 4, 0, 2), (32, 4, 8, 2), (64, 1, 8, 1), (16, 16, 0, 3), (16, 3, 0, 2), (48, 2, 8, 1),
                                             (32, 8, 8, 5), (32, 8, 8, 13)])      # the last case is large enough for the CTA-pair GEMM
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_swinv2_qkv_and_tcgen05_attention(ops, H, heads, shift, B, dtype):
    """csvit_swinv2_qkv (cosine normalisation in the GEMM epilogue) + csvit_swinv2_attn_tc (tcgen05, P as a TMEM operand) against
    fp32 torch math of V2:450-487 on the same 16-bit-rounded activations and weights."""
    g = torch.Generator(device="cuda").manual_seed(7 * H + heads + shift)
    ws, W, C, L = 16, H, heads * 32, 256
    nW = (H // ws) * (W // ws)
    xw = torch.randn(B * H * W, C, device="cuda", generator=g).to(dtype)
    wqkv = (torch.randn(3 * C, C, device="cuda", generator=g) / math.sqrt(C)).to(dtype)
    bqkv = 0.3 * torch.randn(3 * C, device="cuda", generator=g)
    bqkv[C:2 * C] = 0                                                   # V2: the key projection has no bias
    tab = _bias_table(ws, heads, g)
    scale = (math.log(10.0) + 0.8 * torch.randn(heads, device="cuda", generator=g)).clamp(max=math.log(100.0)).exp().contiguous()
    qkv_n = ops.swinv2_qkv(xw, wqkv, bqkv, (scale * 1.4426950408889634).contiguous())
    raw = xw.float() @ wqkv.float().T + bqkv
    q, k, v = raw.view(B * nW, L, 3, heads, 32).permute(2, 0, 3, 1, 4)
    qn, kn = torch.nn.functional.normalize(q, dim=-1), torch.nn.functional.normalize(k, dim=-1)
    # the epilogue: q_hat * log2(e) * scale | k_hat | v, rounded once
    want_n = torch.stack([qn * (scale * 1.4426950408889634).view(1, heads, 1, 1), kn, v], 0).permute(1, 3, 0, 2, 4).reshape(B * H * W, 3 * C)
    assert rel(qkv_n, want_n) < (4e-3 if dtype == torch.bfloat16 else 5e-4)
    out = ops.swinv2_attn_tc(qkv_n, ops.swinv2_bias_log2(tab), B, H, W, heads, shift)
    idx = ops.rel_pos_index(ws).long().reshape(-1)
    s = (qn @ kn.transpose(-1, -2)) * scale.view(1, heads, 1, 1) + tab[:, idx].view(1, heads, L, L)
    if shift:
        m = ops.shift_mask(H, W, ws, shift)
        s = (s.view(B, nW, heads, L, L) + 2 * m[None, :, None]).view(B * nW, heads, L, L)
    ref = (s.softmax(-1) @ v).transpose(1, 2).reshape(B * H * W, C)
    tol = 2e-2 if dtype == torch.bfloat16 else 3e-3
    assert rel(out, ref) < tol, rel(out, ref)
    out_tok = ops.swinv2_attn_tc(qkv_n, ops.swinv2_bias_log2(tab), B, H, W, heads, shift, token_order=True)
    widx = ops.window_index_map(H, W, ws, shift).long()
    want_tok = torch.empty_like(out).view(B, H * W, C)
    want_tok[:, widx] = out.view(B, H * W, C)
    assert torch.equal(out_tok.view(B, H * W, C), want_tok)
    if shift:
        out0 = ops.swinv2_attn_tc(qkv_n, ops.swinv2_bias_log2(tab), B, H, W, heads, shift, mask_repeat=0)
        assert rel(out0, ref) > tol


def test_swinv2_window_attention_rejects_unbuilt_windows(ops):
    qkv = torch.zeros(36, 96, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="not built"):
        ops.swinv2_window_attention(qkv, torch.zeros(1, 121, device="cuda"), torch.ones(1, device="cuda"), 1, 6, 6, 1, 6, 0)


@pytest.mark.parametrize("C", [32, 96, 128, 256, 512, 1024])
@pytest.mark.parametrize("mode", ["none", "identity", "window", "merge"])
@pytest.mark.parametrize("cdt", [torch.bfloat16, torch.float32])
def test_layernorm_post(ops, C, mode, cdt):
    g = torch.Generator(device="cuda").manual_seed(C)
    B, H, ws, shift = 3, 16, 8, 4
    rows = B * H * H
    y = torch.randn(rows, C, device="cuda", generator=g) * 2 + 0.5
    x = torch.randn(rows, C, device="cuda", generator=g)
    gamma = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
    beta = 0.1 * torch.randn(C, device="cuda", generator=g)
    want = x + torch.nn.functional.layer_norm(y, (C,), gamma, beta, 1e-5)
    cm = {"none": ops.COPY_NONE, "identity": ops.COPY_IDENTITY, "window": ops.COPY_WINDOW, "merge": ops.COPY_MERGE2X2}[mode]
    out, copy = ops.layernorm_post(y, x, gamma, beta, 1e-5, copy_mode=cm, copy_dtype=cdt, geom=(H, H, ws, shift))
    assert rel(out, want) < 2e-6
    if mode == "none":
        assert copy is None
        return
    if mode == "identity":
        ref = out
    elif mode == "window":
        idx = ops.window_index_map(H, H, ws, shift).long()
        ref = out.view(B, H * H, C)[:, idx].reshape(rows, C)
    else:
        g4 = out.view(B, H, H, C)
        ref = torch.cat([g4[:, 0::2, 0::2], g4[:, 1::2, 0::2], g4[:, 0::2, 1::2], g4[:, 1::2, 1::2]], -1).reshape(rows // 4, 4 * C)
    assert copy.dtype == cdt and copy.shape == ref.shape
    assert torch.equal(copy, ref.to(cdt))          # the copy is the rounded fp32 output, bit for bit
    # in place on the residual, and the no-residual form (patch embedding / patch merging norms)
    x2 = x.clone()
    out2, _ = ops.layernorm_post(y, x2, gamma, beta, 1e-5, out=x2)
    assert torch.equal(out2, out) and out2.data_ptr() == x2.data_ptr()
    out3, _ = ops.layernorm_post(y, None, gamma, beta, 1e-5)
    assert rel(out3, want - x) < 2e-6


@pytest.mark.parametrize("S", [67, 128, 100])
def test_dense_attention_up_to_128_keys(ops, S):
    """The "encoder" head on a 256^2 SwinV2 backbone attends over 3 + 64 = 67 tokens."""
    g = torch.Generator(device="cuda").manual_seed(S)
    n, heads, Lq = 3, 8, S
    D = heads * 32
    q = torch.randn(n * Lq, D, device="cuda", generator=g)
    k = torch.randn(n * S, D, device="cuda", generator=g)
    v = torch.randn(n * S, D, device="cuda", generator=g)
    scale = math.sqrt(32.0) * 0.1
    out = ops.attention(q, k, v, n, Lq, S, heads, scale)
    qh, kh, vh = (t.view(n, -1, heads, 32).transpose(1, 2) for t in (q, k, v))
    ref = ((qh @ kh.transpose(-1, -2) * scale).softmax(-1) @ vh).transpose(1, 2).reshape(n * Lq, D)
    assert rel(out, ref) < 1e-5


def _backbone(case, precision, tmp):
    from cs_vit.net.swinv2_b200 import load_backbone
    from cs_vit.synthetic import make_random_backbone_dir
    d = make_random_backbone_dir(os.path.join(tmp, f"{case['variant']}_w{case['window']}"), case["variant"], seed=case["weight_seed"],
                                 image_size=case["image_size"], window_size=case["window"])
    return load_backbone(d, precision=precision).cuda()


@pytest.mark.parametrize("name", sorted(V2_CASES))
@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16"])
def test_swinv2_backbone_matches_hf_goldens(name, precision, tmp_path):
    sd, px, gold, case = v2_case(name)
    model = _backbone(case, precision, str(tmp_path))
    with torch.no_grad():
        out, stages = model.forward_features(px.cuda(), normalize=False, return_stages=True)
    # bars of BASELINE.json:north_star - 1e-4 (fp32), 1e-2 (bf16); fp16 operands land in between
    tol = {"fp32": 1e-4, "fp16": 2e-3, "bf16": 1e-2}[precision]
    err = rel(out.cpu(), torch.from_numpy(gold["last_hidden_state"]))
    serr = [rel(t[:, ::case["stage_token_stride"]].cpu(), torch.from_numpy(gold[f"stage{s}"])) for s, t in enumerate(stages)]
    print(f"{name} {precision}: last {err:.2e} stages {['%.1e' % e for e in serr]}")
    assert err < tol and max(serr) < tol
    with torch.no_grad():
        assert torch.equal(model(px.cuda()).last_hidden_state, out)    # the HF seam: forward() on normalised pixels


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("fp16", 3e-3)])
def test_swinv2_backward_matches_hf_gradient_golden(precision, tol, tmp_path):
    """The differentiable SwinV2 path (finetune step with the shipped backbone family): features and the gradient of <features, R>
    w.r.t. every Swinv2Model parameter against HF's autograd (tests/golden/train_swinv2_xs_w16_linear.npz)."""
    import numpy as np
    from test_swinv2_oracle import V2_GRAD_CASES
    from oracle.make_train_goldens import projections
    name = sorted(V2_GRAD_CASES)[0]
    sd, px, gold, case = v2_case(name)
    model = _backbone(case, precision, str(tmp_path))
    model.config.drop_path_rate = 0.0
    model.train()
    for p_ in model.parameters():
        p_.requires_grad_(True)
    feats = model.forward_features(px.cuda(), normalize=False)
    assert feats.requires_grad
    assert rel(feats.detach().cpu(), torch.from_numpy(gold["features"])) < tol
    R = torch.randn(feats.shape, generator=torch.Generator().manual_seed(case["projection_seed"])).cuda()
    (feats * R).sum().backward()
    torch.cuda.synchronize()
    params = dict(model.named_parameters())
    names = [str(n) for n in gold["param_names"]]
    assert set(names) == set(params)
    gnorm = float(np.sqrt((gold["grad_norm"] ** 2).sum()))
    worst = 0.0
    for i, n in enumerate(names):
        g = params[n].grad
        assert g is not None, n
        assert abs(g.double().norm().item() - gold["grad_norm"][i]) < 10 * tol * gold["grad_norm"][i] + 1e-5 * gnorm, n
        assert np.allclose(projections(g.cpu(), n), gold["grad_proj"][i], rtol=0, atol=10 * tol * gold["grad_norm"][i] * 4 + 1e-5 * gnorm), n
        if "grad/" + n in gold and gold["grad_norm"][i] > 1e-4 * gnorm:
            worst = max(worst, rel(g.cpu(), torch.from_numpy(gold["grad/" + n])))
    assert worst < 10 * tol, worst
    # eval / no_grad still takes the tuned inference path, and stochastic depth draws are applied in train mode
    model.config.drop_path_rate = 0.1
    out_train = model.forward_features(px.cuda(), normalize=False)
    assert torch.isfinite(out_train).all()
    model.eval()
    with torch.no_grad():
        out_eval = model.forward_features(px.cuda(), normalize=False)
    assert rel(out_eval.cpu(), torch.from_numpy(gold["features"])) < max(tol, 2e-3)


@pytest.mark.parametrize("layer_type,decorate", [("encoder", "patch"), ("decoder", "query")])
def test_poser_on_swinv2_backbone_matches_composed_oracle(layer_type, decorate, tmp_path):
    """predict_batch at 256^2 on a swinv2 w16 backbone (the shipped configuration family) in fp32 mode against the oracle:
    SwinV2 restatement (pinned to HF) feeding the head restatement (pinned to the reference Poser)."""
    from cs_vit.net import Poser
    from cs_vit.synthetic import SWINV2_VARIANTS, make_inputs, make_random_backbone_dir, randomize_head_
    from cs_vit.utils.mano_standin import SyntheticMANO
    from oracle import head_restated as head
    from oracle import swinv2_restated as v2

    variant = "swinv2_t"
    d = make_random_backbone_dir(str(tmp_path / "v2t"), variant, seed=0, image_size=256, window_size=16)
    torch.manual_seed(0)
    model = Poser(d, image_size=256, mano_layer=SyntheticMANO(), spatial_layer_type=layer_type, persp_decorate=decorate, precision="fp32")
    randomize_head_(model, seed=1)
    model.phase(Poser.TrainingPhase.SPATIAL)
    model.eval()
    inputs = make_inputs(2, 1, 256, seed=5)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    _, depths, heads = SWINV2_VARIANTS[variant]
    opt = head.HeadOptions(num_heads=heads[-1], depths=depths, swin_heads=heads, spatial_layer_type=layer_type,
                           persp_decorate=decorate, phase="spatial")
    bsd = {k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}
    mean = torch.tensor(head.IMAGENET_MEAN)[None, :, None, None]
    std = torch.tensor(head.IMAGENET_STD)[None, :, None, None]

    def features(flat):
        return v2.swinv2_forward((flat - mean) / std, bsd, depths, heads, window=16)

    with torch.no_grad():
        want = head.predict_batch(inputs, sd, opt, SyntheticMANO(), execute_all=False, features_fn=features)
    model = model.cuda()
    with torch.no_grad():
        got = model.predict_batch(*(inputs[k].cuda() for k in ("patches", "square_bboxes", "timestamp", "focal", "princpt")))
    for k in ("joint_cam", "verts_cam", "shape", "root_transl"):
        assert got[k].shape == want[k].shape
        assert rel(got[k].cpu(), want[k]) < 1e-4, (k, rel(got[k].cpu(), want[k]))


def test_finetune_step_on_swinv2_backbone(tmp_path):
    """The SPATIAL finetune phase on the backbone family of every shipped reference configuration, with HF's default
    drop_path_rate 0.1 left in config.json (this used to raise): loss falls over a few AdamW steps, every
    trainable backbone parameter receives a finite gradient."""
    import json
    from cs_vit.net import Poser
    from cs_vit.synthetic import make_inputs, make_random_backbone_dir, randomize_head_
    from cs_vit.train import finetune_step
    from cs_vit.utils.mano_standin import SyntheticMANO
    d = make_random_backbone_dir(str(tmp_path / "v2xs"), "swinv2_xs", seed=0, image_size=256, window_size=16)
    with open(os.path.join(d, "config.json")) as f:
        cfg = json.load(f)
    cfg["drop_path_rate"] = 0.1
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(cfg, f)
    torch.manual_seed(0)
    model = Poser(d, image_size=256, mano_layer=SyntheticMANO(), spatial_layer_type="encoder", persp_decorate="patch", precision="fp16")
    randomize_head_(model, seed=1)
    model.phase(Poser.TrainingPhase.SPATIAL)
    model = model.cuda()
    batch = {k: v.cuda() for k, v in make_inputs(4, 1, 256, seed=3, labels=True).items()}
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=2e-5)
    losses = [finetune_step(model, batch, opt).item() for _ in range(4)]
    assert all(math.isfinite(v) for v in losses) and losses[-1] < losses[0], losses
    grads = [p.grad for n, p in model.backbone.named_parameters()]
    assert all(g is not None and torch.isfinite(g).all() for g in grads)
