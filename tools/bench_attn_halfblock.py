"""BASELINE configs[4]: attention half-block microbench  x -> x + proj(WMSA(shift(LN x)))  across the Swin-B stages,
shifted and unshifted, batch 256: this repo's kernels (LN-gather, QKV GEMM, window attention, out-proj + scatter +
residual) vs the reference PyTorch path (HF SwinLayer's attention half) on the same GPU (fp32 eager, as the reference
runs it) and on the host CPU (small sample).  Prints one JSON line per case; FLOPs = 8NC^2 + 196NC per image."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from transformers.models.swin import modeling_swin as hf
from cs_vit import ops

B = int(os.environ.get("B", "256")); PREC = os.environ.get("PREC", "bf16")
dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[PREC]
peak = 1398.1
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
except Exception:
    pass

def hf_attention_half(layer, x, hw):
    """HF:swin/modeling_swin.py:604-646 (everything before layernorm_after)."""
    H = W = hw; Bn, _, C = x.shape
    h = layer.layernorm_before(x).view(Bn, H, W, C)
    s = layer.shift_size
    if s > 0: h = torch.roll(h, shifts=(-s, -s), dims=(1, 2))
    win = hf.window_partition(h, layer.window_size).view(-1, layer.window_size ** 2, C)
    mask = layer.get_attn_mask(H, W, dtype=x.dtype, device=x.device)
    a = layer.attention(win, mask)[0].view(-1, layer.window_size, layer.window_size, C)
    h = hf.window_reverse(a, layer.window_size, H, W)
    if s > 0: h = torch.roll(h, shifts=(s, s), dims=(1, 2))
    return x + h.view(Bn, H * W, C)

def timeit(fn, it):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it

for stage, (hw, C, heads) in enumerate([(56, 128, 4), (28, 256, 8), (14, 512, 16), (7, 1024, 32)]):
    for shift in ((0, 3) if hw > 7 else (0,)):
        torch.manual_seed(stage * 10 + shift)
        cfg = hf.SwinConfig(window_size=7)
        layer = hf.SwinLayer(cfg, dim=C, input_resolution=(hw, hw), num_heads=heads, shift_size=shift).eval()
        with torch.no_grad():
            layer.attention.self.relative_position_bias_table.normal_(0, 0.02)
        N = hw * hw
        x = torch.randn(B, N, C)
        flops_img = 8 * N * C * C + 196 * N * C
        # --- CPU reference, small sample
        xs = x[:8]
        with torch.inference_mode():
            hf_attention_half(layer, xs, hw); t0 = time.perf_counter(); ref_cpu = hf_attention_half(layer, xs, hw); cpu_ms = (time.perf_counter() - t0) * 1e3
        # --- GPU: reference eager fp32
        lg = layer.cuda(); xg = x.cuda()
        with torch.inference_mode():
            ref_ms = timeit(lambda: hf_attention_half(lg, xg, hw), 5)
            ref = hf_attention_half(lg, xg[:8], hw)
        # --- GPU: this repo
        sa = lg.attention.self
        wo = lg.attention.output.dense.weight.detach().to(dt).contiguous(); bo = lg.attention.output.dense.bias.detach().float()
        g, b_ = lg.layernorm_before.weight.detach().float(), lg.layernorm_before.bias.detach().float()
        rel_index = ops.rel_pos_index(7).long()
        qkvp = (sa.query.weight, sa.key.weight, sa.value.weight, sa.query.bias, sa.key.bias, sa.value.bias)
        if C in ops.ATTN_FUSED_WIDTHS:      # the product's path per stage (cs_vit/net/swin_b200.py::_block)
            pk = ops.pack_attn_fused(*qkvp, sa.relative_position_bias_table, rel_index, dt, g, b_)
            def ours(xin):
                ctx = ops.swin_attn_fused(xin, 1e-5, *pk, xin.shape[0] // N, hw, hw, heads, 7, shift)
                ops.linear(ctx, wo, bo, resid=xin, out=xin)
                return xin
        else:
            wqs, bqs = ops.pack_qkv_prescaled(*qkvp, dt)
            bias_l2 = ops.pack_rel_bias_log2(sa.relative_position_bias_table, rel_index)
            def ours(xin):
                xn = ops.layernorm(xin, g, b_, 1e-5, out_dtype=dt, mode=ops.LN_WINDOW, grid=(hw, hw), ws=7, shift=shift)
                qkv = ops.linear(xn, wqs, bqs, out_dtype=dt)
                ctx = ops.swin_attn_core(qkv, bias_l2, xin.shape[0] // N, hw, hw, heads, 7, shift, token_order=True, q_prescaled=True)
                ops.linear(ctx, wo, bo, resid=xin, out=xin)
                return xin
        x2 = xg.reshape(B * N, C).clone()
        chk = ours(xg[:8].reshape(8 * N, C).clone()).view(8, N, C)
        err = ((chk - ref).norm() / ref.norm()).item()
        our_ms = timeit(lambda: ours(x2), 20)
        print(json.dumps({"stage": stage, "tokens": N, "C": C, "shift": shift, "batch": B, "operands": PREC,
                          "ours_ms": round(our_ms, 4), "ours_img_s": round(B / our_ms * 1e3), "tflops": round(B * flops_img / our_ms / 1e9, 1),
                          "frac_of_sustained_bf16_peak": round(B * flops_img / our_ms / 1e9 / peak, 3),
                          "ref_gpu_fp32_eager_ms": round(ref_ms, 3), "speedup_vs_ref_gpu": round(ref_ms / our_ms, 2),
                          "ref_cpu_ms_per_8img": round(cpu_ms, 1), "cpu_cores": torch.get_num_threads(),
                          "rel_err_vs_ref": float(f"{err:.2e}")}), flush=True)
