"""Torch-tensor front end of the C ABI: argument checking, output allocation, stream hand-off.

PyTorch is used here only for device memory and streams.  Every function requires CUDA tensors and launches
on ``torch.cuda.current_stream()``; nothing falls back to ATen math.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib

F32, BF16, F16 = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
LN_IDENTITY, LN_WINDOW, LN_MERGE2X2 = 0, 1, 2
GEMM_TC, GEMM_SIMT = 0, 1

_DT = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}

# Count of kernel launches issued through this module (bench.py reports it as `gpu_launches`).
launch_count = 0


def _code(dtype: torch.dtype) -> int:
    try:
        return _DT[dtype]
    except KeyError:
        raise TypeError(f"cs_vit kernels take float32, bfloat16 or float16 tensors, got {dtype}") from None


def _dev(*tensors: Optional[torch.Tensor]) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("cs_vit kernels run on CUDA tensors only (there is no CPU fallback)")


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


_profile = None  # {"name": entry point, "events": [(start, stop, flops, bytes)]} while bench.py instruments a kernel


def begin_profile(name: str) -> None:
    """Bracket every launch of C-ABI entry ``name`` with CUDA events on the launch stream (bench.py roofline)."""
    global _profile
    _profile = {"name": name, "events": []}


def end_profile() -> dict:
    global _profile
    prof, _profile = _profile, None
    torch.cuda.synchronize()
    ev = prof["events"]
    return {"launches": len(ev), "ms": sum(a.elapsed_time(b) for a, b, _, _ in ev),
            "flops": float(sum(f for _, _, f, _ in ev)), "bytes": float(sum(b for _, _, _, b in ev))}


def _call(name: str, *args, flops: float = 0.0, nbytes: float = 0.0) -> None:
    global launch_count
    launch_count += 1
    if _profile is not None and _profile["name"] == name:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(getattr(_lib.load(), name)(*args))
        e1.record()
        _profile["events"].append((e0, e1, flops, nbytes))
        return
    _lib.check(getattr(_lib.load(), name)(*args))


def _rows2d(t: torch.Tensor) -> Tuple[int, int, int]:
    """(rows, cols, row pitch in elements) of a 2-D tensor whose last dim is contiguous."""
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"expected a 2-D tensor with contiguous rows, got shape {tuple(t.shape)} strides {t.stride()}")
    return t.shape[0], t.shape[1], t.stride(0)


# ---------------------------------------------------------------------------------------------- integer maps
def window_index_map(H: int, W: int, ws: int, shift: int, device="cuda") -> torch.Tensor:
    out = torch.empty(H * W, dtype=torch.int32, device=device)
    _call("csvit_window_index_map", H, W, ws, shift, out.data_ptr(), _stream())
    return out


def shift_mask(H: int, W: int, ws: int, shift: int, device="cuda") -> torch.Tensor:
    nW, L = (H // ws) * (W // ws), ws * ws
    out = torch.empty(nW, L, L, dtype=torch.float32, device=device)
    _call("csvit_shift_mask", H, W, ws, shift, out.data_ptr(), _stream())
    return out


def rel_pos_index(ws: int, device="cuda") -> torch.Tensor:
    out = torch.empty(ws * ws, ws * ws, dtype=torch.int32, device=device)
    _call("csvit_rel_pos_index", ws, out.data_ptr(), _stream())
    return out


def merge_index_map(H: int, W: int, device="cuda") -> torch.Tensor:
    out = torch.empty((H // 2) * (W // 2), 4, dtype=torch.int32, device=device)
    _call("csvit_merge_index_map", H, W, out.data_ptr(), _stream())
    return out


def expand_rel_bias(table: torch.Tensor, ws: int) -> torch.Tensor:
    _dev(table)
    table = table.contiguous().float()
    heads = table.shape[1]
    out = torch.empty(heads, ws * ws, ws * ws, dtype=torch.float32, device=table.device)
    _call("csvit_expand_rel_bias", table.data_ptr(), out.data_ptr(), heads, ws, _stream())
    return out


def expand_rel_bias_mma(table: torch.Tensor, ws: int = 7) -> torch.Tensor:
    """Bias table in MMA accumulator-fragment order for the 16-bit window-attention kernel."""
    _dev(table)
    table = table.contiguous().float()
    heads = table.shape[1]
    out = torch.empty(heads, 4, 7, 32, 4, dtype=torch.float32, device=table.device)
    _call("csvit_expand_rel_bias_mma", table.data_ptr(), out.data_ptr(), heads, ws, _stream())
    return out


# ---------------------------------------------------------------------------------------------- row kernels
def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, *, out_dtype=torch.float32,
              mode: int = LN_IDENTITY, grid: Tuple[int, int] = (0, 0), ws: int = 0, shift: int = 0,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: fp32 ``[rows_in, C]`` (tokens of all images, row-major).  Returns ``[rows_out, C or 4C]``."""
    _dev(x, gamma, beta, out)
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("layernorm input must be contiguous float32")
    rows_in, C = x.shape
    H, W = grid
    if mode == LN_MERGE2X2:
        rows, width = rows_in // 4, 4 * C
    else:
        rows, width = rows_in, C
    if gamma.numel() != width or beta.numel() != width:
        raise ValueError(f"layernorm affine width {gamma.numel()} != {width}")
    if out is None:
        out = torch.empty(rows, width, dtype=out_dtype, device=x.device)
    _call("csvit_layernorm", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), float(eps), out.data_ptr(), _code(out.dtype),
          out.stride(0), rows, C, mode, H, W, ws, shift, _stream(),
          nbytes=float(rows * width * (4 + out.element_size())))
    return out


def affine_rows(x: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, *, out_dtype=torch.float32) -> torch.Tensor:
    _dev(x, scale, shift)
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("affine_rows input must be contiguous float32")
    C = x.shape[-1]
    out = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    _call("csvit_affine_rows", x.data_ptr(), scale.data_ptr(), shift.data_ptr(), out.data_ptr(), _code(out_dtype),
          x.numel() // C, C, _stream())
    return out


_MEAN = (ctypes.c_float * 3)(0.485, 0.456, 0.406)  # ref:cs_vit/net/ti_poser.py:240-242
_STD = (ctypes.c_float * 3)(0.229, 0.224, 0.225)


def patch_im2col(img: torch.Tensor, *, out_dtype=torch.bfloat16, normalize: bool = True) -> torch.Tensor:
    """img fp32 ``[B,3,S,S]`` in [0,1] -> ``[B*(S/4)^2, 48]`` normalised 4x4 patches."""
    _dev(img)
    if img.dtype != torch.float32 or not img.is_contiguous() or img.dim() != 4 or img.shape[1] != 3:
        raise ValueError("patch_im2col takes contiguous float32 [B,3,S,S]")
    B, _, S, S2 = img.shape
    if S != S2:
        raise ValueError("square images only")
    out = torch.empty(B * (S // 4) ** 2, 48, dtype=out_dtype, device=img.device)
    mean = _MEAN if normalize else (ctypes.c_float * 3)(0, 0, 0)
    std = _STD if normalize else (ctypes.c_float * 3)(1, 1, 1)
    _call("csvit_patch_im2col", img.data_ptr(), out.data_ptr(), _code(out_dtype), B, S, mean, std, _stream())
    return out


# ---------------------------------------------------------------------------------------------- GEMM engine
def linear(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, act: int = ACT_NONE,
           resid: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None,
           scatter: Optional[Tuple[int, int, int, int]] = None, impl: int = GEMM_TC) -> torch.Tensor:
    """``out[orow] = act(a @ w.T + bias) + resid[orow]``; see ``csvit_linear`` in include/csvit.h."""
    _dev(a, w, bias, resid, out)
    M, K, lda = _rows2d(a)
    N, K2, ldw = _rows2d(w)
    if K != K2 or a.dtype != w.dtype:
        raise ValueError(f"linear: A {tuple(a.shape)}/{a.dtype} vs W {tuple(w.shape)}/{w.dtype}")
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype or a.dtype, device=a.device)
    _, No, ldo = _rows2d(out)
    if No != N:
        raise ValueError("linear: output width mismatch")
    ldr = 0
    if resid is not None:
        if resid.dtype != torch.float32:
            raise ValueError("linear: residual must be float32")
        ldr = _rows2d(resid)[2]
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != N):
        raise ValueError("linear: bias must be float32 [N]")
    sH, sW, sws, ssh = scatter if scatter is not None else (0, 0, 0, 0)
    _call("csvit_linear", a.data_ptr(), lda, w.data_ptr(), ldw, _code(a.dtype), M, N, K, _p(bias), act, _p(resid), ldr,
          out.data_ptr(), ldo, _code(out.dtype), sH, sW, sws, ssh, impl, _stream(), flops=2.0 * M * N * K)
    return out


def ln_linear(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, w: torch.Tensor,
              bias: Optional[torch.Tensor] = None, *, act: int = ACT_NONE, mode: int = LN_IDENTITY,
              grid: Tuple[int, int] = (0, 0), ws: int = 0, shift: int = 0) -> torch.Tensor:
    """``act(LayerNorm(x rows) @ w.T + bias)`` in one kernel (see ``csvit_ln_linear``); x fp32 ``[M, C]``, w 16-bit ``[N, C]``."""
    _dev(x, gamma, beta, w, bias)
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("ln_linear input must be contiguous float32")
    M, C = x.shape
    N, C2, ldw = _rows2d(w)
    if C2 != C or w.dtype not in (torch.bfloat16, torch.float16):
        raise ValueError("ln_linear: weight must be 16-bit [N, C]")
    out = torch.empty(M, N, dtype=w.dtype, device=x.device)
    H, W = grid
    _call("csvit_ln_linear", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), float(eps), mode, H, W, ws, shift, w.data_ptr(),
          ldw, _code(w.dtype), M, N, C, _p(bias), act, out.data_ptr(), out.stride(0), _stream(), flops=2.0 * M * N * C)
    return out


LN_LINEAR_WIDTHS = (128, 256, 512)
MLP_FUSED_WIDTHS = (128, 256)


def mlp_fused(xn: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """In place ``x += GELU(xn @ w1.T + b1) @ w2.T + b2`` without materialising the hidden tensor (``csvit_mlp_fused``)."""
    _dev(xn, w1, b1, w2, b2, x)
    M, C, ldxn = _rows2d(xn)
    if w1.shape != (4 * C, C) or w2.shape != (C, 4 * C) or x.shape != (M, C) or x.dtype != torch.float32:
        raise ValueError("mlp_fused: shape mismatch")
    if not (xn.dtype == w1.dtype == w2.dtype) or xn.dtype not in (torch.bfloat16, torch.float16):
        raise ValueError("mlp_fused: xn / w1 / w2 must share a 16-bit dtype")
    _call("csvit_mlp_fused", xn.data_ptr(), ldxn, w1.data_ptr(), w1.stride(0), b1.data_ptr(), w2.data_ptr(), w2.stride(0),
          b2.data_ptr(), x.data_ptr(), x.stride(0), _code(xn.dtype), M, C, _stream(), flops=16.0 * M * C * C)
    return x


# ---------------------------------------------------------------------------------------------- attention
def window_attention(qkv: torch.Tensor, bias_exp: torch.Tensor, B: int, H: int, W: int, heads: int, ws: int,
                     shift: int, bias_mma: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``bias_exp``: ``expand_rel_bias`` table ``[h,L,L]`` (fp32 kernel and the tcgen05 16-bit kernel) or, for backward
    compatibility, an ``expand_rel_bias_mma`` table (5-D) which selects the mma.sync 16-bit kernel."""
    _dev(qkv, bias_exp, bias_mma)
    plain, frag = bias_exp, bias_mma
    if bias_exp.dim() == 5:
        plain, frag = None, bias_exp
    if qkv.dtype != torch.float32 and plain is None and frag is None:
        raise ValueError("window attention needs a bias table")
    rows, C3, ld = _rows2d(qkv)
    C = C3 // 3
    if ld != C3 or rows != B * H * W:
        raise ValueError("window_attention: qkv must be dense [B*H*W, 3C]")
    out = torch.empty(rows, C, dtype=qkv.dtype, device=qkv.device)
    _call("csvit_window_attention", qkv.data_ptr(), _p(plain), _p(frag), out.data_ptr(), _code(qkv.dtype), B, H, W, C,
          heads, ws, shift, _stream(), nbytes=float(qkv.numel() + out.numel()) * qkv.element_size())
    return out


def set_attention_impl(use_tcgen05: bool = True) -> None:
    _lib.check(_lib.load().csvit_set_attention_impl(1 if use_tcgen05 else 0))


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, n_seq: int, Lq: int, S: int, heads: int,
              scale: float) -> torch.Tensor:
    """q ``[n_seq*Lq, D]`` (may be a column slice), k/v ``[n_seq*S, D]``; exact fp32 softmax attention."""
    _dev(q, k, v)
    rq, D, ldq = _rows2d(q)
    rk, _, ldk = _rows2d(k)
    _, _, ldv = _rows2d(v)
    if rq != n_seq * Lq or rk != n_seq * S or D != heads * 32:
        raise ValueError("attention: shape mismatch (head_dim must be 32)")
    out = torch.empty(rq, D, dtype=q.dtype, device=q.device)
    _call("csvit_attention", q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), _code(q.dtype), ldq, ldk, ldv, D,
          n_seq, Lq, S, heads, float(scale), _stream())
    return out


# ---------------------------------------------------------------------------------------------- host-side maps
def host_maps(H: int, W: int, ws: int, shift: int):
    """CPU evaluation of the kernels' closed-form integer maps (same inline functions, compiled for the host).
    Returns ``(window_index_map [H*W] int32, shift_mask [nW,L,L] fp32)`` as CPU tensors; needs no GPU."""
    nW, L = (H // ws) * (W // ws), ws * ws
    idx = torch.empty(H * W, dtype=torch.int32)
    mask = torch.empty(nW, L, L, dtype=torch.float32)
    lib = _lib.load()
    _lib.check(lib.csvit_host_window_index_map(H, W, ws, shift, idx.data_ptr()))
    _lib.check(lib.csvit_host_shift_mask(H, W, ws, shift, mask.data_ptr()))
    return idx, mask


def host_rel_pos_index(ws: int) -> torch.Tensor:
    out = torch.empty(ws * ws, ws * ws, dtype=torch.int32)
    _lib.check(_lib.load().csvit_host_rel_pos_index(ws, out.data_ptr()))
    return out


def host_merge_index_map(H: int, W: int) -> torch.Tensor:
    out = torch.empty((H // 2) * (W // 2), 4, dtype=torch.int32)
    _lib.check(_lib.load().csvit_host_merge_index_map(H, W, out.data_ptr()))
    return out


def set_gemm_tuning(cluster: int = 0, tma_store: int = -1, max_ctas: int = 0, pair: int = -1) -> None:
    """Ablation knobs of the GEMM engine (see ``csvit_set_gemm_tuning``); defaults restore automatic choices."""
    _lib.check(_lib.load().csvit_set_gemm_tuning(cluster, tma_store, max_ctas, pair))
