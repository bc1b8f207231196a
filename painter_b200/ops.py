"""Thin torch-tensor front end over the C ABI (include/painter_b200.h).

Every function takes CUDA tensors, passes raw device pointers + the current CUDA stream to
libpainter_b200.so and returns torch tensors.  PyTorch is only the allocator / stream provider here.
"""
import ctypes

import torch

from . import _lib
from ._lib import (EPI_BF16, EPI_DGELU, EPI_F32, EPI_GELU, EPI_PIXSHUF, EPI_RESID, PkEpilogue, check,
                   lib)

_vp = ctypes.c_void_p


def _ptr(t):
    return _vp(t.data_ptr()) if t is not None else _vp(0)


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def _req(t, dtype, name):
    if not t.is_cuda:
        raise RuntimeError(f"painter_b200: {name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"painter_b200: {name} must be {dtype}, got {t.dtype}")


def gemm(a, b, *, trans_a=False, trans_b=False, kind=EPI_BF16, out=None, out2=None, bias=None,
         aux=None, rowscale=None, rows_per_group=0, alpha=1.0, accumulate=False, pixshuf=None):
    """C[M,N] = A[M,K] . B[N,K]^T on tcgen05 (bf16 in, fp32 accumulate).

    a: [M,K] (or [K,M] when trans_a);  b: [N,K] (or [K,N] when trans_b); last dim contiguous.
    pixshuf = (h, w, p, c, out_tensor[B, h*p, w*p, c]) for EPI_PIXSHUF.
    """
    _req(a, torch.bfloat16, "a")
    _req(b, torch.bfloat16, "b")
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    if trans_a:
        K, M = a.shape
    else:
        M, K = a.shape
    if trans_b:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    assert K == Kb, (a.shape, b.shape, trans_a, trans_b)
    e = PkEpilogue()
    e.kind = kind
    if kind == EPI_PIXSHUF:
        h, w, p, c, out = pixshuf
        e.ps_h, e.ps_w, e.ps_p, e.ps_c = h, w, p, c
        _req(out, torch.bfloat16, "out")
        assert out.is_contiguous()
        e.ldc = N
    else:
        if out is None:
            odt = torch.float32 if kind in (EPI_F32, EPI_RESID) else torch.bfloat16
            out = torch.empty((M, N), dtype=odt, device=a.device)
        assert out.stride(1) == 1 and tuple(out.shape) == (M, N)
        e.ldc = out.stride(0)
    if kind == EPI_GELU:
        if out2 is None:
            out2 = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
        assert out2.stride() == out.stride()
    e.out = out.data_ptr()
    e.out2 = out2.data_ptr() if out2 is not None else None
    if bias is not None:
        _req(bias, torch.float32, "bias")
        assert bias.numel() == N and bias.is_contiguous()
        e.bias = bias.data_ptr()
    if aux is not None:
        assert aux.stride(1) == 1 and tuple(aux.shape) == (M, N)
        _req(aux, torch.float32 if kind == EPI_RESID else torch.bfloat16, "aux")
        e.aux = aux.data_ptr()
        e.ld_aux = aux.stride(0)
    if rowscale is not None:
        _req(rowscale, torch.float32, "rowscale")
        e.rowscale = rowscale.data_ptr()
        e.rows_per_group = rows_per_group
    e.alpha = alpha
    e.accumulate = 1 if accumulate else 0
    check(lib().pk_gemm_bf16(_ptr(a), _ptr(b), M, N, K, a.stride(0), b.stride(0), int(trans_a),
                             int(trans_b), ctypes.byref(e), _stream()), "pk_gemm_bf16")
    if kind == EPI_GELU:
        return out, out2
    return out


def _pad16(n):
    return (n + 15) // 16 * 16


def relpos_table_bf16(table):
    """fp32 [L, 64] rel-pos table -> zero-padded bf16 [pad16(L), 64] (attention kernel operand)."""
    _req(table, torch.float32, "table")
    assert table.dim() == 2 and table.shape[1] == 64 and table.is_contiguous()
    L = table.shape[0]
    out = torch.empty((_pad16(L), 64), dtype=torch.bfloat16, device=table.device)
    check(lib().pk_relpos_table_bf16(_ptr(table), _ptr(out), L, out.shape[0], _stream()), "pk_relpos_table_bf16")
    return out


def attn_fwd(qkv, th, tw, B, heads, h, w, need_lse=True):
    """Fused attention forward. qkv: bf16 [B*h*w, 3*heads*64]; th/tw: padded bf16 tables for an (h, w) grid.
    Returns (out bf16 [B*h*w, heads*64], lse fp32 [B*heads, h*w] in the log2 domain)."""
    _req(qkv, torch.bfloat16, "qkv")
    N, C = h * w, heads * 64
    assert qkv.is_contiguous() and tuple(qkv.shape) == (B * N, 3 * C)
    assert th.shape[0] >= 2 * h - 1 and tw.shape[0] >= 2 * w - 1
    out = torch.empty((B * N, C), dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty((B * heads, N), dtype=torch.float32, device=qkv.device) if need_lse else None
    check(lib().pk_attn_fwd(_ptr(qkv), _ptr(th), _ptr(tw), _ptr(out), _ptr(lse), B, heads, h, w,
                            th.shape[0], tw.shape[0], _stream()), "pk_attn_fwd")
    return out, lse
