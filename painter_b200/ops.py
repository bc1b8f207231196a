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
