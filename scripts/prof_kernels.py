"""One launch of each hot kernel at the BASELINE geometry (batch 8) for `ncu --set full` captures.
Usage (GPU box):  ncu --set full --clock-control none --import-source on -o gpurun_out/prof python scripts/prof_kernels.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from painter_b200 import ops  # noqa: E402

which = sys.argv[1:] or ["gemm", "attn"]
torch.manual_seed(0)
dev = "cuda"
M, C = 12544, 1024
B, heads, h, w = 8, 16, 56, 28
N = h * w

x = (torch.randn(M, C, device=dev)).bfloat16()
x4 = (torch.randn(M, 4 * C, device=dev)).bfloat16()
wqkv = (torch.randn(3 * C, C, device=dev) * 0.02).bfloat16()
wproj = (torch.randn(C, C, device=dev) * 0.02).bfloat16()
wfc1 = (torch.randn(4 * C, C, device=dev) * 0.02).bfloat16()
wfc2 = (torch.randn(C, 4 * C, device=dev) * 0.02).bfloat16()
b3 = torch.randn(3 * C, device=dev)
b1 = torch.randn(C, device=dev)
b4 = torch.randn(4 * C, device=dev)
res = torch.randn(M, C, device=dev)

for rep in range(2):   # first pass warms up (ncu: use -s to skip it)
    if "gemm" in which:
        qkv = ops.gemm(x, wqkv, kind=ops.EPI_BF16, bias=b3)                       # qkv fwd
        ops.gemm(x, wproj, kind=ops.EPI_RESID, bias=b1, aux=res)                   # proj fwd
        z, hh = ops.gemm(x, wfc1, kind=ops.EPI_GELU, bias=b4)                      # fc1 fwd
        ops.gemm(x4, wfc2, kind=ops.EPI_RESID, bias=b1, aux=res)                   # fc2 fwd
        ops.gemm(x, wfc2, trans_b=True, kind=ops.EPI_DGELU, aux=z)                 # fc2 dgrad
        ops.gemm(x4, wfc1, trans_b=True, kind=ops.EPI_F32)                         # fc1 dgrad
        out = torch.zeros(4 * C, C, device=dev)
        ops.gemm(x4, x, trans_a=True, trans_b=True, kind=ops.EPI_F32, out=out, accumulate=2)   # fc1 wgrad
        out = torch.zeros(3 * C, C, device=dev)
        ops.gemm(qkv, x, trans_a=True, trans_b=True, kind=ops.EPI_F32, out=out, accumulate=2)  # qkv wgrad
    if "attn" in which:
        qkv = (torch.randn(B * N, 3 * C, device=dev) * 1.0).bfloat16()
        th = ops.relpos_table_bf16(torch.randn(2 * h - 1, 64, device=dev) * 0.1)
        tw = ops.relpos_table_bf16(torch.randn(2 * w - 1, 64, device=dev) * 0.1)
        o, lse = ops.attn_fwd(qkv, th, tw, B, heads, h, w)
        do = (torch.randn(B * N, C, device=dev) * 0.5).bfloat16()
        ops.attn_bwd(qkv, o, do, lse, th, tw, B, heads, h, w)
    torch.cuda.synchronize()
print("done")
