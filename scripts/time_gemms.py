"""Per-shape timing of the block-level GEMMs of the training step (B = 8: M = 12544 tokens) with rotating operand
copies (working set > L2), CUDA events on the launching stream.  PK_LIB selects a tuning build.
Usage (GPU box): python scripts/time_gemms.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from painter_b200 import ops  # noqa: E402

dev = "cuda"
M, C = 12544, 1024
R = 3   # rotating copies
torch.manual_seed(0)


def rnd(*shape, dtype=torch.bfloat16, scale=1.0):
    return [(torch.randn(*shape, device=dev) * scale).to(dtype) for _ in range(R)]


x, x3, x4 = rnd(M, C), rnd(M, 3 * C), rnd(M, 4 * C)
wqkv, wproj = rnd(3 * C, C, scale=0.02), rnd(C, C, scale=0.02)
wfc1, wfc2 = rnd(4 * C, C, scale=0.02), rnd(C, 4 * C, scale=0.02)
b1, b3, b4 = torch.randn(C, device=dev), torch.randn(3 * C, device=dev), torch.randn(4 * C, device=dev)
res = rnd(M, C, dtype=torch.float32)
z = rnd(M, 4 * C)
g4 = [torch.zeros(4 * C, C, device=dev) for _ in range(R)]
g3 = [torch.zeros(3 * C, C, device=dev) for _ in range(R)]
g1 = [torch.zeros(C, C, device=dev) for _ in range(R)]

cases = [
    ("qkv fwd   N=3072 K=1024 bf16+bias", 2.0 * M * 3 * C * C, lambda i: ops.gemm(x[i], wqkv[i], kind=ops.EPI_BF16, bias=b3)),
    ("proj fwd  N=1024 K=1024 resid", 2.0 * M * C * C, lambda i: ops.gemm(x[i], wproj[i], kind=ops.EPI_RESID, bias=b1, aux=res[i])),
    ("fc1 fwd   N=4096 K=1024 gelu", 2.0 * M * 4 * C * C, lambda i: ops.gemm(x[i], wfc1[i], kind=ops.EPI_GELU, bias=b4)),
    ("fc2 fwd   N=1024 K=4096 resid", 2.0 * M * 4 * C * C, lambda i: ops.gemm(x4[i], wfc2[i], kind=ops.EPI_RESID, bias=b1, aux=res[i])),
    ("fc2 dgrad N=4096 K=1024 dgelu", 2.0 * M * 4 * C * C, lambda i: ops.gemm(x[i], wfc2[i], trans_b=True, kind=ops.EPI_DGELU, aux=z[i])),
    ("fc1 dgrad N=1024 K=4096 bf16", 2.0 * M * 4 * C * C, lambda i: ops.gemm(x4[i], wfc1[i], trans_b=True, kind=ops.EPI_BF16)),
    ("proj dgrad N=1024 K=1024 bf16", 2.0 * M * C * C, lambda i: ops.gemm(x[i], wproj[i], trans_b=True, kind=ops.EPI_BF16)),
    ("qkv dgrad N=1024 K=3072 f32", 2.0 * M * 3 * C * C, lambda i: ops.gemm(x3[i], wqkv[i], trans_b=True, kind=ops.EPI_F32)),
    ("fc1 wgrad 4096x1024 K=12544", 2.0 * M * 4 * C * C, lambda i: ops.gemm(x4[i], x[i], trans_a=True, trans_b=True, kind=ops.EPI_F32, out=g4[i], accumulate=2)),
    ("qkv wgrad 3072x1024 K=12544", 2.0 * M * 3 * C * C, lambda i: ops.gemm(x3[i], x[i], trans_a=True, trans_b=True, kind=ops.EPI_F32, out=g3[i], accumulate=2)),
    ("proj wgrad 1024x1024 K=12544", 2.0 * M * C * C, lambda i: ops.gemm(x[i], x[(i + 1) % R], trans_a=True, trans_b=True, kind=ops.EPI_F32, out=g1[i], accumulate=2)),
]
if os.environ.get("PK_CUBLAS"):
    cases += [
        ("cuBLAS    N=3072 K=1024 (torch.matmul)", 2.0 * M * 3 * C * C, lambda i: torch.matmul(x[i], wqkv[i].t())),
        ("cuBLAS    N=4096 K=1024 (torch.matmul)", 2.0 * M * 4 * C * C, lambda i: torch.matmul(x[i], wfc1[i].t())),
        ("cuBLAS    N=1024 K=4096 (torch.matmul)", 2.0 * M * 4 * C * C, lambda i: torch.matmul(x4[i], wfc2[i].t())),
        ("cuBLAS    N=1024 K=1024 (torch.matmul)", 2.0 * M * C * C, lambda i: torch.matmul(x[i], wproj[i].t())),
    ]

iters = 30
tot = 0.0
for name, flops, fn in cases:
    for i in range(3):
        fn(i % R)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % R)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    if not name.startswith("cuBLAS"):
        tot += us
    print(f"{name:42s} {us:8.1f} us  {flops / us / 1e6:7.0f} TFLOP/s")
print(f"sum over one block's GEMMs (fwd + bwd, ex. fc2 wgrad): {tot:.1f} us")
