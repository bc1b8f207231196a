"""Per-kernel timing of the attention backward (debug skip flags) at the BASELINE geometry."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from painter_b200 import ops, _lib

B, heads, h, w = int(os.environ.get("PK_B", 8)), 16, 56, 28
N, C = h * w, heads * 64
qkv = torch.randn(B * N, 3 * C, device="cuda").bfloat16()
th = ops.relpos_table_bf16(torch.randn(2 * h - 1, 64, device="cuda") * 0.1)
tw = ops.relpos_table_bf16(torch.randn(2 * w - 1, 64, device="cuda") * 0.1)
out, lse = ops.attn_fwd(qkv, th, tw, B, heads, h, w)
do = (torch.randn(B * N, C, device="cuda") * 0.5).bfloat16()


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


L = _lib.lib()
res = {"fwd_ms": timed(lambda: ops.attn_fwd(qkv, th, tw, B, heads, h, w)),
       "fwd_save_ms": timed(lambda: ops.attn_fwd(qkv, th, tw, B, heads, h, w, save_rel=True))}
_, _, rel = ops.attn_fwd(qkv, th, tw, B, heads, h, w, save_rel=True)
for tag, r in (("", None), ("saved_", rel)):     # recomputed bias rows | bias rows kept by the forward
    for name, flag in (("bwd_all", 0), ("bwd_no_dq", 4), ("bwd_no_dkv", 8), ("bwd_neither", 12)):
        L.pk_attn_bwd_debug(flag)
        res[tag + name] = timed(lambda: ops.attn_bwd(qkv, out, do, lse, th, tw, B, heads, h, w, rel=r))
    L.pk_attn_bwd_debug(0)
    res[tag + "dq_ms"] = res[tag + "bwd_all"] - res[tag + "bwd_no_dq"]
    res[tag + "dkv_ms"] = res[tag + "bwd_all"] - res[tag + "bwd_no_dkv"]
print({k: round(v, 4) for k, v in res.items()})
