import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from painter_b200 import ops, _lib
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
from gpu_check_attn import ref_attn
B, heads, h, w = 1, 2, 56, 28
N, C = h * w, heads * 64
torch.manual_seed(0)
qkv = (torch.randn(B * N, 3 * C, device="cuda") * 1.5).bfloat16()
th = ops.relpos_table_bf16(torch.randn(2 * h - 1, 64, device="cuda") * 0.3)
tw = ops.relpos_table_bf16(torch.randn(2 * w - 1, 64, device="cuda") * 0.3)
out, lse = ops.attn_fwd(qkv, th, tw, B, heads, h, w)
dout = (torch.randn(B * N, C, device="cuda") * 0.5).bfloat16()
q32 = qkv.float().requires_grad_(True)
ro, _ = ref_attn(q32, th.float(), tw.float(), B, heads, h, w)
(ro * dout.float()).sum().backward()
g = q32.grad.reshape(B * N, 3, C)
for dbg in (1, 0):
    _lib.lib().pk_attn_bwd_debug(dbg)
    dqkv, dTh, dTw = ops.attn_bwd(qkv, out, dout, lse, th, tw, B, heads, h, w)
    torch.cuda.synchronize()
    d = dqkv.float().reshape(B * N, 3, C)
    dq = d[:, 0]
    bad = ~torch.isfinite(dq)
    print("debug", dbg, "nonfinite", int(bad.sum()), "rows with nonfinite", bad.any(1).nonzero().flatten()[:20].tolist(),
          "cols", bad.any(0).nonzero().flatten()[:20].tolist())
    ok = torch.isfinite(dq)
    err = ((dq - g[:, 0]).abs() * ok).max() / g[:, 0].abs().max()
    print("   max rel err over finite entries", err.item(), " dk", ((d[:,1]-g[:,1]).abs().max()/g[:,1].abs().max()).item())
