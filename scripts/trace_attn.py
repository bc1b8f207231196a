"""Timeline of one attention-forward CTA (clock64 stamps) — debug aid."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from painter_b200 import ops, _lib
B, heads, h, w = 8, 16, 56, 28
N, C = h * w, heads * 64
qkv = torch.randn(B * N, 3 * C, device="cuda").bfloat16()
th = ops.relpos_table_bf16(torch.randn(2 * h - 1, 64, device="cuda") * 0.1)
tw = ops.relpos_table_bf16(torch.randn(2 * w - 1, 64, device="cuda") * 0.1)
for _ in range(3):
    ops.attn_fwd(qkv, th, tw, B, heads, h, w)
buf = torch.zeros(3 * 16 * 8, dtype=torch.int64, device="cuda")
_lib.lib().pk_attn_set_trace(ctypes.c_void_p(buf.data_ptr()))
ops.attn_fwd(qkv, th, tw, B, heads, h, w)
torch.cuda.synchronize()
_lib.lib().pk_attn_set_trace(None)
t = buf.cpu().view(3, 16, 8)
t0 = int(t[1, 0, 0])
names = {0: ["ke_seen", "ve_seen"], 1: ["loop_top", "fa_seen", "fb_seen", "vf_seen", "p_seen"],
         2: ["tile_top", "sa_seen", "sb_seen", "softmax_done", "arrived_p"]}
for j in range(14):
    line = f"tile {j:2d}: "
    for role in (1, 2, 0):
        for e, nm in enumerate(names[role]):
            v = int(t[role, j, e])
            line += f"{'PMS'[role] if False else ['prod','mma','smx'][role]}.{nm}={v - t0 if v else -1:7d} "
    print(line)

# ---- backward kernels ----
out, lse, rel = ops.attn_fwd(qkv, th, tw, B, heads, h, w, save_rel=True)
if os.environ.get("PK_TRACE_RECOMPUTE"):
    rel = None      # trace the pair that recomputes the bias rows in the dQ kernel
do = (torch.randn(B * N, C, device="cuda") * 0.5).bfloat16()
for _ in range(2):
    ops.attn_bwd(qkv, out, do, lse, th, tw, B, heads, h, w, rel=rel)
buf2 = torch.zeros(2 * 2 * 16 * 8, dtype=torch.int64, device="cuda")
_lib.lib().pk_attn_bwd_debug(2 if os.environ.get("PK_TRACE_LAST") else 0)
_lib.lib().pk_attn_bwd_set_trace(ctypes.c_void_p(buf2.data_ptr()))
ops.attn_bwd(qkv, out, do, lse, th, tw, B, heads, h, w, rel=rel)
torch.cuda.synchronize()
_lib.lib().pk_attn_bwd_set_trace(None)
_lib.lib().pk_attn_bwd_debug(0)
t2 = buf2.cpu().view(2, 2, 16, 8)
for kern, kn in ((0, "dq "), (1, "dkv")):
    t0 = int(t2[kern, 0, 0, 0])
    for j in range(14 if kern == 0 else 13):
        m = [int(t2[kern, 0, j, e]) - t0 for e in range(6)]
        sx = [int(t2[kern, 1, j, e]) - t0 for e in range(7)]
        print(f"{kn} it {j:2d}: mma top={m[0]:7d} s_next_issued={m[2]:7d} p_seen={m[3]:7d} dp_next_issued={m[5]:7d} "
              f"acc_issued={m[4]:7d} | smx top={sx[5]:7d} operands_in={sx[6]:7d} wait={sx[0]:7d} bar_s_seen={sx[4]:7d} bar_dp_seen={sx[1]:7d} done={sx[2]:7d} arrived={sx[3]:7d}")

ph = [int(t2[0, 1, 15, e]) - int(t2[0, 0, 0, 0]) for e in range(8)]
print("dq phases (cycles rel. to mma top of it 0): start, delta_done, relw_done, relh_done, loop_done, ep1_done, ep2_done, end:", ph)

print("dq: kf_seen(j) [stamp at tile j slot] and first-S-MMA-issued(j), relative:")
t0 = int(t2[0, 0, 0, 0])
print([ (int(t2[0, 0, j, 1]) - t0, int(t2[0, 1, j, 6]) - t0) for j in range(14)])
