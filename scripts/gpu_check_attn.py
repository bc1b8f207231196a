"""GPU bring-up check of the fused attention forward (each case in a subprocess)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (name, B, heads, h, w)
CASES = [
    ("tiny_8x4", 2, 2, 8, 4),
    ("tiny_4x2", 3, 1, 4, 2),
    ("win_14x14", 4, 2, 14, 14),
    ("win_7x7", 3, 2, 7, 7),
    ("partial_6x28", 2, 2, 6, 28),
    ("one_tile_4x28", 1, 1, 4, 28),
    ("mid_56x28", 1, 2, 56, 28),
    ("h_16x8", 2, 1, 16, 8),
    ("long_112x56", 1, 1, 112, 56),
    ("full_b8", 8, 16, 56, 28),
    ("full_b16", 16, 16, 56, 28),
]


def ref_attn(qkv, th, tw, B, heads, h, w):
    import torch
    N, C = h * w, heads * 64
    q, k, v = qkv.float().reshape(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = (q * 0.125) @ k.transpose(-1, -2)
    idx_h = torch.arange(h, device=qkv.device)[:, None] - torch.arange(h, device=qkv.device)[None, :] + h - 1
    idx_w = torch.arange(w, device=qkv.device)[:, None] - torch.arange(w, device=qkv.device)[None, :] + w - 1
    Rh = th.float()[idx_h]
    Rw = tw.float()[idx_w]
    rq = q.reshape(B, heads, h, w, 64)
    rel_h = torch.einsum("bnhwc,hkc->bnhwk", rq, Rh)
    rel_w = torch.einsum("bnhwc,wkc->bnhwk", rq, Rw)
    s = (s.reshape(B, heads, h, w, h, w) + rel_h[..., :, None] + rel_w[..., None, :]).reshape(B, heads, N, N)
    lse = torch.logsumexp(s, -1) * 1.4426950408889634
    p = s.softmax(-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B * N, C)
    return o, lse.reshape(B * heads, N)


def check_bwd(res, qkv, th, tw, out, lse, B, heads, h, w):
    """Backward vs torch autograd through the fp32 reference (on bf16-rounded inputs)."""
    import torch
    from painter_b200 import ops
    N, C = h * w, heads * 64
    dout = (torch.randn(B * N, C, device=qkv.device) * 0.5).bfloat16()
    dqkv, dTh, dTw = ops.attn_bwd(qkv, out, dout, lse, th, tw, B, heads, h, w)
    torch.cuda.synchronize()
    if B * heads * N * N <= 2 ** 26:
        q32 = qkv.float().requires_grad_(True)
        th32 = th.float().requires_grad_(True)
        tw32 = tw.float().requires_grad_(True)
        ro, _ = ref_attn(q32, th32, tw32, B, heads, h, w)
        (ro * dout.float()).sum().backward()
        def rel(a, b):
            return ((a.float() - b).abs().max() / b.abs().max().clamp_min(1e-9)).item()
        g = q32.grad.reshape(B * N, 3, C)
        d = dqkv.float().reshape(B * N, 3, C)
        res["dq_err"] = rel(d[:, 0], g[:, 0])
        res["dk_err"] = rel(d[:, 1], g[:, 1])
        res["dv_err"] = rel(d[:, 2], g[:, 2])
        res["dTh_err"] = rel(dTh, th32.grad[: 2 * h - 1])
        res["dTw_err"] = rel(dTw, tw32.grad[: 2 * w - 1])
        res["bwd_ok"] = bool(max(res["dq_err"], res["dk_err"], res["dv_err"], res["dTh_err"], res["dTw_err"]) < 3e-2)
        res["ok"] = bool(res.get("ok", True) and res["bwd_ok"])
    else:
        res["bwd_finite"] = bool(torch.isfinite(dqkv.float()).all().item() and torch.isfinite(dTh).all().item())
    if B * heads * N * N >= 2 ** 27:
        for _ in range(3):
            ops.attn_bwd(qkv, out, dout, lse, th, tw, B, heads, h, w)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.attn_bwd(qkv, out, dout, lse, th, tw, B, heads, h, w)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res["bwd_ms"] = ms
        res["bwd_tflops"] = 10.0 * B * heads * N * N * 64 / ms / 1e9


def run_case(name, B, heads, h, w):
    import torch
    from painter_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    N, C = h * w, heads * 64
    qkv = (torch.randn(B * N, 3 * C, device=dev) * 1.5).bfloat16()
    th = ops.relpos_table_bf16(torch.randn(2 * h - 1, 64, device=dev) * 0.3)
    tw = ops.relpos_table_bf16(torch.randn(2 * w - 1, 64, device=dev) * 0.3)
    out, lse = ops.attn_fwd(qkv, th, tw, B, heads, h, w)
    torch.cuda.synchronize()
    res = {"case": name, "B": B, "heads": heads, "h": h, "w": w}
    try:
        _run_fwd_checks(res, qkv, th, tw, out, lse, B, heads, h, w)
    finally:
        pass
    check_bwd(res, qkv, th, tw, out, lse, B, heads, h, w)
    return res


def _run_fwd_checks(res, qkv, th, tw, out, lse, B, heads, h, w):
    import torch
    from painter_b200 import ops
    N, C = h * w, heads * 64
    if B * heads * N * N <= 2 ** 28:
        ro, rl = ref_attn(qkv, th, tw, B, heads, h, w)
        res["out_err"] = ((out.float() - ro).abs().max() / ro.abs().max()).item()
        res["lse_err"] = (lse - rl).abs().max().item()
        res["ok"] = bool(res["out_err"] < 2e-2 and res["lse_err"] < 2e-2)
    else:
        # compare one (b, head) slab against the reference
        ro, rl = ref_attn(qkv[:N], th, tw, 1, heads, h, w)
        res["out_err"] = ((out[:N].float() - ro).abs().max() / ro.abs().max()).item()
        res["lse_err"] = (lse[:heads] - rl).abs().max().item()
        res["ok"] = bool(res["out_err"] < 2e-2 and res["lse_err"] < 2e-2 and torch.isfinite(out.float()).all().item())
    if B * heads * N * N >= 2 ** 27:
        for _ in range(3):
            ops.attn_fwd(qkv, th, tw, B, heads, h, w)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.attn_fwd(qkv, th, tw, B, heads, h, w)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res["ms"] = ms
        res["tflops"] = 4.0 * B * heads * N * N * 64 / ms / 1e9


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--case":
        idx = int(sys.argv[2])
        try:
            r = run_case(*CASES[idx])
        except Exception as ex:  # noqa
            r = {"case": CASES[idx][0], "ok": False, "error": repr(ex)[:600]}
        print("RESULT " + json.dumps(r), flush=True)
        sys.exit(0)
    out = os.path.join(ROOT, "gpurun_out", "attn_check.jsonl")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        for i, c in enumerate(CASES):
            t0 = time.time()
            try:
                p = subprocess.run([sys.executable, __file__, "--case", str(i)], capture_output=True, text=True,
                                   timeout=120)
                line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
                r = json.loads(line[-1][7:]) if line else {"case": c[0], "ok": False, "rc": p.returncode,
                                                            "stderr": p.stderr[-800:], "stdout": p.stdout[-800:]}
            except subprocess.TimeoutExpired:
                r = {"case": c[0], "ok": False, "error": "timeout"}
            r["wall_s"] = round(time.time() - t0, 1)
            f.write(json.dumps(r) + "\n")
            f.flush()
            print(json.dumps(r), flush=True)
