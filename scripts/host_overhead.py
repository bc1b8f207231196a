#!/usr/bin/env python
"""Host-side enqueue time of one training step vs its GPU time (is the step CPU-launch-bound?).
   python scripts/host_overhead.py [--batch 8] [--optimizer pk|torch]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from painter_b200 import _lib, models_painter  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--optimizer", default="torch")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = models_painter.painter_vit_large_patch16_input896x448_win_dec64_8glb_sl1().to(dev).train()
    if a.optimizer == "pk":
        from painter_b200.optim import FusedAdamW
        opt = FusedAdamW(model.parameters(), lr=1e-4, weight_decay=0.05)
    else:
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.05, fused=True)
    batch = [t.to(dev) for t in bench._batch(a.batch, 0)]

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss, _, _ = model(batch[0], batch[1], bool_masked_pos=batch[2], valid=batch[3])
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    host, gpu = [], []
    for _ in range(a.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        t0 = time.perf_counter()
        e0.record()
        step()
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        host.append((t1 - t0) * 1e3)
        gpu.append(e0.elapsed_time(e1))
        n1 = _lib.launch_count()
    print(json.dumps({"batch": a.batch, "optimizer": a.optimizer, "host_enqueue_ms": sorted(host)[len(host) // 2],
                      "gpu_ms": sorted(gpu)[len(gpu) // 2], "pk_launches": n1 - n0}))


if __name__ == "__main__":
    main()
