#!/usr/bin/env python
"""Where does a strictly synchronised training step (engine_train.py: loss.item() + torch.cuda.synchronize() every
step) spend its wall time?  Host timestamps around each phase, with and without the DevicePrefetcher."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from painter_b200 import models_painter  # noqa: E402
from painter_b200.data_utils import DevicePrefetcher  # noqa: E402
from painter_b200.optim import FusedAdamW  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = models_painter.painter_vit_large_patch16_input896x448_win_dec64_8glb_sl1().to(dev).train()
    opt = FusedAdamW(model.parameters(), lr=1e-4, weight_decay=0.05)
    host = [t.pin_memory() for t in bench._batch(8, 0)]
    resident = [t.to(dev) for t in host]

    def step(batch, item_after_fwd, sync_end, T):
        t0 = time.perf_counter()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss, _, _ = model(batch[0], batch[1], bool_masked_pos=batch[2], valid=batch[3])
        t1 = time.perf_counter()
        if item_after_fwd:
            loss.item()
        t2 = time.perf_counter()
        loss.backward()
        t3 = time.perf_counter()
        opt.step()
        opt.zero_grad(set_to_none=True)
        t4 = time.perf_counter()
        if sync_end:
            torch.cuda.synchronize()
        t5 = time.perf_counter()
        T.append([round(1e3 * (b - a), 2) for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4), (t4, t5))])

    for _ in range(3):
        step(resident, False, True, [])
    out = {}
    for name, item, sync_end, pref in (("resident_sync_end", False, True, False), ("resident_item_sync", True, True, False),
                                       ("prefetch_item_sync", True, True, True), ("prefetch_sync_end", False, True, True),
                                       ("plain_to_item_sync", True, True, "plain")):
        T = []
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        if pref is True:
            for batch in DevicePrefetcher((host for _ in range(6)), dev):
                step(batch, item, sync_end, T)
        elif pref == "plain":
            for _ in range(6):
                batch = [t.to(dev, non_blocking=True) for t in host]
                step(batch, item, sync_end, T)
        else:
            for _ in range(6):
                step(resident, item, sync_end, T)
        torch.cuda.synchronize()
        out[name] = {"ms_per_step": round(1e3 * (time.perf_counter() - w0) / 6, 2),
                     "phases_fwd_item_bwd_opt_sync": T[-2:]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
