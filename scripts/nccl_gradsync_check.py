#!/usr/bin/env python
"""2-rank NCCL check of painter_b200.dist_utils.GradSync on the CUDA module: synchronised gradients == mean of the
ranks' local gradients.  Launched by tests/test_gpu_model.py with torch.distributed.run."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import painter_oracle as po  # noqa: E402
from oracle.synth import synth_inputs  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from _common import build_model
    from painter_b200.dist_utils import GradSync, rank_seed
    cfg = po.PainterConfig(img_size=(128, 64), embed_dim=128, num_heads=2, decoder_embed_dim=64)
    model, _ = build_model(cfg, 0)
    model.eval()
    args = [t.cuda() for t in synth_inputs(cfg, 2, rank_seed(11, rank) % 9973, valid_kind="mixed")]
    loss, _, _ = model(*args)
    loss.backward()
    local_g = [p.grad.clone() for p in model.parameters()]
    for p in model.parameters():
        p.grad = None
    GradSync(model, bucket_mb=1)
    loss, _, _ = model(*args)
    loss.backward()
    torch.cuda.synchronize()
    worst = 0.0
    for p, g in zip(model.parameters(), local_g):
        gs = [torch.zeros_like(g) for _ in range(world)]
        dist.all_gather(gs, g)
        want = sum(gs) / world
        worst = max(worst, ((p.grad - want).abs().max() / want.abs().max().clamp_min(1e-12)).item())
    if rank == 0:
        json.dump({"max_rel_err": worst}, open(sys.argv[1], "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
