"""World-size-2 gloo (CPU) checks of the data-parallel host logic: per-rank data seeds, max-over-ranks timing,
gradient averaging == DDP == mean of the per-rank gradients (computed with the tiny oracle model)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import painter_oracle as po
    from oracle.synth import synth_inputs, synth_state_dict
    from painter_b200 import dist_utils as du
    cfg = po.PainterConfig(img_size=(64, 32), embed_dim=64, num_heads=1, decoder_embed_dim=64)
    sd = {k: v.clone().requires_grad_(True) for k, v in synth_state_dict(cfg, 0).items()}
    seed = du.rank_seed(1234, rank)
    imgs, tgts, mask, valid = synth_inputs(cfg, 2, seed)
    loss, _, _ = po.forward(sd, cfg, imgs, tgts, mask, valid)
    loss.backward()
    local = {k: v.grad.clone() for k, v in sd.items()}
    du.average_gradients(list(sd.values()))
    # gather every rank's local gradient of one tensor and compare with the averaged one
    key = "blocks.3.attn.qkv.weight"
    gl = [torch.zeros_like(local[key]) for _ in range(world)]
    dist.all_gather(gl, local[key])
    want = sum(gl) / world
    err = (sd[key].grad - want).abs().max().item()
    tmax = du.max_over_ranks([float(rank + 1), 10.0 - rank])
    if rank == 0:
        torch.save(dict(err=err, tmax=tmax, seeds=[du.rank_seed(1234, r) for r in range(world)],
                        differ=bool((gl[0] - gl[1]).abs().max() > 0)), out)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gradient_average_and_timing(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["err"] < 1e-7
    assert r["tmax"] == [2.0, 10.0]
    assert len(set(r["seeds"])) == 2 and r["differ"]      # ranks really process different batches
