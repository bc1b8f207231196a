"""Host-side helpers of the multi-GPU (data-parallel) path, backend-agnostic so they can be exercised with gloo on
CPU.  The reference shards by batch only (DDP, Painter/main_train.py:340); each rank owns a replica, draws its own
samples (sampler seed offset by rank, main_train.py:228-270) and gradients are averaged by all-reduce."""
import os

import torch
import torch.distributed as dist


def world_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def rank_seed(base_seed: int, rank: int) -> int:
    """Distinct, reproducible data seed per rank (weak scaling: every rank processes its own batch)."""
    return base_seed + 1000003 * rank


def max_over_ranks(values, device=None):
    """Element-wise MAX over ranks of a list of python floats (timings are reported as the slowest rank's)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if t.is_cuda:
            t = t.float()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def average_gradients(params):
    """Reference semantics of DDP's reducer: grad <- mean over ranks (flat bucket, one all-reduce)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= dist.get_world_size()
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
