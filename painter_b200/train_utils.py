"""The optimizer recipe around the hot path, restated for painter_b200 modules:

  param_groups_lrd       Painter/util/lr_decay.py:15-76   layer-wise lr decay groups (BEiT rule), same group order
  adjust_learning_rate   Painter/util/lr_sched.py:9-21    half-cycle cosine after linear warm-up, per iteration
  FusedStep              Painter/util/misc.py:252-278      NativeScalerWithGradNormCount.__call__ (unscale -> clip ->
                                                           step) for FusedAdamW: the global gradient norm is one launch
                                                           over the flat arena, the clip coefficient stays on the device
                                                           and rides into the AdamW kernel - no host sync in the step
"""
import math

import torch

from .optim import FusedAdamW, global_grad_norm


def layer_id_for_vit(name, num_layers):
    """lr_decay.get_layer_id_for_vit (:64-76)."""
    if name in ("cls_token", "pos_embed"):
        return 0
    if name.startswith("patch_embed"):
        return 0
    if name.startswith("blocks"):
        return int(name.split(".")[1]) + 1
    return num_layers


def param_groups_lrd(model, weight_decay=0.05, no_weight_decay_list=(), layer_decay=0.75):
    """Groups {lr_scale, weight_decay, params} exactly as the reference builds them (same keys, same order)."""
    num_layers = len(model.blocks) + 1
    scales = [layer_decay ** (num_layers - i) for i in range(num_layers + 1)]
    groups = {}
    for n, p in model.named_parameters():
        if not p.requires_grad:
            continue
        if p.ndim == 1 or n in no_weight_decay_list:
            tag, wd = "no_decay", 0.0
        else:
            tag, wd = "decay", weight_decay
        lid = layer_id_for_vit(n, num_layers)
        key = "layer_%d_%s" % (lid, tag)
        if key not in groups:
            groups[key] = {"lr_scale": scales[lid], "weight_decay": wd, "params": []}
        groups[key]["params"].append(p)
    return list(groups.values())


def adjust_learning_rate(optimizer, epoch, lr, min_lr=0.0, warmup_epochs=0, epochs=1):
    """lr_sched.adjust_learning_rate with explicit arguments instead of the argparse namespace."""
    if epoch < warmup_epochs:
        cur = lr * epoch / warmup_epochs
    else:
        cur = min_lr + (lr - min_lr) * 0.5 * (1.0 + math.cos(math.pi * (epoch - warmup_epochs) /
                                                              (epochs - warmup_epochs)))
    for g in optimizer.param_groups:
        g["lr"] = cur * g["lr_scale"] if "lr_scale" in g else cur
    return cur


class FusedStep:
    """loss -> backward -> (clip) -> FusedAdamW.step, the sequence of misc.NativeScalerWithGradNormCount.__call__
    without a loss scale (the module computes in bf16 operands / fp32 accumulate whatever the autocast dtype, so
    gradients cannot overflow the way fp16 ones do).  Returns the gradient norm as a 0-dim DEVICE tensor."""

    def __init__(self, optimizer: FusedAdamW):
        self.opt = optimizer

    def __call__(self, loss, clip_grad=None, parameters=None, update_grad=True):
        loss.backward()
        if not update_grad:
            return None
        params = list(parameters) if parameters is not None else [p for g in self.opt.param_groups for p in g["params"]]
        norm = global_grad_norm(params)
        if clip_grad is not None:
            coef = (clip_grad / (norm + 1e-6)).reshape(1).float()      # clip_grad_norm_: min(1, max_norm / (norm + 1e-6))
            self.opt.step(grad_scale=coef, grad_scale_cap=1.0)
        else:
            self.opt.step()
        return norm
