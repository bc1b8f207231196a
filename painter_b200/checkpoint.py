"""Checkpoint compatibility with the reference: the same files, keys and call shapes.

  load_pretrained       Painter/main_train.py:199-224   --finetune: MAE / Painter init with key filtering, strict=False
  save_model            Painter/util/misc.py:296-313    checkpoint-{epoch}.pth = {model, optimizer, epoch, scaler, args}
  load_model            Painter/util/misc.py:316-331    --resume
  auto_load_model       Painter/util/misc.py:333-363    --auto_resume: newest checkpoint-*.pth in output_dir

State-dict keys and shapes of the painter_b200 modules are the reference's (tests/test_cabi.py), so
`painter_vit_large.pth`, `seggpt_vit_large.pth` and `mae_pretrain_vit_large.pth` load unchanged and files written here
load into the reference modules.  Optimizer state: optim.FusedAdamW keeps torch.optim.AdamW's state layout
(`step`, `exp_avg`, `exp_avg_sq` per parameter; same param_groups), so a run can be resumed across the two.
Loading replaces parameter storage in place (load_state_dict copies), which bumps the version counters the bf16
operand caches are keyed on; `engine.invalidate_weight_cache` is called anyway for loaders that write through `.data`.
"""
import glob
import os
from pathlib import Path

import torch

from .engine import invalidate_weight_cache


def _is_main_process():
    import torch.distributed as dist
    return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0


def load_pretrained(model, checkpoint, last_norm_instance=False, verbose=True):
    """main_train.py:199-224: `checkpoint` is a path or an already loaded dict with a 'model' entry.  Keys whose
    shapes differ from the model's (decoder_embed.*, mask_token; norm.* with last_norm_instance) are dropped, the rest
    is loaded with strict=False.  Returns load_state_dict's (missing_keys, unexpected_keys) message."""
    if not isinstance(checkpoint, dict):
        checkpoint = torch.load(checkpoint, map_location="cpu", weights_only=False)
    checkpoint_model = dict(checkpoint["model"])
    state_dict = model.state_dict()
    rm_key_list = ["decoder_embed.weight", "decoder_embed.bias", "mask_token"]
    if last_norm_instance:
        rm_key_list.extend(["norm.weight", "norm.bias"])
    for k in rm_key_list:
        if k in checkpoint_model and k in state_dict and checkpoint_model[k].shape != state_dict[k].shape:
            if verbose:
                print(f"Removing key {k} from pretrained checkpoint")
            del checkpoint_model[k]
    msg = model.load_state_dict(checkpoint_model, strict=False)
    invalidate_weight_cache(model)
    if verbose:
        print(msg)
    return msg


def save_model(args, epoch, model, model_without_ddp, optimizer, loss_scaler):
    """misc.save_model (torch.amp branch): rank 0 writes output_dir/checkpoint-{epoch}.pth."""
    output_dir = Path(args.output_dir)
    path = output_dir / ("checkpoint-%s.pth" % str(epoch))
    to_save = {
        "model": model_without_ddp.state_dict(),
        "optimizer": optimizer.state_dict(),
        "epoch": epoch,
        "scaler": loss_scaler.state_dict() if loss_scaler is not None else {},
        "args": args,
    }
    if _is_main_process():
        os.makedirs(output_dir, exist_ok=True)
        torch.save(to_save, path)
    return path


def load_model(args, model_without_ddp, optimizer, loss_scaler):
    """misc.load_model: resume model (+ optimizer, epoch, scaler unless args.eval) from args.resume."""
    if not getattr(args, "resume", ""):
        return False
    checkpoint = torch.load(args.resume, map_location="cpu", weights_only=False)
    model_without_ddp.load_state_dict(checkpoint["model"])
    invalidate_weight_cache(model_without_ddp)
    print("Resume checkpoint %s" % args.resume)
    if "optimizer" in checkpoint and "epoch" in checkpoint and not (hasattr(args, "eval") and args.eval):
        optimizer.load_state_dict(checkpoint["optimizer"])
        args.start_epoch = checkpoint["epoch"] + 1
        if "scaler" in checkpoint and loss_scaler is not None and checkpoint["scaler"]:
            loss_scaler.load_state_dict(checkpoint["scaler"])
        print("With optim & sched!")
    return True


def auto_load_model(args, model, model_without_ddp, optimizer, loss_scaler):
    """misc.auto_load_model (torch.amp branch): with args.auto_resume and no explicit args.resume, pick the
    checkpoint-<N>.pth with the largest N in args.output_dir, then resume from it."""
    output_dir = Path(args.output_dir)
    if getattr(args, "auto_resume", False) and len(getattr(args, "resume", "")) == 0:
        latest = -1
        for ckpt in glob.glob(os.path.join(output_dir, "checkpoint-*.pth")):
            t = ckpt.split("-")[-1].split(".")[0]
            if t.isdigit():
                latest = max(int(t), latest)
        if latest >= 0:
            args.resume = os.path.join(output_dir, "checkpoint-%d.pth" % latest)
        print("Auto resume checkpoint: %s" % args.resume)
    return load_model(args, model_without_ddp, optimizer, loss_scaler)
