"""Checkpoint compatibility, CPU only: reference-format files written by the UNMODIFIED reference
(util/misc.py save_model / auto_load_model, torch.optim.AdamW state) load into painter_b200 modules + FusedAdamW and
vice versa; the --finetune key filtering of main_train.py:199-224 drops and reports the same keys."""
import types
from functools import partial

import pytest
import torch

from oracle import painter_oracle as po
from oracle import ref_loader
from oracle.synth import synth_state_dict

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")

CFG = po.PainterConfig(img_size=(64, 32), embed_dim=64, num_heads=1, decoder_embed_dim=64)


def _ours():
    from painter_b200 import models_painter
    m = models_painter.Painter(img_size=(64, 32), patch_size=16, embed_dim=64, depth=24, num_heads=1,
                               drop_path_rate=0.1, window_size=2, qkv_bias=True, mlp_ratio=4,
                               norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), window_block_indexes=[],
                               residual_block_indexes=[], use_rel_pos=True, decoder_embed_dim=64)
    return m


def _ref():
    mp = ref_loader.models_painter()
    return mp.Painter(img_size=(64, 32), patch_size=16, embed_dim=64, depth=24, num_heads=1, drop_path_rate=0.1,
                      window_size=2, qkv_bias=True, mlp_ratio=4, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6),
                      window_block_indexes=[], residual_block_indexes=[], use_rel_pos=True, decoder_embed_dim=64)


def _args(tmp_path, **kw):
    return types.SimpleNamespace(output_dir=str(tmp_path), resume="", auto_resume=True, start_epoch=0, **kw)


def _equal_sd(a, b):
    return a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)


def test_reference_checkpoint_resumes_into_painter_b200_and_back(tmp_path):
    from painter_b200 import checkpoint
    from painter_b200.optim import FusedAdamW
    from painter_b200.train_utils import param_groups_lrd
    misc = ref_loader.misc()
    lrd = ref_loader.lr_decay()
    ref = _ref()
    ref.load_state_dict(synth_state_dict(CFG, 3), strict=True)
    opt_ref = torch.optim.AdamW(lrd.param_groups_lrd(ref, 0.05, ref.no_weight_decay(), 0.8), lr=1e-3)
    for p in ref.parameters():                      # one AdamW step so that the optimizer has state to carry
        p.grad = torch.full_like(p, 1e-3)
    opt_ref.step()
    scaler = misc.NativeScalerWithGradNormCount()
    misc.save_model(_args(tmp_path), 4, ref, ref, opt_ref, scaler)          # unmodified reference writer
    # ---- resume into painter_b200 + FusedAdamW
    ours = _ours()
    opt = FusedAdamW(param_groups_lrd(ours, 0.05, ours.no_weight_decay(), 0.8), lr=1e-3)
    a = _args(tmp_path)
    assert checkpoint.auto_load_model(a, ours, ours, opt, scaler)
    assert a.resume.endswith("checkpoint-4.pth") and a.start_epoch == 5
    assert _equal_sd(ours.state_dict(), ref.state_dict())
    assert len(opt.param_groups) == len(opt_ref.param_groups)
    for go, gr in zip(opt.param_groups, opt_ref.param_groups):
        assert go["lr_scale"] == gr["lr_scale"] and go["weight_decay"] == gr["weight_decay"]
        for po_, pr in zip(go["params"], gr["params"]):
            assert torch.equal(opt.state[po_]["exp_avg"], opt_ref.state[pr]["exp_avg"])
            assert torch.equal(opt.state[po_]["exp_avg_sq"], opt_ref.state[pr]["exp_avg_sq"])
            assert int(opt.state[po_]["step"]) == int(opt_ref.state[pr]["step"]) == 1
    # ---- write from painter_b200, resume with the unmodified reference loader
    out2 = tmp_path / "b200"
    checkpoint.save_model(_args(out2), 7, ours, ours, opt, scaler)
    ref2 = _ref()
    opt2 = torch.optim.AdamW(lrd.param_groups_lrd(ref2, 0.05, ref2.no_weight_decay(), 0.8), lr=1e-3)
    a2 = _args(out2)
    # torch >= 2.6 unpickles with weights_only=True by default; the reference's checkpoints carry the argparse
    # namespace ('args'), which its own torch~=1.8 loaded freely: allow-list the namespace type for the reference loader
    with torch.serialization.safe_globals([types.SimpleNamespace]):
        misc.auto_load_model(a2, ref2, ref2, opt2, scaler)
    assert a2.start_epoch == 8 and _equal_sd(ref2.state_dict(), ours.state_dict())
    p0 = opt2.param_groups[0]["params"][0]
    assert torch.equal(opt2.state[p0]["exp_avg"], opt.state[opt.param_groups[0]["params"][0]]["exp_avg"])


def test_finetune_key_filtering_matches_main_train(tmp_path):
    """main_train.py:199-224 on an MAE-style checkpoint: decoder_embed.* / mask_token of other shapes are dropped,
    extra keys are reported as unexpected, everything else loads."""
    from painter_b200 import checkpoint
    sd = synth_state_dict(CFG, 5)
    mae = dict(sd)
    mae["decoder_embed.weight"] = torch.randn(32, 64)           # MAE: Linear(embed_dim, decoder_dim)
    mae["decoder_embed.bias"] = torch.randn(32)
    mae["mask_token"] = torch.randn(1, 1, 32)
    mae["cls_token"] = torch.randn(1, 1, 64)
    del mae["segment_token_x"]
    path = tmp_path / "mae.pth"
    torch.save({"model": mae}, path)
    ours = _ours()
    before = {k: v.clone() for k, v in ours.state_dict().items()}
    msg = checkpoint.load_pretrained(ours, str(path), verbose=False)
    # the reference's inline code, on the reference module
    ref = _ref()
    ref.load_state_dict(before, strict=True)
    ck = torch.load(path, map_location="cpu")["model"]
    state_dict = ref.state_dict()
    for k in ["decoder_embed.weight", "decoder_embed.bias", "mask_token"]:
        if k in ck and ck[k].shape != state_dict[k].shape:
            del ck[k]
    msg_ref = ref.load_state_dict(ck, strict=False)
    assert sorted(msg.missing_keys) == sorted(msg_ref.missing_keys)
    assert sorted(msg.unexpected_keys) == sorted(msg_ref.unexpected_keys) == ["cls_token"]
    assert _equal_sd(ours.state_dict(), ref.state_dict())
    assert torch.equal(ours.state_dict()["mask_token"], before["mask_token"])       # kept its init: shape mismatch
