"""CPU-side checks of the C-ABI boundary: the library builds, loads, and exports every symbol the header
declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "painter_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pk_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from painter_b200 import build, _lib
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_library_reports_version_and_errors():
    from painter_b200 import _lib
    lib = _lib.lib()
    assert lib.pk_version() >= 100
    e = _lib.PkEpilogue()
    # null pointers are rejected on the host before any launch
    rc = lib.pk_gemm_bf16(None, None, 128, 128, 64, 64, 64, 0, 0, ctypes.byref(e), None)
    assert rc != 0 and b"null" in lib.pk_last_error()


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from painter_b200 import ops
    a = torch.zeros(128, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.gemm(a, a)


def test_state_dict_keys_match_reference_layout():
    from oracle.painter_oracle import PainterConfig
    from oracle.synth import param_shapes
    from painter_b200 import models_painter, models_seggpt
    m = models_painter.painter_vit_large_patch16_input896x448_win_dec64_8glb_sl1()
    ref = param_shapes(PainterConfig())
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys()) or set(sd.keys()) == set(ref.keys())
    assert all(tuple(sd[k].shape) == tuple(ref[k]) for k in ref)
    assert sum(p.numel() for p in m.parameters()) == 370721155
    assert all(b.window_size == 0 for b in m.blocks)            # section 0.1: stock factory => 24 global blocks
    assert m.no_weight_decay() == {"pos_embed", "cls_token"}
    assert m.patch_embed.num_patches == 1568 and m.patch_size == 16
    s = models_seggpt.seggpt_vit_large_patch16_input896x448()
    assert set(s.state_dict().keys()) == set(param_shapes(PainterConfig(seggpt=True)).keys())
