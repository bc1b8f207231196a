"""TEST INFRASTRUCTURE: import the UNMODIFIED reference modules from /root/reference through the shim
packages in oracle/shims (timm 0.3.2 symbols, import-only detectron2/fairscale/fvcore, torch._six).

Only usable where /root/reference exists (the authoring container).  Nothing in the product path, the
-m gpu tests, smoke() or bench.py imports this file.
"""
import importlib
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("PAINTER_REFERENCE", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "Painter"))


def _install_shims():
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    if "torch._six" not in sys.modules:  # Painter/util/misc.py:22
        import math
        m = types.ModuleType("torch._six")
        m.inf = math.inf
        sys.modules["torch._six"] = m


def _load(name, path, pkg_dir):
    """Load `path` as module `name` with pkg_dir temporarily first on sys.path (for `util.*`)."""
    _install_shims()
    saved = {k: v for k, v in sys.modules.items() if k == "util" or k.startswith("util.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, pkg_dir)
    try:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(pkg_dir)
        for k in [k for k in sys.modules if k == "util" or k.startswith("util.")]:
            sys.modules["_ref_" + name + "." + k] = sys.modules.pop(k)
        sys.modules.update(saved)
    return mod


_cache = {}


def models_painter():
    if "p" not in _cache:
        d = os.path.join(REF_ROOT, "Painter")
        _cache["p"] = _load("ref_models_painter", os.path.join(d, "models_painter.py"), d)
    return _cache["p"]


def models_seggpt():
    if "s" not in _cache:
        d = os.path.join(REF_ROOT, "SegGPT", "SegGPT_inference")
        _cache["s"] = _load("ref_models_seggpt", os.path.join(d, "models_seggpt.py"), d)
    return _cache["s"]


def masking_generator():
    if "m" not in _cache:
        d = os.path.join(REF_ROOT, "Painter")
        _cache["m"] = _load("ref_masking_generator", os.path.join(d, "util", "masking_generator.py"), d)
    return _cache["m"]
