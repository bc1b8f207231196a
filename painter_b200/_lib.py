"""ctypes binding of libpainter_b200.so (the C ABI declared in include/painter_b200.h).

There is no fallback: if the shared object is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PK_LIB") or os.path.join(_HERE, "libpainter_b200.so")   # PK_LIB: a tuning build

_lib = None


class PkEpilogue(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int),
        ("out", ctypes.c_void_p),
        ("out2", ctypes.c_void_p),
        ("bias", ctypes.c_void_p),
        ("aux", ctypes.c_void_p),
        ("rowscale", ctypes.c_void_p),
        ("ldc", ctypes.c_int),
        ("ld_aux", ctypes.c_int),
        ("rows_per_group", ctypes.c_int),
        ("accumulate", ctypes.c_int),
        ("alpha", ctypes.c_float),
        ("ps_h", ctypes.c_int),
        ("ps_w", ctypes.c_int),
        ("ps_p", ctypes.c_int),
        ("ps_c", ctypes.c_int),
    ]


EPI_BF16, EPI_F32, EPI_GELU, EPI_RESID, EPI_DGELU, EPI_PIXSHUF = range(6)


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m painter_b200.build` "
                "(painter_b200 has no CPU or PyTorch fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.pk_last_error.restype = ctypes.c_char_p
        _lib.pk_launch_count.restype = ctypes.c_longlong
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"painter_b200 {what} failed (rc={rc}): {lib().pk_last_error().decode()}")


def launch_count():
    return int(lib().pk_launch_count())
