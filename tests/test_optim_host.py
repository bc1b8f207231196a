"""Host-side tables of the fused optimizer (no GPU): chunk table and tensor records must match the C ABI layout
(`PkOptTensor`, include/painter_b200.h) and tile every parameter exactly once."""
import ctypes

import numpy as np
import torch

from painter_b200 import _lib, optim


class PkOptTensor(ctypes.Structure):   # as declared in include/painter_b200.h
    _fields_ = [("p", ctypes.c_void_p), ("g", ctypes.c_void_p), ("m", ctypes.c_void_p), ("v", ctypes.c_void_p),
                ("w16", ctypes.c_void_p), ("n", ctypes.c_longlong), ("lr", ctypes.c_float), ("wd", ctypes.c_float)]


def test_record_layout_matches_the_c_struct():
    assert optim._REC.itemsize == ctypes.sizeof(PkOptTensor) == 56
    for name in ("p", "g", "m", "v", "w16", "n", "lr", "wd"):
        assert optim._REC.fields[name][1] == getattr(PkOptTensor, name).offset


def test_chunk_table_tiles_every_tensor_once():
    chunk = _lib.lib().pk_opt_chunk_elems()
    assert chunk > 0 and chunk % 4 == 0
    numels = [3, chunk, chunk + 1, 7104, 5 * chunk - 2, 1]
    tab = optim._chunk_table(numels, torch.device("cpu")).numpy()
    want = [(t, c) for t, n in enumerate(numels) for c in range((n + chunk - 1) // chunk)]
    assert [tuple(r) for r in tab.tolist()] == want
    covered = np.zeros(len(numels), dtype=np.int64)
    for t, c in tab:
        covered[t] += min(chunk, numels[t] - c * chunk)
    assert covered.tolist() == numels


def test_tensor_table_carries_pointers_sizes_and_group_hyperparameters():
    p, g = torch.zeros(10), torch.ones(10)
    m, v = torch.zeros(10), torch.zeros(10)
    w16 = torch.zeros(10, dtype=torch.bfloat16)
    cache = {}
    tab = optim._tensor_table([(p, g, m, v, w16, 3e-3, 0.05)], torch.device("cpu"), cache)
    assert optim._tensor_table([(p, g, m, v, w16, 3e-3, 0.05)], torch.device("cpu"), cache) is tab   # cached
    raw = tab.numpy().tobytes()
    rec = PkOptTensor.from_buffer_copy(raw)
    assert (rec.p, rec.g, rec.m, rec.v, rec.n) == (p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), 10)
    assert rec.w16 == w16.data_ptr()
    assert abs(rec.lr - 3e-3) < 1e-9 and abs(rec.wd - 0.05) < 1e-8
    # a changed learning rate rewrites the SAME device table (a captured training step keeps its address)
    ptr = tab.data_ptr()
    tab2 = optim._tensor_table([(p, g, m, v, w16, 4e-3, 0.05)], torch.device("cpu"), cache)
    assert tab2 is tab and tab2.data_ptr() == ptr
    assert abs(PkOptTensor.from_buffer_copy(tab2.numpy().tobytes()).lr - 4e-3) < 1e-9
    # a different parameter list does not fit the old table: new allocation
    tab3 = optim._tensor_table([(p, g, m, v, w16, 4e-3, 0.05), (p, g, m, v, None, 1e-3, 0.0)], torch.device("cpu"), cache)
    assert tab3 is not tab and tab3.numel() == 2 * tab.numel()
