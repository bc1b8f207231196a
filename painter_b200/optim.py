"""Optimizer step on the device library.

`FusedAdamW` mirrors `torch.optim.AdamW` as the reference builds it (`Painter/main_train.py:344-348`:
`torch.optim.AdamW(param_groups, lr=args.lr, betas=(0.9, 0.999))` over `lr_decay.param_groups_lrd` groups, each with
its own `lr` (after `lr_sched.adjust_learning_rate` multiplies by `lr_scale`) and `weight_decay`): decoupled weight
decay, bias-corrected first/second moments, fp32 state.  One kernel launch updates every parameter
(`pk_adamw_step`); unscaling by the AMP loss scale and gradient clipping ride along as a device-side multiplier,
the way `util/misc.py:252-278` sequences unscale -> clip -> step, without extra passes over the gradients.
The same pass refreshes the bf16 operand copies of the GEMM weights / rel-pos tables (engine.bf16_weight) and, when
the gradients live in the module's flat arena (painter_b200/arena.py), clears them for the next backward - so a
training step launches no cast and no fill kernels.
`global_grad_norm` is `misc.get_grad_norm_` / `clip_grad_norm_`'s 2-norm in one launch (`pk_grad_sumsq`).
"""
import ctypes

import numpy as np
import torch

from ._lib import check, lib

_REC = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("w16", "<u8"), ("n", "<i8"),
                 ("lr", "<f4"), ("wd", "<f4")])
assert _REC.itemsize == 56


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_CHUNK_CACHE = {}


def _chunk_table(numels, device):
    """int32 [nchunks, 2] = (tensor index, chunk index), one row per CUDA block; cached per list of sizes."""
    key = (tuple(numels), str(device))
    tab = _CHUNK_CACHE.get(key)
    if tab is None:
        chunk = lib().pk_opt_chunk_elems()
        counts = np.array([(n + chunk - 1) // chunk for n in numels], dtype=np.int64)
        ti = np.repeat(np.arange(len(numels), dtype=np.int32), counts)
        ci = (np.arange(counts.sum(), dtype=np.int64) - np.repeat(np.cumsum(counts) - counts, counts)).astype(np.int32)
        tab = torch.from_numpy(np.stack([ti, ci], axis=1).copy()).to(device)
        if len(_CHUNK_CACHE) > 16:
            _CHUNK_CACHE.clear()
        _CHUNK_CACHE[key] = tab
    return tab


def _tensor_table(entries, device, cache=None):
    """entries: list of (p, g, m, v, w16, lr, wd) (m / v / w16 may be None) -> device uint8 table of PkOptTensor.
    `cache` (a dict) keeps the device copy while the host records are unchanged (static pointers, lr and wd)."""
    rec = np.zeros(len(entries), dtype=_REC)
    for i, (p, g, m, v, w16, lr, wd) in enumerate(entries):
        rec[i] = (p.data_ptr(), g.data_ptr(), m.data_ptr() if m is not None else 0,
                  v.data_ptr() if v is not None else 0, w16.data_ptr() if w16 is not None else 0, p.numel(), lr, wd)
    if cache is not None:
        raw = rec.tobytes()
        dev_t = cache.get("dev")
        if cache.get("raw") == raw and dev_t.device == torch.device(device):
            return dev_t
        cache["raw"] = raw
        host = torch.from_numpy(rec.view(np.uint8).copy())
        if dev_t is not None and dev_t.device == torch.device(device) and dev_t.numel() == host.numel():
            dev_t.copy_(host)          # same address: a captured training step keeps reading this table
        else:
            cache["dev"] = host.to(device)
        return cache["dev"]
    return torch.from_numpy(rec.view(np.uint8)).to(device)


def _check_fp32(t, what):
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise RuntimeError(f"painter_b200.optim: {what} must be a contiguous fp32 CUDA tensor")


def global_grad_norm(parameters):
    """2-norm of all gradients as a 0-dim device tensor (no host sync): `misc.get_grad_norm_(params, 2.0)`."""
    parameters = list(parameters)
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return torch.zeros((), device="cuda")
    dev = grads[0].device
    ref = getattr(parameters[0], "_pk_arena", None)
    arena = ref() if ref is not None else None
    if arena is not None and len(grads) == len(arena.params) and arena.grads_in_arena():
        grads = [arena.slab]      # every gradient lives in the flat arena (padding is zero): one contiguous pass
    for g in grads:
        _check_fp32(g, "gradient")
    table = _tensor_table([(g, g, None, None, None, 0.0, 0.0) for g in grads], dev)
    chunks = _chunk_table([g.numel() for g in grads], dev)
    out = torch.zeros(1, dtype=torch.float32, device=dev)
    check(lib().pk_grad_sumsq(_ptr(table), _ptr(chunks), chunks.shape[0], _ptr(out), _stream()), "pk_grad_sumsq")
    return out.sqrt_()[0]


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, arena=None):
        """arena: the module's painter_b200.arena.GradArena (optional; found through the parameters otherwise)."""
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._arena = arena
        self._tables = {}

    def _find_arena(self, params):
        if self._arena is not None:
            return self._arena
        for p in params:
            ref = getattr(p, "_pk_arena", None)
            if ref is not None:
                return ref()
        return None

    # ---- CUDA-graph support (painter_b200.train_utils.GraphedTrainStep) -------------------------------------------
    # A captured step cannot take the step count, the bias corrections or the learning rates as launch arguments:
    # graph_prepare() does the host half of step() - state creation, step counters, the per-tensor (pointer, lr, wd)
    # table rewritten IN PLACE, the two bias-correction scalars copied into a fixed device buffer - and
    # graph_launch() issues the one kernel that reads them; the graph captures graph_launch(), every replay is
    # preceded by graph_prepare().
    @torch.no_grad()
    def graph_prepare(self):
        entries, key = [], None
        arena = self._find_arena([p for g in self.param_groups for p in g["params"]])
        for group in self.param_groups:
            for p in group["params"]:
                g = p.grad
                if g is None and arena is not None and arena.owns(p):
                    off, n, shape = arena.offsets[id(p)]   # zero_grad(set_to_none=True) between replays: the graph
                    g = arena.slab[off:off + n].view(shape)  # still writes (and the kernel still clears) this slot
                if g is None:
                    continue
                st = self.state[p]
                if "exp_avg" not in st:
                    _check_fp32(p, "parameter")
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                elif torch.is_tensor(st["step"]):
                    st["step"] = int(st["step"].item())
                st["step"] += 1
                k = (group["betas"][0], group["betas"][1], group["eps"], st["step"])
                if key is None:
                    key = k
                elif k != key:
                    raise RuntimeError("FusedAdamW.graph_prepare: all parameters must share betas, eps and step count")
                _check_fp32(g, "gradient")
                ent = getattr(p, "_pk_bf16", None)
                w16 = ent[2] if ent is not None and ent[1] == p.data_ptr() and ent[2].device == p.device else None
                entries.append((p, g, st["exp_avg"], st["exp_avg_sq"], w16, group["lr"], group["weight_decay"]))
        if not entries:
            raise RuntimeError("FusedAdamW.graph_prepare: no parameter has a gradient (run one eager step first)")
        ids = {id(e[0]) for e in entries}
        zero = 0
        if arena is not None and all(id(p) in ids for p in arena.params):
            lo = arena.slab.data_ptr()
            hi = lo + arena.slab.numel() * 4
            zero = int(all(lo <= e[1].data_ptr() < hi for e in entries))
        dev = entries[0][0].device
        b1, b2, eps, step = key
        gp = self._graph = getattr(self, "_graph", None) or {}
        if "bc_host" not in gp:
            gp["bc_host"] = torch.empty(2, dtype=torch.float32).pin_memory()
            gp["bc_dev"] = torch.empty(2, dtype=torch.float32, device=dev)
        gp["bc_host"][0] = 1.0 / (1.0 - b1 ** step)
        gp["bc_host"][1] = 1.0 / (1.0 - b2 ** step) ** 0.5
        gp["bc_dev"].copy_(gp["bc_host"], non_blocking=True)
        gp.update(chunks=_chunk_table([e[0].numel() for e in entries], dev),
                  table=_tensor_table(entries, dev, self._tables.setdefault("graph", {})),
                  b1=b1, b2=b2, eps=eps, zero=zero, arena=arena, params=[e[0] for e in entries])
        return gp

    @torch.no_grad()
    def graph_launch(self):
        gp = self._graph
        check(lib().pk_adamw_step_graph(_ptr(gp["table"]), _ptr(gp["chunks"]), gp["chunks"].shape[0],
                                        ctypes.c_double(gp["b1"]), ctypes.c_double(gp["b2"]),
                                        ctypes.c_double(gp["eps"]), _ptr(gp["bc_dev"]), None, ctypes.c_float(0.0),
                                        gp["zero"], None, _stream()), "pk_adamw_step_graph")
        if gp["zero"]:
            gp["arena"].mark_clean()

    @torch.no_grad()
    def step(self, closure=None, grad_scale=None, grad_scale_cap=0.0, found_inf=None):
        """grad_scale: optional 1-element fp32 device tensor every gradient is multiplied by (1 / loss scale, or the
        clip coefficient max_norm / (norm + 1e-6)); grad_scale_cap > 0 clamps it from above (clip: cap = 1).
        found_inf: optional 1-element fp32 device tensor; non-zero turns the step into a no-op (GradScaler.step)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        # groups may differ in betas / eps / step count (they do not in the reference): one launch per such class
        classes = {}
        stepped = []
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if "exp_avg" not in st:
                    _check_fp32(p, "parameter")
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                elif torch.is_tensor(st["step"]):   # state loaded from a torch.optim.AdamW checkpoint (misc.load_model)
                    st["step"] = int(st["step"].item())
                st["step"] += 1
                g = p.grad
                if not g.is_contiguous():
                    g = g.contiguous()
                _check_fp32(g, "gradient")
                ent = getattr(p, "_pk_bf16", None)   # bf16 operand copy kept by engine.bf16_weight / bf16_table
                w16 = ent[2] if ent is not None and ent[1] == p.data_ptr() and ent[2].device == p.device else None
                key = (b1, b2, group["eps"], st["step"])
                classes.setdefault(key, []).append((p, g, st["exp_avg"], st["exp_avg_sq"], w16, group["lr"],
                                                    group["weight_decay"]))
                stepped.append(p)
        if not stepped:
            return loss
        arena = self._find_arena(stepped)
        zero = 0
        if arena is not None and found_inf is None:
            # every slot of the arena is consumed by this step -> leave it zeroed for the next backward
            ids = {id(p) for p in stepped}
            if all(id(p) in ids for p in arena.params) and arena.grads_in_arena():
                zero = 1
        for ci, ((b1, b2, eps, step), entries) in enumerate(classes.items()):
            dev = entries[0][0].device
            chunks = _chunk_table([e[0].numel() for e in entries], dev)
            table = _tensor_table(entries, dev, self._tables.setdefault(ci, {}))
            gs = None
            if grad_scale is not None:
                gs = grad_scale.reshape(-1)[:1]
                _check_fp32(gs, "grad_scale")
            fi = None
            if found_inf is not None:
                fi = found_inf.reshape(-1)[:1]
                _check_fp32(fi, "found_inf")
            check(lib().pk_adamw_step(_ptr(table), _ptr(chunks), chunks.shape[0], ctypes.c_double(b1),
                                      ctypes.c_double(b2), ctypes.c_double(eps), step, _ptr(gs),
                                      ctypes.c_float(grad_scale_cap), zero, _ptr(fi), _stream()), "pk_adamw_step")
            # the kernel wrote through raw pointers: tell autograd (and the (version, storage) key of the bf16 operand
            # cache, whose copy the kernel has just refreshed) that the parameters changed, as an in-place op would
            for e in entries:
                p = e[0]
                torch.autograd.graph.increment_version(p)
                if e[4] is not None:
                    p._pk_bf16 = (p._version, p.data_ptr(), e[4])
        if zero:
            arena.mark_clean()
        return loss
