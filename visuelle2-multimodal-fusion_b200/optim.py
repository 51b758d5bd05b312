"""Adafactor as one multi-tensor CUDA step (csrc/adafactor.cu) behind the optimizer interface the reference's
``configure_optimizers`` returns (/root/reference/models/CrossAttnRNN210.py:229-230:
``Adafactor(self.parameters(), scale_parameter=True, relative_step=True, warmup_init=True, lr=None)`` from fairseq;
the same algorithm is transformers.optimization.Adafactor, the checker of tests/test_gpu_adafactor.py).

State keys and shapes are those of the fairseq / transformers implementations (``step``, ``RMS``,
``exp_avg_sq_row``, ``exp_avg_sq_col`` / ``exp_avg_sq``), so optimizer state_dicts are interchangeable.  There is no
CPU path: parameters and gradients must be fp32 CUDA tensors.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib
from ._lib import check, stream

_SMALL_DIM = 4        # AF_DIM of csrc/adafactor.cu
_VEC_CHUNK = 1024
_THREADS = 256
_COLSTRIP = 32      # AF_COLSTRIP


def _strided_small(p):
    """A dense 4-D small-matrix tensor in a non-row-major layout (channels_last convolution weights)."""
    return p.dim() == 4 and _kind(tuple(p.shape)) == 1 and p.is_contiguous(memory_format=torch.channels_last)


def _kind(shape):
    if len(shape) < 2:
        return 0
    return 1 if (shape[-1] <= _SMALL_DIM and shape[-2] <= _SMALL_DIM) else 2


class Adafactor(torch.optim.Optimizer):
    """Drop-in for ``fairseq.optim.adafactor.Adafactor`` / ``transformers.optimization.Adafactor`` restricted to
    what the reference uses: ``beta1=None`` (no first moment) and ``weight_decay=0``."""

    def __init__(self, params, lr=None, eps=(1e-30, 1e-3), clip_threshold=1.0, decay_rate=-0.8, beta1=None,
                 weight_decay=0.0, scale_parameter=True, relative_step=True, warmup_init=False):
        if lr is not None and relative_step:
            raise ValueError("Cannot combine manual `lr` and `relative_step=True` options")
        if warmup_init and not relative_step:
            raise ValueError("`warmup_init=True` requires `relative_step=True`")
        if beta1 is not None or weight_decay != 0.0:
            raise NotImplementedError("the fused Adafactor covers beta1=None, weight_decay=0 (the reference's setting)")
        defaults = dict(lr=lr, eps=eps, clip_threshold=clip_threshold, decay_rate=decay_rate, beta1=beta1,
                        weight_decay=weight_decay, scale_parameter=scale_parameter, relative_step=relative_step,
                        warmup_init=warmup_init)
        super().__init__(params, defaults)
        self._plans = {}          # group index -> plan (rebuilt when the set of parameters with gradients changes)

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._plans = {}

    # ------------------------------------------------------------------------------------------ plan
    def _init_state(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = 0
            shape = tuple(p.shape)
            if len(shape) >= 2:
                st["exp_avg_sq_row"] = torch.zeros(shape[:-1], device=p.device, dtype=torch.float32)
                st["exp_avg_sq_col"] = torch.zeros(shape[:-2] + shape[-1:], device=p.device, dtype=torch.float32)
            else:
                st["exp_avg_sq"] = torch.zeros(shape, device=p.device, dtype=torch.float32)
            st["RMS"] = 0
        return st

    def _build(self, group, active):
        dev = active[0].device
        n = len(active)
        rms = torch.zeros(n, device=dev, dtype=torch.float32)
        descs = (_lib.AfDesc * n)()
        vec, small, rows, cols = [], ([], [], []), [], []
        for i, p in enumerate(active):
            shape = tuple(p.shape)
            kind = _kind(shape)
            if not (p.is_cuda and p.dtype == torch.float32 and (p.is_contiguous() or _strided_small(p))):
                raise RuntimeError("fused Adafactor needs fp32 CUDA parameters, contiguous or channels_last 4-D "
                                   "convolution weights (no CPU fallback)")
            st = self._init_state(p)
            d = descs[i]
            d.p, d.rms, d.numel, d.kind = p.data_ptr(), rms[i:i + 1].data_ptr(), p.numel(), kind
            for k in ("exp_avg_sq_row", "exp_avg_sq_col", "exp_avg_sq"):
                if k in st:
                    t = st[k]
                    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                        st[k] = t = t.to(device=dev, dtype=torch.float32).contiguous()
            st["RMS"] = rms[i]
            if kind == 0:
                d.sq, d.nmat, d.R, d.C = st["exp_avg_sq"].data_ptr(), 0, 0, 0
                e0 = np.arange(0, p.numel(), _VEC_CHUNK, dtype=np.int64)
                cnt = np.minimum(_VEC_CHUNK, p.numel() - e0)
                vec.append(np.stack([np.full_like(e0, i), e0, cnt, np.zeros_like(e0)], 1))
            else:
                R, C = shape[-2], shape[-1]
                nmat = p.numel() // (R * C)
                d.row, d.col = st["exp_avg_sq_row"].data_ptr(), st["exp_avg_sq_col"].data_ptr()
                d.nmat, d.R, d.C = nmat, R, C
                if kind == 1:
                    if p.is_contiguous():
                        d.inner, d.sO, d.sI, d.sR, d.sC = 1, R * C, 0, C, 1
                    else:                       # channels_last [O,I,R,C]
                        d.inner = shape[1]
                        d.sO, d.sI, d.sR, d.sC = p.stride()
                    m0 = np.arange(0, nmat, _THREADS, dtype=np.int64)
                    cnt = np.minimum(_THREADS, nmat - m0)
                    cls = 0 if (R, C) == (1, 1) else (1 if (R, C) == (3, 3) else 2)
                    small[cls].append(np.stack([np.full_like(m0, i), m0, cnt, np.zeros_like(m0)], 1))
                else:
                    mm, rr = np.meshgrid(np.arange(nmat, dtype=np.int64), np.arange(R, dtype=np.int64), indexing="ij")
                    rows.append(np.stack([np.full(mm.size, i, np.int64), mm.ravel(), rr.ravel(),
                                          np.zeros(mm.size, np.int64)], 1))
                    mm, cc = np.meshgrid(np.arange(nmat, dtype=np.int64), np.arange(0, C, _COLSTRIP, dtype=np.int64),
                                         indexing="ij")
                    cols.append(np.stack([np.full(mm.size, i, np.int64), mm.ravel(), cc.ravel(),
                                          np.zeros(mm.size, np.int64)], 1))

        def table(parts):
            if not parts:
                return None, 0
            a = np.concatenate(parts, 0).astype(np.int32)
            return torch.from_numpy(a).to(dev).contiguous(), a.shape[0]

        plan = _lib.AfPlan()
        keep = {"rms": rms}
        raw = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).to(dev)
        keep["descs"] = raw
        # gradient tensors are re-allocated every step (zero_grad(set_to_none=True)): their pointers travel in a small
        # pinned table, four of them in rotation so that an in-flight copy is never overwritten by the host
        keep["grads_host"] = [torch.zeros(n, dtype=torch.int64).pin_memory() for _ in range(4)]
        keep["grads_ev"] = [None] * 4
        keep["turn"] = 0
        keep["grads"] = torch.zeros(n, dtype=torch.int64, device=dev)
        keep["acc"] = torch.zeros(n, 2, dtype=torch.float64, device=dev)
        plan.descs, plan.grads, plan.acc = raw.data_ptr(), keep["grads"].data_ptr(), keep["acc"].data_ptr()
        for name, parts in (("vec", vec), ("row", rows), ("col", cols)):
            t, cnt = table(parts)
            keep[name] = t
            setattr(plan, name + "_units", t.data_ptr() if t is not None else None)
            setattr(plan, {"vec": "n_vec", "row": "n_rows", "col": "n_cols"}[name], cnt)
        for k in range(3):
            t, cnt = table(small[k])
            keep["small%d" % k] = t
            plan.small_units[k] = t.data_ptr() if t is not None else None
            plan.n_small[k] = cnt
        plan.n_desc = n
        plan.eps1, plan.eps2 = float(group["eps"][0]), float(group["eps"][1])
        plan.clip_threshold = float(group["clip_threshold"])
        plan.scale_parameter = 1 if group["scale_parameter"] else 0
        return dict(plan=plan, keep=keep, active=active)      # ``active`` keeps the parameters (and their ids) alive

    # ------------------------------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            active = [p for p in group["params"] if p.grad is not None]
            if not active:
                continue
            # the step count is per parameter (a parameter that starts receiving gradients later is younger); all of them
            # share it in the reference's models, so there is normally ONE sub-group = one multi-tensor step
            by_step = {}
            for p in active:
                st = self.state[p]
                if len(st) == 0:
                    self._init_state(p)
                st["step"] += 1
                by_step.setdefault(st["step"], []).append(p)
            for step, params in by_step.items():
                self._step_subgroup(gi, group, step, params)
        return loss

    def _step_subgroup(self, gi, group, step, active):
        key = (gi,) + tuple(id(p) for p in active)
        pl = self._plans.get(key)
        if pl is None:
            if len(self._plans) > 16:            # sets that keep changing: do not hoard plans
                self._plans.clear()
            pl = self._plans[key] = self._build(group, active)
        k = pl["keep"]
        turn = k["turn"] = (k["turn"] + 1) % 4
        if k["grads_ev"][turn] is not None:
            k["grads_ev"][turn].synchronize()
        gh = k["grads_host"][turn]
        ptrs = []
        for p in active:
            g = p.grad
            if g.stride() != p.stride() or g.dtype != torch.float32 or not g.is_cuda:
                if g.is_sparse:
                    raise RuntimeError("Adafactor does not support sparse gradients.")
                # strides of size-1 dims are arbitrary: two row-major tensors are the same layout whatever they say
                if g.dtype != torch.float32 or not g.is_cuda or not (p.is_contiguous() and g.is_contiguous()):
                    g = g.to(device=p.device, dtype=torch.float32)
                    if not (p.is_contiguous() and g.is_contiguous()):
                        g = torch.empty_like(p).copy_(g)      # the parameter's memory layout (e.g. channels_last)
                    p.grad = g
            ptrs.append(g.data_ptr())
        gh.numpy()[:] = ptrs
        k["grads"].copy_(gh, non_blocking=True)
        k["grads_ev"][turn] = torch.cuda.Event()
        k["grads_ev"][turn].record()
        if group["relative_step"]:
            min_step = 1e-6 * step if group["warmup_init"] else 1e-2
            rel = min(min_step, 1.0 / math.sqrt(step))
        else:
            rel = float(group["lr"])
        beta2t = 1.0 - math.pow(step, group["decay_rate"])
        check(_lib.lib().v2f_adafactor_step(ctypes.byref(pl["plan"]), beta2t, rel, stream()), "v2f_adafactor_step")
        # fairseq writes group["lr"] = _get_lr(...) for every parameter, so the value the reference prints/logs in
        # validation_epoch_end is the LAST parameter's: rel * max(eps2, RMS(p)) (device scalar, read lazily)
        if group["relative_step"]:
            group["lr"] = (k["rms"][-1].clamp_min(float(group["eps"][1])) * rel) if group["scale_parameter"] else rel
