"""torch.autograd.Function wrappers over the GTM-family row operators of the C ABI
(include/v2f.h, csrc/gtm_ops.cu).  Same rules as functional.py: PyTorch owns memory, streams and
the autograd graph; every arithmetic operation runs in libv2f_b200.so; nothing falls back to CPU.
"""
import ctypes

import torch

from . import _lib
from ._lib import check, ptr, stream
from .functional import _c, _f32, gemm, colsum, linear, keep_mask, MaskMul, dropout  # noqa: F401


def _L():
    return _lib.lib()


# --------------------------------------------------------------------------- add + LayerNorm
class _AddLayerNorm(torch.autograd.Function):
    """y = LayerNorm(x + a*m) over the last dim (a, m optional)."""

    @staticmethod
    def forward(ctx, x, a, m, gamma, beta, eps):
        x = _c(x)
        a = _c(a) if a is not None else None
        m = _c(m) if m is not None else None
        gamma, beta = _c(gamma), _c(beta)
        D = x.shape[-1]
        M = x.numel() // D
        y = torch.empty_like(x)
        xhat = torch.empty_like(x)
        rstd = _f32(M, device=x.device)
        check(_L().v2f_add_ln_fwd(M, D, ptr(x), ptr(a, allow_none=True), ptr(m, allow_none=True), ptr(gamma),
                                  ptr(beta), float(eps), ptr(y), ptr(xhat), ptr(rstd), stream()), "v2f_add_ln_fwd")
        ctx.save_for_backward(xhat, rstd, gamma, m)
        ctx.has_a = a is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        xhat, rstd, gamma, m = ctx.saved_tensors
        dy = _c(dy)
        D = xhat.shape[-1]
        M = xhat.numel() // D
        dx = torch.empty_like(xhat)
        da = torch.empty_like(xhat) if (ctx.has_a and m is not None) else None
        nblk = _L().v2f_add_ln_bwd_blocks(M)
        part = _f32(nblk, 2, D, device=dy.device)
        dgb = _f32(2, D, device=dy.device)
        check(_L().v2f_add_ln_bwd(M, D, ptr(dy), ptr(xhat), ptr(rstd), ptr(gamma), ptr(m, allow_none=True),
                                  ptr(dx), ptr(da, allow_none=True), ptr(part), ptr(dgb), stream()),
              "v2f_add_ln_bwd")
        if ctx.has_a and da is None:
            da = dx                      # no mask: both branches receive the same gradient
        return dx, (da if ctx.has_a else None), None, dgb[0], dgb[1], None


def add_layer_norm(x, a, m, gamma, beta, eps=1e-5):
    return _AddLayerNorm.apply(x, a, m, gamma, beta, eps)


# --------------------------------------------------------------------------- BatchNorm1d
class _BatchNorm1d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, run_mean, run_var, training, momentum, eps):
        x, gamma, beta = _c(x), _c(gamma), _c(beta)
        B, D = x.shape
        y = torch.empty_like(x)
        mean, rstd = _f32(D, device=x.device), _f32(D, device=x.device)
        check(_L().v2f_bn1d_fwd(B, D, ptr(x), ptr(gamma), ptr(beta), ptr(run_mean), ptr(run_var),
                                1 if training else 0, float(momentum), float(eps), ptr(y), ptr(mean), ptr(rstd),
                                stream()), "v2f_bn1d_fwd")
        ctx.save_for_backward(x, gamma, mean, rstd)
        ctx.training = training
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        dy = _c(dy)
        B, D = x.shape
        dx = torch.empty_like(x)
        dg, db = _f32(D, device=x.device), _f32(D, device=x.device)
        check(_L().v2f_bn1d_bwd(B, D, ptr(x), ptr(dy), ptr(gamma), ptr(mean), ptr(rstd), 1 if ctx.training else 0,
                                ptr(dx), ptr(dg), ptr(db), stream()), "v2f_bn1d_bwd")
        return dx, dg, db, None, None, None, None, None


class _BatchNorm1dSync(torch.autograd.Function):
    """Train-mode BatchNorm1d over a batch SHARDED across the ranks of ``group``: the statistics are those of the global
    batch, as in the reference's single-process run (BatchNorm1d under batch sharding).  Two tiny
    all-reduces of [2,D] doubles per direction; every rank must call it (same count, same order)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, run_mean, run_var, momentum, eps, group):
        import torch.distributed as dist
        x = _c(x)
        B, D = x.shape
        dev = x.device
        sums = torch.empty(2, D, device=dev, dtype=torch.float64)
        check(_L().v2f_bn1d_stats(B, D, ptr(x), sums.data_ptr(), stream()), "v2f_bn1d_stats")
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        btot = float(B * dist.get_world_size(group))       # equal shards (ddp.shard_batch raises otherwise)
        y = torch.empty_like(x)
        mean, rstd = _f32(D, device=dev), _f32(D, device=dev)
        check(_L().v2f_bn1d_apply(B, D, ptr(x), ptr(gamma), ptr(beta), sums.data_ptr(), btot, ptr(run_mean),
                                  ptr(run_var), float(momentum), float(eps), ptr(y), ptr(mean), ptr(rstd), stream()),
              "v2f_bn1d_apply")
        ctx.save_for_backward(x, gamma, mean, rstd)
        ctx.group, ctx.btot = group, btot
        return y

    @staticmethod
    def backward(ctx, dy):
        import torch.distributed as dist
        x, gamma, mean, rstd = ctx.saved_tensors
        dy = _c(dy)
        B, D = x.shape
        dev = x.device
        sums = torch.empty(2, D, device=dev, dtype=torch.float64)
        dg, db = _f32(D, device=dev), _f32(D, device=dev)
        check(_L().v2f_bn1d_bwd_stats(B, D, ptr(x), ptr(dy), ptr(mean), ptr(rstd), sums.data_ptr(), ptr(dg), ptr(db),
                                      stream()), "v2f_bn1d_bwd_stats")
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=ctx.group)
        dx = torch.empty_like(x)
        check(_L().v2f_bn1d_bwd_apply(B, D, ptr(x), ptr(dy), ptr(gamma), ptr(mean), ptr(rstd), sums.data_ptr(),
                                      ctx.btot, ptr(dx), stream()), "v2f_bn1d_bwd_apply")
        return dx, dg, db, None, None, None, None, None


def _sync_group(bn):
    """The process group a BatchNorm1d was marked with by ddp.sync_batchnorm1d (None: per-rank statistics)."""
    g = getattr(bn, "v2f_sync_group", None)
    if g is None:
        return None
    import torch.distributed as dist
    if not dist.is_initialized():
        return None
    group = None if g == "world" else g
    return g if dist.get_world_size(group) > 1 else None


def batch_norm1d(x, bn, training):
    """``bn``: an nn.BatchNorm1d used as parameter / buffer container (running stats updated in place)."""
    if training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    mom = 0.1 if bn.momentum is None else bn.momentum
    g = _sync_group(bn) if training else None
    if g is not None:
        return _BatchNorm1dSync.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, mom, bn.eps,
                                      None if g == "world" else g)
    return _BatchNorm1d.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, training, mom, bn.eps)


# --------------------------------------------------------------------------- gates / elementwise
class _Gate(torch.autograd.Function):
    """mode 0: x*sigmoid(g); mode 1: x + x*sigmoid(g)."""

    @staticmethod
    def forward(ctx, x, g, mode):
        x, g = _c(x), _c(g)
        assert x.shape == g.shape
        out = torch.empty_like(x)
        check(_L().v2f_gate_fwd(x.numel(), ptr(x), ptr(g), mode, ptr(out), stream()), "v2f_gate_fwd")
        ctx.save_for_backward(x, g)
        ctx.mode = mode
        return out

    @staticmethod
    def backward(ctx, dout):
        x, g = ctx.saved_tensors
        dout = _c(dout)
        dx, dg = torch.empty_like(x), torch.empty_like(g)
        check(_L().v2f_gate_bwd(x.numel(), ptr(x), ptr(g), ptr(dout), ctx.mode, ptr(dx), ptr(dg), stream()),
              "v2f_gate_bwd")
        return dx, dg, None


def gate(x, g, residual=False):
    return _Gate.apply(x, g, 1 if residual else 0)


class _Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _c(a), _c(b)
        assert a.shape == b.shape
        out = torch.empty_like(a)
        check(_L().v2f_add_f32(a.numel(), ptr(a), ptr(b), ptr(out), stream()), "v2f_add_f32")
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


def add(a, b):
    return _Add.apply(a, b)


class _Relu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        y = torch.empty_like(x)
        check(_L().v2f_relu_fwd(x.numel(), ptr(x), ptr(y), stream()), "v2f_relu_fwd")
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = _c(dy)
        g = torch.empty_like(dy)
        check(_L().v2f_relu_bwd(dy.numel(), ptr(dy), ptr(y), ptr(g), stream()), "v2f_relu_bwd")
        return g


def relu(x):
    return _Relu.apply(x)


class _AddBcast(torch.autograd.Function):
    """x [R, ...] + p [...] broadcast over the leading dim (p is a buffer: no gradient)."""

    @staticmethod
    def forward(ctx, x, p):
        x, p = _c(x), _c(p)
        n = p.numel()
        rows = x.numel() // n
        out = torch.empty_like(x)
        check(_L().v2f_add_bcast(rows, n, ptr(x), ptr(p), ptr(out), stream()), "v2f_add_bcast")
        return out

    @staticmethod
    def backward(ctx, g):
        return g, None


def add_bcast(x, p):
    return _AddBcast.apply(x, p)


def copy2d(rows, cols, src, src_off, lds, dst, dst_off, ldd):
    check(_L().v2f_copy2d(rows, cols, src.data_ptr() + 4 * src_off, lds, dst.data_ptr() + 4 * dst_off, ldd,
                          stream()), "v2f_copy2d")


class _ConcatCols(torch.autograd.Function):
    """torch.cat(tensors, dim=1) for 2-D fp32 tensors, through v2f_copy2d."""

    @staticmethod
    def forward(ctx, *ts):
        ts = [_c(t) for t in ts]
        rows = ts[0].shape[0]
        widths = [t.shape[1] for t in ts]
        total = sum(widths)
        out = _f32(rows, total, device=ts[0].device)
        off = 0
        for t, w in zip(ts, widths):
            ptr(t)
            copy2d(rows, w, t, 0, w, out, off, total)
            off += w
        ctx.widths = widths
        return out

    @staticmethod
    def backward(ctx, g):
        g = _c(g)
        rows, total = g.shape
        outs, off = [], 0
        for i, w in enumerate(ctx.widths):
            if ctx.needs_input_grad[i]:
                d = _f32(rows, w, device=g.device)
                copy2d(rows, w, g, off, total, d, 0, w)
                outs.append(d)
            else:
                outs.append(None)
            off += w
        return tuple(outs)


def concat_cols(*ts):
    return _ConcatCols.apply(*ts)


class _TakeStep(torch.autograd.Function):
    """x [N,L,D] -> x[:, t, :] (contiguous); the gradient is scattered into zeros."""

    @staticmethod
    def forward(ctx, x, t):
        x = _c(x)
        N, L, D = x.shape
        t = t % L
        out = _f32(N, D, device=x.device)
        ptr(x)
        copy2d(N, D, x, t * D, L * D, out, 0, D)
        ctx.dims = (N, L, D, t)
        return out

    @staticmethod
    def backward(ctx, g):
        N, L, D, t = ctx.dims
        g = _c(g)
        dx = _f32(N, L, D, device=g.device, zero=True)
        copy2d(N, D, g, 0, D, dx, t * D, L * D)
        return dx, None


def take_step(x, t):
    return _TakeStep.apply(x, t)


class _PutStep0(torch.autograd.Function):
    """tgt [N,T,D] = zeros with tgt[:,0,:] = x (the autoregressive decoder input, GTM_Visuelle2.py:251-252)."""

    @staticmethod
    def forward(ctx, x, T):
        x = _c(x)
        N, D = x.shape
        out = _f32(N, T, D, device=x.device, zero=True)
        ptr(x)
        copy2d(N, D, x, 0, D, out, 0, T * D)
        ctx.dims = (N, T, D)
        return out

    @staticmethod
    def backward(ctx, g):
        N, T, D = ctx.dims
        g = _c(g)
        dx = _f32(N, D, device=g.device)
        copy2d(N, D, g, 0, T * D, dx, 0, D)
        return dx, None


def put_step0(x, T):
    return _PutStep0.apply(x, T)


class _RepeatRows(torch.autograd.Function):
    """repeat_interleave(W, dim=0) of x [B, ...]."""

    @staticmethod
    def forward(ctx, x, W):
        x = _c(x)
        B = x.shape[0]
        D = x.numel() // B
        out = _f32(B * W, *x.shape[1:], device=x.device)
        check(_L().v2f_repeat_rows(B, W, D, ptr(x), ptr(out), stream()), "v2f_repeat_rows")
        ctx.dims = (B, W, D, tuple(x.shape))
        return out

    @staticmethod
    def backward(ctx, g):
        B, W, D, shape = ctx.dims
        g = _c(g)
        dx = _f32(*shape, device=g.device)
        check(_L().v2f_fold_rows(B, W, D, ptr(g), ptr(dx), stream()), "v2f_fold_rows")
        return dx, None


def repeat_rows(x, W):
    return x if W == 1 else _RepeatRows.apply(x, W)


# --------------------------------------------------------------------------- static encoders
class _Gather4(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t0, t1, t2, t3, idx, drop):
        tabs = [_c(t) for t in (t0, t1, t2, t3)]
        idx = _c(idx)
        B, E = idx.shape[1], tabs[0].shape[1]
        out = _f32(B, 4, E, device=idx.device)
        arr = (ctypes.c_void_p * 4)(*[ptr(t) for t in tabs])
        check(_L().v2f_gather4_fwd(B, E, arr, ptr(idx, torch.int64), ptr(drop, allow_none=True), ptr(out),
                                   stream()), "v2f_gather4_fwd")
        ctx.save_for_backward(idx, drop)
        ctx.rows, ctx.E = [t.shape[0] for t in tabs], E
        return out

    @staticmethod
    def backward(ctx, dout):
        idx, drop = ctx.saved_tensors
        dout = _c(dout)
        B, E = idx.shape[1], ctx.E
        dt = [_f32(r, E, device=dout.device) for r in ctx.rows]
        arr = (ctypes.c_void_p * 4)(*[ptr(t) for t in dt])
        rows = (ctypes.c_int * 4)(*ctx.rows)
        check(_L().v2f_gather4_bwd(B, E, ptr(idx, torch.int64), ptr(drop, allow_none=True), ptr(dout), rows, arr,
                                   stream()), "v2f_gather4_bwd")
        return dt[0], dt[1], dt[2], dt[3], None, None


def gather4(tables, cat, col, fab, store, p_drop, training):
    """AttributeEncoder of the GTM family -> [B,4,E] (dropout folded into the kernel)."""
    idx = torch.stack([cat, col, fab, store], 0).to(torch.int64)
    B, E = idx.shape[1], tables[0].shape[1]
    drop = keep_mask((B, 4, E), p_drop, training, idx.device)
    return _Gather4.apply(tables[0], tables[1], tables[2], tables[3], idx, drop)


class _Feat4(torch.autograd.Function):
    """out [B,4,E]: the four Linear(1->E) of DummyEmbedder / TemporalEmbedder, concatenated."""

    @staticmethod
    def forward(ctx, temporal, Wt, bt):
        temporal, Wt, bt = _c(temporal), _c(Wt), _c(bt)
        B, E = temporal.shape[0], Wt.shape[1]
        out = _f32(B, 4, E, device=temporal.device)
        check(_L().v2f_feat4_fwd(B, E, ptr(temporal), ptr(Wt), ptr(bt), ptr(out), stream()), "v2f_feat4_fwd")
        ctx.save_for_backward(temporal)
        ctx.E = E
        return out

    @staticmethod
    def backward(ctx, dout):
        (temporal,) = ctx.saved_tensors
        dout = _c(dout)
        B, E = temporal.shape[0], ctx.E
        dWt, dbt = _f32(4, E, device=dout.device), _f32(4, E, device=dout.device)
        check(_L().v2f_feat4_bwd(B, E, ptr(temporal), ptr(dout), ptr(dWt), ptr(dbt), stream()), "v2f_feat4_bwd")
        return None, dWt, dbt


def feat4(temporal, linears):
    """``linears``: the four nn.Linear(1, E) in (day, week, month, year) order."""
    Wt = torch.stack([m.weight[:, 0] for m in linears], 0)
    bt = torch.stack([m.bias for m in linears], 0)
    return _Feat4.apply(temporal.float(), Wt, bt)


class _MeanPool(torch.autograd.Function):
    """Global average over the spatial positions of the trunk's feature map [B,C,h,w] -> [B,C] fp32.
    Accepts NCHW-contiguous or channels_last storage, fp32 or bf16; the gradient comes back in the same
    storage so the (unreplaced) torchvision backward consumes it directly."""

    @staticmethod
    def forward(ctx, feat):
        B, C, h, w = feat.shape
        L = h * w
        if feat.dtype not in (torch.float32, torch.bfloat16):
            feat = feat.float()
        if feat.is_contiguous():
            layout = 0
        elif feat.is_contiguous(memory_format=torch.channels_last):
            layout = 1
        else:
            feat = feat.contiguous()
            layout = 0
        if not feat.is_cuda:
            raise RuntimeError("visuelle2-multimodal-fusion_b200 runs on CUDA only (got a CPU tensor)")
        kind = 1 if feat.dtype == torch.float32 else 0
        out = _f32(B, C, device=feat.device)
        check(_L().v2f_meanpool_fwd(B, L, C, feat.data_ptr(), layout, kind, ptr(out), stream()), "v2f_meanpool_fwd")
        ctx.cfg = (B, C, h, w, layout, kind, feat.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, C, h, w, layout, kind, dtype = ctx.cfg
        dout = _c(dout)
        dx = torch.empty((B, C, h, w), device=dout.device, dtype=dtype,
                         memory_format=torch.channels_last if layout == 1 else torch.contiguous_format)
        check(_L().v2f_meanpool_bwd(B, h * w, C, ptr(dout), layout, kind, dx.data_ptr(), stream()),
              "v2f_meanpool_bwd")
        return dx


def mean_pool(feat):
    return _MeanPool.apply(feat)


# --------------------------------------------------------------------------- attention pieces
class _SdpaKV(torch.autograd.Function):
    """Cross-attention core: q [N,Lq,D] against a packed key|value projection kv [N,Lk,2D]."""

    @staticmethod
    def forward(ctx, q, kv, heads, mask, drop, scale):
        q, kv = _c(q), _c(kv)
        N, Lq, D = q.shape
        Lk = kv.shape[1]
        hd = D // heads
        o = _f32(N, Lq, D, device=q.device)
        P = _f32(N, heads, Lq, Lk, device=q.device)
        kb = ptr(kv)
        check(_L().v2f_sdpa_fwd(N, heads, Lq, Lk, hd, ptr(q), D, Lq * D, kb, 2 * D, Lk * 2 * D, kb + 4 * D, 2 * D,
                                Lk * 2 * D, ptr(o), D, Lq * D, ptr(mask, allow_none=True),
                                ptr(drop, allow_none=True), ptr(P), float(scale), stream()), "v2f_sdpa_fwd")
        ctx.save_for_backward(q, kv, P, drop)
        ctx.heads, ctx.scale = heads, scale
        return o

    @staticmethod
    def backward(ctx, dO):
        q, kv, P, drop = ctx.saved_tensors
        N, Lq, D = q.shape
        Lk = kv.shape[1]
        heads = ctx.heads
        hd = D // heads
        dO = _c(dO)
        dq = torch.empty_like(q)
        dkv = torch.empty_like(kv)
        kb, db = ptr(kv), ptr(dkv)
        check(_L().v2f_sdpa_bwd(N, heads, Lq, Lk, hd, ptr(q), D, Lq * D, kb, 2 * D, Lk * 2 * D, kb + 4 * D, 2 * D,
                                Lk * 2 * D, ptr(dO), D, Lq * D, ptr(drop, allow_none=True), ptr(P),
                                ptr(dq), D, Lq * D, db, 2 * D, Lk * 2 * D, db + 4 * D, 2 * D, Lk * 2 * D,
                                float(ctx.scale), stream()), "v2f_sdpa_bwd")
        return dq, dkv, None, None, None, None


def sdpa_kv(q, kv, heads, mask=None, drop=None):
    return _SdpaKV.apply(q, kv, heads, mask, drop, (q.shape[-1] // heads) ** -0.5)


class _CrossProj(torch.autograd.Function):
    """The packed in-projection of nn.MultiheadAttention used as cross-attention: q = x Wq^T + bq from
    the target rows, kv = mem [Wk;Wv]^T + [bk;bv] from the memory rows.  One weight-gradient tensor."""

    @staticmethod
    def forward(ctx, x, mem, W, b):
        x, mem, W, b = _c(x), _c(mem), _c(W), _c(b)
        D = W.shape[1]
        Mx, Mm = x.numel() // D, mem.numel() // D
        q = _f32(*x.shape[:-1], D, device=x.device)
        kv = _f32(*mem.shape[:-1], 2 * D, device=x.device)
        for t in (x, mem, W, b):
            ptr(t)
        gemm(0, 1, Mx, D, D, x, D, W, D, q, D, bias=b[:D])
        gemm(0, 1, Mm, 2 * D, D, mem, D, W, D, kv, 2 * D, bias=b[D:], b_off=D * D)
        ctx.save_for_backward(x, mem, W)
        return q, kv

    @staticmethod
    def backward(ctx, dq, dkv):
        x, mem, W = ctx.saved_tensors
        dq, dkv = _c(dq), _c(dkv)
        D = W.shape[1]
        Mx, Mm = x.numel() // D, mem.numel() // D
        dx = torch.empty_like(x)
        dmem = torch.empty_like(mem)
        dW = torch.empty_like(W)
        db = _f32(3 * D, device=x.device)
        gemm(0, 0, Mx, D, D, dq, D, W, D, dx, D)
        gemm(0, 0, Mm, D, 2 * D, dkv, 2 * D, W, D, dmem, D, b_off=D * D)
        gemm(1, 0, D, D, Mx, dq, D, x, D, dW, D)
        gemm(1, 0, 2 * D, D, Mm, dkv, 2 * D, mem, D, dW, D, c_off=D * D)
        check(_L().v2f_colsum_f32(Mx, D, ptr(dq), D, db.data_ptr(), 0.0, stream()), "v2f_colsum_f32")
        check(_L().v2f_colsum_f32(Mm, 2 * D, ptr(dkv), 2 * D, db.data_ptr() + 4 * D, 0.0, stream()),
              "v2f_colsum_f32")
        return dx, dmem, dW, db


def cross_proj(x, mem, W, b):
    return _CrossProj.apply(x, mem, W, b)
