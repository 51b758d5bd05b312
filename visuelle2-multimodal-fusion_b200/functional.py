"""torch.autograd.Function wrappers, one per fused region, over the C ABI (include/v2f.h).

PyTorch is plumbing here: it owns device memory (caller-allocated workspaces), the stream and
the autograd graph between fused regions.  All arithmetic runs in libv2f_b200.so.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import DecodeParams, check, ptr, stream


_PRECISION = "fp32"
_TF32_ROUND = False      # tensor-core mode: round tf32 operands to nearest instead of the hardware truncation


def set_precision(mode):
    """'fp32': exact CUDA-core GEMMs (1e-5 contract).  'bf16': tcgen05 tensor-core GEMMs -- bf16
    operands where the data already is bf16 (backbone features), tf32 elsewhere (2e-2 contract)."""
    global _PRECISION
    if mode not in ("fp32", "bf16"):
        raise ValueError(mode)
    _PRECISION = mode


def get_precision():
    return _PRECISION


class precision:
    """Context: precision mode, and (tensor-core mode) whether tf32 operands are rounded to nearest
    (v2f_gemm_tc act bit 2): unbiased, for the small cancellation-prone GEMMs of the GTM family."""

    def __init__(self, mode, round_tf32=False):
        self.mode = mode
        self.round_tf32 = round_tf32

    def __enter__(self):
        global _TF32_ROUND
        self.prev = (_PRECISION, _TF32_ROUND)
        set_precision(self.mode)
        _TF32_ROUND = self.round_tf32

    def __exit__(self, *a):
        global _TF32_ROUND
        set_precision(self.prev[0])
        _TF32_ROUND = self.prev[1]


def _rnd():
    return 4 if _TF32_ROUND else 0


def _tc():
    return _PRECISION == "bf16"


def _f32(*shape, device, zero=False):
    return (torch.zeros if zero else torch.empty)(*shape, device=device, dtype=torch.float32)


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


# --------------------------------------------------------------------------- raw calls
def gemm(ta, tb, M, N, K, A, lda, B, ldb, C, ldc, bias=None, beta=0.0, act=0, batch=1, sA=0, sB=0, sC=0,
         a_off=0, b_off=0, c_off=0):
    """C = op(A) op(B) (+bias) (+beta C); offsets are in elements."""
    check(_lib.lib().v2f_gemm_f32(ta, tb, M, N, K, A.data_ptr() + 4 * a_off, lda, sA,
                                  B.data_ptr() + 4 * b_off, ldb, sB, C.data_ptr() + 4 * c_off, ldc, sC,
                                  batch, ptr(bias, allow_none=True), float(beta), act, stream()), "v2f_gemm_f32")


def colsum(X, M, N, ldx, out, x_off=0):
    check(_lib.lib().v2f_colsum_f32(M, N, X.data_ptr() + 4 * x_off, ldx, ptr(out), 0.0, stream()), "v2f_colsum_f32")


KIND_BF16, KIND_TF32 = 0, 1
TC_MIN_MACS = 1 << 24


def gemm_tc(kind, M, N, K, A, lda, B, ldb, C, ldc, bias=None, beta=0.0, act=0, splits=1, a_off=0, b_off=0, c_off=0):
    """C[M,N] = A[M,K] B[N,K]^T on tcgen05 tensor cores (kind: bf16 or fp32-as-tf32 operands)."""
    es = 2 if kind == KIND_BF16 else 4
    check(_lib.lib().v2f_gemm_tc(kind, M, N, K, A.data_ptr() + es * a_off, lda, B.data_ptr() + es * b_off, ldb,
                                 C.data_ptr() + 4 * c_off, ldc, ptr(bias, allow_none=True), float(beta), act,
                                 splits, stream()), "v2f_gemm_tc")


def gemm_tc_batched(kind, M, N, K, A, lda, sA, B, ldb, sB, C, ldc, sC, batch, bias=None, beta=0.0, act=0, splits=1):
    check(_lib.lib().v2f_gemm_tc_batched(kind, M, N, K, A.data_ptr(), lda, sA, B.data_ptr(), ldb, sB, C.data_ptr(),
                                         ldc, sC, batch, ptr(bias, allow_none=True), float(beta), act, splits,
                                         stream()), "v2f_gemm_tc_batched")


def cast_bf16(x):
    x = _c(x)
    out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    check(_lib.lib().v2f_cast_bf16(x.numel(), ptr(x), out.data_ptr(), stream()), "v2f_cast_bf16")
    return out


def transpose2d(x, out_dtype=None, pad=False):
    """[rows, cols] -> [cols, rows], optionally converting fp32 <-> bf16.  ``pad``: round the row pitch of
    the result up to 16 bytes (TMA requirement) and return ``(buffer, pitch)``; the tail is never read."""
    assert x.dim() == 2 and x.stride(1) == 1
    out_dtype = out_dtype or x.dtype
    kinds = {torch.bfloat16: 0, torch.float32: 1}
    rows, cols = x.shape
    q = 8 if out_dtype == torch.bfloat16 else 4
    pitch = (rows + q - 1) // q * q if pad else rows
    out = torch.empty(cols, pitch, device=x.device, dtype=out_dtype)
    check(_lib.lib().v2f_transpose(rows, cols, x.data_ptr(), x.stride(0), kinds[x.dtype], out.data_ptr(), pitch,
                                   kinds[out_dtype], stream()), "v2f_transpose")
    return (out, pitch) if pad else out


# --------------------------------------------------------------------------- linear
def _splits_for(M, N, K, kb):
    """split-K so that a weight-gradient product with few output tiles still fills the 148 SMs."""
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    s = 1
    while tiles * s < 96 and (K // kb) // (s * 2) >= 8 and s < 16:
        s *= 2
    return s


class _Linear(torch.autograd.Function):
    """y = x W^T + b over the last dim (nn.Linear).  fp32 mode: exact CUDA-core GEMM.  bf16 mode:
    tcgen05 GEMM, bf16 operands if x already is bf16 (backbone features) else tf32."""

    @staticmethod
    def forward(ctx, x, W, b, act):
        x = _c(x)
        W = _c(W)
        K = x.shape[-1]
        M = x.numel() // K
        N = W.shape[0]
        y = _f32(*x.shape[:-1], N, device=x.device)
        ptr(W)
        q = 8 if x.dtype == torch.bfloat16 else 4
        # below ~16 M multiply-adds a product is launch-latency bound on either path: keep it exact
        tc = _tc() and K % q == 0 and K >= 16 and N % 4 == 0 and M * N * K >= TC_MIN_MACS
        if x.dtype == torch.bfloat16 and not tc:
            x = x.float()
        if tc:
            kind = KIND_BF16 if x.dtype == torch.bfloat16 else KIND_TF32
            Wk = cast_bf16(W) if kind == KIND_BF16 else W
            gemm_tc(kind, M, N, K, x, K, Wk, K, y, N, bias=b, act=act | _rnd())
        else:
            ptr(x)
            gemm(0, 1, M, N, K, x, K, W, K, y, N, bias=b, act=act)
        ctx.save_for_backward(x, W, y if act else None)
        ctx.has_bias = b is not None
        ctx.act, ctx.tc, ctx.rnd = act, tc, _rnd()
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, y = ctx.saved_tensors
        dy = _c(dy)
        if ctx.act:
            g = torch.empty_like(dy)
            check(_lib.lib().v2f_relu_bwd(dy.numel(), ptr(dy), ptr(y), ptr(g), stream()), "v2f_relu_bwd")
            dy = g
        K = x.shape[-1]
        M = x.numel() // K
        N = W.shape[0]
        dx = dW = db = None
        x2, dy2 = x.view(M, K), dy.view(M, N)
        if ctx.needs_input_grad[0]:
            if ctx.tc:
                WT = transpose2d(W)                                   # [K,N]
                dx = torch.empty(x.shape, device=x.device, dtype=x.dtype)
                gemm_tc(KIND_TF32, M, K, N, dy2, N, WT, N, dx, K,
                        act=(2 if x.dtype == torch.bfloat16 else 0) | ctx.rnd)
            else:
                dx = torch.empty_like(x)
                gemm(0, 0, M, K, N, dy, N, W, K, dx, K)
        if ctx.needs_input_grad[1]:
            if ctx.tc:
                bf = x.dtype == torch.bfloat16
                dt = torch.bfloat16 if bf else torch.float32
                dyT, pm = transpose2d(dy2, dt, pad=True)              # [N,Mp]
                xT, _ = transpose2d(x2, dt, pad=True)                 # [K,Mp]
                splits = _splits_for(N, K, M, 64 if bf else 32)
                dW = (torch.zeros if splits > 1 else torch.empty)(W.shape, device=W.device, dtype=torch.float32)
                gemm_tc(KIND_BF16 if bf else KIND_TF32, N, K, M, dyT, pm, xT, pm, dW, K, act=ctx.rnd, splits=splits)
            else:
                dW = torch.empty_like(W)
                gemm(1, 0, N, K, M, dy, N, x, K, dW, K)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _f32(N, device=x.device)
            colsum(dy, M, N, N, db)
        return dx, dW, db, None


def linear(x, W, b=None, act=0):
    return _Linear.apply(x, W, b, act)


class MaskMul(torch.autograd.Function):
    """out = x * m (dropout keep-mask, already scaled)."""

    @staticmethod
    def forward(ctx, x, m):
        x, m = _c(x), _c(m)
        out = torch.empty_like(x)
        check(_lib.lib().v2f_mul_f32(x.numel(), ptr(x), ptr(m), ptr(out), stream()), "v2f_mul_f32")
        ctx.save_for_backward(m)
        return out

    @staticmethod
    def backward(ctx, g):
        (m,) = ctx.saved_tensors
        g = _c(g)
        out = torch.empty_like(g)
        check(_lib.lib().v2f_mul_f32(g.numel(), ptr(g), ptr(m), ptr(out), stream()), "v2f_mul_f32")
        return out, None


def keep_mask(shape, p, training, device):
    """Scaled Bernoulli keep-mask drawn from torch's CUDA generator, or None when inactive."""
    if not training or p <= 0.0:
        return None
    return (torch.rand(shape, device=device) >= p).float().div_(1.0 - p)


class _Dropout(torch.autograd.Function):
    """nn.Dropout with the Bernoulli decisions generated inside the kernel from a 128-bit key in device memory (two
    int64 drawn from torch's CUDA generator: one tiny launch, graph-safe); the backward replays the same key."""

    @staticmethod
    def forward(ctx, x, key, p):
        x = _c(x)
        out = torch.empty_like(x)
        check(_lib.lib().v2f_dropout(x.numel(), ptr(x), key.data_ptr(), float(p), ptr(out), stream()), "v2f_dropout")
        ctx.save_for_backward(key)
        ctx.p = p
        return out

    @staticmethod
    def backward(ctx, g):
        (key,) = ctx.saved_tensors
        g = _c(g)
        out = torch.empty_like(g)
        check(_lib.lib().v2f_dropout(g.numel(), ptr(g), key.data_ptr(), float(ctx.p), ptr(out), stream()), "v2f_dropout")
        return out, None, None


def dropout(x, p, training):
    if not training or p <= 0.0:
        return x
    key = torch.randint(-(1 << 62), 1 << 62, (2,), device=x.device, dtype=torch.int64)
    return _Dropout.apply(x, key, p)


# --------------------------------------------------------------------------- trend_linear re-association
class _TrendProj(torch.autograd.Function):
    """P[b,j,:] = W_tl[:, jE:(j+1)E] @ V[b,j,:]  (SURVEY.md 8a identity 2; trend_linear,
    models/CrossAttnRNN210.py:128,196) as one batched GEMM over the 52 positions."""

    @staticmethod
    def forward(ctx, V, W):
        V, W = _c(V), _c(W)
        B, L, E = V.shape
        Eo = W.shape[0]
        assert W.shape[1] == L * E
        P = _f32(B, L, Eo, device=V.device)
        ptr(V), ptr(W)
        tc = _tc() and E % 4 == 0 and Eo % 4 == 0 and E >= 16
        if tc:
            gemm_tc_batched(KIND_TF32, B, Eo, E, V, L * E, E, W, L * E, E, P, L * Eo, Eo, L)
        else:
            gemm(0, 1, B, Eo, E, V, L * E, W, L * E, P, L * Eo, batch=L, sA=E, sB=E, sC=Eo)
        ctx.save_for_backward(V, W)
        ctx.tc = tc
        return P

    @staticmethod
    def backward(ctx, dP):
        V, W = ctx.saved_tensors
        dP = _c(dP)
        B, L, E = V.shape
        Eo = W.shape[0]
        dV = torch.empty_like(V)
        dW = torch.empty_like(W)
        if ctx.tc:
            WT = transpose2d(W)                                           # [L*E, Eo]; block j = W_j^T
            gemm_tc_batched(KIND_TF32, B, E, Eo, dP, L * Eo, Eo, WT, Eo, E * Eo, dV, L * E, E, L)
            dPT, pb = transpose2d(dP.view(B, L * Eo), pad=True)           # [L*Eo, Bp]
            VT, _ = transpose2d(V.view(B, L * E), pad=True)               # [L*E, Bp]
            gemm_tc_batched(KIND_TF32, Eo, E, B, dPT, pb, Eo * pb, VT, pb, E * pb, dW, L * E, E, L)
        else:
            gemm(0, 0, B, E, Eo, dP, L * Eo, W, L * E, dV, L * E, batch=L, sA=Eo, sB=E, sC=E)
            gemm(1, 0, Eo, E, B, dP, L * Eo, V, L * E, dW, L * E, batch=L, sA=Eo, sB=E, sC=E)
        return dV, dW


def trend_proj(V, W):
    return _TrendProj.apply(V, W)


# --------------------------------------------------------------------------- GRU over a sequence
class _GruSeq(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, h0, w_ih, w_hh, b_ih, b_hh):
        x, h0, w_ih, w_hh, b_ih, b_hh = map(_c, (x, h0, w_ih, w_hh, b_ih, b_hh))
        N, L, I = x.shape
        H = w_hh.shape[1]
        dev = x.device
        out = _f32(N, L, H, device=dev)
        GI = _f32(N, L, 3 * H, device=dev)
        GH = _f32(N, 3 * H, device=dev)
        RZN = _f32(L, N, 3 * H, device=dev)
        GHN = _f32(L, N, H, device=dev)
        prec = 1 if (_tc() and H % 4 == 0 and N * 3 * H * H >= TC_MIN_MACS) else 0
        check(_lib.lib().v2f_gru_seq_fwd(N, L, I, H, ptr(x), ptr(h0), ptr(w_ih), ptr(w_hh), ptr(b_ih),
                                         ptr(b_hh), ptr(out), ptr(GI), ptr(GH), ptr(RZN), ptr(GHN), prec,
                                         stream()), "v2f_gru_seq_fwd")
        ctx.save_for_backward(x, h0, w_ih, w_hh, out, RZN, GHN)
        ctx.prec = prec
        return out

    @staticmethod
    def backward(ctx, dOut):
        x, h0, w_ih, w_hh, out, RZN, GHN = ctx.saved_tensors
        N, L, I = x.shape
        H = w_hh.shape[1]
        dev = x.device
        dOut = _c(dOut)
        dh = _f32(N, H, device=dev)
        DGI = _f32(N, L, 3 * H, device=dev)
        DGH = _f32(L, N, 3 * H, device=dev)
        Hprev = _f32(L, N, H, device=dev)
        dx = _f32(N, L, I, device=dev) if ctx.needs_input_grad[0] else None
        dh0 = _f32(N, H, device=dev) if ctx.needs_input_grad[1] else None
        dw_ih, dw_hh = torch.empty_like(w_ih), torch.empty_like(w_hh)
        db_ih, db_hh = _f32(3 * H, device=dev), _f32(3 * H, device=dev)
        w_hhT = ws = None
        ws_n = 0
        if ctx.prec:
            w_hhT = _f32(H, 3 * H, device=dev)
            ws_n = (3 * H + H) * ((N * L + 3) // 4 * 4)
            ws = _f32(ws_n, device=dev)
        check(_lib.lib().v2f_gru_seq_bwd(N, L, I, H, ptr(x), ptr(h0), ptr(w_ih), ptr(w_hh), ptr(out),
                                         ptr(RZN), ptr(GHN), ptr(dOut), None, ptr(dh), ptr(DGI), ptr(DGH),
                                         ptr(Hprev), ptr(dx, allow_none=True), ptr(dh0, allow_none=True),
                                         ptr(dw_ih), ptr(dw_hh), ptr(db_ih), ptr(db_hh),
                                         ptr(w_hhT, allow_none=True), ptr(ws, allow_none=True), ws_n, ctx.prec,
                                         stream()), "v2f_gru_seq_bwd")
        return dx, dh0, dw_ih, dw_hh, db_ih, db_hh


def gru_seq(x, h0, w_ih, w_hh, b_ih, b_hh):
    """All hidden states [N,L,H] of a batch-first single-layer GRU (h0 [N,H])."""
    return _GruSeq.apply(x, h0, w_ih, w_hh, b_ih, b_hh)


# --------------------------------------------------------------------------- attention core
class _Sdpa(torch.autograd.Function):
    """softmax(scale q k^T + mask) (*drop) v per (batch, head); q [B,Lq,heads*hd] etc. (views with a
    last-dim stride of 1 and arbitrary row / batch strides are accepted, e.g. slices of a packed QKV)."""

    @staticmethod
    def forward(ctx, q, k, v, heads, mask, drop, scale):
        B, Lq, D = q.shape
        Lk = k.shape[1]
        hd = D // heads
        for t in (q, k, v):
            assert t.stride(2) == 1 and t.dtype == torch.float32 and t.is_cuda
        o = _f32(B, Lq, D, device=q.device)
        P = _f32(B, heads, Lq, Lk, device=q.device)
        check(_lib.lib().v2f_sdpa_fwd(B, heads, Lq, Lk, hd, q.data_ptr(), q.stride(1), q.stride(0),
                                      k.data_ptr(), k.stride(1), k.stride(0), v.data_ptr(), v.stride(1),
                                      v.stride(0), ptr(o), D, Lq * D, ptr(mask, allow_none=True),
                                      ptr(drop, allow_none=True), ptr(P), float(scale), stream()),
              "v2f_sdpa_fwd")
        ctx.save_for_backward(q, k, v, P, drop)
        ctx.heads, ctx.scale = heads, scale
        return o

    @staticmethod
    def backward(ctx, dO):
        q, k, v, P, drop = ctx.saved_tensors
        B, Lq, D = q.shape
        Lk = k.shape[1]
        heads = ctx.heads
        hd = D // heads
        dO = _c(dO)
        dq = _f32(B, Lq, D, device=q.device)
        dk = _f32(B, Lk, D, device=q.device)
        dv = _f32(B, Lk, D, device=q.device)
        check(_lib.lib().v2f_sdpa_bwd(B, heads, Lq, Lk, hd, q.data_ptr(), q.stride(1), q.stride(0),
                                      k.data_ptr(), k.stride(1), k.stride(0), v.data_ptr(), v.stride(1),
                                      v.stride(0), ptr(dO), D, Lq * D, ptr(drop, allow_none=True), ptr(P),
                                      ptr(dq), D, Lq * D, ptr(dk), D, Lk * D, ptr(dv), D, Lk * D,
                                      float(ctx.scale), stream()), "v2f_sdpa_bwd")
        return dq, dk, dv, None, None, None, None


def sdpa(q, k, v, heads, mask=None, drop=None, scale=None):
    if scale is None:
        scale = (q.shape[-1] // heads) ** -0.5
    return _Sdpa.apply(q, k, v, heads, mask, drop, scale)


class _SdpaPacked(torch.autograd.Function):
    """Self-attention core on a packed projection qkv [B,L,3E] (q | k | v); the gradient comes back
    packed as well, so no slice/concat traffic surrounds the kernel."""

    @staticmethod
    def forward(ctx, qkv, heads, mask, drop, scale):
        qkv = _c(qkv)
        B, L, D3 = qkv.shape
        D = D3 // 3
        hd = D // heads
        o = _f32(B, L, D, device=qkv.device)
        P = _f32(B, heads, L, L, device=qkv.device)
        base = ptr(qkv)
        check(_lib.lib().v2f_sdpa_fwd(B, heads, L, L, hd, base, D3, L * D3, base + 4 * D, D3, L * D3,
                                      base + 8 * D, D3, L * D3, ptr(o), D, L * D,
                                      ptr(mask, allow_none=True), ptr(drop, allow_none=True), ptr(P),
                                      float(scale), stream()), "v2f_sdpa_fwd")
        ctx.save_for_backward(qkv, P, drop)
        ctx.heads, ctx.scale = heads, scale
        return o

    @staticmethod
    def backward(ctx, dO):
        qkv, P, drop = ctx.saved_tensors
        B, L, D3 = qkv.shape
        D = D3 // 3
        heads = ctx.heads
        hd = D // heads
        dO = _c(dO)
        dqkv = torch.empty_like(qkv)
        base, dbase = ptr(qkv), ptr(dqkv)
        check(_lib.lib().v2f_sdpa_bwd(B, heads, L, L, hd, base, D3, L * D3, base + 4 * D, D3, L * D3,
                                      base + 8 * D, D3, L * D3, ptr(dO), D, L * D,
                                      ptr(drop, allow_none=True), ptr(P), dbase, D3, L * D3,
                                      dbase + 4 * D, D3, L * D3, dbase + 8 * D, D3, L * D3,
                                      float(ctx.scale), stream()), "v2f_sdpa_bwd")
        return dqkv, None, None, None, None


def mha_self(x, in_w, in_b, out_w, out_b, heads, p_drop, training, mask=None):
    """nn.MultiheadAttention(x,x,x) on batch-first x [B,L,E] (models/CrossAttnRNN210.py:176-179)."""
    B, L, E = x.shape
    qkv = linear(x, in_w, in_b)                       # [B,L,3E]
    drop = keep_mask((B, heads, L, L), p_drop, training, x.device)
    o = _SdpaPacked.apply(qkv, heads, mask, drop, (E // heads) ** -0.5)
    return linear(o, out_w, out_b)


# --------------------------------------------------------------------------- embedders
class _Embed(torch.autograd.Function):
    """(date, attributes) [B,2,E]; Wt,bt [4,E]; tables: 4 tensors; idx [4,B] int64; drop [B,8,E]|None."""

    @staticmethod
    def forward(ctx, temporal, Wt, bt, t0, t1, t2, t3, idx, drop):
        temporal, Wt, bt, idx = _c(temporal), _c(Wt), _c(bt), _c(idx)
        tabs = [_c(t) for t in (t0, t1, t2, t3)]
        B, E = temporal.shape[0], Wt.shape[1]
        out = _f32(B, 2, E, device=temporal.device)
        arr = (ctypes.c_void_p * 4)(*[ptr(t) for t in tabs])
        check(_lib.lib().v2f_embed_fwd(B, E, ptr(temporal), ptr(Wt), ptr(bt), arr, ptr(idx, torch.int64),
                                       ptr(drop, allow_none=True), ptr(out), stream()), "v2f_embed_fwd")
        ctx.save_for_backward(temporal, idx, drop)
        ctx.rows = [t.shape[0] for t in tabs]
        ctx.E = E
        return out

    @staticmethod
    def backward(ctx, dout):
        temporal, idx, drop = ctx.saved_tensors
        dout = _c(dout)
        B, E = temporal.shape[0], ctx.E
        dev = temporal.device
        dWt, dbt = _f32(4, E, device=dev), _f32(4, E, device=dev)
        dt = [_f32(r, E, device=dev) for r in ctx.rows]
        arr = (ctypes.c_void_p * 4)(*[ptr(t) for t in dt])
        rows = (ctypes.c_int * 4)(*ctx.rows)
        check(_lib.lib().v2f_embed_bwd(B, E, ptr(temporal), ptr(idx, torch.int64), ptr(drop, allow_none=True),
                                       ptr(dout), rows, ptr(dWt), ptr(dbt), arr, stream()), "v2f_embed_bwd")
        return None, dWt, dbt, dt[0], dt[1], dt[2], dt[3], None, None


def embed(temporal, Wt, bt, tables, idx, drop=None):
    return _Embed.apply(temporal, Wt, bt, tables[0], tables[1], tables[2], tables[3], idx, drop)


# --------------------------------------------------------------------------- fused decoder
VARIANT_210, VARIANT_21, VARIANT_DEMAND = 0, 1, 2
STREAM_ATTENTION = True      # TMA-staged streaming attention kernels when the dims allow (E % 256 == 0)
# whole decode loop as one cooperative launch (csrc/decode_persist.cu) when the dims allow; V2F_PERSISTENT_DECODE=0 is
# the A/B switch of bench.py / tools
PERSISTENT_DECODE = os.environ.get("V2F_PERSISTENT_DECODE", "1") != "0"
TEAM_DECODE = os.environ.get("V2F_TEAM_DECODE", "1") != "0"    # A/B switch: 0 keeps the decoder on decode_persist.cu


KEEP_LAST_PERSIST_WS = False   # profiling: remember the persistent decoder's workspace (phase stamps live in it)
_last_persist = []
PERSIST_PHASES = ("P1 S-product", "P2 attention sweep", "P2b combine", "P3 HC-product", "P4 multimodal attention",
                  "P5/P6 embedder + GRU gates")
TEAM_PHASES = ("P1 S-product", "P2 attention sweep", "P3 HC-product", "P4 multimodal attention", "P5 GI-product + GRU gates")


def persist_phase_times():
    """Per-phase times (us, mean over the T steps, CTA 0's %globaltimer stamps) of the most recent persistent decode
    forward run with the stamps enabled and KEEP_LAST_PERSIST_WS: {phase: (work_us, barrier_us)}."""
    if not _last_persist:
        return None
    team = len(_last_persist) == 3
    ws, (N, E, H, T, Li, Lt) = _last_persist[:2]
    if team:
        off = _lib.lib().v2f_decode_team_stamps_offset(N, _last_persist[2], T, Li, Lt) // 4
        names = TEAM_PHASES
    else:
        off = _lib.lib().v2f_decode_persist_stamps_offset(N, E, H) // 4
        names = PERSIST_PHASES
    torch.cuda.synchronize()
    ns = 32 if team else 16
    st = ws[off:off + 2 * T * ns].cpu().view(torch.int64).view(T, ns)
    out = {}
    for k, name in enumerate(names):
        work = sum(int(st[t, 2 * k + 1]) - int(st[t, 2 * k]) for t in range(T)) / T / 1e3
        wait = sum(int(st[t, 2 * k + 2]) - int(st[t, 2 * k + 1]) for t in range(T)) / T / 1e3
        out[name] = (work, wait)
    if team:      # inside P2 (CTA 0, thread 0): pass 1 | wait for the other warps | softmax | pass 2 | tail (combine + writes)
        marks = [2, 11, 12, 13, 14, 3]
        names2 = ["P2.a energies pass", "P2.b barrier", "P2.c softmax", "P2.d context pass", "P2.e combine"]
        for (a, b), name in zip(zip(marks[:-1], marks[1:]), names2):
            out[name] = (sum(int(st[t, b]) - int(st[t, a]) for t in range(T)) / T / 1e3, 0.0)
        for a_, b_, name in ((0, 20, "P1.m first chunk landed (MMA thread)"), (20, 21, "P1.m last chunk landed"),
                             (21, 22, "P1.m MMAs issued + commits"), (22, 17, "P1.m commit -> epilogue sees mma_done")):
            out[name] = (sum(int(st[t, b_]) - int(st[t, a_]) for t in range(T)) / T / 1e3, 0.0)
        marks = [0, 16, 17, 18, 19, 1]
        names1 = ["P1.a TMA issue", "P1.b loads + MMA", "P1.c TMEM -> smem", "P1.d barrier", "P1.e S stores"]
        for (a, b), name in zip(zip(marks[:-1], marks[1:]), names1):
            out[name] = (sum(int(st[t, b]) - int(st[t, a]) for t in range(T)) / T / 1e3, 0.0)
    return out


class _Decode(torch.autograd.Function):
    """The fused recurrent-attention decoder (v2f_decode_fwd / v2f_decode_bwd)."""

    @staticmethod
    def forward(ctx, cfg, Himg, Vimg, Htr, Ptr, Mst, HMst, h0, x0, y,
                Wd_img, Wd_tr, Wd_mm, w_img, w_tr, w_mm, b_img, b_tr, b_mm, b_tl, We_mm, W_me, b_me,
                W_ih, W_hh, b_ih, b_hh, w_fc, b_fc):
        variant, W, T, tf_mask, mod_mask = cfg
        dev = Himg.device
        Himg, Htr, Ptr, Mst, HMst, h0 = map(_c, (Himg, Htr, Ptr, Mst, HMst, h0))
        Vimg = Himg if variant == VARIANT_DEMAND else _c(Vimg)
        B, Li, E = Himg.shape
        Lt = Htr.shape[1]
        N = B * W
        H = h0.shape[1]
        gru = variant != VARIANT_21
        G = 3 * H if gru else 0
        with torch.no_grad():
            if gru:
                Wcat = torch.cat([Wd_img, Wd_tr, Wd_mm, W_hh], 0).contiguous()
                bcat = torch.cat([b_hh.new_zeros(3 * E), b_hh]).contiguous()
                W_ihc = W_ih[:, :E].contiguous()
                w_x = W_ih[:, E].contiguous()
                b_ih_c = _c(b_ih)
            else:
                Wcat = torch.cat([Wd_img, Wd_tr, Wd_mm], 0).contiguous()
                bcat = Wcat.new_zeros(3 * E)
                W_ihc = w_x = b_ih_c = None
            w_att = torch.cat([w_img.reshape(1, E), w_tr.reshape(1, E), w_mm.reshape(1, E)], 0).contiguous()
            beta_att = torch.cat([b_img.reshape(1), b_tr.reshape(1), b_mm.reshape(1)]).contiguous()
            w_fc_c = w_fc.reshape(-1).contiguous()
        full = (mod_mask & 0b1010) == 0b1010
        p = DecodeParams()
        p.N, p.B, p.W, p.E, p.H, p.Li, p.Lt, p.T = N, B, W, E, H, Li, Lt, T
        tf_dev = tf_mask if torch.is_tensor(tf_mask) else None      # device-resident bits (CUDA-graph replay)
        if tf_dev is not None:
            tf_mask = 0
        p.variant, p.mod_mask, p.tf_mask = variant, mod_mask, tf_mask if y is not None else 0
        prec = 1 if (_tc() and E % 4 == 0 and H % 4 == 0) else 0
        p.precision = prec
        keep = dict(Himg=Himg, Vimg=Vimg, Htr=Htr, Ptr=Ptr, Mst=Mst, HMst=HMst, h0=h0,
                    x0=_c(x0) if x0 is not None else None, y=_c(y) if y is not None else None,
                    Wcat=Wcat, bcat=bcat, w_att=w_att, beta_att=beta_att, b_tl=_c(b_tl), We_mm=_c(We_mm),
                    W_me=_c(W_me), b_me=_c(b_me), W_ihc=W_ihc, w_x=w_x, b_ih=b_ih_c, w_fc=w_fc_c, b_fc=_c(b_fc))
        z = not full     # disabled modalities leave slots unwritten: zero them so 0 * garbage cannot appear
        keep.update(
            yhat=_f32(N, T, device=dev), h_all=_f32(T + 1, N, H, device=dev),
            S_all=_f32(T, N, 3 * E + G, device=dev), alpha_img=_f32(T, N, Li, device=dev, zero=z),
            alpha_tr=_f32(T, N, Lt, device=dev, zero=z), alpha_mm=_f32(T, N, 4, device=dev),
            C=_f32(T, N, 2, E, device=dev, zero=z), HC=_f32(T, N, 2, E, device=dev, zero=z),
            U=_f32(T, N, E, device=dev), CTX=_f32(T, N, E, device=dev),
            GI=_f32(N, max(3 * H, 1), device=dev), RZN=_f32(T, N, max(3 * H, 1), device=dev),
            xin=_f32(T + 1, N, device=dev, zero=True))
        if STREAM_ATTENTION and E % 256 == 0 and E <= 1024:
            keep["attn_ws"] = _f32(N * ((Li + 7) // 8 + (Lt + 7) // 8) * (2 * E + 2), device=dev)
            if PERSISTENT_DECODE and gru and E in (256, 512):
                keep["persist_ws"] = _f32(_lib.lib().v2f_decode_persist_ws_floats(N, E, H, T), device=dev)
                if KEEP_LAST_PERSIST_WS:
                    _last_persist[:] = [keep["persist_ws"], (N, E, H, T, Li, Lt)]
                # row-team tcgen05 kernel (csrc/decode_team.cu): default dims, tensor-core mode, <= 128 rows per launch
                if TEAM_DECODE and prec and E == 512 and H == 512 and N <= 128 and full:
                    p.team_ws_floats = _lib.lib().v2f_decode_team_ws_floats(N, B, T, Li, Lt)
                    keep["team_ws"] = _f32(p.team_ws_floats, device=dev)
                    if KEEP_LAST_PERSIST_WS:
                        _last_persist[:] = [keep["team_ws"], (N, E, H, T, Li, Lt), B]
        for k, v in keep.items():
            setattr(p, k, ptr(v, allow_none=True))
        p.tf_mask_dev = ptr(tf_dev, torch.int32, allow_none=True)
        check(_lib.lib().v2f_decode_fwd(ctypes.byref(p), stream()), "v2f_decode_fwd")
        ctx.keep, ctx.dims = keep, (variant, W, T, mod_mask, N, B, E, H, Li, Lt, G)
        ctx.tf_mask, ctx.prec, ctx.tf_dev = p.tf_mask, prec, tf_dev
        yhat, a_img, a_mm = keep["yhat"], keep["alpha_img"], keep["alpha_mm"]
        ctx.mark_non_differentiable(a_img, a_mm)
        return yhat, a_img, a_mm

    @staticmethod
    def backward(ctx, dY, _da_img, _da_mm):
        variant, W, T, mod_mask, N, B, E, H, Li, Lt, G = ctx.dims
        keep = ctx.keep
        dev = dY.device
        gru = variant != VARIANT_21
        full = (mod_mask & 0b1010) == 0b1010
        z = not full
        p = DecodeParams()
        p.N, p.B, p.W, p.E, p.H, p.Li, p.Lt, p.T = N, B, W, E, H, Li, Lt, T
        p.variant, p.mod_mask, p.tf_mask = variant, mod_mask, ctx.tf_mask
        p.precision = ctx.prec
        H3 = max(3 * H, 1)
        g = dict(
            dY=_c(dY), dh=_f32(N, H, device=dev, zero=True), DScat=_f32(T, N, 3 * E + G, device=dev, zero=z),
            DGI=_f32(T, N, H3, device=dev), DCTX=_f32(T, N, E, device=dev), dU=_f32(N, E, device=dev),
            DHC=_f32(T, N, 2, E, device=dev), DC=_f32(T, N, 2, E, device=dev),
            DE_img=_f32(T, N, Li, device=dev, zero=z), DE_tr=_f32(T, N, Lt, device=dev, zero=z),
            DYH=_f32(T, N, device=dev), dxn=_f32(N, device=dev, zero=True),
            dw_acc=_f32(N, 3, E, device=dev, zero=True), dMst_acc=_f32(N, 2, E, device=dev, zero=True),
            dHMst_acc=_f32(N, 2, E, device=dev, zero=True),
            dHimg=_f32(B, Li, E, device=dev, zero=z),
            dVimg=_f32(B, Li, E, device=dev, zero=z) if variant != VARIANT_DEMAND else None,
            dHtr=_f32(B, Lt, E, device=dev, zero=z), dPtr=_f32(B, Lt, E, device=dev, zero=z),
            dMst=_f32(B, 2, E, device=dev), dHMst=_f32(B, 2, E, device=dev),
            dWcat=_f32(3 * E + G, H, device=dev), dbcat=_f32(3 * E + G, device=dev),
            dw_att=_f32(3, E, device=dev), db_tl=_f32(E, device=dev), dWe_mm=_f32(E, E, device=dev),
            dW_me=_f32(E, E, device=dev), db_me=_f32(E, device=dev),
            dW_ihc=_f32(H3, E, device=dev), dw_x=_f32(H3, device=dev), db_ih=_f32(H3, device=dev),
            dw_fc=_f32(H if gru else E, device=dev), db_fc=_f32(1, device=dev))
        if ctx.prec:
            ldS = 3 * E + G
            tn4, rows = (T * N + 3) // 4 * 4, (2 * T * N + 3) // 4 * 4
            ws_n = max((ldS + H) * tn4, (3 * H + E) * tn4, 2 * E * rows) + 64
            if "team_ws" in keep:      # scratch of the row-team persistent backward (csrc/decode_team.cu)
                ws_n = max(ws_n, _lib.lib().v2f_decode_team_bwd_ws_floats(N, T))
                p.team_ws_floats = keep["team_ws"].numel()
            g.update(WcatT=_f32(H, ldS, device=dev), W_ihcT=_f32(E, H3, device=dev), W_meT=_f32(E, E, device=dev),
                     We_mmT=_f32(E, E, device=dev), ws=_f32(ws_n, device=dev))
            p.ws_floats = ws_n
        for k, v in keep.items():
            setattr(p, k, ptr(v, allow_none=True))
        for k, v in g.items():
            setattr(p, k, ptr(v, allow_none=True))
        p.tf_mask_dev = ptr(ctx.tf_dev, torch.int32, allow_none=True)
        check(_lib.lib().v2f_decode_bwd(ctypes.byref(p), stream()), "v2f_decode_bwd")
        dWcat, dw_att = g["dWcat"], g["dw_att"]
        zero1 = dWcat.new_zeros(1)
        if gru:
            dW_ih = torch.cat([g["dW_ihc"], g["dw_x"].unsqueeze(1)], 1)
            dW_hh, db_hh, db_ih = dWcat[3 * E:], g["dbcat"][3 * E:], g["db_ih"]
            dx0 = None
        else:
            dW_ih = dW_hh = db_hh = db_ih = dx0 = None
        dw_fc = g["dw_fc"].reshape(1, -1)
        return (None, g["dHimg"], g["dVimg"], g["dHtr"], g["dPtr"], g["dMst"], g["dHMst"], g["dh"], dx0, None,
                dWcat[:E], dWcat[E:2 * E], dWcat[2 * E:3 * E], dw_att[0:1], dw_att[1:2], dw_att[2:3],
                zero1, zero1.clone(), zero1.clone(), g["db_tl"], g["dWe_mm"], g["dW_me"], g["db_me"],
                dW_ih, dW_hh, db_ih, db_hh, dw_fc, g["db_fc"])


def decode(variant, W, T, tf_mask, mod_mask, Himg, Vimg, Htr, Ptr, Mst, HMst, h0, x0, y, weights):
    """``weights``: dict with the 19 parameter tensors named as in _Decode.forward."""
    order = ["Wd_img", "Wd_tr", "Wd_mm", "w_img", "w_tr", "w_mm", "b_img", "b_tr", "b_mm", "b_tl", "We_mm",
             "W_me", "b_me", "W_ih", "W_hh", "b_ih", "b_hh", "w_fc", "b_fc"]
    return _Decode.apply((variant, W, T, tf_mask, mod_mask), Himg, Vimg, Htr, Ptr, Mst, HMst, h0, x0, y,
                         *[weights.get(k) for k in order])
