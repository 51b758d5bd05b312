"""GPU unit parity of the C-ABI primitives against plain torch / the oracle's explicit math."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max() / max(float(b.abs().max()), 1e-30))


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 512, 512), (37, 19, 53), (1, 7, 3), (200, 1, 1280), (64, 64, 16)])
def test_gemm(ta, tb, M, N, K):
    from visuelle2_multimodal_fusion_b200 import functional as Fv
    g = torch.Generator().manual_seed(M * 131 + N * 7 + K)
    A = torch.randn((K, M) if ta else (M, K), generator=g).cuda()
    B = torch.randn((N, K) if tb else (K, N), generator=g).cuda()
    bias = torch.randn(N, generator=g).cuda()
    C0 = torch.randn(M, N, generator=g).cuda()
    C = C0.clone()
    Fv.gemm(ta, tb, M, N, K, A, A.shape[1], B, B.shape[1], C, N, bias=bias, beta=1.0)
    ref = (A.t() if ta else A).double() @ (B.t() if tb else B).double() + bias.double() + C0.double()
    assert _rel(C, ref) < TOL


def test_gemm_batched_strided():
    from visuelle2_multimodal_fusion_b200 import functional as Fv
    g = torch.Generator().manual_seed(0)
    B_, L, E, Eo = 5, 7, 32, 24
    V = torch.randn(B_, L, E, generator=g).cuda().requires_grad_(True)
    W = torch.randn(Eo, L * E, generator=g).cuda().requires_grad_(True)
    P = Fv.trend_proj(V, W)
    ref = torch.einsum("ble,ole->blo", V.double(), W.double().view(Eo, L, E))
    assert _rel(P, ref) < TOL
    # the identity the re-association rests on: trend_linear(vec(alpha_j V_j)) = sum_j alpha_j P_j
    alpha = torch.rand(B_, L, generator=g).cuda()
    lhs = (alpha.unsqueeze(2) * V).reshape(B_, -1) @ W.t()
    rhs = (alpha.unsqueeze(2) * P).sum(1)
    assert _rel(rhs, lhs) < TOL
    dP = torch.randn(B_, L, Eo, generator=g).cuda()
    P.backward(dP)
    Vd, Wd = V.detach().double().requires_grad_(True), W.detach().double().requires_grad_(True)
    torch.einsum("ble,ole->blo", Vd, Wd.view(Eo, L, E)).backward(dP.double())
    assert _rel(V.grad, Vd.grad) < TOL and _rel(W.grad, Wd.grad) < TOL


@pytest.mark.parametrize("N,L,I,H", [(5, 7, 3, 48), (128, 52, 3, 512), (9, 2, 1, 64)])
def test_gru_seq(N, L, I, H):
    from oracle import rnn as orc
    from visuelle2_multimodal_fusion_b200 import functional as Fv
    g = torch.Generator().manual_seed(N + L)
    k = 1 / math.sqrt(H)
    P = {"weight_ih_l0": (torch.rand(3 * H, I, generator=g) * 2 - 1) * k,
         "weight_hh_l0": (torch.rand(3 * H, H, generator=g) * 2 - 1) * k,
         "bias_ih_l0": (torch.rand(3 * H, generator=g) * 2 - 1) * k,
         "bias_hh_l0": (torch.rand(3 * H, generator=g) * 2 - 1) * k}
    x = torch.rand(N, L, I, generator=g)
    h0 = torch.randn(N, H, generator=g) * 0.1
    dOut = torch.randn(N, L, H, generator=g)
    Pc = {k_: v.clone().requires_grad_(True) for k_, v in P.items()}
    xc, hc = x.clone().requires_grad_(True), h0.clone().requires_grad_(True)
    out_ref, _ = orc.gru_seq(xc, hc, Pc, "")
    out_ref.backward(dOut)
    Pg = {k_: v.cuda().requires_grad_(True) for k_, v in P.items()}
    xg, hg = x.cuda().requires_grad_(True), h0.cuda().requires_grad_(True)
    out = Fv.gru_seq(xg, hg, Pg["weight_ih_l0"], Pg["weight_hh_l0"], Pg["bias_ih_l0"], Pg["bias_hh_l0"])
    out.backward(dOut.cuda())
    assert _rel(out, out_ref) < TOL
    assert _rel(xg.grad, xc.grad) < TOL and _rel(hg.grad, hc.grad) < TOL
    for k_ in P:
        assert _rel(Pg[k_].grad, Pc[k_].grad) < TOL, k_


@pytest.mark.parametrize("B,L,E,heads,masked", [(3, 52, 32, 4, False), (4, 52, 512, 4, False), (2, 52, 64, 4, True)])
def test_mha_self(B, L, E, heads, masked):
    from oracle import rnn as orc
    from visuelle2_multimodal_fusion_b200 import functional as Fv
    g = torch.Generator().manual_seed(B + E)
    k = 1 / math.sqrt(E)
    P = {"in_proj_weight": torch.randn(3 * E, E, generator=g) * k, "in_proj_bias": torch.randn(3 * E, generator=g) * 0.1,
         "out_proj.weight": torch.randn(E, E, generator=g) * k, "out_proj.bias": torch.randn(E, generator=g) * 0.1}
    x = torch.randn(B, L, E, generator=g)
    mask = None
    if masked:
        mask = torch.full((L, L), float("-inf"))
        for i in range(0, L, 4):
            mask[i:i + 4, i:i + 4] = 0.0
    dO = torch.randn(B, L, E, generator=g)
    Pc = {k_: v.clone().requires_grad_(True) for k_, v in P.items()}
    xc = x.clone().requires_grad_(True)
    ref = orc.mha_self(xc.permute(1, 0, 2), Pc, "", heads, 0.0, False, attn_mask=mask).permute(1, 0, 2)
    ref.backward(dO)
    Pg = {k_: v.cuda().requires_grad_(True) for k_, v in P.items()}
    xg = x.cuda().requires_grad_(True)
    out = Fv.mha_self(xg, Pg["in_proj_weight"], Pg["in_proj_bias"], Pg["out_proj.weight"], Pg["out_proj.bias"],
                      heads, 0.0, False, mask=mask.cuda() if masked else None)
    out.backward(dO.cuda())
    assert _rel(out, ref) < TOL
    assert _rel(xg.grad, xc.grad) < TOL
    for k_ in P:
        assert _rel(Pg[k_].grad, Pc[k_].grad) < TOL, k_


def test_embed_with_mask_and_linear_relu():
    from visuelle2_multimodal_fusion_b200 import functional as Fv
    g = torch.Generator().manual_seed(1)
    B, E = 9, 32
    rows = [5, 3, 7, 11]
    tables = [torch.randn(r, E, generator=g).cuda().requires_grad_(True) for r in rows]
    idx = torch.stack([torch.randint(0, r, (B,), generator=g) for r in rows]).cuda()
    temporal = torch.rand(B, 4, generator=g).cuda()
    Wt = torch.randn(4, E, generator=g).cuda().requires_grad_(True)
    bt = torch.randn(4, E, generator=g).cuda().requires_grad_(True)
    drop = ((torch.rand(B, 8, E, generator=g) > 0.3).float() / 0.7).cuda()
    out = Fv.embed(temporal, Wt, bt, tables, idx, drop)
    d = sum(drop[:, k] * (temporal[:, k:k + 1] * Wt[k] + bt[k]) for k in range(4))
    a = sum(drop[:, 4 + k] * tables[k][idx[k]] for k in range(4))
    ref = torch.stack([d, a], 1)
    assert _rel(out, ref) < TOL
    dout = torch.randn(B, 2, E, generator=g).cuda()
    gr = torch.autograd.grad(ref, [Wt, bt] + tables, dout, retain_graph=True)
    gm = torch.autograd.grad(out, [Wt, bt] + tables, dout)
    for x, y in zip(gm, gr):
        assert _rel(x, y) < TOL
    # linear + relu epilogue
    x = torch.randn(17, 40, generator=g).cuda().requires_grad_(True)
    W = torch.randn(24, 40, generator=g).cuda().requires_grad_(True)
    b = torch.randn(24, generator=g).cuda().requires_grad_(True)
    y = Fv.linear(x, W, b, act=1)
    yr = torch.relu(x @ W.t() + b)
    dy = torch.randn_like(y)
    gm = torch.autograd.grad(y, [x, W, b], dy)
    gr = torch.autograd.grad(yr, [x, W, b], dy)
    assert _rel(y, yr) < TOL
    for p, q in zip(gm, gr):
        assert _rel(p, q) < TOL


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 3072, 512), (12800, 512, 2048), (300, 200, 520),
                                   (77, 48, 40), (256, 512, 512), (128, 1536, 1536)])
def test_gemm_tc(kind, M, N, K):
    """tcgen05 GEMM vs an fp64 matmul of the operands as the tensor core sees them
    (bf16-rounded, resp. tf32-truncated): only the fp32 accumulation order differs."""
    from visuelle2_multimodal_fusion_b200 import functional as Fv
    g = torch.Generator().manual_seed(M + N + K + kind)
    A = torch.randn(M, K, generator=g).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    bias = torch.randn(N, generator=g).cuda()
    C0 = torch.randn(M, N, generator=g).cuda()
    C = C0.clone()
    if kind == 0:
        Ab, Bb = Fv.cast_bf16(A), Fv.cast_bf16(B)
        assert torch.equal(Ab, A.bfloat16()) and torch.equal(Bb, B.bfloat16())
        Fv.gemm_tc(0, M, N, K, Ab, K, Bb, K, C, N, bias=bias, beta=1.0)
        Ar, Br = Ab.double(), Bb.double()
    else:
        Fv.gemm_tc(1, M, N, K, A, K, B, K, C, N, bias=bias, beta=1.0)
        trunc = lambda t: (t.view(torch.int32) & ~0x1FFF).view(torch.float32).double()
        Ar, Br = trunc(A), trunc(B)
        # act bit 2: operands rounded to nearest tf32 (ties away from zero, cvt.rna) instead of truncated
        C2 = C0.clone()
        Fv.gemm_tc(1, M, N, K, A, K, B, K, C2, N, bias=bias, beta=1.0, act=4)
        rna = lambda t: ((t.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32).double()
        torch.cuda.synchronize()
        assert _rel(C2, rna(A) @ rna(B).t() + bias.double() + C0.double()) < 2e-5
    torch.cuda.synchronize()
    ref = Ar @ Br.t() + bias.double() + C0.double()
    assert _rel(C, ref) < 2e-5


def test_gemm_tc_splitk_and_transpose():
    from visuelle2_multimodal_fusion_b200 import functional as Fv
    g = torch.Generator().manual_seed(5)
    M, N, K = 512, 2048, 12800
    dy = torch.randn(K, M, generator=g).cuda()          # [rows, M]
    x = torch.randn(K, N, generator=g).cuda().bfloat16()  # [rows, N] bf16
    dyT = Fv.transpose2d(dy, torch.bfloat16)             # [M, K]
    xT = Fv.transpose2d(x)                               # [N, K]
    assert torch.equal(dyT, dy.t().contiguous().bfloat16()) and torch.equal(xT, x.t().contiguous())
    C = torch.zeros(M, N, device="cuda")
    Fv.gemm_tc(0, M, N, K, dyT, K, xT, K, C, N, splits=4)
    torch.cuda.synchronize()
    ref = dyT.double() @ xT.double().t()
    assert _rel(C, ref) < 2e-5


@pytest.mark.parametrize("N,L,I,H", [(128, 52, 3, 512), (40, 5, 1, 64), (130, 3, 3, 32), (1280, 2, 1, 512), (300, 9, 2, 256)])
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_persistent_gru_equals_per_step_path(N, L, I, H, prec):
    """The one-launch cooperative GRU (W_hh resident in shared memory, csrc/gru_persist.cu) against the
    GEMM + gate kernel per step path of the same library: outputs and every gradient."""
    from visuelle2_multimodal_fusion_b200 import _lib
    from visuelle2_multimodal_fusion_b200 import functional as Fv
    g = torch.Generator().manual_seed(N + L + H)
    k = 1 / math.sqrt(H)
    mk = lambda *s: ((torch.rand(*s, generator=g) * 2 - 1) * k).cuda().requires_grad_(True)
    x = torch.randn(N, L, I, generator=g).cuda().requires_grad_(True)
    h0 = (torch.randn(N, H, generator=g) * 0.3).cuda().requires_grad_(True)
    P = [mk(3 * H, I), mk(3 * H, H), mk(3 * H), mk(3 * H)]
    d = torch.randn(N, L, H, generator=g).cuda()
    res = {}
    for on in (1, 0):
        _lib.lib().v2f_gru_persistent_enable(on)
        try:
            with Fv.precision(prec):
                out = Fv.gru_seq(x, h0, *P)
            out.backward(d)
            res[on] = [out.detach().clone()] + [t.grad.clone() for t in [x, h0] + P]
            for t in [x, h0] + P:
                t.grad = None
        finally:
            _lib.lib().v2f_gru_persistent_enable(1)
    for a, b in zip(res[1], res[0]):
        assert _rel(a, b) < (2e-5 if prec == "fp32" else 5e-3)     # tf32 (rounded vs truncated operands) in bf16 mode


def test_fused_dropout_statistics_and_backward_replays_the_mask():
    """v2f_dropout: keep rate 1 - p, kept values scaled by 1 / (1 - p), decisions a function of (key, index) only --
    the backward pass regenerates exactly the forward's mask; a different key gives a different mask."""
    import visuelle2_multimodal_fusion_b200.functional as Fv
    torch.manual_seed(3)
    x = (torch.rand(257, 1031, device="cuda") + 0.5).requires_grad_(True)       # odd sizes: tail elements
    for p in (0.1, 0.2, 0.5):
        y = Fv.dropout(x, p, True)
        kept = y != 0
        rate = float(kept.float().mean())
        assert abs(rate - (1 - p)) < 4 * (p * (1 - p) / x.numel()) ** 0.5 + 1e-4, (p, rate)
        assert torch.allclose(y[kept], x.detach()[kept] / (1 - p), rtol=1e-6)
        (g,) = torch.autograd.grad(y, x, torch.ones_like(y))
        assert torch.equal(g != 0, kept) and torch.allclose(g[kept], torch.full_like(g[kept], 1 / (1 - p)))
        y2 = Fv.dropout(x, p, True)
        assert float(((y2 != 0) != kept).float().mean()) > 0.05          # fresh key, fresh mask
        # no run of identical decisions along rows / columns (counter layout sanity)
        assert abs(float((kept[:, 1:] & kept[:, :-1]).float().mean()) - (1 - p) ** 2) < 0.01
    assert Fv.dropout(x, 0.3, False) is x
