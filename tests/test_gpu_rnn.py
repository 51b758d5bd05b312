"""GPU parity of the CUDA path (through the C ABI) against the reference-generated goldens and
against the oracle at the reference's default dims.  fp32 contract: rel err <= 1e-5."""
import pytest
import torch

from helpers import compare_blob, load_golden

pytestmark = pytest.mark.gpu
TOL = 1e-5
RNN_CASES = ["rnn210_small", "rnn210_notf", "rnn21_small", "demand_small", "demand_notf"]


TOL_TC = 2e-2   # bf16/tf32 tensor-core path (BASELINE.json north_star)


@pytest.mark.parametrize("name", RNN_CASES)
def test_cuda_tensorcore_path_matches_reference_golden(name):
    rows = compare_blob(load_golden(name), TOL_TC, precision="bf16")
    bad = [r for r in rows if not r[3]]
    assert not bad, "\n".join(f"{w}: rel={e:.3e} scale={s:.3e}" for w, e, s, _ in bad)


@pytest.mark.parametrize("name", RNN_CASES)
def test_cuda_matches_reference_golden(name):
    rows = compare_blob(load_golden(name), TOL)
    bad = [r for r in rows if not r[3]]
    assert not bad, "\n".join(f"{w}: rel={e:.3e} scale={s:.3e}" for w, e, s, _ in bad)


@pytest.mark.parametrize("precision,tol", [("fp32", TOL), ("bf16", TOL_TC)])
@pytest.mark.parametrize("model,T", [("CrossAttnRNN210", 10), ("CrossAttnRNNDemand", 12), ("CrossAttnRNN21", 1)])
def test_cuda_matches_oracle_default_dims(model, T, precision, tol):
    """E=A=H=512, Li=100, Lt=52 (train_dl.py:197-199) at a small batch; oracle on CPU is the checker."""
    from helpers import oracle_vs_cuda_default_dims
    oracle_vs_cuda_default_dims(model, T, precision, tol, B=4)


@pytest.mark.parametrize("model,T", [("CrossAttnRNN210", 10), ("CrossAttnRNNDemand", 12)])
def test_cuda_tensorcore_path_matches_oracle_default_dims_64_rows(model, T):
    """The same at 64 rows: the size class where the persistent decoder runs with full row groups."""
    from helpers import oracle_vs_cuda_default_dims
    oracle_vs_cuda_default_dims(model, T, "bf16", TOL_TC, B=64)


def test_streaming_attention_equals_simple_kernels():
    """The TMA-staged streaming attention (148-way split + partial-softmax combine) against the
    simple per-(row, modality) kernels, B=40 rows at the default dims."""
    import visuelle2_multimodal_fusion_b200.functional as Fv
    import visuelle2_multimodal_fusion_b200.synth as synth
    import torch.nn as nn
    import visuelle2_multimodal_fusion_b200.models.modules as mods
    from visuelle2_multimodal_fusion_b200.models import CrossAttnRNN210
    orig = mods.resnet101_trunk
    mods.resnet101_trunk = lambda: nn.Identity()
    try:
        torch.manual_seed(0)
        cat_d, col_d, fab_d = synth.label_dicts()
        m = CrossAttnRNN210.CrossAttnRNN(512, 512, 512, cat_d, col_d, fab_d, synth.STORE_N, 3).cuda().eval()
    finally:
        mods.resnet101_trunk = orig
    data, feat = synth.make_batch(40, out_len=10, seed=3, feat_hw=10)
    data = tuple(t.cuda() for t in data)
    res = {}
    for flag in (True, False):
        Fv.STREAM_ATTENTION = flag
        try:
            f = feat.cuda().clone().requires_grad_(True)
            torch.manual_seed(9)
            out, _ = m(*data, f)
            out.square().mean().backward()
            res[flag] = (out.detach().clone(), f.grad.clone(),
                         m.img_attention.encoder_linear.weight.grad.clone(), m.trend_linear.weight.grad.clone())
            m.zero_grad(set_to_none=True)
        finally:
            Fv.STREAM_ATTENTION = True
    for a, b in zip(res[True], res[False]):
        assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max()) + 1e-9


@pytest.mark.parametrize("model", ["CrossAttnRNN210", "CrossAttnRNNDemand"])
def test_cuda_graph_replay_equals_eager_step(model):
    """graphs.GraphedTrainStep: captured forward+loss+backward, replayed with fresh batches and fresh
    teacher-forcing draws, equals the eager training_step/backward (eval-mode dropout, same host RNG state)."""
    import torch.nn as nn
    import visuelle2_multimodal_fusion_b200.synth as synth
    import visuelle2_multimodal_fusion_b200.models.modules as mods
    from visuelle2_multimodal_fusion_b200.graphs import GraphedTrainStep
    from visuelle2_multimodal_fusion_b200.models import CrossAttnRNN210, CrossAttnRNNDemand
    orig = mods.resnet101_trunk
    mods.resnet101_trunk = lambda: nn.Identity()
    try:
        torch.manual_seed(0)
        cat_d, col_d, fab_d = synth.label_dicts()
        if model == "CrossAttnRNN210":
            m = CrossAttnRNN210.CrossAttnRNN(64, 64, 64, cat_d, col_d, fab_d, synth.STORE_N, 3, out_len=10)
        else:
            m = CrossAttnRNNDemand.CrossAttnRNN(64, 64, 3, 64, cat_d, col_d, fab_d, synth.STORE_N, True, True, True,
                                                True, out_len=12, use_teacher_forcing=True)
    finally:
        mods.resnet101_trunk = orig
    m = m.cuda().eval()
    m.use_teacher_forcing = True
    demand = model == "CrossAttnRNNDemand"

    def batch(seed):
        data, feat = synth.make_batch(16, out_len=10, demand=demand, seed=seed, feat_hw=4)
        return tuple(t.cuda() for t in data), feat.cuda()

    step = GraphedTrainStep(m, batch(1))
    got = []
    for seed in (2, 3, 4):
        torch.manual_seed(100 + seed)
        loss = step(batch(seed))
        got.append((loss.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}))
    step.release()
    for (gl, gg), seed in zip(got, (2, 3, 4)):
        torch.manual_seed(100 + seed)
        for p in m.parameters():
            p.grad = None
        loss = m.training_step(batch(seed), 0)
        loss.backward()
        assert float((loss - gl).abs()) <= 1e-6 * float(loss.abs()) + 1e-9
        for k, p in m.named_parameters():
            if p.grad is None:
                assert k not in gg
                continue
            d = float((p.grad - gg[k]).abs().max())
            assert d <= 1e-5 * float(p.grad.abs().max()) + 1e-9, (k, d)


def _head_model(model, E=512, T=None, H=None):
    import torch.nn as nn
    import visuelle2_multimodal_fusion_b200.synth as synth
    import visuelle2_multimodal_fusion_b200.models.modules as mods
    from visuelle2_multimodal_fusion_b200.models import CrossAttnRNN210, CrossAttnRNNDemand
    orig = mods.resnet101_trunk
    mods.resnet101_trunk = lambda: nn.Identity()
    try:
        torch.manual_seed(0)
        cat_d, col_d, fab_d = synth.label_dicts()
        if model == "CrossAttnRNN210":
            m = CrossAttnRNN210.CrossAttnRNN(E, E, H or E, cat_d, col_d, fab_d, synth.STORE_N, 3, out_len=T or 10)
        else:
            m = CrossAttnRNNDemand.CrossAttnRNN(E, E, 3, H or E, cat_d, col_d, fab_d, synth.STORE_N, True, True, True,
                                                True, out_len=T or 12, use_teacher_forcing=True)
    finally:
        mods.resnet101_trunk = orig
    return m.cuda().eval()


# tensor-core mode: the two paths are two DIFFERENT approximations (tf32 products per step vs bf16 operands and bf16 tiles
# in the persistent kernels), each within 2e-2 of the reference (tests/test_gpu_fullsize.py, the oracle tests above):
# against each other they are compared at twice that
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16", 4e-2)])
@pytest.mark.parametrize("model,B,E,T", [("CrossAttnRNN210", 128, 512, 10), ("CrossAttnRNNDemand", 128, 512, 12),
                                         ("CrossAttnRNN210", 40, 512, 10), ("CrossAttnRNN210", 150, 512, 10),
                                         ("CrossAttnRNN210", 24, 256, 10), ("CrossAttnRNN210", 1, 512, 10),
                                         ("CrossAttnRNN210", 9, 512, 5), ("CrossAttnRNNDemand", 3, 256, 12)])
def test_persistent_decoder_equals_step_per_launch_path(model, B, E, T, precision, tol):
    """csrc/decode_persist.cu (one cooperative launch for the whole horizon, weights resident in shared memory)
    against the step-per-launch path of rnn_decode.cu on the same inputs: forecasts, attention maps and every
    gradient (the backward consumes the activations the forward path saved)."""
    import visuelle2_multimodal_fusion_b200.functional as Fv
    import visuelle2_multimodal_fusion_b200.synth as synth
    from visuelle2_multimodal_fusion_b200 import _lib
    demand = model == "CrossAttnRNNDemand"
    m = _head_model(model, E, T)                  # T = 5: six windows per item (rows n index items as n // W)
    m.precision = precision
    m.use_teacher_forcing = True
    data, feat = synth.make_batch(B, out_len=T if not demand else 10, demand=demand, seed=5, feat_hw=10)
    data = tuple(t.cuda() for t in data)
    res = {}
    launches = {}
    for flag in (True, False):
        Fv.PERSISTENT_DECODE = flag
        try:
            f = feat.cuda().clone().requires_grad_(True)
            torch.manual_seed(9)
            n0 = _lib.launch_count()
            out = m(*data, f)[0]
            launches[flag] = _lib.launch_count() - n0
            out.square().mean().backward()
            res[flag] = [out.detach().clone(), f.grad.clone()] + \
                [p.grad.clone() for _, p in sorted(m.named_parameters()) if p.grad is not None]
            names = ["out", "grad_feat"] + [k for k, p in sorted(m.named_parameters()) if p.grad is not None]
            m.zero_grad(set_to_none=True)
        finally:
            Fv.PERSISTENT_DECODE = True
    assert launches[True] < launches[False] - 4 * T, launches      # the loop really collapsed into one launch
    # decoder_fc.bias sums d loss / d yhat over every row and step; with the test's loss (mean of squares of zero-mean
    # forecasts) that sum cancels to ~5 % of the sum of magnitudes, so it is compared at the scale of the summands
    fc_floor = tol * 2.0 * float(res[False][0].abs().mean())
    for k, a, b in zip(names, res[True], res[False]):
        floor = 1e-6 if k.endswith("attn_linear.bias") else (fc_floor if k == "decoder_fc.bias" else 1e-9)
        assert float((a - b).abs().max()) <= tol * float(b.abs().max()) + floor, (k, float((a - b).abs().max()),
                                                                                 float(b.abs().max()))


@pytest.mark.parametrize("model", ["CrossAttnRNN210", "CrossAttnRNNDemand"])
def test_graphed_forecast_equals_eager_eval_forward(model):
    """graphs.GraphedForecast: the reference's no-grad forecast loop (forecast_dl.py:123-171) replayed from a CUDA
    graph gives the eager eval forward's numbers and leaves the host RNG where the eager call leaves it."""
    import visuelle2_multimodal_fusion_b200.synth as synth
    from visuelle2_multimodal_fusion_b200.graphs import GraphedForecast
    demand = model == "CrossAttnRNNDemand"
    m = _head_model(model, 256)
    m.on_validation_epoch_start()

    def inputs(seed, B=6):
        data, feat = synth.make_batch(B, out_len=10, demand=demand, seed=seed, feat_hw=10)
        return tuple(t.cuda() for t in data) + (feat.cuda(),)

    fc = GraphedForecast(m, inputs(1))
    for seed in (2, 3):
        torch.manual_seed(40 + seed)
        got = fc(inputs(seed))[0].clone()
        after_graph = torch.rand(1)
        torch.manual_seed(40 + seed)
        with torch.no_grad():
            want = m(*inputs(seed))[0]
        after_eager = torch.rand(1)
        assert torch.equal(after_graph, after_eager)
        assert float((got - want).abs().max()) <= 1e-6 * float(want.abs().max()) + 1e-9
    torch.manual_seed(7)
    small = fc(inputs(5, B=3))[0]                 # other batch size: eager fallback
    assert small.shape[0] == 3


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_full_size_properties_of_the_decoder(precision, tol):
    """BASELINE.json's full size (B = 128 items, E = A = H = 512, 100 image positions, 52 trend steps, 12 steps), where
    the CPU oracle is too slow: properties that hold at any size.  (1) every attention map the Demand forward returns
    is a distribution; (2) items are independent: permuting the batch permutes the forecasts and the feature-map
    gradient (the work split of the streaming sweep moves with the row order, so to rounding, not bit for bit);
    (3) the loss gradient of an item that does not enter the loss is exactly zero."""
    import visuelle2_multimodal_fusion_b200.synth as synth
    m = _head_model("CrossAttnRNNDemand", 512)
    m.precision = precision
    m.on_validation_epoch_start()
    B = 128
    data, feat = synth.make_batch(B, out_len=10, demand=True, seed=17, feat_hw=10)
    data = tuple(t.cuda() for t in data)
    feat = feat.cuda()

    def run(d, f, weight):
        f = f.clone().requires_grad_(True)
        torch.manual_seed(3)
        out, ia, ma = m(*d, f)
        (out.squeeze(-1) * weight).square().sum().backward()
        g = f.grad.clone()
        m.zero_grad(set_to_none=True)
        return out.detach(), torch.stack(ia), torch.stack(ma), g

    w = torch.ones(B, 1, device="cuda")
    w[5] = 0.0                                     # item 5 does not enter the loss
    out, ia, ma, g = run(data, feat, w)
    assert ia.shape == (12, B, 100) and ma.shape == (12, B, 4)
    assert float((ia.sum(-1) - 1).abs().max()) < 1e-5 and float((ma.sum(-1) - 1).abs().max()) < 1e-5
    assert float(ia.min()) >= 0.0 and float(ma.min()) >= 0.0
    assert float(g[5].abs().max()) == 0.0 and float(g[4].abs().max()) > 0.0
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(2)).cuda()
    out_p, ia_p, ma_p, g_p = run(tuple(t[perm] for t in data), feat[perm], w[perm])
    assert float((out_p - out[perm]).abs().max()) <= tol * float(out.abs().max())
    assert float((ia_p - ia[:, perm]).abs().max()) <= tol
    assert float((g_p - g[perm]).abs().max()) <= tol * float(g.abs().max())


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16", 2e-2)])
@pytest.mark.parametrize("model,E,H", [("CrossAttnRNN210", 512, 256), ("CrossAttnRNN210", 256, 512),
                                       ("CrossAttnRNNDemand", 512, 256)])
def test_persistent_decoder_with_hidden_dim_different_from_embedding_dim(model, E, H, precision, tol):
    """hidden_dim != embedding_dim (the reference's constructors take them separately, train_dl.py:197-199 merely sets
    both to 512): the persistent decoder against the step-per-launch path."""
    import visuelle2_multimodal_fusion_b200.functional as Fv
    import visuelle2_multimodal_fusion_b200.synth as synth
    from visuelle2_multimodal_fusion_b200 import _lib
    demand = model == "CrossAttnRNNDemand"
    m = _head_model(model, E, None, H)
    m.precision = precision
    m.use_teacher_forcing = True
    data, feat = synth.make_batch(20, out_len=10, demand=demand, seed=6, feat_hw=10)
    data = tuple(t.cuda() for t in data)
    res, launches = {}, {}
    for flag in (True, False):
        Fv.PERSISTENT_DECODE = flag
        try:
            f = feat.cuda().clone().requires_grad_(True)
            torch.manual_seed(9)
            n0 = _lib.launch_count()
            out = m(*data, f)[0]
            launches[flag] = _lib.launch_count() - n0
            out.square().mean().backward()
            res[flag] = [out.detach().clone(), f.grad.clone()] + \
                [p.grad.clone() for _, p in sorted(m.named_parameters()) if p.grad is not None]
            names = ["out", "grad_feat"] + [k for k, p in sorted(m.named_parameters()) if p.grad is not None]
            m.zero_grad(set_to_none=True)
        finally:
            Fv.PERSISTENT_DECODE = True
    assert launches[True] < launches[False] - 40, launches
    # decoder_fc.bias sums d loss / d yhat over every row and step; with the test's loss (mean of squares of zero-mean
    # forecasts) that sum cancels to ~5 % of the sum of magnitudes, so it is compared at the scale of the summands
    fc_floor = tol * 2.0 * float(res[False][0].abs().mean())
    for k, a, b in zip(names, res[True], res[False]):
        floor = 1e-6 if k.endswith("attn_linear.bias") else (fc_floor if k == "decoder_fc.bias" else 1e-9)
        assert float((a - b).abs().max()) <= tol * float(b.abs().max()) + floor, (k, float((a - b).abs().max()),
                                                                                 float(b.abs().max()))
