"""GPU: the multi-tensor Adafactor step (csrc/adafactor.cu through optim.Adafactor) against
transformers.optimization.Adafactor on CPU -- the port of the fairseq optimizer the reference's
configure_optimizers builds (models/CrossAttnRNN210.py:229-230) -- on identical parameters and gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(7,), (1000,), (2500,), (33, 65), (512, 1300), (1536, 513), (8, 4, 3, 3), (16, 8, 1, 1), (5, 3, 4, 3),
          (2, 3, 40, 50), (1, 512), (28, 32)]


def _close(a, b, tol, what):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    d = float((a - b).abs().max())
    s = float(b.abs().max())
    assert d <= tol * s + 1e-30, f"{what}: max|diff| {d:.3e} scale {s:.3e}"


@pytest.mark.parametrize("kwargs", [dict(scale_parameter=True, relative_step=True, warmup_init=True, lr=None),
                                    dict(scale_parameter=False, relative_step=False, warmup_init=False, lr=1e-3),
                                    dict(scale_parameter=True, relative_step=True, warmup_init=False, lr=None)])
def test_fused_adafactor_matches_transformers(kwargs):
    from transformers.optimization import Adafactor as Ref
    from visuelle2_multimodal_fusion_b200.optim import Adafactor
    g = torch.Generator().manual_seed(5)
    ref_p = [torch.nn.Parameter(torch.randn(s, generator=g) * (0.02 + 0.3 * (i % 3))) for i, s in enumerate(SHAPES)]
    ref_p.append(torch.nn.Parameter(torch.randn(9, generator=g)))            # never receives a gradient
    # 4-D tensors with more than one channel: channels_last on the device, like the convolution weights of the bf16 trunk
    cl = lambda t: t.contiguous(memory_format=torch.channels_last) if (t.dim() == 4 and t.shape[1] > 1 and
                                                                       1 < t.shape[2] * t.shape[3] <= 16) else t
    our_p = [torch.nn.Parameter(cl(p.detach().clone().cuda())) for p in ref_p]
    assert any(not p.is_contiguous() for p in our_p)
    ref, ours = Ref(ref_p, **kwargs), Adafactor(our_p, **kwargs)
    for step in range(6):
        for i, (rp, op) in enumerate(zip(ref_p[:-1], our_p[:-1])):
            scale = 10.0 ** ((step % 3) - 2)                                   # exercises the update clipping
            gr = torch.randn(rp.shape, generator=g) * scale
            rp.grad = gr
            op.grad = gr.clone().cuda() if (step + i) % 2 else cl(gr.clone().cuda())   # either layout may arrive
        ref.step()
        ours.step()
        ref.zero_grad()
        ours.zero_grad()
    torch.cuda.synchronize()
    for i, (rp, op) in enumerate(zip(ref_p, our_p)):
        _close(op, rp, 2e-6, f"param {tuple(rp.shape)}")
        if i == len(ref_p) - 1:
            assert len(ours.state[op]) == 0
            continue
        rs, os_ = ref.state[rp], ours.state[op]
        assert os_["step"] == rs["step"] == 6
        # the reference takes ||p|| in fp32 (6e-6 off for 665k elements); the kernel accumulates it in double
        _close(os_["RMS"], torch.as_tensor(rs["RMS"]), 2e-5, f"RMS {tuple(rp.shape)}")
        for k in ("exp_avg_sq_row", "exp_avg_sq_col", "exp_avg_sq"):
            assert (k in rs) == (k in os_)
            if k in rs:
                assert os_[k].shape == rs[k].shape
                _close(os_[k], rs[k], 5e-6, f"{k} {tuple(rp.shape)}")


def test_fused_adafactor_state_dict_round_trip_with_reference_optimizer():
    """optimizer checkpoints are interchangeable: 2 reference steps -> state_dict -> 2 fused steps == 4 reference steps."""
    from transformers.optimization import Adafactor as Ref
    from visuelle2_multimodal_fusion_b200.optim import Adafactor
    kw = dict(scale_parameter=True, relative_step=True, warmup_init=True, lr=None)
    g = torch.Generator().manual_seed(8)
    shapes = [(40,), (24, 36), (6, 5, 3, 3)]
    ref_p = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    grads = [[torch.randn(s, generator=g) for s in shapes] for _ in range(4)]
    ref = Ref(ref_p, **kw)
    for t in range(2):
        for p, gr in zip(ref_p, grads[t]):
            p.grad = gr.clone()
        ref.step()
    our_p = [torch.nn.Parameter(p.detach().clone().cuda()) for p in ref_p]
    ours = Adafactor(our_p, **kw)
    ours.load_state_dict(ref.state_dict())
    for t in range(2, 4):
        for p, op, gr in zip(ref_p, our_p, grads[t]):
            p.grad = gr.clone()
            op.grad = gr.clone().cuda()
        ref.step()
        ours.step()
    for rp, op in zip(ref_p, our_p):
        _close(op, rp, 2e-6, f"param {tuple(rp.shape)}")


def test_fused_adafactor_parameters_with_different_ages():
    """A parameter that starts receiving gradients later has its own step count (relative step size, decay): the fused
    optimizer steps it as its own sub-group, as the per-parameter loop of the reference does."""
    from transformers.optimization import Adafactor as Ref
    from visuelle2_multimodal_fusion_b200.optim import Adafactor
    kw = dict(scale_parameter=True, relative_step=True, warmup_init=True, lr=None)
    g = torch.Generator().manual_seed(11)
    shapes = [(50,), (20, 30), (4, 3, 3, 3)]
    ref_p = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone().cuda()) for p in ref_p]
    ref, ours = Ref(ref_p, **kw), Adafactor(our_p, **kw)
    for t in range(5):
        for i, (rp, op) in enumerate(zip(ref_p, our_p)):
            if i == 1 and t < 2:               # the matrix joins at step 3
                rp.grad = op.grad = None
                continue
            gr = torch.randn(rp.shape, generator=g)
            rp.grad, op.grad = gr, gr.clone().cuda()
        ref.step()
        ours.step()
    assert [ours.state[p]["step"] for p in our_p] == [5, 3, 5] == [ref.state[p]["step"] for p in ref_p]
    for rp, op in zip(ref_p, our_p):
        _close(op, rp, 2e-6, f"param {tuple(rp.shape)}")


def test_fused_adafactor_matches_the_oracle():
    """CUDA path vs oracle/optim.py (numpy restatement, pinned to transformers' Adafactor on CPU in
    tests/test_oracle_golden.py) on the same parameters and gradients."""
    from oracle import optim as oo
    from visuelle2_multimodal_fusion_b200.optim import Adafactor
    kw = dict(scale_parameter=True, relative_step=True, warmup_init=True, lr=None)
    g = torch.Generator().manual_seed(21)
    shapes = [(300,), (64, 48), (16, 8, 3, 3), (32, 16, 1, 1), (700, 33)]
    init = [torch.randn(s, generator=g) * 0.2 for s in shapes]
    ref = [t.numpy().copy() for t in init]
    states = [oo.adafactor_init(p) for p in ref]
    our_p = [torch.nn.Parameter(t.clone().cuda()) for t in init]
    ours = Adafactor(our_p, **kw)
    for t in range(4):
        grads = [torch.randn(s, generator=g) * 10.0 ** (t % 3 - 2) for s in shapes]
        for p, gr, st in zip(ref, grads, states):
            oo.adafactor_step(p, gr.numpy(), st, **kw)
        for op, gr in zip(our_p, grads):
            op.grad = gr.clone().cuda()
        ours.step()
    for rp, op, st in zip(ref, our_p, states):
        _close(op, torch.from_numpy(rp), 2e-6, f"param {rp.shape}")
        os_ = ours.state[op]
        for k in ("exp_avg_sq_row", "exp_avg_sq_col", "exp_avg_sq"):
            if k in st:
                _close(os_[k], torch.from_numpy(st[k]), 5e-6, f"{k} {rp.shape}")
