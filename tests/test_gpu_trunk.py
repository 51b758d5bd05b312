"""GPU parity of the fused BatchNorm2d(+add)(+ReLU) sweeps (csrc/bn_act.cu) and of the fused trunk walk
against torch's own BatchNorm2d / torchvision forward on the same bf16 channels_last inputs."""
import copy
import warnings

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
CL = torch.channels_last


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / max(float(b.double().abs().max()), 1e-30))


@pytest.mark.parametrize("N,C,H,W", [(4, 64, 9, 7), (8, 256, 19, 19), (3, 2048, 10, 10), (2, 192, 5, 5), (16, 64, 75, 75)])
@pytest.mark.parametrize("relu,res", [(True, False), (False, False), (True, True)])
@pytest.mark.parametrize("training", [True, False])
def test_bn_act_matches_torch(N, C, H, W, relu, res, training):
    from visuelle2_multimodal_fusion_b200 import trunk
    torch.manual_seed(N * C + H)
    bn = nn.BatchNorm2d(C).cuda()
    ref = nn.BatchNorm2d(C).cuda().double()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_()
        bn.running_mean.uniform_(-0.3, 0.3)
        bn.running_var.uniform_(0.5, 1.5)
        ref.load_state_dict({k: (v.double() if v.is_floating_point() else v.clone()) for k, v in bn.state_dict().items()})
    bn.train(training)
    ref.train(training)
    x = (torch.randn(N, C, H, W, device="cuda") * 1.7 + 0.4).bfloat16().contiguous(memory_format=CL).requires_grad_(True)
    r = torch.randn(N, C, H, W, device="cuda").bfloat16().contiguous(memory_format=CL).requires_grad_(True) if res else None
    dy = torch.randn(N, C, H, W, device="cuda").bfloat16().contiguous(memory_format=CL)
    y = trunk.bn_act(x, bn, relu=relu, res=r)
    assert y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=CL)
    y.backward(dy)
    xd = x.detach().double().requires_grad_(True)
    rd = r.detach().double().requires_grad_(True) if res else None
    z = ref(xd) + (rd if res else 0.0)
    yr = torch.relu(z) if relu else z
    yr.backward(dy.double())
    assert _rel(y, yr) < 8e-3                     # one bf16 ulp of the output
    assert _rel(bn.running_mean, ref.running_mean) < 1e-5 and _rel(bn.running_var, ref.running_var) < 1e-5
    assert int(bn.num_batches_tracked) == int(ref.num_batches_tracked)
    assert _rel(x.grad, xd.grad) < 1.2e-2
    assert _rel(bn.weight.grad, ref.weight.grad) < 5e-3 and _rel(bn.bias.grad, ref.bias.grad) < 5e-3
    if res:
        assert _rel(r.grad, rd.grad) < 8e-3


def _cos(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float(a @ b / (a.norm() * b.norm() + 1e-30))


def _trunks():
    import copy
    import warnings
    from visuelle2_multimodal_fusion_b200.models._base import resnet101_trunk
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        truth = resnet101_trunk().cuda().train()
    return truth, copy.deepcopy(truth).to(memory_format=CL), copy.deepcopy(truth).to(memory_format=CL)


def test_every_bottleneck_matches_torchvision_on_identical_inputs():
    """Each of the 33 Bottleneck blocks, fed the SAME bf16 input: fused walk vs torchvision's forward under
    autocast -- outputs within bf16 rounding, input / parameter gradients aligned."""
    from visuelle2_multimodal_fusion_b200 import trunk
    truth, cnn, ref = _trunks()
    x = torch.randn(8, 3, 299, 299, device="cuda")
    mt, mc, mr = list(truth.children()), list(cnn.children()), list(ref.children())
    with torch.no_grad():
        a = mt[3](mt[2](mt[1](mt[0](x))))
    n = 0
    for li in range(4, 8):
        for bi in range(len(mt[li])):
            xin = a.detach().bfloat16().contiguous(memory_format=CL)
            x1, x2 = xin.clone().requires_grad_(True), xin.clone().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y1 = trunk._bottleneck(mc[li][bi], x1)
                y2 = mr[li][bi](x2)
            g = torch.randn_like(y2)
            y1.backward(g)
            y2.backward(g)
            assert _rel(y1, y2) < 3e-2 and _cos(y1, y2) > 0.9999, (li, bi, _rel(y1, y2), _cos(y1, y2))
            assert _cos(x1.grad, x2.grad) > 0.999, (li, bi, _cos(x1.grad, x2.grad))
            for (k, p), (_, q) in zip(mc[li][bi].named_parameters(), mr[li][bi].named_parameters()):
                if q.grad is None:
                    assert p.grad is None, k
                    continue
                assert _cos(p.grad, q.grad) > 0.995, (li, bi, k, _cos(p.grad, q.grad))
                p.grad = q.grad = None
            with torch.no_grad():
                a = mt[li][bi](a)
            n += 1
    assert n == 33


def test_fused_trunk_matches_torchvision_forward_backward():
    """Whole ResNet-101 trunk, same weights, three executions: fp32 torchvision (the truth), torchvision under
    bf16 autocast, and the fused walk.  A 101-layer random-init network in train mode amplifies bf16 rounding
    chaotically (torch's own bf16 run ends at cos 0.63 against fp32), so the whole-trunk criterion is relative:
    the fused walk must be at least as close to the fp32 truth as torch's bf16 execution is."""
    from visuelle2_multimodal_fusion_b200 import trunk
    truth, cnn, ref = _trunks()
    assert trunk.supported(cnn)
    x = torch.randn(8, 3, 299, 299, device="cuda")
    g = torch.randn(8, 2048, 10, 10, device="cuda")
    ft = truth(x)
    ft.backward(g)
    f = trunk.forward(cnn, x)
    f.backward(g.bfloat16())
    with torch.autocast("cuda", dtype=torch.bfloat16):
        fr = ref(x.contiguous(memory_format=CL))
    fr.backward(g.bfloat16())
    assert f.shape == fr.shape == (8, 2048, 10, 10) and f.dtype == torch.bfloat16
    c_fused, c_torch = _cos(f, ft), _cos(fr, ft)
    print(f"feature map cos vs fp32: fused {c_fused:.5f}  torch-bf16 {c_torch:.5f}")
    assert c_fused >= c_torch - 5e-3
    st, sd, sr = truth.state_dict(), cnn.state_dict(), ref.state_dict()
    for k in sd:
        if k.endswith("running_mean") or k.endswith("running_var"):
            e_f, e_t = _rel(sd[k], st[k]), _rel(sr[k], st[k])
            assert e_f <= 1.5 * e_t + 2e-2, (k, e_f, e_t)
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(st[k]) == 1
    # gradients of the whole chaotic stack are decorrelated from the fp32 truth for BOTH bf16 executions
    # (cos ~ 0.03), so they are checked block by block above; here only which parameters receive one
    pt = dict(truth.named_parameters())
    for n, p in cnn.named_parameters():
        assert (p.grad is None) == (pt[n].grad is None), n
        assert p.grad is None or bool(torch.isfinite(p.grad).all()), n
    # frozen part (conv1 .. layer2) gets no gradient, as in the reference
    assert all(p.grad is None for n, p in cnn.named_parameters() if n.split(".")[0] in ("0", "1", "4", "5"))


@pytest.mark.parametrize("N,H,W", [(3, 150, 150), (2, 9, 7), (4, 16, 16)])
@pytest.mark.parametrize("training", [True, False])
def test_fused_stem_matches_torch(N, H, W, training):
    """conv -> BN -> ReLU -> maxpool(3,2,1) in one sweep (frozen stem) vs torch on the same bf16 conv output."""
    from visuelle2_multimodal_fusion_b200 import trunk
    torch.manual_seed(H)
    conv = nn.Conv2d(3, 64, 7, 2, 3, bias=False).cuda().to(memory_format=CL)
    bn = nn.BatchNorm2d(64).cuda()
    pool = nn.MaxPool2d(3, 2, 1)
    for p in list(conv.parameters()) + list(bn.parameters()):
        p.requires_grad = False
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_()
    bn.train(training)
    import copy
    ref = copy.deepcopy(bn).double()
    x = torch.randn(N, 3, 2 * H, 2 * W, device="cuda").contiguous(memory_format=CL)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = trunk.stem(conv, bn, pool, x)
        c = conv(x)
    yr = pool(torch.relu(ref(c.double())))
    assert y.shape == yr.shape and y.dtype == torch.bfloat16
    assert _rel(y, yr) < 8e-3
    assert _rel(bn.running_mean, ref.running_mean) < 1e-5 and _rel(bn.running_var, ref.running_var) < 1e-5
    assert int(bn.num_batches_tracked) == int(ref.num_batches_tracked)


@pytest.mark.parametrize("res", [False, True])
def test_bn_act_fork_sums_the_two_consumer_gradients(res):
    """bn_act_fork returns one storage twice; the two upstream gradients are summed inside the backward sweep."""
    from visuelle2_multimodal_fusion_b200 import trunk
    torch.manual_seed(3)
    N, C, H, W = 6, 256, 7, 5
    bn = nn.BatchNorm2d(C).cuda().train()
    mk = lambda: torch.randn(N, C, H, W, device="cuda").bfloat16().contiguous(memory_format=CL)
    x = mk().requires_grad_(True)
    r = mk().requires_grad_(True) if res else None
    g1, g2 = mk(), mk()
    y1, y2 = trunk.bn_act_fork(x, bn, relu=True, res=r)
    assert y1.data_ptr() == y2.data_ptr()
    (y1.float() * g1.float()).sum().backward(retain_graph=True)          # only the first edge
    only_first = x.grad.clone()
    x.grad = None
    bn.zero_grad()
    if res:
        r.grad = None
    ((y1.float() * g1.float()).sum() + (y2.float() * g2.float()).sum()).backward()
    fork = [x.grad.clone(), bn.weight.grad.clone(), bn.bias.grad.clone()] + ([r.grad.clone()] if res else [])
    x2 = x.detach().clone().requires_grad_(True)
    r2 = r.detach().clone().requires_grad_(True) if res else None
    bn.zero_grad()
    y = trunk.bn_act(x2, bn, relu=True, res=r2)
    y.backward((g1.float() + g2.float()).bfloat16())
    plain = [x2.grad, bn.weight.grad, bn.bias.grad] + ([r2.grad] if res else [])
    for a, b in zip(fork, plain):
        assert _rel(a, b) < 1.2e-2                     # bf16 rounding of (g1+g2) happens at a different point
    y.backward(g1, inputs=[x2]) if False else None
    x3 = x.detach().clone().requires_grad_(True)
    trunk.bn_act(x3, bn, relu=True, res=r.detach() if res else None).backward(g1)
    assert _rel(only_first, x3.grad) < 1e-6


def test_frozen_weight_cache_does_not_outlive_its_parameter():
    """trunk._bf16_weights caches the bf16 copies of frozen convolution weights; a second model built after the
    first was freed must never be served the first model's copies (its parameters may reuse the freed ids)."""
    import gc
    import torch.nn as nn
    from visuelle2_multimodal_fusion_b200 import trunk
    for rep in range(6):
        cin, cout = (8, 16) if rep % 2 == 0 else (32, 8)
        convs = [nn.Conv2d(cin, cout, 3, bias=False).cuda() for _ in range(40)]
        for c in convs:
            c.weight.requires_grad_(False)
        got = trunk._bf16_weights(convs)
        for c in convs:
            assert got[c].shape == c.weight.shape and got[c].dtype == torch.bfloat16
            assert torch.equal(got[c], c.weight.to(torch.bfloat16))
        again = trunk._bf16_weights(convs)
        assert all(again[c] is got[c] for c in convs)                  # cached while the parameter is unchanged
        with torch.no_grad():
            convs[0].weight.add_(1.0)                                  # in-place update bumps the version
        assert torch.equal(trunk._bf16_weights(convs)[convs[0]], convs[0].weight.to(torch.bfloat16))
        del convs, got, again
        gc.collect()


def test_full_model_training_with_fused_trunk_tracks_the_autocast_trunk():
    """End-to-end criterion for the fused bf16 trunk: the full CrossAttnRNN210 (random-init ResNet-101 + head, bf16
    mode) trained for 30 Adafactor steps twice from the same initial state -- once with the fused BatchNorm/add/ReLU
    sweeps, once with the unmodified torchvision modules under bf16 autocast (same cuDNN convolutions).  The two loss
    curves must agree like two bf16 executions of the same network do."""
    import visuelle2_multimodal_fusion_b200.synth as synth
    from visuelle2_multimodal_fusion_b200.models.CrossAttnRNN210 import CrossAttnRNN
    from oracle.refshim import zero_dropout
    cat_d, col_d, fab_d = synth.label_dicts()
    steps, B = 30, 8
    batches = []
    for i in range(4):
        data, im = synth.make_batch(B, out_len=10, seed=300 + i, dense_sales=True)
        batches.append((tuple(t.cuda() for t in data), im.cuda()))
    curves = {}
    state0 = None
    for fused in (True, False):
        torch.manual_seed(77)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = CrossAttnRNN(512, 512, 512, cat_d, col_d, fab_d, synth.STORE_N, 3).cuda()
        if state0 is None:
            state0 = copy.deepcopy(m.state_dict())
        else:
            m.load_state_dict(state0)
        zero_dropout(m).train()
        m.on_train_epoch_start()
        m.image_encoder.use_bf16_backbone(True)
        m.image_encoder.fused_trunk = fused
        m.precision = "bf16"
        opt = m.configure_optimizers()[0]
        losses = []
        for s in range(steps):
            torch.manual_seed(4000 + s)
            loss = m.training_step(batches[s % 4], s)
            opt.zero_grad()
            loss.backward()
            opt.step()
            losses.append(float(loss))
        curves[fused] = torch.tensor(losses)
        del m, opt
    a, b = curves[True], curves[False]
    rel = ((a - b).abs() / b.abs().clamp_min(1e-6))
    print(f"fused vs autocast trunk, 30 steps: loss {float(a[0]):.5f}->{float(a[-1]):.5f} vs {float(b[0]):.5f}->{float(b[-1]):.5f}; "
          f"rel diff mean {float(rel.mean()):.4f} max {float(rel.max()):.4f}")
    assert float(rel.mean()) < 0.02 and float(rel.max()) < 0.06, (rel.mean(), rel.max())


def test_eval_forward_sees_optimizer_updates_of_trainable_trunk_weights():
    """optim.Adafactor updates parameters from a CUDA kernel (no Tensor._version bump); a no_grad
    forward after the step must use the NEW layer3/layer4 weights, not a cached bf16 copy of the old ones."""
    import visuelle2_multimodal_fusion_b200.synth as synth
    from visuelle2_multimodal_fusion_b200.models.CrossAttnRNN210 import CrossAttnRNN
    cat_d, col_d, fab_d = synth.label_dicts()
    torch.manual_seed(5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = CrossAttnRNN(64, 64, 64, cat_d, col_d, fab_d, synth.STORE_N, 3).cuda()
    m.image_encoder.use_bf16_backbone(True)
    m.precision = "bf16"
    data, im = synth.make_batch(4, out_len=10, seed=1)
    batch = (tuple(t.cuda() for t in data), im.cuda())
    trunk_params = [p for n, p in m.image_encoder.cnn.named_parameters() if p.requires_grad and p.dim() == 4]
    opt = m.configure_optimizers()[0]

    def feat():
        m.eval()
        with torch.no_grad():
            f = m.image_encoder(batch[1]).float().clone()
        return f

    f0 = feat()                               # first no_grad pass (what Lightning's sanity check does)
    m.train()
    for s in range(3):
        loss = m.training_step(batch, s)
        opt.zero_grad()
        loss.backward()
        opt.step()
    w_before = trunk_params[-1].detach().clone()
    f1 = feat()
    # reference: the same eval forward with every cache dropped
    from visuelle2_multimodal_fusion_b200 import trunk
    trunk._frozen_cache.clear() if hasattr(trunk._frozen_cache, "clear") else None
    f2 = feat()
    assert torch.equal(trunk_params[-1].detach(), w_before)
    assert float((f1 - f2).abs().max()) == 0.0, "eval forward depends on a stale weight cache"
    assert float((f1 - f0).abs().max()) > 0.0, "eval output did not change after three optimizer steps"


# ------------------------------------------------------------------ stem convolution on tcgen05 (csrc/stem_conv.cu)
@pytest.mark.parametrize("nchw", [False, True])
@pytest.mark.parametrize("N,H,W,in_bf16", [(3, 299, 299, False), (2, 299, 299, True), (2, 300, 300, False),
                                           (5, 256, 272, False), (1, 261, 341, True)])
def test_stem_conv_tcgen05_matches_exact_convolution(N, H, W, in_bf16, nchw):
    """conv1 (7x7, stride 2, pad 3, 3 -> 64) as the tcgen05 implicit GEMM: every output equals the exact (fp64)
    convolution of the bf16-rounded operands to within one bf16 rounding of the output, and the per-CTA partial
    statistics of the epilogue are the sums of the STORED values.  Odd / even sizes, partial last tile (N*OH*OW not a
    multiple of 128), image boundaries inside a tile, fp32 and bf16 inputs."""
    from visuelle2_multimodal_fusion_b200 import _lib, trunk
    torch.manual_seed(N * H + W)
    conv = nn.Conv2d(3, 64, 7, 2, 3, bias=False).cuda()
    x = torch.randn(N, 3, H, W, device="cuda") * 1.3 + 0.2
    if not nchw:
        x = x.contiguous(memory_format=CL)
    if in_bf16:
        x = x.bfloat16()
    assert x.is_contiguous() if nchw else x.is_contiguous(memory_format=CL)
    nblk = _lib.lib().v2f_stem_conv_blocks(N, H, W, 1 if in_bf16 else 0, 1 if nchw else 0)
    assert nblk > 0
    OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty((N, 64, OH, OW), device="cuda", dtype=torch.bfloat16, memory_format=CL)
    part = torch.full((nblk, 2, 64), float("nan"), device="cuda")
    pk = trunk._stem_packed_weight(conv)
    _lib.check(_lib.lib().v2f_stem_conv_fwd(N, H, W, x.data_ptr(), 1 if in_bf16 else 0, 1 if nchw else 0, pk.data_ptr(),
                                            y.data_ptr(), part.data_ptr(), _lib.stream()), "v2f_stem_conv_fwd")
    ref = torch.nn.functional.conv2d(x.bfloat16().double(), conv.weight.detach().bfloat16().double(), None, 2, 3)
    err = (y.double() - ref).abs()
    tol = ref.abs() * 2.0 ** -8 + 1e-3          # half an ulp of bf16 + fp32 accumulation noise near zero
    assert bool((err <= tol).all()), float((err - tol).max())
    s = y.double().sum((0, 2, 3))
    q = (y.double() ** 2).sum((0, 2, 3))
    ps = part.double().sum(0)
    assert _rel(ps[0], s) < 1e-5 and _rel(ps[1], q) < 1e-5
    # without statistics (eval): same outputs
    y2 = torch.empty_like(y)
    _lib.check(_lib.lib().v2f_stem_conv_fwd(N, H, W, x.data_ptr(), 1 if in_bf16 else 0, 1 if nchw else 0, pk.data_ptr(),
                                            y2.data_ptr(), None, _lib.stream()), "v2f_stem_conv_fwd")
    assert torch.equal(y, y2)


def test_stem_conv_unsupported_shapes_keep_the_library_convolution():
    from visuelle2_multimodal_fusion_b200 import _lib
    assert _lib.lib().v2f_stem_conv_blocks(4, 64, 64, 0, 0) == 0          # OW < 128
    assert _lib.lib().v2f_stem_conv_blocks(4, 299, 400, 0, 1) == 0        # 3 W > 1024


@pytest.mark.parametrize("training", [True, False])
def test_stem_with_tcgen05_conv_equals_stem_with_library_conv(training):
    """trunk.stem (conv1 -> bn1 -> relu -> maxpool) with conv1 on csrc/stem_conv.cu + statistics from its epilogue vs
    the same stem with the library convolution + statistics sweep: outputs within bf16 rounding, running statistics
    equal to fp32 noise."""
    import torchvision
    from visuelle2_multimodal_fusion_b200 import trunk
    torch.manual_seed(11)
    net = torchvision.models.resnet101(weights=None).cuda().train(training)
    for p in net.parameters():
        p.requires_grad_(False)
    with torch.no_grad():
        net.bn1.running_mean.uniform_(-0.2, 0.2)
        net.bn1.running_var.uniform_(0.5, 1.5)
    x = torch.randn(6, 3, 299, 299, device="cuda")          # NCHW, as a DataLoader collates it
    outs, stats = [], []
    state = copy.deepcopy(net.bn1.state_dict())
    for flag in (True, False):
        net.bn1.load_state_dict(state)
        trunk.STEM_CONV_TC = flag
        trunk._w16 = trunk._bf16_weights([net.conv1])
        try:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                outs.append(trunk.stem(net.conv1, net.bn1, net.maxpool, x).float())
        finally:
            trunk._w16 = None
            trunk.STEM_CONV_TC = True
        stats.append((net.bn1.running_mean.clone(), net.bn1.running_var.clone(), int(net.bn1.num_batches_tracked)))
    assert outs[0].shape == outs[1].shape == (6, 64, 75, 75)
    assert _rel(outs[0], outs[1]) < 1.6e-2 and _cos(outs[0], outs[1]) > 0.99999
    assert _rel(stats[0][0], stats[1][0]) < 1e-4 and _rel(stats[0][1], stats[1][1]) < 1e-4 and stats[0][2] == stats[1][2]
