"""CPU: the C-ABI library loads and exports every symbol include/v2f.h declares; the drop-in modules
keep the reference's state_dict keys / shapes; the product refuses to compute without CUDA."""
import ctypes
import os
import re

import pytest
import torch

from helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "v2f.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(v2f_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from visuelle2_multimodal_fusion_b200 import build
    lib = ctypes.CDLL(build.build())
    names = _declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/v2f.h but not exported"
    assert lib.v2f_version() == 4


def test_decode_params_struct_matches_header():
    """Field order of the ctypes mirror == field order of the C struct."""
    from visuelle2_multimodal_fusion_b200._lib import DecodeParams
    src = open(os.path.join(ROOT, "include", "v2f.h")).read()
    body = src[src.index("typedef struct v2f_decode_params {"):src.index("} v2f_decode_params;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S).split("{", 1)[1]
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(const\s+)?(unsigned|int|float|long long)\s*", "", decl)
        fields += [f.strip().lstrip("*").strip() for f in decl.split(",")]
    assert fields == [f[0] for f in DecodeParams._fields_]
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as d:                      # and the same size as the C compiler gives it
        c = os.path.join(d, "s.c")
        open(c, "w").write('#include <stdio.h>\n#include "v2f.h"\n'
                           'int main(void){printf("%zu", sizeof(v2f_decode_params));return 0;}\n')
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", os.path.join(d, "s")])
        assert int(subprocess.check_output([os.path.join(d, "s")]).decode()) == ctypes.sizeof(DecodeParams)


@pytest.mark.parametrize("name", ["rnn210_small", "rnn21_small", "demand_small", "gtm_demand_eval", "gtm_ar_eval",
                                  "v1_demand_train", "v2_demand_train", "v3_demand_train", "v4_demand_train",
                                  "m4ft_demand_train"])
def test_state_dict_keys_match_reference(name):
    from helpers import product_model
    blob = load_golden(name)
    m = product_model(blob, device="cpu")
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    ref = {k: tuple(v.shape) for k, v in blob["state"].items()}
    assert mine == ref


def test_no_cpu_fallback():
    from helpers import product_model, product_run
    blob = load_golden("rnn210_notf")
    m = product_model(blob, device="cpu")
    with pytest.raises(RuntimeError, match="CUDA only"):
        product_run(m, blob, device="cpu")


def test_same_seed_same_init_as_reference():
    """Construction order mirrors the reference => identical default initialisation under a seed."""
    from oracle import refshim
    if not refshim.reference_available():
        pytest.skip("reference tree not mounted")
    import torch.nn as nn
    import visuelle2_multimodal_fusion_b200.synth as synth
    import visuelle2_multimodal_fusion_b200.models.modules as mods
    from visuelle2_multimodal_fusion_b200.models import CrossAttnRNN210
    cat_d, col_d, fab_d = synth.label_dicts()
    ref_mod = refshim.load_reference_module("CrossAttnRNN210")
    import torchvision.models as tvm
    orig_tv = tvm.resnet101
    tvm.resnet101 = lambda *a, **k: nn.Sequential(nn.Identity(), nn.Identity(), nn.Identity())
    orig = mods.resnet101_trunk
    mods.resnet101_trunk = lambda: nn.Identity()
    try:
        torch.manual_seed(21)
        r = ref_mod.CrossAttnRNN(32, 32, 48, cat_d, col_d, fab_d, synth.STORE_N, 3)
        torch.manual_seed(21)
        m = CrossAttnRNN210.CrossAttnRNN(32, 32, 48, cat_d, col_d, fab_d, synth.STORE_N, 3)
    finally:
        tvm.resnet101 = orig_tv
        mods.resnet101_trunk = orig
    rs, ms = r.state_dict(), m.state_dict()
    assert set(rs) == set(ms)
    for k in rs:
        assert torch.equal(rs[k], ms[k]), k


@pytest.mark.parametrize("variant", ["gtm", "v1", "v2", "v3", "v4", "m4ft"])
def test_gtm_family_same_seed_same_init_as_reference(variant):
    """GTM family: same constructor arguments + same seed => bit-identical parameters and buffers."""
    from oracle import refshim
    if not refshim.reference_available():
        pytest.skip("reference tree not mounted")
    import visuelle2_multimodal_fusion_b200.synth as synth
    from helpers import _restore_gtm_trunk, gtm_product_ctor
    from oracle.make_golden_gtm import build_reference
    cat_d, col_d, fab_d = synth.label_dicts()
    import torch.nn as nn
    import torchvision.models as tvm
    refshim.load_reference_module("GTM_Visuelle2")           # installs the shims once
    orig_tv = tvm.resnet101
    fake = lambda *a, **k: nn.Sequential(nn.Identity(), nn.Identity(), nn.Identity())   # consumes no RNG
    fake._v2f_patched = True
    tvm.resnet101 = fake
    try:
        torch.manual_seed(5)
        r = build_reference(variant, 8, 16, 12, 4, False, "image")
    finally:
        tvm.resnet101 = orig_tv
    try:
        ctor = gtm_product_ctor(variant)
        torch.manual_seed(5)
        m = ctor(8, 16, 12, 4, 1, 1, 1, cat_d, col_d, fab_d, synth.STORE_N, 52, 3, 0, use_encoder_mask=1,
                 autoregressive=False)
    finally:
        _restore_gtm_trunk()
    rs = {k: v for k, v in r.state_dict().items() if not k.startswith("image_encoder.cnn")}
    ms = {k: v for k, v in m.state_dict().items() if not k.startswith("image_encoder.cnn")}
    assert set(rs) == set(ms)
    for k in rs:
        assert torch.equal(rs[k], ms[k]), k


def test_fused_adafactor_interface_and_no_cpu_fallback():
    """optim.Adafactor keeps the constructor contract of the fairseq / transformers optimizer and refuses CPU tensors."""
    import pytest
    import torch
    from visuelle2_multimodal_fusion_b200.optim import Adafactor
    p = torch.nn.Parameter(torch.zeros(4, 4))
    with pytest.raises(ValueError):
        Adafactor([p], lr=1e-3, relative_step=True)
    with pytest.raises(ValueError):
        Adafactor([p], lr=1e-3, relative_step=False, warmup_init=True)
    with pytest.raises(NotImplementedError):
        Adafactor([p], beta1=0.9)
    opt = Adafactor([p], scale_parameter=True, relative_step=True, warmup_init=True, lr=None)
    assert opt.param_groups[0]["lr"] is None
    opt.step()                                   # no gradient anywhere: nothing to do
    p.grad = torch.ones(4, 4)
    with pytest.raises(RuntimeError):
        opt.step()                               # CPU parameter: there is no fallback


def test_persistent_decoder_column_ownership_is_a_partition():
    """Host logic of csrc/decode_persist.cu (no GPU): over the CTAs of a grid, the owned ranges tile every output
    column of the three weight-stationary products exactly once, and the B200 grid (148 SMs) fits the shared-memory
    envelope (n1 <= 24, n3 <= 8, n5 <= 16, nu <= 4) at the reference's dims."""
    from visuelle2_multimodal_fusion_b200 import build
    lib = ctypes.CDLL(build.build())
    out = (ctypes.c_int * 11)()
    for G, E, H in [(148, 512, 512), (148, 256, 256), (132, 512, 512), (160, 512, 448), (2, 256, 64)]:
        s_cols, units, ctx_cols = [], [], []
        hc = {0: [], 1: []}
        for c in range(G):
            assert lib.v2f_decode_persist_ownership(c, G, E, H, out) == 0
            a_lo, na, u_lo, nu, n1, m3, e_lo3, n3, x_lo, nx, n5 = list(out)
            assert n1 == na + 3 * nu and n5 == nx + 3 * nu and m3 == c % 2
            s_cols += list(range(a_lo, a_lo + na))
            units += list(range(u_lo, u_lo + nu))
            ctx_cols += list(range(x_lo, x_lo + nx))
            hc[m3] += list(range(e_lo3, e_lo3 + n3))
            if (G, E, H) == (148, 512, 512):
                assert n1 <= 24 and n3 <= 8 and n5 <= 16 and 1 <= nu <= 4
        assert s_cols == list(range(3 * E)) and units == list(range(H)) and ctx_cols == list(range(E))
        assert hc[0] == list(range(E)) and hc[1] == list(range(E))
    assert lib.v2f_decode_persist_ownership(5, 1, 512, 512, out) != 0          # bad grid
    assert lib.v2f_decode_persist_ws_floats(128, 512, 512, 10) > 3 * 512 * 512


def test_batched_weight_cast_host_logic():
    """trunk._bf16_weights / _CastWeights (host logic, runs on CPU tensors too): trainable weights are cast by one
    multi-tensor copy whose backward returns fp32 gradients equal to the per-convolution ``weight.to(bf16)`` path that
    autocast takes; frozen weights are cached until they change."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    from visuelle2_multimodal_fusion_b200 import trunk
    torch.manual_seed(0)
    convs = [nn.Conv2d(4, 6, 3, padding=1, bias=False), nn.Conv2d(6, 5, 1, bias=False), nn.Conv2d(5, 3, 3, bias=False)]
    convs[1].weight.requires_grad_(False)                        # a frozen one in the middle
    x = torch.randn(2, 4, 8, 8).to(torch.bfloat16)

    def run(weights):
        y = x
        for c, w in zip(convs, weights):
            y = F.conv2d(y, w, None, c.stride, c.padding, c.dilation, c.groups)
        return y.float().square().sum()

    w16 = trunk._bf16_weights(convs)
    assert all(w16[c].dtype == torch.bfloat16 and w16[c].shape == c.weight.shape for c in convs)
    assert w16[convs[1]] is trunk._bf16_weights(convs)[convs[1]]            # frozen: cached
    run([w16[c] for c in convs]).backward()
    got = [c.weight.grad.clone() if c.weight.grad is not None else None for c in convs]
    for c in convs:
        c.weight.grad = None
    run([c.weight.to(torch.bfloat16) for c in convs]).backward()
    for c, g in zip(convs, got):
        if not c.weight.requires_grad:
            assert g is None and c.weight.grad is None
        else:
            assert g.dtype == torch.float32 and torch.equal(g, c.weight.grad)
    with torch.no_grad():
        assert all(not w.requires_grad for w in trunk._bf16_weights(convs).values())   # inference: plain cached copies


@pytest.mark.parametrize("cname,mirror", [("v2f_af_desc", "AfDesc"), ("v2f_adafactor_plan", "AfPlan")])
def test_adafactor_structs_match_header(cname, mirror):
    """Field order AND size of the ctypes mirrors of the optimizer's ABI structs == the C declarations (the size is
    checked by compiling a two-line C program against include/v2f.h with the host compiler)."""
    import subprocess
    import tempfile
    from visuelle2_multimodal_fusion_b200 import _lib
    cls = getattr(_lib, mirror)
    src = open(os.path.join(ROOT, "include", "v2f.h")).read()
    body = src[src.index("typedef struct %s {" % cname):src.index("} %s;" % cname)]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S).split("{", 1)[1]
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(const\s+)?(v2f_af_desc|unsigned|int|float|double|long long|void)\s*", "", decl)
        fields += [re.sub(r"\[\d+\]$", "", f.strip().lstrip("*").strip()) for f in decl.split(",")]
    assert fields == [f[0] for f in cls._fields_]
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write('#include <stdio.h>\n#include "v2f.h"\nint main(void){printf("%%zu", sizeof(%s));return 0;}\n' % cname)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        assert int(subprocess.check_output([exe]).decode()) == ctypes.sizeof(cls)


def test_lazy_weight_casts_run_their_backward_with_their_own_layers():
    """trunk._LazyCasts: the multi-tensor cast of a group of trainable convolution weights is created at the group's
    FIRST USE, so the autograd engine (which runs ready nodes in reverse order of creation) produces the group's fp32
    weight gradients right after the group's earliest layer -- not after the whole backward, where every gradient
    bucket of the trunk would be all-reduced in a tail (ddp.GradReducer)."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    from visuelle2_multimodal_fusion_b200 import trunk
    torch.manual_seed(1)
    convs = [nn.Conv2d(3, 4, 3, padding=1, bias=False) for _ in range(1)] + \
            [nn.Conv2d(4, 4, 3, padding=1, bias=False) for _ in range(5)]
    old = trunk.CAST_GROUP_ELEMS
    trunk.CAST_GROUP_ELEMS = 2 * convs[1].weight.numel()          # groups of two layers
    try:
        events = []
        for i, c in enumerate(convs):
            c.weight.register_post_accumulate_grad_hook(lambda p, i=i: events.append(("wgrad", i)))
        w16 = trunk._bf16_weights(convs)
        assert isinstance(w16, trunk._LazyCasts) and len(w16) == 0          # nothing cast yet
        x = torch.randn(2, 3, 6, 6).to(torch.bfloat16).requires_grad_(True)
        y = x
        for i, c in enumerate(convs):
            y = F.conv2d(y, w16.get(c), None, c.stride, c.padding)
            y.register_hook(lambda g, i=i: events.append(("act", i)))        # fires when layer i+1's backward is done
        y.float().square().sum().backward()
    finally:
        trunk.CAST_GROUP_ELEMS = old
    pos = {e: k for k, e in enumerate(events)}
    groups = {id(g): [convs.index(c) for c in g] for g in w16.group_of.values()}.values()
    assert len(groups) >= 3
    for g in groups:
        m = min(g)
        if m == 0:
            continue
        # ("act", i) fires when layer i's backward is about to start: every weight gradient of the group is out after
        # its earliest layer m has run and BEFORE layer m-1's backward starts
        assert all(pos[("act", m)] < pos[("wgrad", j)] < pos[("act", m - 1)] for j in g), (g, events)
    ref = [c.weight.grad.clone() for c in convs]
    for c in convs:
        c.weight.grad = None
    y = x
    for c in convs:
        y = F.conv2d(y, c.weight.to(torch.bfloat16), None, c.stride, c.padding)
    y.float().square().sum().backward()
    for c, g in zip(convs, ref):
        assert torch.equal(g, c.weight.grad)


def test_stem_weight_packing_is_the_k_layout_of_the_header():
    """trunk._stem_packed_weight: wpk[o, kh*24 + kw*3 + c] = w[o,c,kh,kw], zero elsewhere (include/v2f.h,
    v2f_stem_conv_fwd) -- and a GEMM over patches gathered the way csrc/stem_conv.cu gathers them (whole rows staged
    with the zero padding baked in at offset 9, 24-value windows starting at 6*ow) IS the 7x7 / stride 2 / pad 3
    convolution."""
    import torch
    import torch.nn as nn
    from visuelle2_multimodal_fusion_b200 import trunk
    torch.manual_seed(2)
    conv = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
    pk = trunk._stem_packed_weight(conv)
    assert pk.shape == (64, 192) and pk.dtype == torch.bfloat16
    w = conv.weight.detach().to(torch.bfloat16)
    for o, c, kh, kw in [(0, 0, 0, 0), (5, 2, 3, 6), (63, 1, 6, 2)]:
        assert pk[o, kh * 24 + kw * 3 + c] == w[o, c, kh, kw]
    mask = torch.ones(192, dtype=torch.bool)
    for kh in range(7):
        mask[kh * 24:kh * 24 + 21] = False
    assert bool((pk[:, mask] == 0).all())
    N, H, W = 1, 15, 17
    x = torch.randn(N, 3, H, W).to(torch.bfloat16)
    OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    rowlen = max(6 * (OW - 1) + 24, 9 + 3 * W)
    rows = torch.zeros(N * H + 1, rowlen, dtype=torch.float64)
    rows[:N * H, 9:9 + 3 * W] = x.double().permute(0, 2, 3, 1).reshape(N * H, 3 * W)
    out = torch.zeros(N, 64, OH, OW, dtype=torch.float64)
    for oh in range(OH):
        for ow in range(OW):
            a = torch.zeros(192, dtype=torch.float64)
            for kc in range(21):
                kh, c = divmod(kc, 3)
                ih = 2 * oh - 3 + kh
                src = rows[-1] if (ih < 0 or ih >= H) else rows[ih]
                a[kc * 8:kc * 8 + 8] = src[6 * ow + 8 * c:6 * ow + 8 * c + 8]
            out[0, :, oh, ow] = pk.double() @ a
    ref = torch.nn.functional.conv2d(x.double(), w.double(), None, 2, 3)
    assert float((out - ref).abs().max()) < 1e-9
