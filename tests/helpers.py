"""Shared helpers for the parity tests (oracle side + comparison metric)."""
import os

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def rel_err(a, b):
    """max|a-b| / max(|b|, tiny), per tensor."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def assert_close(a, b, tol, what, floor=1e-7):
    """rel_err <= tol, except that a tensor whose reference magnitude is numerical noise
    (e.g. the softmax-invariant attn_linear.bias gradient) is compared absolutely against ``floor``."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    diff = float((a - b).abs().max()) if a.numel() else 0.0
    scale = float(b.abs().max()) if b.numel() else 0.0
    assert diff <= tol * scale + floor, f"{what}: max|diff|={diff:.3e} scale={scale:.3e} rel={diff / max(scale, 1e-30):.3e} > {tol}"


def noise_only_grads(blob):
    """Parameters whose gradient is exactly zero in exact arithmetic (both the reference's autograd and
    ours return rounding noise there): a bias added right before train-mode BatchNorm (GTM fusion) and the
    attn_linear biases feeding a softmax (RNN family).  Compared absolutely."""
    names = set()
    if blob["model"] == "GTM:gtm" and blob["cfg"]["mode"] != "eval":
        names |= {"image_encoder.projection.bias", "dummy_encoder.dummy_fusion.bias"}
        names |= {f"dummy_encoder.{n}_emb.bias" for n in ("day", "week", "month", "year")}
    if blob["model"] == "GTM:m4ft" and blob["cfg"]["mode"] != "eval":
        # every static embedding reaches the decoder only through train-mode BatchNorm1d layers (M4FT fusion blocks)
        names |= {"text_encoder.proj.bias", "temporal_encoder.proj.bias", "image_encoder.final_proj.bias",
                  "image_encoder.projection.bias", "fusion_network.fusion_temp_text.net.4.bias",
                  "fusion_network.fusion_text_vis.net.4.bias"}
        names |= {f"temporal_encoder.{n}_emb.bias" for n in ("day", "week", "month", "year")}
    names |= {k for k in blob["grads"] if k.endswith("attn_linear.bias")}
    # a key bias shifts every score of a softmax row equally (Proposed_model_v2's separate k_proj)
    names |= {k for k in blob["state"] if k.endswith("k_proj.bias")}
    return names


def oracle_run(blob, requires_grad=True):
    """Run the oracle on a golden blob's inputs/weights; returns (out, extras, P, feat)."""
    from oracle import rnn
    P = {k: v.clone().requires_grad_(requires_grad and v.is_floating_point()) for k, v in blob["state"].items()}
    inp = blob["inputs"]
    feat = inp["feat"].clone().requires_grad_(requires_grad)
    cfg = blob["cfg"]
    model = blob["model"]
    extras = {}
    if model == "CrossAttnRNN210":
        out, _ = rnn.rnn210_forward(P, inp["X"], inp["y"], inp["cat"], inp["col"], inp["fab"], inp["store"],
                                    inp["temporal"], inp["gtrends"], feat, out_len=cfg["T"],
                                    use_teacher_forcing=cfg["tf"], tf_mask=blob["tf_mask"])
        loss = torch.nn.functional.mse_loss(inp["y"].reshape(out.shape), out)
    elif model == "CrossAttnRNN21":
        out, _ = rnn.rnn21_forward(P, inp["X"], inp["y"], inp["cat"], inp["col"], inp["fab"], inp["store"],
                                   inp["temporal"], inp["gtrends"], feat)
        loss = torch.nn.functional.mse_loss(inp["y"], out)
    elif model == "CrossAttnRNNDemand":
        out, ia, ma = rnn.demand_forward(P, inp["ts"], inp["cat"], inp["col"], inp["fab"], inp["store"],
                                         inp["temporal"], inp["gtrends"], feat, out_len=cfg["T"],
                                         use_teacher_forcing=cfg["tf"], tf_mask=blob["tf_mask"])
        extras = dict(img_alphas=torch.stack(ia), mm_alphas=torch.stack(ma))
        loss = torch.nn.functional.mse_loss(inp["ts"], out.squeeze())
    elif model.startswith("GTM:"):
        from oracle import gtm
        train = cfg["mode"] != "eval"
        out, _ = gtm.gtm_family_forward(model[4:], P, inp["item_sales"], inp["cat"], inp["col"], inp["fab"],
                                        inp["store"], inp["temporal"], inp["gtrends"], feat, output_len=cfg["T"],
                                        heads=cfg["heads"], autoregressive=cfg["autoregressive"], training=train,
                                        drop=False, query_modality=cfg["query_modality"])
        loss = torch.nn.functional.mse_loss(inp["y"].reshape(-1), out.reshape(-1))
    else:
        raise KeyError(model)
    return out, loss, extras, P, feat


# ------------------------------------------------------------------ product side (CUDA only)
_GTM_TRUNK_PATCH = []


def gtm_product_ctor(variant):
    """The GTM-family drop-in class for a variant, with the backbone factory patched to identity."""
    import importlib
    import torch.nn as nn
    modname, clsname = {"gtm": ("GTM_Visuelle2", "GTM_Visuelle2"), "v1": ("Proposed_model", "GatedMultimodal_Visuelle2"),
                        "v2": ("Proposed_model_v2", "GatedMultimodal_Visuelle2"),
                        "v3": ("Proposed_model_v3", "TARG_M4FT_Visuelle2"),
                        "v4": ("Proposed_model_v4", "GatedMultimodal_Visuelle2"),
                        "m4ft": ("M4FT_Visuelle2", "M4FT_Visuelle2")}[variant]
    import visuelle2_multimodal_fusion_b200.models._gtm as g
    v3 = importlib.import_module("visuelle2_multimodal_fusion_b200.models.Proposed_model_v3")
    for holder in (g, v3):
        _GTM_TRUNK_PATCH.append((holder, holder.resnet101_trunk))
        holder.resnet101_trunk = lambda: nn.Identity()
    mod = importlib.import_module("visuelle2_multimodal_fusion_b200.models." + modname)
    return getattr(mod, clsname)


def _restore_gtm_trunk():
    while _GTM_TRUNK_PATCH:
        holder, fn = _GTM_TRUNK_PATCH.pop()
        holder.resnet101_trunk = fn


def product_model(blob, device="cuda"):
    """The drop-in module for a golden blob, backbone replaced by identity, weights loaded."""
    import torch.nn as nn
    import visuelle2_multimodal_fusion_b200.synth as synth
    from visuelle2_multimodal_fusion_b200.models import CrossAttnRNN21, CrossAttnRNN210, CrossAttnRNNDemand
    import visuelle2_multimodal_fusion_b200.models.modules as mods
    cfg, name = blob["cfg"], blob["model"]
    cat_d, col_d, fab_d = synth.label_dicts()
    orig = mods.resnet101_trunk
    mods.resnet101_trunk = lambda: nn.Identity()       # goldens feed feature maps, not images
    try:
        if name == "CrossAttnRNN210":
            m = CrossAttnRNN210.CrossAttnRNN(cfg["E"], cfg["E"], cfg["H"], cat_d, col_d, fab_d, synth.STORE_N, 3,
                                             out_len=cfg["T"], use_teacher_forcing=cfg["tf"])
        elif name == "CrossAttnRNN21":
            m = CrossAttnRNN21.CrossAttnRNN(cfg["E"], cfg["E"], cfg["H"], cat_d, col_d, fab_d, synth.STORE_N, 3)
        elif name == "CrossAttnRNNDemand":
            m = CrossAttnRNNDemand.CrossAttnRNN(cfg["E"], cfg["E"], 3, cfg["H"], cat_d, col_d, fab_d,
                                                synth.STORE_N, True, True, True, True, out_len=cfg["T"],
                                                use_teacher_forcing=cfg["tf"])
        elif name.startswith("GTM:"):
            m = gtm_product_ctor(name[4:])(cfg["E"], cfg["H"], cfg["T"], cfg["heads"], 1, 1, 1, cat_d, col_d, fab_d,
                                           synth.STORE_N, 52, 3, 0, use_encoder_mask=1,
                                           autoregressive=cfg["autoregressive"],
                                           **(dict(query_modality=cfg["query_modality"]) if name == "GTM:v3" else {}))
        else:
            raise KeyError(name)
    finally:
        mods.resnet101_trunk = orig
        _restore_gtm_trunk()
    missing, unexpected = m.load_state_dict(blob["state"], strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("image_encoder.cnn") for k in missing), missing
    m = m.to(device).eval()
    if name.startswith("GTM:") and cfg["mode"] != "eval":
        m.train()            # goldens of mode train_nodrop: BatchNorm batch statistics, every dropout p = 0
        for mod in m.modules():
            if isinstance(mod, nn.Dropout):
                mod.p = 0.0
            if isinstance(mod, nn.MultiheadAttention):
                mod.dropout = 0.0
    return m


def product_run(m, blob, device="cuda"):
    """forward + training loss + backward of the drop-in module on a blob's inputs."""
    import torch.nn.functional as F
    inp = {k: v.to(device) for k, v in blob["inputs"].items()}
    feat = inp["feat"].clone().requires_grad_(True)
    torch.manual_seed(blob["cfg"]["seed"] + 1)       # host teacher-forcing draws, as in make_golden
    name = blob["model"]
    extras = {}
    if name.startswith("GTM:"):
        out, _ = m(inp["item_sales"], inp["cat"], inp["col"], inp["fab"], inp["store"], inp["temporal"],
                   inp["gtrends"], feat)
        loss = F.mse_loss(inp["y"].reshape(-1), out.reshape(-1))
    elif name == "CrossAttnRNNDemand":
        out, ia, ma = m(inp["ts"], inp["cat"], inp["col"], inp["fab"], inp["store"], inp["temporal"],
                        inp["gtrends"], feat)
        extras = dict(img_alphas=torch.stack(ia), mm_alphas=torch.stack(ma))
        loss = F.mse_loss(inp["ts"], out.squeeze())
    else:
        out, _ = m(inp["X"], inp["y"], inp["cat"], inp["col"], inp["fab"], inp["store"], inp["temporal"],
                   inp["gtrends"], feat)
        y = inp["y"]
        loss = F.mse_loss(y.reshape(out.shape) if name == "CrossAttnRNN210" else y, out)
    loss.backward()
    grads = {k: p.grad for k, p in m.named_parameters()}
    return out, loss, extras, grads, feat.grad


def compare_blob(blob, tol, report=None, precision="fp32"):
    """Compare the CUDA product against a golden blob; returns list of (what, rel_err, ok)."""
    m = product_model(blob)
    m.precision = precision
    out, loss, extras, grads, gfeat = product_run(m, blob)
    rows = []

    def cmp(what, a, b):
        a = a.detach().double().cpu()
        b = b.detach().double().cpu()
        diff = float((a - b).abs().max())
        scale = float(b.abs().max())
        ok = (a.shape == b.shape) and diff <= tol * scale + (1e-7 if tol < 1e-3 else 2e-6)
        rows.append((what, diff / max(scale, 1e-30), scale, ok))

    cmp("out", out, blob["out"])
    cmp("loss", loss, blob["loss"])
    for k, v in extras.items():
        cmp(k, v, blob[k])
    cmp("grad_feat", gfeat, blob["grad_feat"])
    noisy = noise_only_grads(blob)
    for k, g in blob["grads"].items():
        if k.startswith("image_encoder.cnn"):
            continue
        mine = grads.get(k)
        if k in noisy and g is not None and mine is not None:
            d = float((mine.detach().double().cpu() - g.double()).abs().max())
            rows.append(("grad:" + k + " (noise-only)", d, float(g.abs().max()), d <= (2e-5 if tol < 1e-3 else 1e-3)))
            continue
        if g is None:
            rows.append(("grad:" + k + " (none)", 0.0, 0.0, mine is None or float(mine.abs().max()) == 0.0))
        elif mine is None:
            rows.append(("grad:" + k + " MISSING", float("inf"), 0.0, False))
        else:
            cmp("grad:" + k, mine, g)
    return rows


def oracle_vs_cuda_default_dims(model, T, precision, tol, B=4, E=512, H=512, seed=11):
    """One forward + loss + backward of a CrossAttnRNN drop-in at the reference's default dims (E=A=H=512, Li=100,
    Lt=52, train_dl.py:197-199; feature maps in) on cuda:0 through the C ABI, compared with the oracle (CPU) on
    outputs, loss, feature-map gradient and every parameter gradient.  Returns the number of tensors compared."""
    import torch.nn as nn
    import visuelle2_multimodal_fusion_b200.models.modules as mods
    import visuelle2_multimodal_fusion_b200.synth as synth
    torch.manual_seed(3)
    cfg = dict(E=E, A=E, H=H, T=T, B=B, tf=True, seed=seed)
    blob = dict(model=model, cfg=cfg, state={})
    orig = mods.resnet101_trunk
    mods.resnet101_trunk = lambda: nn.Identity()
    try:
        from visuelle2_multimodal_fusion_b200.models import CrossAttnRNN21, CrossAttnRNN210, CrossAttnRNNDemand
        cat_d, col_d, fab_d = synth.label_dicts()
        # random weights from the product module's own (reference-identical) initialisation
        if model == "CrossAttnRNN210":
            m = CrossAttnRNN210.CrossAttnRNN(E, E, H, cat_d, col_d, fab_d, synth.STORE_N, 3, out_len=T)
        elif model == "CrossAttnRNN21":
            m = CrossAttnRNN21.CrossAttnRNN(E, E, H, cat_d, col_d, fab_d, synth.STORE_N, 3)
        else:
            m = CrossAttnRNNDemand.CrossAttnRNN(E, E, 3, H, cat_d, col_d, fab_d, synth.STORE_N, True, True, True,
                                                True, out_len=T, use_teacher_forcing=True)
    finally:
        mods.resnet101_trunk = orig
    blob["state"] = {k: v.detach().clone() for k, v in m.state_dict().items()}
    demand = model == "CrossAttnRNNDemand"
    data, feat = synth.make_batch(B, out_len=(1 if model == "CrossAttnRNN21" else 10), demand=demand, seed=seed,
                                  feat_hw=10)
    keys = ["ts", "cat", "col", "fab", "store", "temporal", "gtrends"] if demand else \
        ["X", "y", "cat", "col", "fab", "store", "temporal", "gtrends"]
    blob["inputs"] = dict(zip(keys, data), feat=feat)
    torch.manual_seed(cfg["seed"] + 1)
    blob["tf_mask"] = [bool(torch.rand(1) < 0.5) for _ in range(T)]
    o_out, o_loss, o_extras, P, o_feat = oracle_run(blob)
    o_loss.backward()
    m = m.cuda().eval()
    m.precision = precision
    out, loss, extras, grads, gfeat = product_run(m, blob)
    assert_close(out, o_out, tol, "out")
    assert_close(loss, o_loss, tol, "loss")
    assert_close(gfeat, o_feat.grad, tol, "grad_feat")
    n = 3
    for k, p in P.items():
        if p.grad is None:
            continue
        assert grads.get(k) is not None, k
        # attn_linear.bias feeds a softmax, so its true gradient is exactly 0 (the kernel returns 0,
        # autograd returns rounding noise): compare those absolutely
        floor = 1e-6 if k.endswith("attn_linear.bias") else max(1e-6 * float(p.grad.abs().max() + 1e-3), 1e-6 * tol)
        assert_close(grads[k], p.grad, tol, "grad:" + k, floor=floor)
        n += 1
    return n


# ------------------------------------------------------------------ full-size fixtures (oracle/make_golden_full.py)
def sample_index(numel, k=4096):
    if numel <= k:
        return torch.arange(numel)
    return torch.linspace(0, numel - 1, k).long().unique()


def full_model(blob, device="cpu"):
    """The drop-in module of a full_*.pt fixture: constructed under the fixture's seed (backbone = identity), with the
    per-tensor checksums of the reference's state verified -- the fixture does not carry the 80 MB of weights."""
    import torch.nn as nn
    import visuelle2_multimodal_fusion_b200.models.modules as mods
    import visuelle2_multimodal_fusion_b200.synth as synth
    from visuelle2_multimodal_fusion_b200.models import CrossAttnRNN21, CrossAttnRNN210, CrossAttnRNNDemand
    cfg, kind = blob["cfg"], blob["kind"]
    E, H, T = cfg["E"], cfg["H"], cfg["T"]
    cat_d, col_d, fab_d = synth.label_dicts()
    orig = mods.resnet101_trunk
    mods.resnet101_trunk = lambda: nn.Identity()
    try:
        torch.manual_seed(cfg["seed"])
        if kind == "rnn210":
            m = CrossAttnRNN210.CrossAttnRNN(E, E, H, cat_d, col_d, fab_d, synth.STORE_N, 3, out_len=T,
                                             use_teacher_forcing=cfg["tf"], teacher_forcing_ratio=0.5)
        elif kind == "rnn21":
            m = CrossAttnRNN21.CrossAttnRNN(E, E, H, cat_d, col_d, fab_d, synth.STORE_N, 3)
        else:
            m = CrossAttnRNNDemand.CrossAttnRNN(E, E, 3, H, cat_d, col_d, fab_d, synth.STORE_N, True, True, True, True,
                                                out_len=T, use_teacher_forcing=cfg["tf"], teacher_forcing_ratio=0.5)
    finally:
        mods.resnet101_trunk = orig
    sd = m.state_dict()
    for k, (s, a) in blob["checksum"].items():
        v = sd[k].double()
        assert abs(float(v.sum()) - s) <= 1e-9 * max(a, 1.0) and abs(float(v.abs().sum()) - a) <= 1e-9 * max(a, 1.0), \
            f"same-seed initialisation differs from the reference's for {k}"
    return m.to(device)


def full_inputs(blob):
    import visuelle2_multimodal_fusion_b200.synth as synth
    cfg, kind = blob["cfg"], blob["kind"]
    data, feat = synth.make_batch(cfg["B"], out_len=(1 if kind == "rnn21" else 10), demand=(kind == "demand"),
                                  seed=cfg["seed"] + 1, feat_hw=cfg["hw"])
    if kind == "demand":
        data = (data[0][:, :cfg["T"]].contiguous(),) + data[1:]
    return data, feat


def full_compare(blob, out, loss, extras, grads, gfeat, tol):
    """outputs / loss / attention maps in full; every gradient by L2 norm and by the strided sample."""
    assert_close(out, blob["out"], tol, "out")
    assert_close(loss, blob["loss"], tol, "loss")
    for k, v in blob["extras"].items():
        assert_close(extras[k], v, tol, k, floor=1e-6)
    n = 2

    def one(what, mine, ref):
        nonlocal n
        assert mine is not None, what + " missing"
        mine = mine.detach().double().cpu().reshape(-1)
        # absolute floor: a gradient that is zero up to rounding in the reference (e.g. trend attn_linear.weight when all
        # 52 projected trend rows are nearly equal: |g| ~ 1e-13) is compared absolutely, at a floor that scales with
        # the precision contract
        floor = 1e-6 if what.endswith("attn_linear.bias") else max(1e-6 * (ref["absmax"] + 1e-3), 1e-6 * tol)
        d = float((mine[sample_index(mine.numel())] - ref["sample"].double()).abs().max())
        assert d <= tol * ref["absmax"] + floor, f"{what}: sample max|diff|={d:.3e} absmax={ref['absmax']:.3e} tol={tol}"
        dn = abs(float(mine.norm()) - ref["norm"])
        assert dn <= tol * ref["norm"] + floor * mine.numel() ** 0.5, f"{what}: norm {float(mine.norm()):.6e} vs {ref['norm']:.6e}"
        n += 1

    one("grad_feat", gfeat, blob["grad_feat"])
    for k, ref in blob["grads"].items():
        if ref is None:
            assert grads.get(k) is None or float(grads[k].abs().max()) == 0.0, k
        else:
            one("grad:" + k, grads.get(k), ref)
    return n


def full_run(m, blob, device):
    """forward + training loss + backward of module ``m`` (product on CUDA) on the fixture's regenerated inputs."""
    import torch.nn.functional as F
    data, feat = full_inputs(blob)
    data = tuple(t.to(device) for t in data)
    feat = feat.to(device).requires_grad_(True)
    kind = blob["kind"]
    torch.manual_seed(blob["cfg"]["seed"] + 2)       # the host teacher-forcing draws of make_golden_full
    extras = {}
    if kind == "demand":
        out, ia, ma = m(*data, feat)
        extras = dict(img_alphas=torch.stack(ia), mm_alphas=torch.stack(ma))
        loss = F.mse_loss(data[0], out.squeeze())
    else:
        out, _ = m(*data, feat)
        y = data[1]
        loss = F.mse_loss(y.reshape(out.shape) if kind == "rnn210" else y, out)
    loss.backward()
    return out, loss, extras, {k: p.grad for k, p in m.named_parameters()}, feat.grad
