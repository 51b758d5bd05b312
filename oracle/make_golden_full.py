"""Full-size fixtures from the UNMODIFIED reference (build container only).  TEST INFRASTRUCTURE ONLY.

The small goldens (E=32) only reach the exact fallback kernels; the kernels the benchmark times -- persistent
decoder, streaming attention, tcgen05 GEMMs, persistent GRU -- engage at the reference's default dims
E=A=H=512, Li=100, Lt=52 (train_dl.py:197-199).  This script runs the reference modules there (B=8; feature maps in,
backbone = identity; eval mode like the small goldens) and stores what a test needs WITHOUT the 80 MB of head
weights: the construction seed (the drop-in classes initialise bit-identically under a seed,
tests/test_boundary_cpu.py) plus a per-tensor checksum of the state to prove it, the input seed (synth.make_batch
regenerates the batch), the teacher-forcing draws, the full outputs / attention maps / loss, and for every
gradient its shape, L2 norm, max |g| and a strided sample of up to 4096 values.

Usage:  python -m oracle.make_golden_full [rnn210 demand rnn21]
"""
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refshim  # noqa: E402
from oracle.make_golden import GOLDEN_DIR  # noqa: E402

SAMPLE = 4096


def sample_index(numel, k=SAMPLE):
    """Deterministic strided sample positions (shared with tests/test_gpu_fullsize.py)."""
    if numel <= k:
        return torch.arange(numel)
    return torch.linspace(0, numel - 1, k).long().unique()


def summarize(t):
    t = t.detach().double().reshape(-1)
    return dict(shape=None, norm=float(t.norm()), absmax=float(t.abs().max()), sample=t[sample_index(t.numel())].float())


def checksum(state):
    return {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in state.items() if v.is_floating_point()}


def build_reference(kind, E, H, T, seed, tf=True):
    """The reference module with the backbone factory patched to identity BEFORE construction, so that the head's
    initialisation is a function of the seed alone (the same patch the product side applies)."""
    import torchvision.models as tvm
    import visuelle2_multimodal_fusion_b200.synth as synth
    name = {"rnn210": "CrossAttnRNN210", "rnn21": "CrossAttnRNN21", "demand": "CrossAttnRNNDemand"}[kind]
    mod = refshim.load_reference_module(name)
    cat_d, col_d, fab_d = synth.label_dicts()
    orig = tvm.resnet101
    tvm.resnet101 = lambda *a, **k: nn.Sequential(nn.Identity(), nn.Identity(), nn.Identity())
    try:
        torch.manual_seed(seed)
        if kind == "rnn210":
            m = mod.CrossAttnRNN(E, E, H, cat_d, col_d, fab_d, synth.STORE_N, 3, out_len=T, use_teacher_forcing=tf,
                                 teacher_forcing_ratio=0.5)
        elif kind == "rnn21":
            m = mod.CrossAttnRNN(E, E, H, cat_d, col_d, fab_d, synth.STORE_N, 3, out_len=1)
        else:
            m = mod.CrossAttnRNN(E, E, 3, H, cat_d, col_d, fab_d, synth.STORE_N, True, True, True, True, out_len=T,
                                 use_teacher_forcing=tf, teacher_forcing_ratio=0.5)
    finally:
        tvm.resnet101 = orig
    return refshim.strip_backbone(m)


def case(kind, B=8, E=512, H=512, hw=10, seed=31):
    import visuelle2_multimodal_fusion_b200.synth as synth
    T = {"rnn210": 10, "demand": 12, "rnn21": 1}[kind]
    m = build_reference(kind, E, H, T, seed).eval()
    demand = kind == "demand"
    data, feat = synth.make_batch(B, out_len=(1 if kind == "rnn21" else 10), demand=demand, seed=seed + 1, feat_hw=hw)
    feat.requires_grad_(True)
    torch.manual_seed(seed + 2)
    tf_mask = [bool(torch.rand(1) < 0.5) for _ in range(T)] if kind != "rnn21" else None
    torch.manual_seed(seed + 2)
    extras = {}
    if demand:
        ts = data[0][:, :T].contiguous()
        out, ia, ma = m(ts, *data[1:], feat)
        loss = F.mse_loss(ts, out.squeeze())
        extras = dict(img_alphas=torch.stack([a.detach() for a in ia]), mm_alphas=torch.stack([a.detach() for a in ma]))
    else:
        out, _ = m(*data, feat)
        y = data[1]
        loss = F.mse_loss(y.reshape(out.shape) if kind == "rnn210" else y, out)
    loss.backward()
    grads = {}
    for k, p in m.named_parameters():
        if p.grad is None:
            grads[k] = None
        else:
            grads[k] = summarize(p.grad)
            grads[k]["shape"] = tuple(p.shape)
    gf = summarize(feat.grad)
    gf["shape"] = tuple(feat.shape)
    return dict(kind=kind, cfg=dict(B=B, E=E, H=H, T=T, hw=hw, seed=seed, tf=True), checksum=checksum(m.state_dict()),
                tf_mask=tf_mask, out=out.detach(), loss=loss.detach(), extras=extras, grads=grads, grad_feat=gf)


if __name__ == "__main__":
    if not refshim.reference_available():
        raise SystemExit("reference tree not mounted; fixtures can only be generated in the build container")
    torch.backends.mha.set_fastpath_enabled(False)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for kind in (sys.argv[1:] or ["rnn210", "demand", "rnn21"]):
        blob = case(kind)
        path = os.path.join(GOLDEN_DIR, f"full_{kind}.pt")
        torch.save(blob, path)
        print(f"{path}: loss {float(blob['loss']):.6f}, {len(blob['grads'])} gradients, {os.path.getsize(path) / 1e6:.2f} MB")
