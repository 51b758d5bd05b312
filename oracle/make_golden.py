"""Generate tests/golden/*.pt by running the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Usage:  python -m oracle.make_golden [case ...]

Every fixture holds: the constructor config, the head ``state_dict`` (the torchvision backbone
is replaced by identity, so ``images`` is a small feature map ``[B,2048,h,w]``), the inputs, the
teacher-forcing decisions the host RNG produced, the reference outputs, the training loss
(``training_step`` formula) and the gradients autograd gives for every parameter and for the
feature map.  Mode is ``eval()`` (dropout off, BatchNorm running stats) unless the case says
``train_nodrop`` (train() with every dropout p set to 0: BatchNorm batch statistics).
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refshim  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _synth():
    import visuelle2_multimodal_fusion_b200.synth as synth
    return synth


def _grads(model, feat):
    g = {k: (p.grad.clone() if p.grad is not None else None) for k, p in model.named_parameters()}
    return g, feat.grad.clone()


def _head_state(model):
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def _tf_draws(seed, n, ratio):
    torch.manual_seed(seed)
    return [bool(torch.rand(1) < ratio) for _ in range(n)]


def case_rnn210(name, B, E, H, T, hw, tf, seed):
    synth = _synth()
    mod = refshim.load_reference_module("CrossAttnRNN210")
    cat_d, col_d, fab_d = synth.label_dicts()
    torch.manual_seed(seed)
    m = mod.CrossAttnRNN(E, E, H, cat_d, col_d, fab_d, synth.STORE_N, 3, out_len=T,
                         use_teacher_forcing=tf, teacher_forcing_ratio=0.5)
    refshim.strip_backbone(m).eval()
    (X, y, cat, col, fab, store, temporal, gt), feat = synth.make_batch(B, out_len=T, seed=seed, feat_hw=hw)
    feat.requires_grad_(True)
    tf_mask = _tf_draws(seed + 1, T, 0.5) if tf else None
    torch.manual_seed(seed + 1)
    out, _ = m(X, y, cat, col, fab, store, temporal, gt, feat)
    loss = F.mse_loss(y.reshape(out.shape), out)
    loss.backward()
    grads, gfeat = _grads(m, feat)
    return dict(model="CrossAttnRNN210", cfg=dict(E=E, A=E, H=H, T=T, B=B, tf=tf, seed=seed), state=_head_state(m),
                inputs=dict(X=X, y=y, cat=cat, col=col, fab=fab, store=store, temporal=temporal,
                            gtrends=gt, feat=feat.detach()),
                tf_mask=tf_mask, out=out.detach(), loss=loss.detach(), grads=grads, grad_feat=gfeat)


def case_rnn21(name, B, E, H, hw, seed):
    synth = _synth()
    mod = refshim.load_reference_module("CrossAttnRNN21")
    cat_d, col_d, fab_d = synth.label_dicts()
    torch.manual_seed(seed)
    m = mod.CrossAttnRNN(E, E, H, cat_d, col_d, fab_d, synth.STORE_N, 3, out_len=1)
    refshim.strip_backbone(m).eval()
    (X, y, cat, col, fab, store, temporal, gt), feat = synth.make_batch(B, out_len=1, seed=seed, feat_hw=hw)
    feat.requires_grad_(True)
    out, _ = m(X, y, cat, col, fab, store, temporal, gt, feat)
    loss = F.mse_loss(y, out)
    loss.backward()
    grads, gfeat = _grads(m, feat)
    return dict(model="CrossAttnRNN21", cfg=dict(E=E, A=E, H=H, T=1, B=B, seed=seed), state=_head_state(m),
                inputs=dict(X=X, y=y, cat=cat, col=col, fab=fab, store=store, temporal=temporal,
                            gtrends=gt, feat=feat.detach()),
                tf_mask=None, out=out.detach(), loss=loss.detach(), grads=grads, grad_feat=gfeat)


def case_demand(name, B, E, H, T, hw, tf, seed):
    synth = _synth()
    mod = refshim.load_reference_module("CrossAttnRNNDemand")
    cat_d, col_d, fab_d = synth.label_dicts()
    torch.manual_seed(seed)
    m = mod.CrossAttnRNN(E, E, 3, H, cat_d, col_d, fab_d, synth.STORE_N, True, True, True, True,
                         out_len=T, use_teacher_forcing=tf, teacher_forcing_ratio=0.5)
    refshim.strip_backbone(m).eval()
    (ts, cat, col, fab, store, temporal, gt), feat = synth.make_batch(B, demand=True, seed=seed, feat_hw=hw)
    ts = ts[:, :T].contiguous()
    feat.requires_grad_(True)
    tf_mask = _tf_draws(seed + 1, T, 0.5)          # Demand draws every step, even in eval
    torch.manual_seed(seed + 1)
    out, img_a, mm_a = m(ts, cat, col, fab, store, temporal, gt, feat)
    loss = F.mse_loss(ts, out.squeeze())
    loss.backward()
    grads, gfeat = _grads(m, feat)
    return dict(model="CrossAttnRNNDemand", cfg=dict(E=E, A=E, H=H, T=T, B=B, tf=tf, seed=seed), state=_head_state(m),
                inputs=dict(ts=ts, cat=cat, col=col, fab=fab, store=store, temporal=temporal,
                            gtrends=gt, feat=feat.detach()),
                tf_mask=tf_mask, out=out.detach(), img_alphas=torch.stack([a.detach() for a in img_a]),
                mm_alphas=torch.stack([a.detach() for a in mm_a]), loss=loss.detach(), grads=grads,
                grad_feat=gfeat)


CASES = {
    # name: (fn, kwargs)
    "rnn210_small": (case_rnn210, dict(B=3, E=32, H=48, T=10, hw=3, tf=True, seed=21)),
    "rnn210_notf": (case_rnn210, dict(B=2, E=32, H=32, T=4, hw=2, tf=False, seed=5)),
    "rnn21_small": (case_rnn21, dict(B=3, E=32, H=48, hw=3, seed=22)),
    "demand_small": (case_demand, dict(B=3, E=32, H=48, T=12, hw=3, tf=True, seed=23)),
    "demand_notf": (case_demand, dict(B=2, E=32, H=32, T=5, hw=2, tf=False, seed=7)),
}


def main(argv):
    if not refshim.reference_available():
        raise SystemExit("reference tree not mounted; goldens can only be generated in the build container")
    torch.set_num_threads(1)
    torch.backends.mha.set_fastpath_enabled(False)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    try:
        from oracle import make_golden_gtm
        CASES.update(make_golden_gtm.CASES)
    except ImportError:
        pass
    names = argv or list(CASES)
    for nm in names:
        fn, kw = CASES[nm]
        blob = fn(nm, **kw)
        path = os.path.join(GOLDEN_DIR, nm + ".pt")
        torch.save(blob, path)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main(sys.argv[1:])
