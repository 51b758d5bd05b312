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
    else:
        raise KeyError(model)
    return out, loss, extras, P, feat
