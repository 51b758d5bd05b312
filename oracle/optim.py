"""Oracle (CPU, explicit math) for the optimizer step and the image transform.  TEST INFRASTRUCTURE ONLY.

Adafactor: the reference's ``configure_optimizers`` (/root/reference/models/CrossAttnRNN210.py:229-230,
/root/reference/models/GTM_Visuelle2.py:264-266) builds ``fairseq.optim.adafactor.Adafactor(params,
scale_parameter=True, relative_step=True, warmup_init=True, lr=None)``.  fairseq is a third-party dependency that is
NOT in /root/reference and not in this image (the README pins no version); its algorithm (Shazeer & Stern 2018,
"Adafactor: Adaptive Learning Rates with Sublinear Memory Cost", as implemented in fairseq/optim/adafactor.py and
ported verbatim to transformers.optimization.Adafactor) is restated here in numpy from the published equations:

    step += 1 ; RMS = ||p||_2 / sqrt(numel)
    rel  = min(1e-6 * step if warmup_init else 1e-2, 1 / sqrt(step))        (relative_step)
    lr   = max(eps2, RMS) * rel                                             (scale_parameter)
    beta = 1 - step ** decay_rate                                           (decay_rate = -0.8)
    upd  = g * g + eps1
    ndim >= 2:  row <- beta row + (1 - beta) mean(upd, -1) ; col <- beta col + (1 - beta) mean(upd, -2)
                u = g * rsqrt(row / mean(row, -1, keepdim)) [..., None] * rsqrt(col) [..., None, :]
    ndim <  2:  sq  <- beta sq + (1 - beta) upd ; u = g * rsqrt(sq)
    u <- u / max(1, RMS(u) / clip_threshold) ; p <- p - lr * u              (beta1 = None, weight_decay = 0)

Pinned (tests/test_oracle_golden.py::test_adafactor_oracle_matches_transformers, CPU) against
transformers.optimization.Adafactor, the port of the fairseq optimizer that IS in the image.

Image transform: /root/reference/dataset_fusion.py:50-57 -- ``ToTensor()`` (uint8 HWC -> float CHW / 255) followed by
``Normalize(mean, std)``; restated in numpy and pinned against torchvision on CPU in the same test file.
"""
import math

import numpy as np


def adafactor_init(p):
    st = {"step": 0, "RMS": 0.0}
    if p.ndim >= 2:
        st["exp_avg_sq_row"] = np.zeros(p.shape[:-1], np.float32)
        st["exp_avg_sq_col"] = np.zeros(p.shape[:-2] + p.shape[-1:], np.float32)
    else:
        st["exp_avg_sq"] = np.zeros(p.shape, np.float32)
    return st


def adafactor_step(p, g, st, lr=None, eps=(1e-30, 1e-3), clip_threshold=1.0, decay_rate=-0.8, scale_parameter=True,
                   relative_step=True, warmup_init=False):
    """One step on numpy float32 arrays ``p`` (updated in place and returned), ``g``; ``st`` from adafactor_init."""
    f32 = np.float32
    st["step"] += 1
    step = st["step"]
    rms = f32(np.sqrt(np.sum(p.astype(np.float64) ** 2)) / math.sqrt(p.size))
    st["RMS"] = rms
    if relative_step:
        rel = min(1e-6 * step if warmup_init else 1e-2, 1.0 / math.sqrt(step))
    else:
        rel = lr
    scale = max(f32(eps[1]), rms) if scale_parameter else f32(1.0)
    lr_t = f32(scale * rel)
    beta = f32(1.0 - math.pow(step, decay_rate))
    omb = f32(1.0 - (1.0 - math.pow(step, decay_rate)))
    upd = g * g + f32(eps[0])
    if p.ndim >= 2:
        row, col = st["exp_avg_sq_row"], st["exp_avg_sq_col"]
        row[...] = row * beta + upd.mean(-1, dtype=np.float32) * omb
        col[...] = col * beta + upd.mean(-2, dtype=np.float32) * omb
        r = 1.0 / np.sqrt(row / row.mean(-1, keepdims=True, dtype=np.float32))
        c = 1.0 / np.sqrt(col)
        u = g * (r[..., None] * c[..., None, :]).astype(np.float32)
    else:
        sq = st["exp_avg_sq"]
        sq[...] = sq * beta + upd * omb
        u = g * (1.0 / np.sqrt(sq)).astype(np.float32)
    rms_u = f32(np.sqrt(np.sum(u.astype(np.float64) ** 2)) / math.sqrt(u.size))
    u = u / max(f32(1.0), rms_u / f32(clip_threshold))
    p -= (u * lr_t).astype(np.float32)
    return p


def normalize_uint8(u8, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """uint8 [B,H,W,C] -> float32 [B,C,H,W]: Normalize(mean, std)(ToTensor(img)) per item (dataset_fusion.py:50-57)."""
    x = u8.astype(np.float32) / np.float32(255.0)
    x = (x - np.asarray(mean, np.float32)) / np.asarray(std, np.float32)
    return np.ascontiguousarray(x.transpose(0, 3, 1, 2))
