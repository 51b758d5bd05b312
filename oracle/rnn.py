"""Oracle (CPU, explicit math) for the CrossAttnRNN family.  TEST INFRASTRUCTURE ONLY.

Restates, as written, the forward passes of
  * CrossAttnRNN210  /root/reference/models/CrossAttnRNN210.py:143-227
  * CrossAttnRNN21   /root/reference/models/CrossAttnRNN21.py:137-211
  * CrossAttnRNNDemand /root/reference/models/CrossAttnRNNDemand.py:247-349
on a flat ``dict`` of tensors keyed by the reference ``state_dict`` names.
The third-party pieces the reference delegates to torch (nn.GRU,
nn.MultiheadAttention, nn.Embedding, nn.Linear, softmax) are written out from
their published equations so the oracle does not depend on the modules the
product replaces.  Backward is obtained with autograd over these plain ops.

Nothing is hoisted or re-associated here: the per-step ``encoder_linear`` and
``trend_linear`` recomputation is kept exactly as the reference does it, so the
oracle doubles as the honest CPU baseline (``bench.py --impl reference``).

The ResNet-101 backbone (torchvision, third party for both the reference and the
product) is not restated: ``feat`` is the backbone output ``[B,2048,h,w]``.

Pinned against fixtures produced by the unmodified reference: see
``oracle/make_golden.py`` and ``tests/test_oracle_golden.py``.
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- #
# third-party building blocks, from their equations
# --------------------------------------------------------------------------- #
def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """torch.nn.GRU cell, gate order (r, z, n); used at CrossAttnRNN210.py:15,123,135."""
    hid = h.shape[-1]
    gi = x @ w_ih.t() + b_ih
    gh = h @ w_hh.t() + b_hh
    r = torch.sigmoid(gi[:, :hid] + gh[:, :hid])
    z = torch.sigmoid(gi[:, hid:2 * hid] + gh[:, hid:2 * hid])
    n = torch.tanh(gi[:, 2 * hid:] + r * gh[:, 2 * hid:])
    return (1.0 - z) * n + z * h


def gru_seq(x, h0, P, prefix):
    """batch_first single-layer GRU over ``x [N,L,I]``; returns (all outputs [N,L,H], last h)."""
    w_ih, w_hh = P[prefix + "weight_ih_l0"], P[prefix + "weight_hh_l0"]
    b_ih, b_hh = P[prefix + "bias_ih_l0"], P[prefix + "bias_hh_l0"]
    h = h0
    outs = []
    for t in range(x.shape[1]):
        h = gru_cell(x[:, t], h, w_ih, w_hh, b_ih, b_hh)
        outs.append(h)
    return torch.stack(outs, dim=1), h


def mha_self(x, P, prefix, heads, dropout_p, training, attn_mask=None):
    """nn.MultiheadAttention(q=k=v=x) on ``x [L,N,E]`` (seq-first), packed in-proj.

    CrossAttnRNN210.py:126,176-179.  Returns ``[L,N,E]`` (averaged weights discarded there)."""
    L, N, E = x.shape
    hd = E // heads
    qkv = x @ P[prefix + "in_proj_weight"].t() + P[prefix + "in_proj_bias"]
    q, k, v = qkv.split(E, dim=-1)

    def split_heads(t):  # [L,N,E] -> [N*heads, L, hd]
        return t.reshape(L, N * heads, hd).transpose(0, 1)

    q, k, v = split_heads(q), split_heads(k), split_heads(v)
    scores = (q * (1.0 / math.sqrt(hd))) @ k.transpose(1, 2)
    if attn_mask is not None:
        scores = scores + attn_mask
    a = torch.softmax(scores, dim=-1)
    a = F.dropout(a, dropout_p, training)
    o = (a @ v).transpose(0, 1).reshape(L, N, E)
    return o @ P[prefix + "out_proj.weight"].t() + P[prefix + "out_proj.bias"]


def additive_attention(enc, h_dec, P, prefix, weighted_by_projection):
    """Bahdanau attention, CrossAttnRNN210.py:83-89 (returns alpha*enc) and
    CrossAttnRNNDemand.py:134-149 (returns alpha*h_j)."""
    h_j = enc @ P[prefix + "encoder_linear.weight"].t()
    s_i = h_dec @ P[prefix + "decoder_linear.weight"].t()
    energy = torch.tanh(h_j + s_i.unsqueeze(1)) @ P[prefix + "attn_linear.weight"].t()
    energy = energy.squeeze(2) + P[prefix + "attn_linear.bias"]
    alpha = torch.softmax(energy, dim=1)
    src = h_j if weighted_by_projection else enc
    return alpha.unsqueeze(2) * src, alpha


# --------------------------------------------------------------------------- #
# static encoders
# --------------------------------------------------------------------------- #
def image_encoder(feat, P, training):
    """ImageEncoder minus the torchvision backbone, CrossAttnRNN210.py:69-72."""
    x = feat.flatten(2).permute(0, 2, 1)
    x = x @ P["image_encoder.fc.weight"].t() + P["image_encoder.fc.bias"]
    return F.dropout(x, 0.1, training)


def trend_encoder(gtrends, P, training):
    """TSEmbedder, CrossAttnRNN210.py:23-24 on ``gtrends.permute(0,2,1)``."""
    x = gtrends.permute(0, 2, 1)
    hid = P["trend_encoder.ts_embedder.weight_hh_l0"].shape[1]
    h0 = x.new_zeros(x.shape[0], hid)
    out, _ = gru_seq(x, h0, P, "trend_encoder.ts_embedder.")
    return F.dropout(out, 0.1, training)


def temporal_encoder(temporal, P, training, day_only):
    """TemporalFeatureEncoder.  ``day_only`` reproduces the Demand copy, which feeds all four
    features through ``day_embedding`` (CrossAttnRNNDemand.py:61-64); 21/210 use the four
    distinct linears (CrossAttnRNN210.py:51-56)."""
    names = ["day", "week", "month", "year"]
    total = 0
    for k, nm in enumerate(names):
        src = "day" if day_only else nm
        w = P[f"temp_encoder.{src}_embedding.weight"]
        b = P[f"temp_encoder.{src}_embedding.bias"]
        total = total + F.dropout(temporal[:, k:k + 1] @ w.t() + b, 0.1, training)
    return total


def attribute_encoder(cat, col, fab, store, P, training):
    """AttributeEncoder, CrossAttnRNN210.py:35-40."""
    total = 0
    for nm, idx in (("cat", cat), ("col", col), ("fab", fab), ("store", store)):
        total = total + F.dropout(P[f"attribute_encoder.{nm}_embedder.weight"][idx], 0.1, training)
    return total


# --------------------------------------------------------------------------- #
# model forwards
# --------------------------------------------------------------------------- #
def _flatten_windows(X, y):
    num_windows = 1
    if X.dim() == 3:
        bs, num_windows, hist = X.shape
        X = X.reshape(bs * num_windows, hist)
        if y is not None:
            y = y.reshape(bs * num_windows, -1)
    else:
        bs = X.shape[0]
    if X.dim() == 2:
        X = X.unsqueeze(-1)
    return X, y, bs, num_windows


def _static_encode(P, cat, col, fab, store, temporal, gtrends, feat, num_windows, training, day_only):
    img = image_encoder(feat, P, training)
    gt = trend_encoder(gtrends, P, training)
    dummy = temporal_encoder(temporal, P, training, day_only)
    attr = attribute_encoder(cat, col, fab, store, P, training)
    if num_windows > 1:
        img = img.repeat_interleave(num_windows, dim=0)
        gt = gt.repeat_interleave(num_windows, dim=0)
        dummy = dummy.repeat_interleave(num_windows, dim=0)
        attr = attr.repeat_interleave(num_windows, dim=0)
    gt = mha_self(gt.permute(1, 0, 2), P, "ts_self_attention.", 4, 0.1, training)  # [52,N,E]
    return img, gt.permute(1, 0, 2), dummy, attr


def _fuse_step(P, img, gt, dummy, attr, h, by_proj, modal=(True, True, True)):
    """One execution of the three attentions + multimodal embedder (the loop body
    CrossAttnRNN210.py:192-208 / CrossAttnRNNDemand.py:286-333)."""
    use_img, use_att, use_trends = modal
    n = dummy.shape[0]
    rows = [dummy]
    a_img = a_tr = None
    if use_img:
        w_img, a_img = additive_attention(img, h, P, "img_attention.", by_proj)
        rows.append(w_img.sum(1))
    if use_att:
        rows.append(attr)
    if use_trends:
        w_tr, a_tr = additive_attention(gt, h, P, "ts_attention.", by_proj)
        rows.append(w_tr.reshape(n, -1) @ P["trend_linear.weight"].t() + P["trend_linear.bias"])
    mm_in = torch.stack(rows, dim=1)
    w_mm, a_mm = additive_attention(mm_in, h, P, "multimodal_attention.", by_proj)
    ctx = (mm_in + w_mm).sum(1) @ P["multimodal_embedder.weight"].t() + P["multimodal_embedder.bias"]
    return ctx, a_img, a_tr, a_mm


def rnn210_forward(P, X, y, cat, col, fab, store, temporal, gtrends, feat, *, out_len=10,
                   use_teacher_forcing=True, teacher_forcing_ratio=0.5, training=False,
                   tf_mask=None):
    """CrossAttnRNN210.py:143-227.  ``tf_mask``: optional list of T bools replacing the host
    ``torch.rand(1)`` draws (same order, drawn only when teacher forcing is on and y is given)."""
    X, y, bs, nw = _flatten_windows(X, y)
    img, gt, dummy, attr = _static_encode(P, cat, col, fab, store, temporal, gtrends, feat, nw,
                                          training, day_only=False)
    hid = P["sales_encoder_gru.weight_hh_l0"].shape[1]
    _, h = gru_seq(X, X.new_zeros(X.shape[0], hid), P, "sales_encoder_gru.")
    x_in = X[:, -1, :]
    outs = []
    for t in range(out_len):
        ctx, _, _, _ = _fuse_step(P, img, gt, dummy, attr, h, by_proj=False)
        h = gru_cell(torch.cat([ctx, x_in], dim=1), h, P["decoder_gru.weight_ih_l0"],
                     P["decoder_gru.weight_hh_l0"], P["decoder_gru.bias_ih_l0"],
                     P["decoder_gru.bias_hh_l0"])
        pred = h @ P["decoder_fc.weight"].t() + P["decoder_fc.bias"]
        outs.append(pred)
        if use_teacher_forcing and y is not None:
            forced = bool(tf_mask[t]) if tf_mask is not None else bool(torch.rand(1) < teacher_forcing_ratio)
            x_in = y[:, t:t + 1] if forced else pred
        else:
            x_in = pred
    return torch.cat(outs, dim=1), None


def rnn21_forward(P, X, y, cat, col, fab, store, temporal, gtrends, feat, *, training=False):
    """CrossAttnRNN21.py:137-211: the fusion executed once on the sales-GRU state, MLP head."""
    X, y, bs, nw = _flatten_windows(X, y)
    img, gt, dummy, attr = _static_encode(P, cat, col, fab, store, temporal, gtrends, feat, nw,
                                          training, day_only=False)
    hid = P["sales_encoder_gru.weight_hh_l0"].shape[1]
    _, h = gru_seq(X, X.new_zeros(X.shape[0], hid), P, "sales_encoder_gru.")
    ctx, _, _, _ = _fuse_step(P, img, gt, dummy, attr, h, by_proj=False)
    pred = ctx @ P["decoder_fc.weight"].t() + P["decoder_fc.bias"]
    return pred.view(bs, nw, 1), None


def demand_forward(P, ts, cat, col, fab, store, temporal, gtrends, feat, *, out_len=12,
                   use_teacher_forcing=False, teacher_forcing_ratio=0.5, training=False,
                   use_img=True, use_att=True, use_trends=True, tf_mask=None):
    """CrossAttnRNNDemand.py:247-349.  The host RNG is consumed once per step even in eval
    (``:343-345``) unless ``tf_mask`` is supplied."""
    bs = ts.shape[0]
    gt = trend_encoder(gtrends, P, training)
    img = image_encoder(feat, P, training)
    dummy = temporal_encoder(temporal, P, training, day_only=True)
    attr = attribute_encoder(cat, col, fab, store, P, training)
    if use_trends:
        gt = mha_self(gt.permute(1, 0, 2), P, "ts_self_attention.", 4, 0.1, training).permute(1, 0, 2)
    hid = P["decoder.weight_hh_l0"].shape[1]
    h = ts.new_zeros(bs, hid)
    x_in = ts.new_zeros(bs, 1)
    outs, img_alphas, mm_alphas = [], [], []
    for t in range(out_len):
        ctx, a_img, _, a_mm = _fuse_step(P, img, gt, dummy, attr, h, by_proj=True,
                                         modal=(use_img, use_att, use_trends))
        if a_img is not None:
            img_alphas.append(a_img)
        mm_alphas.append(a_mm)
        h = gru_cell(torch.cat([ctx, x_in], dim=1), h, P["decoder.weight_ih_l0"],
                     P["decoder.weight_hh_l0"], P["decoder.bias_ih_l0"], P["decoder.bias_hh_l0"])
        pred = h @ P["decoder_fc.weight"].t() + P["decoder_fc.bias"]
        outs.append(pred)
        x_in = pred
        forced = bool(tf_mask[t]) if tf_mask is not None else bool(torch.rand(1) < teacher_forcing_ratio)
        if use_teacher_forcing and forced and ts is not None:
            x_in = ts[:, t:t + 1]
    return torch.stack(outs, dim=1), img_alphas, mm_alphas


# --------------------------------------------------------------------------- #
# metrics (utils.py:4-11; CrossAttnRNN210.py:271-273; CrossAttnRNNDemand.py:416-422)
# --------------------------------------------------------------------------- #
def mae_wape(gt, pred, abs_denominator=True, norm_scalar=53.0):
    mae = torch.mean(torch.abs(gt - pred)) * norm_scalar
    den = torch.sum(torch.abs(gt)) if abs_denominator else torch.sum(gt)
    wape = 100.0 * torch.sum(torch.abs(gt - pred)) / den
    return mae, wape
