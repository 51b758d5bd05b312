"""Oracle (CPU, explicit math) for the GTM family.  TEST INFRASTRUCTURE ONLY.

Restates, as written, the forwards of
  * GTM_Visuelle2               /root/reference/models/GTM_Visuelle2.py:215-262
  * Proposed_model (v1)         /root/reference/models/Proposed_model.py:309-359
  * Proposed_model_v2           /root/reference/models/Proposed_model_v2.py:802-847
  * Proposed_model_v3 (TARG)    /root/reference/models/Proposed_model_v3.py:284-327
  * Proposed_model_v4           /root/reference/models/Proposed_model_v4.py:245-289
  * M4FT_Visuelle2              /root/reference/models/M4FT_Visuelle2.py:252-300
on a flat dict of tensors keyed by the reference state_dict names.  The torch building blocks the
reference delegates to (nn.TransformerEncoderLayer / DecoderLayer post-LN with ReLU,
nn.MultiheadAttention, nn.LayerNorm, nn.BatchNorm1d, nn.GRU, nn.Embedding, 1x1 nn.Conv2d +
AdaptiveAvgPool2d) are written from their published equations.  ``feat`` is the output of the
torchvision ResNet-101 trunk (third party on both sides) ``[B,2048,h,w]``.

``training`` selects dropout on/off AND BatchNorm batch statistics; ``drop`` (default = training)
lets the tests ask for batch statistics without dropout (the reference's ``train()`` with p=0).
Pinned against fixtures produced by the unmodified reference: oracle/make_golden_gtm.py.
"""
import math

import torch
import torch.nn.functional as F

from .rnn import gru_seq


# --------------------------------------------------------------------------- building blocks
def layer_norm(x, P, prefix, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * P[prefix + "weight"] + P[prefix + "bias"]


def batch_norm1d(x, P, prefix, batch_stats, eps=1e-5):
    """nn.BatchNorm1d forward: batch statistics (biased variance) in train mode, running stats in eval."""
    if batch_stats:
        mu = x.mean(0)
        var = ((x - mu) ** 2).mean(0)
    else:
        mu, var = P[prefix + "running_mean"], P[prefix + "running_var"]
    return (x - mu) / torch.sqrt(var + eps) * P[prefix + "weight"] + P[prefix + "bias"]


def lin(x, P, prefix):
    y = x @ P[prefix + "weight"].t()
    b = P.get(prefix + "bias")
    return y if b is None else y + b


def attention(q, k, v, heads, mask, p_drop, drop):
    """softmax(q k^T / sqrt(hd) + mask) v on seq-first tensors q [Lq,N,D], k/v [Lk,N,D]."""
    Lq, N, D = q.shape
    Lk = k.shape[0]
    hd = D // heads
    sh = lambda t, L: t.reshape(L, N * heads, hd).transpose(0, 1)
    qh, kh, vh = sh(q, Lq), sh(k, Lk), sh(v, Lk)
    s = (qh * (hd ** -0.5)) @ kh.transpose(1, 2)
    if mask is not None:
        s = s + mask
    a = F.dropout(torch.softmax(s, -1), p_drop, drop)
    return (a @ vh).transpose(0, 1).reshape(Lq, N, D), qh


def mha(q_in, kv_in, P, prefix, heads, mask, p_drop, drop):
    """nn.MultiheadAttention with packed in_proj (query from q_in, key=value from kv_in)."""
    D = q_in.shape[-1]
    W, b = P[prefix + "in_proj_weight"], P[prefix + "in_proj_bias"]
    q = q_in @ W[:D].t() + b[:D]
    k = kv_in @ W[D:2 * D].t() + b[D:2 * D]
    v = kv_in @ W[2 * D:].t() + b[2 * D:]
    o, _ = attention(q, k, v, heads, mask, p_drop, drop)
    return lin(o, P, prefix + "out_proj.")


def encoder_layer(x, P, prefix, heads, mask, p_drop, drop):
    """nn.TransformerEncoderLayer (post-LN, ReLU, default dim_feedforward)."""
    a = mha(x, x, P, prefix + "self_attn.", heads, mask, p_drop, drop)
    x = layer_norm(x + F.dropout(a, p_drop, drop), P, prefix + "norm1.")
    f = lin(F.dropout(F.relu(lin(x, P, prefix + "linear1.")), p_drop, drop), P, prefix + "linear2.")
    return layer_norm(x + F.dropout(f, p_drop, drop), P, prefix + "norm2.")


def decoder_layer(tgt, mem, P, prefix, heads, tgt_mask, p_drop, drop):
    """nn.TransformerDecoderLayer (post-LN, ReLU)."""
    a = mha(tgt, tgt, P, prefix + "self_attn.", heads, tgt_mask, p_drop, drop)
    tgt = layer_norm(tgt + F.dropout(a, p_drop, drop), P, prefix + "norm1.")
    a = mha(tgt, mem, P, prefix + "multihead_attn.", heads, None, p_drop, drop)
    tgt = layer_norm(tgt + F.dropout(a, p_drop, drop), P, prefix + "norm2.")
    f = lin(F.dropout(F.relu(lin(tgt, P, prefix + "linear1.")), p_drop, drop), P, prefix + "linear2.")
    return layer_norm(tgt + F.dropout(f, p_drop, drop), P, prefix + "norm3.")


def positional_encoding(L, D):
    """PositionalEncoding.pe, GTM_Visuelle2.py:18-24."""
    pe = torch.zeros(L, D)
    pos = torch.arange(0, L, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, D, 2).float() * (-math.log(10000.0) / D))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(1)


def encoder_mask(size, horizon):
    """Block-diagonal additive mask with blocks of gcd(size, horizon), GTM_Visuelle2.py:57-64."""
    split = math.gcd(size, horizon)
    m = torch.full((size, size), float("-inf"))
    for i in range(0, size, split):
        m[i:i + split, i:i + split] = 0.0
    return m


def causal_mask(size):
    return torch.triu(torch.full((size, size), float("-inf")), diagonal=1)


# --------------------------------------------------------------------------- shared encoders
def gtrend_encoder(gtrends, P, prefix, heads, horizon, use_mask, drop, lin_name="input_linear.module.",
                   enc_name="encoder.layers.", layer_fn=encoder_layer, n_layers=2):
    """GTrendEmbedder, GTM_Visuelle2.py:46-74: Linear(3->D) -> +PE -> dropout .1 -> 2 encoder layers (p=.2)."""
    x = lin(gtrends.permute(0, 2, 1), P, prefix + lin_name).permute(1, 0, 2)   # [52,B,D]
    L, _, D = x.shape
    x = F.dropout(x + positional_encoding(L, D), 0.1, drop)
    mask = encoder_mask(L, horizon) if use_mask == 1 else None
    for i in range(n_layers):
        x = layer_fn(x, P, f"{prefix}{enc_name}{i}.", heads, mask, 0.2, drop)
    return x


def text_encoder(cat, col, fab, store, P, drop):
    """AttributeEncoder, GTM_Visuelle2.py:81-96 -> [B,4,E]."""
    e = torch.stack([P["text_encoder.cat_emb.weight"][cat], P["text_encoder.col_emb.weight"][col],
                     P["text_encoder.fab_emb.weight"][fab], P["text_encoder.store_emb.weight"][store]], 1)
    return F.dropout(e, 0.1, drop)


def image_projection(feat, P):
    """ImageEncoder minus the trunk, GTM_Visuelle2.py:119-126: 1x1 conv then global average pool."""
    W = P["image_encoder.projection.weight"].flatten(1)        # [E,2048]
    x = torch.einsum("bchw,ec->behw", feat, W) + P["image_encoder.projection.bias"][None, :, None, None]
    return x.mean((2, 3))


def four_linears(temporal, P, prefix):
    return torch.cat([temporal[:, k:k + 1] @ P[f"{prefix}{n}_emb.weight"].t() + P[f"{prefix}{n}_emb.bias"]
                      for k, n in enumerate(["day", "week", "month", "year"])], 1)


def dummy_encoder(temporal, P, drop):
    """DummyEmbedder, GTM_Visuelle2.py:129-145."""
    return F.dropout(lin(four_linears(temporal, P, "dummy_encoder."), P, "dummy_encoder.dummy_fusion."), 0.2, drop)


def sales_encoder(item_sales, P, drop):
    """SalesEncoder, GTM_Visuelle2.py:99-107 -> all GRU outputs with dropout .1."""
    hid = P["sales_encoder.gru.weight_hh_l0"].shape[1]
    out, _ = gru_seq(item_sales, item_sales.new_zeros(item_sales.shape[0], hid), P, "sales_encoder.gru.")
    return F.dropout(out, 0.1, drop)


# --------------------------------------------------------------------------- fusion networks
def fusion_gtm(h_img, h_text, h_dummy, P, bn_batch, drop):
    """GTMFusionNetwork, GTM_Visuelle2.py:151-172."""
    x = torch.cat([h_img, h_text.flatten(1), h_dummy], 1)
    x = batch_norm1d(x, P, "fusion_network.feature_fusion.0.", bn_batch)
    x = F.dropout(F.relu(lin(x, P, "fusion_network.feature_fusion.1.")), 0.2, drop)
    return lin(x, P, "fusion_network.feature_fusion.4.")


def fusion_v4(h_img, h_text, h_dummy, P, bn_batch, drop):
    """TextGuidedFusionNetwork, Proposed_model_v4.py:152-198 (built with dropout=0.1, :224)."""
    t = h_text.flatten(1)
    g_i = torch.sigmoid(lin(torch.cat([t, h_img], 1), P, "fusion_network.img_gate_fc."))
    g_d = torch.sigmoid(lin(torch.cat([t, h_dummy], 1), P, "fusion_network.dummy_gate_fc."))
    x = torch.cat([h_img + h_img * g_i, t, h_dummy + h_dummy * g_d], 1)
    x = layer_norm(lin(x, P, "fusion_network.fusion_fc.0."), P, "fusion_network.fusion_fc.1.")
    return F.dropout(F.relu(x), 0.1, drop)


def fusion_v1(h_img, h_text, h_dummy, P, bn_batch, drop):
    """ResidualGatedFusionNetwork, Proposed_model.py:141-188."""
    def block(x, name):
        g = torch.sigmoid(lin(x, P, f"fusion_network.{name}.gate_fc."))
        return layer_norm(x + x * g, P, f"fusion_network.{name}.norm.")
    x = torch.cat([block(h_img, "img_gate"), block(h_text.flatten(1), "text_gate"), block(h_dummy, "dummy_gate")], 1)
    return F.dropout(F.relu(lin(x, P, "fusion_network.fusion_fc.0.")), 0.2, drop)


def fusion_v2(h_img, h_text, h_dummy, P, bn_batch, drop):
    """PureGatedFusionNetwork, Proposed_model_v2.py:604-637."""
    x = torch.cat([h_img, h_text.flatten(1), h_dummy], 1)
    x = x + x * torch.sigmoid(lin(x, P, "fusion_network.gate_fc."))
    return F.dropout(F.relu(lin(x, P, "fusion_network.fusion_fc.0.")), 0.2, drop)


def fusion_v3(e_temp, e_text, e_vis, P, bn_batch, drop, query_modality):
    """TARGFusionNetwork, Proposed_model_v3.py:175-236."""
    Q, C1, C2 = {"text": (e_text, e_vis, e_temp), "image": (e_vis, e_text, e_temp),
                 "temporal": (e_temp, e_text, e_vis)}[query_modality]
    f1 = C1 * torch.sigmoid(lin(torch.cat([Q, C1], 1), P, "fusion_network.gate_fc1."))
    f2 = C2 * torch.sigmoid(lin(torch.cat([Q, C2], 1), P, "fusion_network.gate_fc2."))
    x = Q + f1 + f2
    x = batch_norm1d(x, P, "fusion_network.fusion_final.net.0.", bn_batch)
    x = F.dropout(F.relu(lin(x, P, "fusion_network.fusion_final.net.1.")), 0.2, drop)
    return lin(x, P, "fusion_network.fusion_final.net.4.")


def _fusion_block(x, P, prefix, bn_batch, drop):
    """FusionBlock, M4FT_Visuelle2.py:161-173 / Proposed_model_v3.py:160-172."""
    x = batch_norm1d(x, P, prefix + "net.0.", bn_batch)
    x = F.dropout(F.relu(lin(x, P, prefix + "net.1.")), 0.2, drop)
    return lin(x, P, prefix + "net.4.")


def fusion_m4ft(e_temp, e_text, e_vis, P, bn_batch, drop):
    """M4FTFusionNetwork, M4FT_Visuelle2.py:175-202."""
    out_tt = _fusion_block(e_temp + e_text, P, "fusion_network.fusion_temp_text.", bn_batch, drop)
    out_tv = _fusion_block(e_text + e_vis, P, "fusion_network.fusion_text_vis.", bn_batch, drop)
    return _fusion_block(out_tt + out_tv + e_temp + e_text + e_vis, P, "fusion_network.fusion_final.", bn_batch, drop)


# --------------------------------------------------------------------------- v1 / v2 custom layers
def decoder_layer_v1(tgt, mem, P, prefix, heads, tgt_mask, p_drop, drop):
    """GatedTransformerDecoderLayer, Proposed_model.py:226-262 (cross-attn output * sigmoid(gate_proj(query)))."""
    a = mha(tgt, tgt, P, prefix + "self_attn.", heads, tgt_mask, p_drop, drop)
    tgt = layer_norm(tgt + F.dropout(a, p_drop, drop), P, prefix + "norm1.")
    a = mha(tgt, mem, P, prefix + "cross_attn.mha.", heads, None, p_drop, drop)
    a = F.dropout(a * torch.sigmoid(lin(tgt, P, prefix + "cross_attn.gate_proj.")), p_drop, drop)
    tgt = layer_norm(tgt + a, P, prefix + "norm2.")
    f = lin(F.dropout(F.relu(lin(tgt, P, prefix + "linear1.")), p_drop, drop), P, prefix + "linear2.")
    return layer_norm(tgt + F.dropout(f, p_drop, drop), P, prefix + "norm3.")


def _qkv_attention(q_in, kv_in, P, prefix, heads, mask, p_drop, drop):
    q, k, v = lin(q_in, P, prefix + "q_proj."), lin(kv_in, P, prefix + "k_proj."), lin(kv_in, P, prefix + "v_proj.")
    return attention(q, k, v, heads, mask, p_drop, drop)


def decoder_layer_v2(tgt, mem, P, prefix, heads, tgt_mask, p_drop, drop):
    """Proposed_model_v2.py:713-741: cross-attention = PureGatedMultiheadAttention (:546-602)."""
    a = mha(tgt, tgt, P, prefix + "self_attn.", heads, tgt_mask, p_drop, drop)
    tgt = layer_norm(tgt + F.dropout(a, p_drop, drop), P, prefix + "norm1.")
    o, _ = _qkv_attention(tgt, mem, P, prefix + "cross_attn.", heads, None, p_drop, drop)
    o = lin(o * torch.sigmoid(lin(tgt, P, prefix + "cross_attn.gate_proj.")), P, prefix + "cross_attn.out_proj.")
    tgt = layer_norm(tgt + o, P, prefix + "norm2.")
    f = lin(F.dropout(F.relu(lin(tgt, P, prefix + "linear1.")), p_drop, drop), P, prefix + "linear2.")
    return layer_norm(tgt + F.dropout(f, p_drop, drop), P, prefix + "norm3.")


def encoder_layer_v2(x, P, prefix, heads, mask, p_drop, drop):
    """Proposed_model_v2.py:692-711: self-attention = HeadSpecificGatedAttention (:643-690), a per-head
    gate sigmoid(Linear_hd(q_head)) on the attention output before out_proj."""
    L, N, D = x.shape
    hd = D // heads
    o, qh = _qkv_attention(x, x, P, prefix + "self_attn.", heads, mask, p_drop, drop)
    oh = o.reshape(L, N * heads, hd).transpose(0, 1)                     # [N*heads, L, hd]
    oh = oh * torch.sigmoid(lin(qh, P, prefix + "self_attn.gate_proj."))
    o = lin(oh.transpose(0, 1).reshape(L, N, D), P, prefix + "self_attn.out_proj.")
    x = layer_norm(x + F.dropout(o, p_drop, drop), P, prefix + "norm1.")
    f = lin(F.dropout(F.relu(lin(x, P, prefix + "linear1.")), p_drop, drop), P, prefix + "linear2.")
    return layer_norm(x + F.dropout(f, p_drop, drop), P, prefix + "norm2.")


# --------------------------------------------------------------------------- model forward
def gtm_family_forward(variant, P, item_sales, cat, col, fab, store, temporal, gtrends, feat, *, output_len,
                       heads, num_layers=1, use_encoder_mask=1, autoregressive=False, training=False,
                       drop=None, query_modality="image"):
    """``variant`` in {'gtm','v1','v2','v3','v4','m4ft'}.  Returns (forecast [N, output_len], None)."""
    drop = training if drop is None else drop
    if item_sales.dim() == 3:
        bs, splits, window = item_sales.shape
    else:
        bs, window = item_sales.shape
        splits = 1
        item_sales = item_sales.unsqueeze(1)
    if variant == "v2":
        enc = gtrend_encoder(gtrends, P, "", heads, output_len, use_encoder_mask, drop,
                             lin_name="gtrend_input_linear.module.", enc_name="gtrend_encoder.layers.",
                             layer_fn=encoder_layer_v2)
    else:
        enc = gtrend_encoder(gtrends, P, "gtrend_encoder.", 4, output_len, use_encoder_mask, drop)
    if variant in ("v3", "m4ft"):
        e_text = F.dropout(lin(torch.cat([P["text_encoder.cat_emb.weight"][cat], P["text_encoder.col_emb.weight"][col],
                                          P["text_encoder.fab_emb.weight"][fab],
                                          P["text_encoder.store_emb.weight"][store]], 1), P, "text_encoder.proj."),
                           0.1, drop)
        e_vis = lin(image_projection(feat, P), P, "image_encoder.final_proj.")
        e_temp = F.dropout(lin(four_linears(temporal, P, "temporal_encoder."), P, "temporal_encoder.proj."), 0.2, drop)
        statics = [e_temp, e_text, e_vis]
    else:
        statics = [image_projection(feat, P), text_encoder(cat, col, fab, store, P, drop),
                   dummy_encoder(temporal, P, drop)]
    if splits > 1:
        enc = enc.repeat_interleave(splits, dim=1)
        statics = [s.repeat_interleave(splits, dim=0) for s in statics]
    h_sales = sales_encoder(item_sales.reshape(bs * splits, window, 1), P, drop)
    if variant == "m4ft":
        ctx = fusion_m4ft(*statics, P, training, drop)
    elif variant == "v3":
        ctx = fusion_v3(*statics, P, training, drop, query_modality)
    else:
        ctx = {"gtm": fusion_gtm, "v1": fusion_v1, "v2": fusion_v2, "v4": fusion_v4}[variant](*statics, P, training, drop)
    dec_in = h_sales[:, -1, :] + ctx
    layer = {"v1": decoder_layer_v1, "v2": decoder_layer_v2}.get(variant, decoder_layer)
    if autoregressive:
        tgt = torch.zeros(output_len, dec_in.shape[0], dec_in.shape[1])
        tgt = torch.cat([dec_in.unsqueeze(0), tgt[1:]], 0)
        tgt = F.dropout(tgt + positional_encoding(output_len, dec_in.shape[1]), 0.1, drop)
        tmask = causal_mask(output_len)
    else:
        tgt, tmask = dec_in.unsqueeze(0), None
    for i in range(num_layers):
        tgt = layer(tgt, enc, P, f"decoder.layers.{i}.", heads, tmask, 0.1, drop)
    out = F.dropout(lin(tgt, P, "decoder_fc.0."), 0.2, drop)
    return out.transpose(0, 1).reshape(bs * splits, output_len), None
