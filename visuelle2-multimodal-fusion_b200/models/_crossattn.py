"""Common body of the three CrossAttnRNN drop-ins: hoisted static encoding + fused decoder call."""
import torch

from .. import functional as Fv
from .modules import decoder_weights, static_embed


def draw_teacher_forcing(n_steps, ratio):
    """``torch.rand(1) < ratio`` once per step on the host CPU generator, in the reference's order
    (models/CrossAttnRNN210.py:216-217), packed into a bit mask for the kernel."""
    mask = 0
    for t in range(n_steps):
        if bool(torch.rand(1) < ratio):
            mask |= 1 << t
    return mask


def encode_static(m, categories, colors, fabrics, stores, temporal_features, gtrends, images, by_proj,
                  use_trends=True):
    """Everything that depends only on the item, computed once per item (SURVEY.md 8a identities
    1-4): image map V, trend sequence Vtr, (date, attributes) and their step-invariant projections."""
    V = m.image_encoder(images)                                           # [B,Li,E]
    G = m.trend_encoder(gtrends.permute(0, 2, 1).float())                 # [B,52,E]
    Mst = static_embed(m.temp_encoder, m.attribute_encoder, temporal_features, categories, colors,
                       fabrics, stores, m.training)                       # [B,2,E]
    a = m.ts_self_attention
    Vtr = Fv.mha_self(G, a.in_proj_weight, a.in_proj_bias, a.out_proj.weight, a.out_proj.bias,
                      a.num_heads, a.dropout, m.training) if use_trends else G
    Himg = Fv.linear(V, m.img_attention.encoder_linear.weight)
    Htr = Fv.linear(Vtr, m.ts_attention.encoder_linear.weight)
    HMst = Fv.linear(Mst, m.multimodal_attention.encoder_linear.weight)
    Ptr = Fv.trend_proj(Htr if by_proj else Vtr, m.trend_linear.weight)
    return V, Himg, Htr, Ptr, Mst, HMst


def flatten_windows(X, y):
    num_windows = 1
    if X.dim() == 3:
        bs, num_windows, hist_len = X.shape
        X = X.reshape(bs * num_windows, hist_len)
        if y is not None:
            y = y.reshape(bs * num_windows, -1)
    else:
        bs = X.shape[0]
    if X.dim() == 2:
        X = X.unsqueeze(-1)
    return X.float().contiguous(), (y.float().contiguous() if y is not None else None), bs, num_windows


def sales_state(m, X):
    g = m.sales_encoder_gru
    h0 = X.new_zeros(X.shape[0], g.hidden_size)
    out = Fv.gru_seq(X, h0, g.weight_ih_l0, g.weight_hh_l0, g.bias_ih_l0, g.bias_hh_l0)
    return out[:, -1, :]


def run_decoder(m, variant, W, T, tf_mask, mod_mask, tiles, h0, x0, y, gru, fc):
    V, Himg, Htr, Ptr, Mst, HMst = tiles
    return Fv.decode(variant, W, T, tf_mask, mod_mask, Himg, V, Htr, Ptr, Mst, HMst, h0, x0, y,
                     decoder_weights(m, gru, fc))
