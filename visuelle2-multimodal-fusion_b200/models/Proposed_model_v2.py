"""Drop-in for the reference's ``models/Proposed_model_v2.py`` (v2: head-specific gated encoder
self-attention, post-concat gated cross-attention, soft-gated fusion; gate biases start at +2).
Surface: ``/root/reference/models/Proposed_model_v2.py:546-898``."""
import torch.nn as nn

from .. import functional as Fv
from .. import functional_gtm as Fg
from ._gtm import (AttributeEncoder, DummyEmbedder, GTMFamilyBase, ImageEncoder, LayerStack, PositionalEncoding,
                   SalesEncoder, TimeDistributed, _add_norm, _ffn, _mha_self, encoder_mask, make_decoder_fc)


class _QKVGated(nn.Module):
    """q/k/v/out projections + a sigmoid gate projection whose bias starts at +2 (parameter container)."""

    def __init__(self, embed_dim, num_heads, gate_dim, dropout=0.1):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == embed_dim, "embed_dim must be divisible by num_heads"
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.k_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        self.gate_proj = nn.Linear(gate_dim, gate_dim)
        nn.init.constant_(self.gate_proj.bias, 2.0)
        self.dropout = nn.Dropout(dropout)
        self.scale = self.head_dim ** -0.5


class PureGatedMultiheadAttention(_QKVGated):
    """Cross attention: out_proj(attn * sigmoid(gate_proj(query)))  (Proposed_model_v2.py:546-602)."""

    def __init__(self, embed_dim, num_heads, dropout=0.1):
        super().__init__(embed_dim, num_heads, embed_dim, dropout)


class HeadSpecificGatedAttention(_QKVGated):
    """Encoder self attention with a per-head gate sigmoid(Linear_hd(q_head)) (Proposed_model_v2.py:643-690)."""

    def __init__(self, embed_dim, num_heads, dropout=0.1):
        super().__init__(embed_dim, num_heads, embed_dim // num_heads, dropout)


def _lin(x, m):
    return Fv.linear(x, m.weight, m.bias)


def _attend(attn, q, k, v, mask, training):
    N, Lq, _ = q.shape
    drop = Fv.keep_mask((N, attn.num_heads, Lq, k.shape[1]), attn.dropout.p, training, q.device)
    return Fv.sdpa(q, k, v, attn.num_heads, mask, drop, attn.scale)


class PureGatedFusionNetwork(nn.Module):
    """x + x*sigmoid(Linear(x)) on the concatenated statics, Linear -> ReLU -> Dropout
    (Proposed_model_v2.py:604-637)."""

    def __init__(self, embedding_dim, hidden_dim, dropout=0.2):
        super().__init__()
        self.img_dim = embedding_dim
        self.text_dim = embedding_dim * 4
        self.dummy_dim = embedding_dim
        total_dim = self.img_dim + self.text_dim + self.dummy_dim
        self.gate_fc = nn.Linear(total_dim, total_dim)
        nn.init.constant_(self.gate_fc.bias, 2.0)
        self.fusion_fc = nn.Sequential(nn.Linear(total_dim, hidden_dim), nn.ReLU(), nn.Dropout(dropout))

    def forward(self, img_encoding, text_encoding, dummy_encoding):
        x = Fg.concat_cols(img_encoding, text_encoding.flatten(1), dummy_encoding)
        x = Fg.gate(x, _lin(x, self.gate_fc), residual=True)
        fc = self.fusion_fc[0]
        return Fv.dropout(Fv.linear(x, fc.weight, fc.bias, act=1), self.fusion_fc[2].p, self.training)


class GatedTransformerEncoderLayer(nn.Module):
    """Parameter container (Proposed_model_v2.py:692-711)."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1):
        super().__init__()
        self.self_attn = HeadSpecificGatedAttention(d_model, nhead, dropout=dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout2 = nn.Dropout(dropout)


class GatedTransformerDecoderLayer(nn.Module):
    """Parameter container (Proposed_model_v2.py:713-741)."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.cross_attn = PureGatedMultiheadAttention(d_model, nhead, dropout=dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.dropout2 = nn.Dropout(dropout)
        self.dropout3 = nn.Dropout(dropout)


class GatedMultimodal_Visuelle2(GTMFamilyBase):
    def __init__(self, embedding_dim, hidden_dim, output_dim, num_heads, num_layers, use_text, use_img,
                 cat_dict, col_dict, fab_dict, store_num, trend_len, num_trends, gpu_num, use_encoder_mask=1,
                 autoregressive=False):
        super().__init__()
        self._init_common(embedding_dim, hidden_dim, output_dim, gpu_num, autoregressive)
        self.save_hyperparameters()
        self.sales_encoder = SalesEncoder(input_dim=1, embedding_dim=hidden_dim)
        self.text_encoder = AttributeEncoder(len(cat_dict) + 1, len(col_dict) + 1, len(fab_dict) + 1, store_num + 1,
                                             embedding_dim)
        self.image_encoder = ImageEncoder(embedding_dim)
        self.dummy_encoder = DummyEmbedder(embedding_dim)
        self.gtrend_input_linear = TimeDistributed(nn.Linear(num_trends, hidden_dim))
        self.gtrend_pos_embedding = PositionalEncoding(hidden_dim, max_len=trend_len)
        enc_layer = GatedTransformerEncoderLayer(d_model=hidden_dim, nhead=num_heads, dropout=0.2)
        self.gtrend_encoder = LayerStack(enc_layer, 2)
        self.use_encoder_mask = use_encoder_mask
        self.trend_len = trend_len
        self.fusion_network = PureGatedFusionNetwork(embedding_dim, hidden_dim)
        self.decoder_linear = TimeDistributed(nn.Linear(1, hidden_dim))
        dec_layer = GatedTransformerDecoderLayer(d_model=hidden_dim, nhead=num_heads, dim_feedforward=hidden_dim * 4,
                                                 dropout=0.1)
        if autoregressive:
            self.pos_encoder = PositionalEncoding(hidden_dim, max_len=12)
        self.decoder = LayerStack(dec_layer, num_layers)
        self.decoder_fc = make_decoder_fc(hidden_dim, self.output_len, autoregressive)

    def _trend_memory(self, gtrends):
        tr = self.training
        x = self.gtrend_input_linear(gtrends.permute(0, 2, 1).float().contiguous())
        x = self.gtrend_pos_embedding(x)
        mask = encoder_mask(x.shape[1], self.output_len, x.device) if self.use_encoder_mask == 1 else None
        for layer in self.gtrend_encoder.layers:
            at = layer.self_attn
            q = _lin(x, at.q_proj)
            o = _attend(at, q, _lin(x, at.k_proj), _lin(x, at.v_proj), mask, tr)
            B, L, D = q.shape
            qh = q.view(B, L, at.num_heads, at.head_dim)                 # the gate reads the (unscaled) per-head query
            o = Fg.gate(o, _lin(qh, at.gate_proj).view(B, L, D))
            x = _add_norm(x, _lin(o, at.out_proj), layer.norm1, layer.dropout1.p, tr)
            x = _add_norm(x, _ffn(x, layer, tr), layer.norm2, layer.dropout2.p, tr)
        return x

    def _decoder_layer(self, x, memory, W, layer, tgt_mask):
        tr = self.training
        a = _mha_self(x, layer.self_attn, tgt_mask, tr)
        x = _add_norm(x, a, layer.norm1, layer.dropout1.p, tr)
        ca = layer.cross_attn
        k = Fg.repeat_rows(_lin(memory, ca.k_proj), W)
        v = Fg.repeat_rows(_lin(memory, ca.v_proj), W)
        o = _attend(ca, _lin(x, ca.q_proj), k, v, None, tr)
        o = _lin(Fg.gate(o, _lin(x, ca.gate_proj)), ca.out_proj)
        x = _add_norm(x, o, layer.norm2, 0.0, tr)                        # no dropout on this branch (:737-738)
        return _add_norm(x, _ffn(x, layer, tr), layer.norm3, layer.dropout3.p, tr)
