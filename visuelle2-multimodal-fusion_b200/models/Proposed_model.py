"""Drop-in for the reference's ``models/Proposed_model.py`` (v1: residual gated fusion network +
gated cross-attention decoder).  Surface: ``/root/reference/models/Proposed_model.py:141-417``."""
import torch.nn as nn

from .. import functional as Fv
from .. import functional_gtm as Fg
from ._gtm import (AttributeEncoder, DummyEmbedder, GTMFamilyBase, GTrendEmbedder, ImageEncoder, LayerStack,
                   PositionalEncoding, SalesEncoder, TimeDistributed, _add_norm, _cross_mha, _ffn, _mha_self,
                   make_decoder_fc)


class GatedResidualBlock(nn.Module):
    """LayerNorm(x + x * sigmoid(Linear(x)))  (Proposed_model.py:141-155)."""

    def __init__(self, input_dim):
        super().__init__()
        self.gate_fc = nn.Linear(input_dim, input_dim)
        self.norm = nn.LayerNorm(input_dim)

    def forward(self, x):
        g = Fg.gate(x, Fv.linear(x, self.gate_fc.weight, self.gate_fc.bias), residual=True)
        return Fg.add_layer_norm(g, None, None, self.norm.weight, self.norm.bias, self.norm.eps)


class ResidualGatedFusionNetwork(nn.Module):
    """Per-modality gated residual blocks, concat, Linear -> ReLU -> Dropout (Proposed_model.py:157-188)."""

    def __init__(self, embedding_dim, hidden_dim, dropout=0.2):
        super().__init__()
        self.img_dim = embedding_dim
        self.text_dim = embedding_dim * 4
        self.dummy_dim = embedding_dim
        self.img_gate = GatedResidualBlock(self.img_dim)
        self.text_gate = GatedResidualBlock(self.text_dim)
        self.dummy_gate = GatedResidualBlock(self.dummy_dim)
        total_dim = self.img_dim + self.text_dim + self.dummy_dim
        self.fusion_fc = nn.Sequential(nn.Linear(total_dim, hidden_dim), nn.ReLU(), nn.Dropout(dropout))

    def forward(self, img_encoding, text_encoding, dummy_encoding):
        x = Fg.concat_cols(self.img_gate(img_encoding), self.text_gate(text_encoding.flatten(1)),
                           self.dummy_gate(dummy_encoding))
        fc = self.fusion_fc[0]
        return Fv.dropout(Fv.linear(x, fc.weight, fc.bias, act=1), self.fusion_fc[2].p, self.training)


class GatedCrossAttention(nn.Module):
    """Parameter container of the query-gated cross attention (Proposed_model.py:194-224); ``norm`` is
    constructed but unused there too."""

    def __init__(self, d_model, nhead, dropout=0.1):
        super().__init__()
        self.mha = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.gate_proj = nn.Linear(d_model, d_model)
        self.dropout = nn.Dropout(dropout)
        self.norm = nn.LayerNorm(d_model)


class GatedTransformerDecoderLayer(nn.Module):
    """Parameter container (Proposed_model.py:226-262)."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.cross_attn = GatedCrossAttention(d_model, nhead, dropout=dropout)
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
        self.gtrend_encoder = GTrendEmbedder(output_dim, hidden_dim, use_encoder_mask, trend_len, num_trends, gpu_num)
        self.sales_encoder = SalesEncoder(input_dim=1, embedding_dim=hidden_dim)
        self.text_encoder = AttributeEncoder(len(cat_dict) + 1, len(col_dict) + 1, len(fab_dict) + 1, store_num + 1,
                                             embedding_dim)
        self.image_encoder = ImageEncoder(embedding_dim)
        self.dummy_encoder = DummyEmbedder(embedding_dim)
        self.fusion_network = ResidualGatedFusionNetwork(embedding_dim, hidden_dim)
        self.decoder_linear = TimeDistributed(nn.Linear(1, hidden_dim))
        layer = GatedTransformerDecoderLayer(d_model=hidden_dim, nhead=num_heads, dim_feedforward=hidden_dim * 4,
                                             dropout=0.1)
        if autoregressive:
            self.pos_encoder = PositionalEncoding(hidden_dim, max_len=12)
        self.decoder = LayerStack(layer, num_layers)
        self.decoder_fc = make_decoder_fc(hidden_dim, self.output_len, autoregressive)

    def _decoder_layer(self, x, memory, W, layer, tgt_mask):
        tr = self.training
        a = _mha_self(x, layer.self_attn, tgt_mask, tr)
        x = _add_norm(x, a, layer.norm1, layer.dropout1.p, tr)
        ca = layer.cross_attn

        def kv_of(attn, xq):
            q, kv = Fg.cross_proj(xq, memory, attn.in_proj_weight, attn.in_proj_bias)
            return q, Fg.repeat_rows(kv, W)

        a = _cross_mha(x, kv_of, ca.mha, tr)
        a = Fg.gate(a, Fv.linear(x, ca.gate_proj.weight, ca.gate_proj.bias))
        x = _add_norm(x, a, layer.norm2, ca.dropout.p, tr)          # dropout(gated) then residual + norm2
        return _add_norm(x, _ffn(x, layer, tr), layer.norm3, layer.dropout3.p, tr)
