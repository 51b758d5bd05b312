"""Fused sub-modules (the public home the reference's dead ``models/modules.py`` names).

Each class owns parameters with the reference's names and shapes -- the stock torch layers are
instantiated as *parameter containers* so default initialisation, RNG consumption and state_dict
keys are identical to the reference -- but ``forward`` runs the CUDA kernels of this package.
Semantics that differ between the reference's per-file copies are constructor flags.
"""
import torch
import torch.nn as nn

from .. import functional as Fv
from .. import trunk
from ._base import resnet101_trunk


class TSEmbedder(nn.Module):
    """52-step GRU over the Google-Trends series + dropout (models/CrossAttnRNN210.py:12-24)."""

    def __init__(self, input_dim, embedding_dim):
        super().__init__()
        self.ts_embedder = nn.GRU(input_size=input_dim, hidden_size=embedding_dim, num_layers=1, batch_first=True)
        self.dropout = nn.Dropout(0.1)

    def forward(self, x):
        g = self.ts_embedder
        h0 = x.new_zeros(x.shape[0], g.hidden_size)
        out = Fv.gru_seq(x.contiguous(), h0, g.weight_ih_l0, g.weight_hh_l0, g.bias_ih_l0, g.bias_hh_l0)
        return Fv.dropout(out, self.dropout.p, self.training)


class AttributeEncoder(nn.Module):
    """Parameter container; evaluated together with TemporalFeatureEncoder by ``static_embed``."""

    def __init__(self, num_cat, num_col, num_fab, num_store, embedding_dim):
        super().__init__()
        self.cat_embedder = nn.Embedding(num_cat, embedding_dim)
        self.col_embedder = nn.Embedding(num_col, embedding_dim)
        self.fab_embedder = nn.Embedding(num_fab, embedding_dim)
        self.store_embedder = nn.Embedding(num_store, embedding_dim)
        self.dropout = nn.Dropout(0.1)

    def tables(self):
        return [self.cat_embedder.weight, self.col_embedder.weight, self.fab_embedder.weight,
                self.store_embedder.weight]


class TemporalFeatureEncoder(nn.Module):
    """``day_only`` reproduces the Demand copy that routes all four features through
    ``day_embedding`` (models/CrossAttnRNNDemand.py:61-64): week/month/year then get no gradient."""

    def __init__(self, embedding_dim, day_only=False):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.day_embedding = nn.Linear(1, embedding_dim)
        self.week_embedding = nn.Linear(1, embedding_dim)
        self.month_embedding = nn.Linear(1, embedding_dim)
        self.year_embedding = nn.Linear(1, embedding_dim)
        self.dropout = nn.Dropout(0.1)
        self.day_only = day_only

    def packed(self):
        mods = [self.day_embedding] * 4 if self.day_only else \
            [self.day_embedding, self.week_embedding, self.month_embedding, self.year_embedding]
        Wt = torch.stack([m.weight[:, 0] for m in mods], 0)
        bt = torch.stack([m.bias for m in mods], 0)
        return Wt, bt


def static_embed(temp_enc, attr_enc, temporal, cat, col, fab, store, training):
    """(date, attributes) -> [B,2,E] in one kernel (models/CrossAttnRNN210.py:35-40,51-56)."""
    Wt, bt = temp_enc.packed()
    idx = torch.stack([cat, col, fab, store], 0).to(torch.int64)
    B, E = temporal.shape[0], Wt.shape[1]
    drop = Fv.keep_mask((B, 8, E), temp_enc.dropout.p, training, temporal.device)
    return Fv.embed(temporal.contiguous().float(), Wt, bt, attr_enc.tables(), idx, drop)


class ImageEncoder(nn.Module):
    """torchvision ResNet-101 trunk (kept as is) -> fused ``fc`` 2048->E + dropout over the
    ``h*w`` positions (models/CrossAttnRNN210.py:58-72).  Accepts images ``[B,3,H,W]``."""

    def __init__(self, embedding_dim=300):
        super().__init__()
        self.cnn = resnet101_trunk()
        self.fc = nn.Linear(2048, embedding_dim)
        self.dropout = nn.Dropout(0.1)
        self.backbone_dtype = None       # torch.bfloat16 -> run the torchvision trunk under autocast

    def use_bf16_backbone(self, on=True):
        """bf16 autocast + channels_last for the (unreplaced) torchvision/cuDNN trunk."""
        self.backbone_dtype = torch.bfloat16 if on else None
        self.cnn.to(memory_format=torch.channels_last if on else torch.contiguous_format)
        self.fused_trunk = bool(on) and trunk.supported(self.cnn)    # BN/add/ReLU sweeps of csrc/bn_act.cu
        return self

    def forward(self, x):
        if self.backbone_dtype is not None and x.dim() == 4 and x.shape[1] == 3:
            if getattr(self, "fused_trunk", False):
                feat = trunk.forward(self.cnn, x)
            else:
                with torch.autocast("cuda", dtype=self.backbone_dtype):
                    feat = self.cnn(x.contiguous(memory_format=torch.channels_last))
        else:
            feat = self.cnn(x)
        B, C = feat.shape[0], feat.shape[1]
        rows = feat.permute(0, 2, 3, 1).reshape(B, -1, C)      # view when the trunk ran channels_last
        v = Fv.linear(rows, self.fc.weight, self.fc.bias)          # bf16 rows feed the bf16 tcgen05 GEMM directly
        return Fv.dropout(v, self.dropout.p, self.training)


class AdditiveAttention(nn.Module):
    """Bahdanau attention parameters (models/CrossAttnRNN210.py:74-89).  Inside the models the
    three attentions of a decode step run fused in ``v2f_decode_*``; this class carries the weights."""

    def __init__(self, encoder_dim, decoder_dim, attention_dim, weighted_by_projection=False):
        super().__init__()
        self.encoder_dim = encoder_dim
        self.encoder_linear = nn.Linear(encoder_dim, attention_dim, bias=False)
        self.decoder_linear = nn.Linear(decoder_dim, attention_dim, bias=False)
        self.attn_linear = nn.Linear(attention_dim, 1)
        self.tanh = nn.Tanh()
        self.softmax = nn.Softmax(dim=1)
        self.weighted_by_projection = weighted_by_projection


def decoder_weights(m, gru, fc):
    """Collect the tensors ``functional.decode`` needs from a CrossAttnRNN-family module."""
    w = dict(Wd_img=m.img_attention.decoder_linear.weight, Wd_tr=m.ts_attention.decoder_linear.weight,
             Wd_mm=m.multimodal_attention.decoder_linear.weight,
             w_img=m.img_attention.attn_linear.weight, w_tr=m.ts_attention.attn_linear.weight,
             w_mm=m.multimodal_attention.attn_linear.weight,
             b_img=m.img_attention.attn_linear.bias, b_tr=m.ts_attention.attn_linear.bias,
             b_mm=m.multimodal_attention.attn_linear.bias, b_tl=m.trend_linear.bias,
             We_mm=m.multimodal_attention.encoder_linear.weight,
             W_me=m.multimodal_embedder.weight, b_me=m.multimodal_embedder.bias,
             w_fc=fc.weight, b_fc=fc.bias)
    if gru is not None:
        w.update(W_ih=gru.weight_ih_l0, W_hh=gru.weight_hh_l0, b_ih=gru.bias_ih_l0, b_hh=gru.bias_hh_l0)
    return w
