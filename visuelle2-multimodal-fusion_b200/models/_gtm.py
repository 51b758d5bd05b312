"""Shared body of the GTM-family drop-ins (GTM_Visuelle2, Proposed_model v1-v4).

The classes below own parameters with the reference's names, shapes and construction order (stock
torch layers are instantiated as *parameter containers*, so default initialisation, RNG consumption
and state_dict keys match the reference), while every ``forward`` runs kernels of libv2f_b200.so
through ``functional`` / ``functional_gtm``.  Internally activations are batch-first ``[B,L,D]``
(the reference is sequence-first; the arithmetic is layout independent).

Reference: /root/reference/models/GTM_Visuelle2.py (helpers :13-145, forward :215-262, hooks :264-312)
and the Proposed_model*.py siblings cited per class.
"""
import copy
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as Fv
from .. import functional_gtm as Fg
from .. import trunk
from ._base import LightningBase, make_adafactor, resnet101_trunk


# --------------------------------------------------------------------------- helpers (parameter containers)
class PositionalEncoding(nn.Module):
    """sin/cos table registered as buffer ``pe`` [max_len,1,d] (GTM_Visuelle2.py:13-28)."""

    def __init__(self, d_model, dropout=0.1, max_len=52):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0).transpose(0, 1))

    def forward(self, x):
        """x batch-first [B,L,D]: + pe[:L] then dropout."""
        L = x.shape[1]
        x = Fg.add_bcast(x, self.pe[:L, 0, :])
        return Fv.dropout(x, self.dropout.p, self.training)


class TimeDistributed(nn.Module):
    """Holds ``module`` (state_dict key ``*.module.*``); applied over the last dim (GTM_Visuelle2.py:30-44)."""

    def __init__(self, module, batch_first=True):
        super().__init__()
        self.module = module
        self.batch_first = batch_first

    def forward(self, x):
        return Fv.linear(x, self.module.weight, self.module.bias)


_MASKS = {}


def encoder_mask(size, horizon, device):
    """Additive block-diagonal mask, blocks of gcd(size, horizon) (GTM_Visuelle2.py:57-64); built once per
    (size, horizon, device) instead of on every forward, and on the module's device (not 'cuda:'+gpu_num)."""
    key = ("enc", size, horizon, str(device))
    if key not in _MASKS:
        split = math.gcd(size, horizon)
        m = torch.full((size, size), float("-inf"))
        for i in range(0, size, split):
            m[i:i + split, i:i + split] = 0.0
        _MASKS[key] = m.to(device)
    return _MASKS[key]


def causal_mask(size, device):
    key = ("causal", size, str(device))
    if key not in _MASKS:
        _MASKS[key] = torch.triu(torch.full((size, size), float("-inf")), diagonal=1).to(device)
    return _MASKS[key]


# --------------------------------------------------------------------------- transformer layers on the kernels
def _ffn(x, layer, training):
    p = layer.dropout.p
    f = Fv.linear(x, layer.linear1.weight, layer.linear1.bias, act=1)
    f = Fv.dropout(f, p, training)
    return Fv.linear(f, layer.linear2.weight, layer.linear2.bias)


def _mha_self(x, attn, mask, training):
    return Fv.mha_self(x, attn.in_proj_weight, attn.in_proj_bias, attn.out_proj.weight, attn.out_proj.bias,
                       attn.num_heads, attn.dropout, training, mask)


def _add_norm(x, a, norm, p, training):
    m = Fv.keep_mask(a.shape, p, training, a.device)
    return Fg.add_layer_norm(x, a, m, norm.weight, norm.bias, norm.eps)


def encoder_layer(x, layer, mask, training):
    """nn.TransformerEncoderLayer, post-LN, ReLU (GTM_Visuelle2.py:52-53)."""
    a = _mha_self(x, layer.self_attn, mask, training)
    x = _add_norm(x, a, layer.norm1, layer.dropout1.p, training)
    return _add_norm(x, _ffn(x, layer, training), layer.norm2, layer.dropout2.p, training)


def _cross_mha(x, kv_of, attn, training):
    """nn.MultiheadAttention(query=x, key=value=memory); ``kv_of(attn)`` yields the packed, window-repeated
    key|value projection [N,Lk,2D] and the query projection."""
    q, kv = kv_of(attn, x)
    N, Lq, D = q.shape
    drop = Fv.keep_mask((N, attn.num_heads, Lq, kv.shape[1]), attn.dropout, training, x.device)
    o = Fg.sdpa_kv(q, kv, attn.num_heads, None, drop)
    return Fv.linear(o, attn.out_proj.weight, attn.out_proj.bias)


def decoder_layer(x, memory, W, layer, tgt_mask, training):
    """nn.TransformerDecoderLayer, post-LN, ReLU (GTM_Visuelle2.py:200-202).  ``memory`` [B,Lk,D] per item;
    rows of ``x`` belong to item n // W."""
    a = _mha_self(x, layer.self_attn, tgt_mask, training)
    x = _add_norm(x, a, layer.norm1, layer.dropout1.p, training)

    def kv_of(attn, xq):
        q, kv = Fg.cross_proj(xq, memory, attn.in_proj_weight, attn.in_proj_bias)
        return q, Fg.repeat_rows(kv, W)

    a = _cross_mha(x, kv_of, layer.multihead_attn, training)
    x = _add_norm(x, a, layer.norm2, layer.dropout2.p, training)
    return _add_norm(x, _ffn(x, layer, training), layer.norm3, layer.dropout3.p, training)


# --------------------------------------------------------------------------- encoders
class GTrendEmbedder(nn.Module):
    """Linear(num_trends->D) + positional encoding + 2 post-LN encoder layers (GTM_Visuelle2.py:46-74)."""

    def __init__(self, forecast_horizon, embedding_dim, use_mask, trend_len, num_trends, gpu_num):
        super().__init__()
        self.forecast_horizon = forecast_horizon
        self.input_linear = TimeDistributed(nn.Linear(num_trends, embedding_dim))
        self.pos_embedding = PositionalEncoding(embedding_dim, max_len=trend_len)
        layer = nn.TransformerEncoderLayer(d_model=embedding_dim, nhead=4, dropout=0.2)
        self.encoder = nn.TransformerEncoder(layer, num_layers=2)
        self.use_mask = use_mask
        self.gpu_num = gpu_num

    def forward(self, gtrends):
        x = self.input_linear(gtrends.permute(0, 2, 1).float().contiguous())      # [B,52,D]
        x = self.pos_embedding(x)
        mask = encoder_mask(x.shape[1], self.forecast_horizon, x.device) if self.use_mask == 1 else None
        for layer in self.encoder.layers:
            x = encoder_layer(x, layer, mask, self.training)
        return x


class AttributeEncoder(nn.Module):
    """Four embedding tables stacked to [B,4,E] + dropout .1 (GTM_Visuelle2.py:81-96)."""

    def __init__(self, num_cat, num_col, num_fab, num_store, embedding_dim):
        super().__init__()
        self.cat_emb = nn.Embedding(num_cat, embedding_dim)
        self.col_emb = nn.Embedding(num_col, embedding_dim)
        self.fab_emb = nn.Embedding(num_fab, embedding_dim)
        self.store_emb = nn.Embedding(num_store, embedding_dim)
        self.dropout = nn.Dropout(0.1)

    def tables(self):
        return [self.cat_emb.weight, self.col_emb.weight, self.fab_emb.weight, self.store_emb.weight]

    def forward(self, cat, col, fab, store):
        return Fg.gather4(self.tables(), cat, col, fab, store, self.dropout.p, self.training)


class SalesEncoder(nn.Module):
    """GRU over the observed sales window, all outputs + dropout .1 (GTM_Visuelle2.py:99-107)."""

    def __init__(self, input_dim, embedding_dim):
        super().__init__()
        self.gru = nn.GRU(input_dim, embedding_dim, batch_first=True)
        self.dropout = nn.Dropout(0.1)

    def forward(self, x):
        g = self.gru
        h0 = x.new_zeros(x.shape[0], g.hidden_size)
        out = Fv.gru_seq(x.contiguous(), h0, g.weight_ih_l0, g.weight_hh_l0, g.bias_ih_l0, g.bias_hh_l0)
        return Fv.dropout(out, self.dropout.p, self.training)


class ImageEncoder(nn.Module):
    """torchvision ResNet-101 trunk (kept) -> global average pool -> 1x1 projection as a GEMM on the
    pooled vector (GTM_Visuelle2.py:110-126; pool and 1x1 conv commute, SURVEY.md 8a identity 5)."""

    def __init__(self, embedding_dim=512):
        super().__init__()
        self.cnn = resnet101_trunk()
        self.projection = nn.Conv2d(2048, embedding_dim, kernel_size=1)
        self.pool = nn.AdaptiveAvgPool2d((1, 1))
        self.backbone_dtype = None

    def use_bf16_backbone(self, on=True):
        self.backbone_dtype = torch.bfloat16 if on else None
        self.cnn.to(memory_format=torch.channels_last if on else torch.contiguous_format)
        self.fused_trunk = bool(on) and trunk.supported(self.cnn)    # BN/add/ReLU sweeps of csrc/bn_act.cu
        return self

    def trunk(self, x):
        if self.backbone_dtype is not None and x.dim() == 4 and x.shape[1] == 3:
            if getattr(self, "fused_trunk", False):
                return trunk.forward(self.cnn, x)
            with torch.autocast("cuda", dtype=self.backbone_dtype):
                return self.cnn(x.contiguous(memory_format=torch.channels_last))
        return self.cnn(x)

    def forward(self, x):
        pooled = Fg.mean_pool(self.trunk(x))                                       # [B,2048] fp32
        return Fv.linear(pooled, self.projection.weight.flatten(1), self.projection.bias)


class DummyEmbedder(nn.Module):
    """Four Linear(1->E) concatenated -> Linear(4E->E) -> dropout .2 (GTM_Visuelle2.py:129-145)."""

    def __init__(self, embedding_dim):
        super().__init__()
        self.day_emb = nn.Linear(1, embedding_dim)
        self.week_emb = nn.Linear(1, embedding_dim)
        self.month_emb = nn.Linear(1, embedding_dim)
        self.year_emb = nn.Linear(1, embedding_dim)
        self.dummy_fusion = nn.Linear(embedding_dim * 4, embedding_dim)
        self.dropout = nn.Dropout(0.2)

    def forward(self, temporal_features):
        f = Fg.feat4(temporal_features, [self.day_emb, self.week_emb, self.month_emb, self.year_emb])
        f = Fv.linear(f.flatten(1), self.dummy_fusion.weight, self.dummy_fusion.bias)
        return Fv.dropout(f, self.dropout.p, self.training)


def clones(layer, n):
    """torch-1.8 ``_get_clones``: the reference builds custom layers inside nn.TransformerEncoder/Decoder,
    which torch >= 2 rejects; a plain ``layers`` ModuleList keeps the state_dict keys (``*.layers.i.*``)."""
    return nn.ModuleList([copy.deepcopy(layer) for _ in range(n)])


class LayerStack(nn.Module):
    def __init__(self, layer, num_layers):
        super().__init__()
        self.layers = clones(layer, num_layers)


# --------------------------------------------------------------------------- model base
class GTMFamilyBase(LightningBase):
    """Forward skeleton + Lightning hooks shared by the five models (GTM_Visuelle2.py:215-312)."""

    precision = "fp32"

    def _init_common(self, embedding_dim, hidden_dim, output_dim, gpu_num, autoregressive):
        self.hidden_dim = hidden_dim
        self.embedding_dim = embedding_dim
        self.output_len = output_dim
        self.gpu_num = gpu_num
        self.autoregressive = autoregressive

    # -- pieces the subclasses provide
    def _trend_memory(self, gtrends):
        return self.gtrend_encoder(gtrends)

    def _statics(self, cat, col, fab, store, temporal, images):
        return [self.image_encoder(images), self.text_encoder(cat, col, fab, store),
                self.dummy_encoder(temporal)]

    def _decoder_layer(self, x, memory, W, layer, tgt_mask):
        return decoder_layer(x, memory, W, layer, tgt_mask, self.training)

    # -- forward
    def forward(self, item_sales, category, color, fabric, store, temporal_features, gtrends, images):
        with Fv.precision(self.precision, round_tf32=True):
            return self._forward(item_sales, category, color, fabric, store, temporal_features, gtrends, images)

    def _forward(self, item_sales, category, color, fabric, store, temporal_features, gtrends, images):
        if item_sales.dim() == 3:
            bs, W, window = item_sales.shape
        else:
            bs, window = item_sales.shape
            W = 1
        memory = self._trend_memory(gtrends)                                      # [B,52,D], per item
        statics = self._statics(category, color, fabric, store, temporal_features, images)
        statics = [Fg.repeat_rows(s, W) for s in statics]
        sales = item_sales.reshape(bs * W, window, 1).float().contiguous()
        h_sales = self.sales_encoder(sales)
        ctx = self.fusion_network(*statics)
        dec_in = Fg.add(Fg.take_step(h_sales, -1), ctx)                           # [N,D]
        if self.autoregressive == 1:
            T = self.output_len
            x = self.pos_encoder(Fg.put_step0(dec_in, T))
            tgt_mask = causal_mask(T, x.device)
        else:
            x, tgt_mask = dec_in.unsqueeze(1), None
        for layer in self.decoder.layers:
            x = self._decoder_layer(x, memory, W, layer, tgt_mask)
        fc, dp = self.decoder_fc[0], self.decoder_fc[1]
        out = Fv.dropout(Fv.linear(x, fc.weight, fc.bias), dp.p, self.training)   # the forecast itself is dropped
        return out.reshape(bs * W, self.output_len), None

    # -- Lightning hooks
    def configure_optimizers(self):
        return [make_adafactor(self.parameters())]

    def _unpack(self, batch):
        data, images = batch
        if len(data) == 8:
            item_sales, y, cat, col, fab, store, temporal, gtrends = data
        else:
            y, cat, col, fab, store, temporal, gtrends = data
            item_sales = torch.zeros(y.shape[0], 1, 2, device=self.device)
        forecast, _ = self.forward(item_sales, cat, col, fab, store, temporal, gtrends, images)
        return y, forecast

    def training_step(self, train_batch, batch_idx):
        y, forecast = self._unpack(train_batch)
        loss = F.mse_loss(y.reshape(-1), forecast.reshape(-1))
        self.log("train_loss", loss)
        return loss

    def validation_step(self, test_batch, batch_idx):
        y, forecast = self._unpack(test_batch)
        return y.reshape(-1), forecast.reshape(-1)

    def validation_epoch_end(self, val_step_outputs):
        gt = torch.cat([x[0] for x in val_step_outputs]).view(-1)
        pred = torch.cat([x[1] for x in val_step_outputs]).view(-1)
        loss = F.mse_loss(gt, pred)
        mae = F.l1_loss(gt * 53, pred * 53)
        wape = 100 * torch.sum(torch.abs(gt * 53 - pred * 53)) / torch.sum(gt * 53)
        self.log("val_mae", mae)
        self.log("val_wWAPE", wape)
        self.log("val_loss", loss)
        lr = self.optimizers().param_groups[0]["lr"]
        print(f"Validation MAE: {mae.detach().cpu().numpy()}, Validation WAPE: {wape.detach().cpu().numpy()}, LR: {lr}")


def make_decoder_fc(hidden_dim, output_len, autoregressive):
    return nn.Sequential(nn.Linear(hidden_dim, output_len if not autoregressive else 1), nn.Dropout(0.2))
