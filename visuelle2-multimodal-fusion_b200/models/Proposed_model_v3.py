"""Drop-in for the reference's ``models/Proposed_model_v3.py`` (v3: TARG, target-anchored residual
gating over text / image / temporal embeddings).  Surface: ``/root/reference/models/Proposed_model_v3.py:83-376``."""
import torch.nn as nn

from .. import functional as Fv
from .. import functional_gtm as Fg
from ._base import resnet101_trunk
from ._gtm import (GTMFamilyBase, GTrendEmbedder, ImageEncoder as _GTMImageEncoder, PositionalEncoding,
                   SalesEncoder, TimeDistributed, make_decoder_fc)


class AttributeEncoder(nn.Module):
    """Four embeddings concatenated -> Linear(4E->H) -> dropout .1 (Proposed_model_v3.py:83-103)."""

    def __init__(self, num_cat, num_col, num_fab, num_store, embedding_dim, hidden_dim):
        super().__init__()
        self.cat_emb = nn.Embedding(num_cat, embedding_dim)
        self.col_emb = nn.Embedding(num_col, embedding_dim)
        self.fab_emb = nn.Embedding(num_fab, embedding_dim)
        self.store_emb = nn.Embedding(num_store, embedding_dim)
        self.proj = nn.Linear(embedding_dim * 4, hidden_dim)
        self.dropout = nn.Dropout(0.1)

    def forward(self, cat, col, fab, store):
        tabs = [self.cat_emb.weight, self.col_emb.weight, self.fab_emb.weight, self.store_emb.weight]
        e = Fg.gather4(tabs, cat, col, fab, store, 0.0, False).flatten(1)
        return Fv.dropout(Fv.linear(e, self.proj.weight, self.proj.bias), self.dropout.p, self.training)


class ImageEncoder(_GTMImageEncoder):
    """Trunk -> pool -> 1x1 projection -> Linear(E->H) (Proposed_model_v3.py:105-125)."""

    def __init__(self, embedding_dim, hidden_dim):
        nn.Module.__init__(self)
        self.cnn = resnet101_trunk()
        self.projection = nn.Conv2d(2048, embedding_dim, kernel_size=1)
        self.pool = nn.AdaptiveAvgPool2d((1, 1))
        self.final_proj = nn.Linear(embedding_dim, hidden_dim)
        self.backbone_dtype = None

    def forward(self, x):
        return Fv.linear(super().forward(x), self.final_proj.weight, self.final_proj.bias)


class TemporalEmbedder(nn.Module):
    """Four Linear(1->E) concatenated -> Linear(4E->H) -> dropout .2 (Proposed_model_v3.py:127-145)."""

    def __init__(self, embedding_dim, hidden_dim):
        super().__init__()
        self.day_emb = nn.Linear(1, embedding_dim)
        self.week_emb = nn.Linear(1, embedding_dim)
        self.month_emb = nn.Linear(1, embedding_dim)
        self.year_emb = nn.Linear(1, embedding_dim)
        self.proj = nn.Linear(embedding_dim * 4, hidden_dim)
        self.dropout = nn.Dropout(0.2)

    def forward(self, temporal_features):
        f = Fg.feat4(temporal_features, [self.day_emb, self.week_emb, self.month_emb, self.year_emb])
        return Fv.dropout(Fv.linear(f.flatten(1), self.proj.weight, self.proj.bias), self.dropout.p, self.training)


class FusionBlock(nn.Module):
    """BatchNorm1d -> Linear -> ReLU -> Dropout -> Linear (Proposed_model_v3.py:160-172)."""

    def __init__(self, hidden_dim, dropout=0.2):
        super().__init__()
        self.net = nn.Sequential(nn.BatchNorm1d(hidden_dim), nn.Linear(hidden_dim, hidden_dim), nn.ReLU(),
                                 nn.Dropout(dropout), nn.Linear(hidden_dim, hidden_dim))

    def forward(self, x):
        n = self.net
        x = Fg.batch_norm1d(x, n[0], self.training)
        x = Fv.dropout(Fv.linear(x, n[1].weight, n[1].bias, act=1), n[3].p, self.training)
        return Fv.linear(x, n[4].weight, n[4].bias)


class TARGFusionNetwork(nn.Module):
    """Q + C1*sigmoid(fc1[Q;C1]) + C2*sigmoid(fc2[Q;C2]) -> FusionBlock (Proposed_model_v3.py:175-236)."""

    def __init__(self, hidden_dim, query_modality="text", dropout=0.2):
        super().__init__()
        self.query_modality = query_modality
        self.hidden_dim = hidden_dim
        self.gate_fc1 = nn.Linear(hidden_dim * 2, hidden_dim)
        self.gate_fc2 = nn.Linear(hidden_dim * 2, hidden_dim)
        nn.init.constant_(self.gate_fc1.bias, 0.0)
        nn.init.constant_(self.gate_fc2.bias, 0.0)
        self.fusion_final = FusionBlock(hidden_dim, dropout)

    def forward(self, e_temp, e_text, e_vis):
        if self.query_modality == "text":
            Q, C1, C2 = e_text, e_vis, e_temp
        elif self.query_modality == "image":
            Q, C1, C2 = e_vis, e_text, e_temp
        elif self.query_modality == "temporal":
            Q, C1, C2 = e_temp, e_text, e_vis
        else:
            raise ValueError(f"Unknown query modality: {self.query_modality}")
        f1 = Fg.gate(C1, Fv.linear(Fg.concat_cols(Q, C1), self.gate_fc1.weight, self.gate_fc1.bias))
        f2 = Fg.gate(C2, Fv.linear(Fg.concat_cols(Q, C2), self.gate_fc2.weight, self.gate_fc2.bias))
        return self.fusion_final(Fg.add(Fg.add(Q, f1), f2))


class TARG_M4FT_Visuelle2(GTMFamilyBase):
    def __init__(self, embedding_dim, hidden_dim, output_dim, num_heads, num_layers, use_text, use_img,
                 cat_dict, col_dict, fab_dict, store_num, trend_len, num_trends, gpu_num, query_modality="image",
                 use_encoder_mask=1, autoregressive=False):
        super().__init__()
        self._init_common(embedding_dim, hidden_dim, output_dim, gpu_num, autoregressive)
        self.save_hyperparameters()
        self.gtrend_encoder = GTrendEmbedder(output_dim, hidden_dim, use_encoder_mask, trend_len, num_trends, gpu_num)
        self.sales_encoder = SalesEncoder(input_dim=1, embedding_dim=hidden_dim)
        self.text_encoder = AttributeEncoder(len(cat_dict) + 1, len(col_dict) + 1, len(fab_dict) + 1, store_num + 1,
                                             embedding_dim, hidden_dim)
        self.image_encoder = ImageEncoder(embedding_dim, hidden_dim)
        self.temporal_encoder = TemporalEmbedder(embedding_dim, hidden_dim)
        self.fusion_network = TARGFusionNetwork(hidden_dim, query_modality=query_modality)
        self.decoder_linear = TimeDistributed(nn.Linear(1, hidden_dim))
        layer = nn.TransformerDecoderLayer(d_model=hidden_dim, nhead=num_heads, dim_feedforward=hidden_dim * 4,
                                           dropout=0.1)
        if autoregressive:
            self.pos_encoder = PositionalEncoding(hidden_dim, max_len=12)
        self.decoder = nn.TransformerDecoder(layer, num_layers)
        self.decoder_fc = make_decoder_fc(hidden_dim, self.output_len, autoregressive)

    def _statics(self, cat, col, fab, store, temporal, images):
        return [self.temporal_encoder(temporal), self.text_encoder(cat, col, fab, store), self.image_encoder(images)]
