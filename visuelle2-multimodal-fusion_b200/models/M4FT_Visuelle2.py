"""Drop-in for the reference's ``models/M4FT_Visuelle2.py`` (summation-based hierarchical fusion of the temporal,
text and vision embeddings through three FusionBlocks; same encoders / decoder as Proposed_model_v3).
Surface: ``/root/reference/models/M4FT_Visuelle2.py:161-349``."""
import torch.nn as nn

from .. import functional_gtm as Fg
from ._gtm import GTMFamilyBase, GTrendEmbedder, PositionalEncoding, SalesEncoder, TimeDistributed, make_decoder_fc
from .Proposed_model_v3 import AttributeEncoder, FusionBlock, ImageEncoder, TemporalEmbedder


class M4FTFusionNetwork(nn.Module):
    """fusion_final(F_tt(temp+text) + F_tv(text+vis) + temp + text + vis)  (M4FT_Visuelle2.py:175-202)."""

    def __init__(self, hidden_dim, dropout=0.2):
        super().__init__()
        self.fusion_temp_text = FusionBlock(hidden_dim, dropout)
        self.fusion_text_vis = FusionBlock(hidden_dim, dropout)
        self.fusion_final = FusionBlock(hidden_dim, dropout)

    def forward(self, e_temp, e_text, e_vis):
        out_tt = self.fusion_temp_text(Fg.add(e_temp, e_text))
        out_tv = self.fusion_text_vis(Fg.add(e_text, e_vis))
        total = Fg.add(Fg.add(Fg.add(Fg.add(out_tt, out_tv), e_temp), e_text), e_vis)
        return self.fusion_final(total)


class M4FT_Visuelle2(GTMFamilyBase):
    def __init__(self, embedding_dim, hidden_dim, output_dim, num_heads, num_layers, use_text, use_img,
                 cat_dict, col_dict, fab_dict, store_num, trend_len, num_trends, gpu_num, use_encoder_mask=1,
                 autoregressive=False):
        super().__init__()
        self._init_common(embedding_dim, hidden_dim, output_dim, gpu_num, autoregressive)
        self.save_hyperparameters()
        self.gtrend_encoder = GTrendEmbedder(output_dim, hidden_dim, use_encoder_mask, trend_len, num_trends, gpu_num)
        self.sales_encoder = SalesEncoder(input_dim=1, embedding_dim=hidden_dim)
        self.text_encoder = AttributeEncoder(len(cat_dict) + 1, len(col_dict) + 1, len(fab_dict) + 1, store_num + 1,
                                             embedding_dim, hidden_dim)
        self.image_encoder = ImageEncoder(embedding_dim, hidden_dim)
        self.temporal_encoder = TemporalEmbedder(embedding_dim, hidden_dim)
        self.fusion_network = M4FTFusionNetwork(hidden_dim)
        self.decoder_linear = TimeDistributed(nn.Linear(1, hidden_dim))
        layer = nn.TransformerDecoderLayer(d_model=hidden_dim, nhead=num_heads, dim_feedforward=hidden_dim * 4,
                                           dropout=0.1)
        if autoregressive:
            self.pos_encoder = PositionalEncoding(hidden_dim, max_len=12)
        self.decoder = nn.TransformerDecoder(layer, num_layers)
        self.decoder_fc = make_decoder_fc(hidden_dim, self.output_len, autoregressive)

    def _statics(self, cat, col, fab, store, temporal, images):
        return [self.temporal_encoder(temporal), self.text_encoder(cat, col, fab, store), self.image_encoder(images)]
