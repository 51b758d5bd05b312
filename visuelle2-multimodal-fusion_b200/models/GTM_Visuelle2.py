"""Drop-in for the reference's ``models/GTM_Visuelle2.py`` (GTM transformer fusion: image + text +
Google-Trends).  Same constructor, state_dict keys, forward signature and Lightning hooks as
``/root/reference/models/GTM_Visuelle2.py:178-312``; the body runs on libv2f_b200.so."""
import torch.nn as nn

from .. import functional as Fv
from .. import functional_gtm as Fg
from ._gtm import (AttributeEncoder, DummyEmbedder, GTMFamilyBase, GTrendEmbedder, ImageEncoder,
                   PositionalEncoding, SalesEncoder, TimeDistributed, make_decoder_fc)


class GTMFusionNetwork(nn.Module):
    """BatchNorm1d(6E) -> Linear(6E,6E, no bias) -> ReLU -> Dropout -> Linear(6E,H)
    (GTM_Visuelle2.py:151-172)."""

    def __init__(self, embedding_dim, hidden_dim, dropout=0.2):
        super().__init__()
        input_dim = embedding_dim * 6
        self.feature_fusion = nn.Sequential(
            nn.BatchNorm1d(input_dim),
            nn.Linear(input_dim, input_dim, bias=False),
            nn.ReLU(),
            nn.Dropout(dropout),
            nn.Linear(input_dim, hidden_dim))

    def forward(self, img_encoding, text_encoding, dummy_encoding):
        ff = self.feature_fusion
        x = Fg.concat_cols(img_encoding, text_encoding.flatten(1), dummy_encoding)
        x = Fg.batch_norm1d(x, ff[0], self.training)
        x = Fv.linear(x, ff[1].weight, None, act=1)
        x = Fv.dropout(x, ff[3].p, self.training)
        return Fv.linear(x, ff[4].weight, ff[4].bias)


class GTM_Visuelle2(GTMFamilyBase):
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
        self.fusion_network = GTMFusionNetwork(embedding_dim, hidden_dim)
        self.decoder_linear = TimeDistributed(nn.Linear(1, hidden_dim))      # constructed, never called (:199)
        layer = nn.TransformerDecoderLayer(d_model=hidden_dim, nhead=num_heads, dim_feedforward=hidden_dim * 4,
                                           dropout=0.1)
        if autoregressive:
            self.pos_encoder = PositionalEncoding(hidden_dim, max_len=12)
        self.decoder = nn.TransformerDecoder(layer, num_layers)
        self.decoder_fc = make_decoder_fc(hidden_dim, self.output_len, autoregressive)
