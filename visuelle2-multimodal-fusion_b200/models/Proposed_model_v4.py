"""Drop-in for the reference's ``models/Proposed_model_v4.py`` (standard GTM encoder/decoder +
text-guided gated fusion).  Surface: ``/root/reference/models/Proposed_model_v4.py:152-340``."""
import torch.nn as nn

from .. import functional as Fv
from .. import functional_gtm as Fg
from ._gtm import (AttributeEncoder, DummyEmbedder, GTMFamilyBase, GTrendEmbedder, ImageEncoder,
                   PositionalEncoding, SalesEncoder, TimeDistributed, make_decoder_fc)


class TextGuidedFusionNetwork(nn.Module):
    """Text anchors two sigmoid gates (image, temporal); gated = x + x*gate; Linear -> LayerNorm -> ReLU ->
    Dropout (Proposed_model_v4.py:152-198)."""

    def __init__(self, embedding_dim, hidden_dim, dropout=0.2):
        super().__init__()
        self.img_dim = embedding_dim
        self.text_dim = embedding_dim * 4
        self.dummy_dim = embedding_dim
        self.img_gate_fc = nn.Linear(self.text_dim + self.img_dim, self.img_dim)
        self.dummy_gate_fc = nn.Linear(self.text_dim + self.dummy_dim, self.dummy_dim)
        nn.init.constant_(self.img_gate_fc.bias, 0.0)
        nn.init.constant_(self.dummy_gate_fc.bias, 0.0)
        total_dim = self.img_dim + self.text_dim + self.dummy_dim
        self.fusion_fc = nn.Sequential(nn.Linear(total_dim, hidden_dim), nn.LayerNorm(hidden_dim), nn.ReLU(),
                                       nn.Dropout(dropout))

    def forward(self, img_encoding, text_encoding, dummy_encoding):
        t = text_encoding.flatten(1)
        g_i = Fv.linear(Fg.concat_cols(t, img_encoding), self.img_gate_fc.weight, self.img_gate_fc.bias)
        g_d = Fv.linear(Fg.concat_cols(t, dummy_encoding), self.dummy_gate_fc.weight, self.dummy_gate_fc.bias)
        x = Fg.concat_cols(Fg.gate(img_encoding, g_i, residual=True), t, Fg.gate(dummy_encoding, g_d, residual=True))
        fc, ln = self.fusion_fc[0], self.fusion_fc[1]
        x = Fg.add_layer_norm(Fv.linear(x, fc.weight, fc.bias), None, None, ln.weight, ln.bias, ln.eps)
        return Fv.dropout(Fg.relu(x), self.fusion_fc[3].p, self.training)


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
        self.fusion_network = TextGuidedFusionNetwork(embedding_dim, hidden_dim, dropout=0.1)
        self.decoder_linear = TimeDistributed(nn.Linear(1, hidden_dim))
        layer = nn.TransformerDecoderLayer(d_model=hidden_dim, nhead=num_heads, dim_feedforward=hidden_dim * 4,
                                           dropout=0.1)
        self.decoder = nn.TransformerDecoder(layer, num_layers)
        if autoregressive:
            self.pos_encoder = PositionalEncoding(hidden_dim, max_len=12)
        self.decoder_fc = make_decoder_fc(hidden_dim, self.output_len, autoregressive)
