"""Drop-in for the reference's ``models/CrossAttnRNN21.py`` (SO-fore2-1).

``/root/reference/models/CrossAttnRNN21.py:94-248``: the same encoders and the three attentions
executed once on the sales-GRU state, MLP head ``decoder_fc: E -> 1``.  The item encodings are
computed once per item and indexed by ``row // num_windows`` inside the kernel instead of being
materialised ``num_windows`` times (``repeat_interleave``, :166-170)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as Fv
from ._base import LightningBase, current_lr, make_adafactor, wape_mae
from ._crossattn import encode_static, flatten_windows, run_decoder, sales_state
from .modules import AdditiveAttention, AttributeEncoder, ImageEncoder, TemporalFeatureEncoder, TSEmbedder


class CrossAttnRNN(LightningBase):
    def __init__(self, attention_dim, embedding_dim, hidden_dim, cat_dict, col_dict, fab_dict, store_num,
                 num_trends, use_img=True, out_len=1):
        super().__init__()
        self.save_hyperparameters()
        if attention_dim != embedding_dim:
            raise ValueError("attention_dim must equal embedding_dim (as the reference implicitly requires)")
        self.out_len = out_len
        self.hidden_dim = hidden_dim
        self.embedding_dim = embedding_dim
        self.use_img = use_img
        self.image_encoder = ImageEncoder(embedding_dim)
        self.trend_encoder = TSEmbedder(num_trends, embedding_dim)
        self.temp_encoder = TemporalFeatureEncoder(embedding_dim)
        self.attribute_encoder = AttributeEncoder(len(cat_dict) + 1, len(col_dict) + 1, len(fab_dict) + 1,
                                                  store_num + 1, embedding_dim)
        self.sales_encoder_gru = nn.GRU(input_size=1, hidden_size=hidden_dim, batch_first=True)
        self.ts_self_attention = nn.MultiheadAttention(embedding_dim, num_heads=4, dropout=0.1)
        self.ts_attention = AdditiveAttention(embedding_dim, hidden_dim, attention_dim)
        self.trend_linear = nn.Linear(52 * attention_dim, embedding_dim)
        self.img_attention = AdditiveAttention(embedding_dim, hidden_dim, attention_dim)
        self.multimodal_attention = AdditiveAttention(embedding_dim, hidden_dim, attention_dim)
        self.multimodal_embedder = nn.Linear(embedding_dim, embedding_dim)
        self.decoder_fc = nn.Linear(embedding_dim, 1)

    precision = "fp32"     # "bf16": tcgen05 tensor-core GEMMs (2e-2 contract), see functional.set_precision

    def forward(self, X, y, categories, colors, fabrics, stores, temporal_features, gtrends, images):
        with Fv.precision(self.precision):
            return self._forward(X, y, categories, colors, fabrics, stores, temporal_features, gtrends, images)

    def _forward(self, X, y, categories, colors, fabrics, stores, temporal_features, gtrends, images):
        X, y, bs, num_windows = flatten_windows(X, y)
        tiles = encode_static(self, categories, colors, fabrics, stores, temporal_features, gtrends,
                              images, by_proj=False)
        h0 = sales_state(self, X)
        yhat, _, _ = run_decoder(self, Fv.VARIANT_21, num_windows, 1, 0, 0b1111, tiles, h0, None, None,
                                 None, self.decoder_fc)
        return yhat.view(bs, num_windows, 1), None

    def configure_optimizers(self):
        return [make_adafactor(self.parameters())]

    def _step(self, batch):
        (X, y, cat, col, fab, store, temp, gtrend), images = batch
        forecasts, _ = self.forward(X, y, cat, col, fab, store, temp, gtrend, images)
        return y, forecasts

    def training_step(self, batch, batch_idx):
        y, forecasts = self._step(batch)
        loss = F.mse_loss(y, forecasts)
        self.log("train_loss", loss)
        return loss

    def validation_step(self, batch, batch_idx):
        return self._step(batch)

    def validation_epoch_end(self, outputs):
        gt = torch.cat([o[0] for o in outputs])
        pred = torch.cat([o[1] for o in outputs])
        mae, wape = wape_mae(gt, pred, abs_den=True)
        self.log("val_mae", mae)
        self.log("val_wWAPE", wape)
        print(f"Validation MAE: {mae:.4f}, WAPE: {wape:.4f}, LR: {current_lr(self):.8f}")
