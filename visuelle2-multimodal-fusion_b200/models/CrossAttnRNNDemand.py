"""Drop-in for the reference's ``models/CrossAttnRNNDemand.py`` (new-product 12-week demand).

``/root/reference/models/CrossAttnRNNDemand.py:184-435``.  Differences from the 210 copy that are
kept: attention returns ``alpha * h_j`` (projected) so contexts live in attention space; all four
date features go through ``day_embedding``; zero initial state and input; the host RNG is
consumed once per step even in eval; ``gate`` (GatingFCN) exists but is unused; forward returns
``(outputs[B,T,1], img_alphas, multimodal_alphas)``."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as Fv
from ._base import LightningBase, make_adafactor
from ._crossattn import encode_static, run_decoder
from .modules import AdditiveAttention, AttributeEncoder, ImageEncoder, TemporalFeatureEncoder, TSEmbedder


class GatingFCN(nn.Module):
    """Constructed by the reference (:167-180, :234) but never called; kept for the state_dict."""

    def __init__(self, input_dim):
        super().__init__()
        self.fc = nn.Linear(input_dim, input_dim, bias=False)
        self.gate_fn = nn.Sigmoid()


class CrossAttnRNN(LightningBase):
    def __init__(self, attention_dim, embedding_dim, num_trends, hidden_dim, cat_dict, col_dict, fab_dict,
                 store_num, use_img, use_att, use_date, use_trends, out_len=12, use_teacher_forcing=False,
                 teacher_forcing_ratio=0.5):
        super().__init__()
        if attention_dim != embedding_dim:
            raise ValueError("attention_dim must equal embedding_dim (as the reference implicitly requires)")
        self.teacher_forcing_ratio = teacher_forcing_ratio
        self.use_teacher_forcing = use_teacher_forcing
        self.out_len = out_len
        self.hidden_dim = hidden_dim
        self.embedding_dim = embedding_dim
        self.use_img = use_img
        self.use_att = use_att
        self.use_date = use_date
        self.use_trends = use_trends
        self.trend_encoder = TSEmbedder(num_trends, embedding_dim)
        self.temp_encoder = TemporalFeatureEncoder(embedding_dim, day_only=True)
        self.image_encoder = ImageEncoder(embedding_dim=embedding_dim)
        self.attribute_encoder = AttributeEncoder(len(cat_dict) + 1, len(col_dict) + 1, len(fab_dict) + 1,
                                                  store_num + 1, embedding_dim)
        self.ts_self_attention = nn.MultiheadAttention(embedding_dim, num_heads=4, dropout=0.1)
        self.ts_attention = AdditiveAttention(embedding_dim, hidden_dim, attention_dim, True)
        self.trend_linear = nn.Linear(52 * attention_dim, embedding_dim)
        self.img_attention = AdditiveAttention(embedding_dim, hidden_dim, attention_dim, True)
        self.multimodal_attention = AdditiveAttention(embedding_dim, hidden_dim, attention_dim, True)
        self.multimodal_embedder = nn.Linear(embedding_dim, embedding_dim)
        self.gate = GatingFCN(attention_dim)
        self.decoder = nn.GRU(input_size=embedding_dim + 1, hidden_size=hidden_dim, num_layers=1,
                              batch_first=True)
        self.decoder_fc = nn.Linear(hidden_dim, 1)
        self.save_hyperparameters()

    precision = "fp32"     # "bf16": tcgen05 tensor-core GEMMs (2e-2 contract), see functional.set_precision
    # int32[1] CUDA tensor, installed by graphs.Graphed* ONLY around their warm-up / capture; None in every eager call
    _tf_mask_dev = None

    @staticmethod
    def tf_targets_given(inputs):
        """``ts`` of a positional ``forward`` input tuple."""
        return inputs[0] is not None

    def draw_tf_mask(self, has_y=True):
        """One host draw per step, always -- even in eval (reference :343-345); the bits are only honoured when
        teacher forcing is on."""
        draws = 0
        for t in range(self.out_len):
            if bool(torch.rand(1) < self.teacher_forcing_ratio):
                draws |= 1 << t
        return draws if (self.use_teacher_forcing and has_y) else 0

    def forward(self, ts, categories, colors, fabrics, stores, temporal_features, gtrends, images):
        with Fv.precision(self.precision):
            return self._forward(ts, categories, colors, fabrics, stores, temporal_features, gtrends, images)

    def _forward(self, ts, categories, colors, fabrics, stores, temporal_features, gtrends, images):
        bs = ts.shape[0]
        tiles = encode_static(self, categories, colors, fabrics, stores, temporal_features, gtrends,
                              images, by_proj=True, use_trends=bool(self.use_trends))
        T = self.out_len
        tf_mask = self._tf_mask_dev if self._tf_mask_dev is not None else self.draw_tf_mask(ts is not None)
        mod_mask = 1 | (2 if self.use_img else 0) | (4 if self.use_att else 0) | (8 if self.use_trends else 0)
        h0 = ts.new_zeros(bs, self.hidden_dim, dtype=torch.float32)
        x0 = ts.new_zeros(bs, dtype=torch.float32)
        y = ts.float().contiguous() if ts is not None else None
        yhat, a_img, a_mm = run_decoder(self, Fv.VARIANT_DEMAND, 1, T, tf_mask, mod_mask, tiles, h0, x0, y,
                                        self.decoder, self.decoder_fc)
        n_mod = bin(mod_mask).count("1")
        img_alphas = [a_img[t] for t in range(T)] if self.use_img else []
        keep = [k for k in range(4) if (mod_mask >> k) & 1]
        mm_alphas = [a_mm[t][:, keep] if n_mod < 4 else a_mm[t] for t in range(T)]
        return yhat.unsqueeze(-1), img_alphas, mm_alphas

    def configure_optimizers(self):
        return [make_adafactor(self.parameters())]

    def on_train_epoch_start(self):
        self.use_teacher_forcing = True

    def on_validation_epoch_start(self):
        self.use_teacher_forcing = False

    def _step(self, batch):
        (ts, categories, colors, fabrics, stores, temporal_features, gtrends), images = batch
        forecasted_sales, _, _ = self.forward(ts, categories, colors, fabrics, stores, temporal_features,
                                              gtrends, images)
        return ts, forecasted_sales

    def training_step(self, train_batch, batch_idx):
        ts, forecasted_sales = self._step(train_batch)
        loss = F.mse_loss(ts, forecasted_sales.squeeze())
        self.log("train_loss", loss)
        return loss

    def validation_step(self, test_batch, batch_idx):
        return self._step(test_batch)

    def validation_epoch_end(self, val_step_outputs):
        item_sales = torch.vstack([o[0] for o in val_step_outputs]).squeeze()
        forecasted_sales = torch.vstack([o[1] for o in val_step_outputs]).squeeze()
        loss = F.mse_loss(item_sales, forecasted_sales)
        mae = F.l1_loss(item_sales * 53, forecasted_sales * 53)
        wape = 100 * torch.sum(torch.abs(item_sales * 53 - forecasted_sales * 53)) / torch.sum(item_sales * 53)
        self.log("val_mae", mae)
        self.log("val_wWAPE", wape)
        self.log("val_loss", loss)
        print("Validation MAE:", mae.detach().cpu().numpy(), "Validation WAPE:", wape.detach().cpu().numpy(),
              "LR:", self.optimizers().param_groups[0]["lr"])
