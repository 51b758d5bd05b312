"""GTM-family fixtures from the UNMODIFIED reference (build container only).  TEST INFRASTRUCTURE ONLY.

Same blob format as make_golden.py.  Two modes per model: ``eval`` (dropout off, BatchNorm running
stats) and ``train_nodrop`` (train() with every dropout p = 0: BatchNorm batch statistics, and the
reference's quirk that decoder_fc's Dropout scales the forecast is then the identity).
Proposed_model (v1) / Proposed_model_v2 need the torch-1.8 container loop (refshim.torch18_stack).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import refshim

MODELS = {"gtm": ("GTM_Visuelle2", "GTM_Visuelle2"), "v1": ("Proposed_model", "GatedMultimodal_Visuelle2"),
          "v2": ("Proposed_model_v2", "GatedMultimodal_Visuelle2"), "v3": ("Proposed_model_v3", "TARG_M4FT_Visuelle2"),
          "v4": ("Proposed_model_v4", "GatedMultimodal_Visuelle2"), "m4ft": ("M4FT_Visuelle2", "M4FT_Visuelle2")}


def build_reference(variant, E, H, out_len, heads, autoregressive, query_modality="image"):
    import visuelle2_multimodal_fusion_b200.synth as synth
    modname, clsname = MODELS[variant]
    mod = refshim.load_reference_module(modname)
    cat_d, col_d, fab_d = synth.label_dicts()
    kw = dict(query_modality=query_modality) if variant == "v3" else {}
    ctor = getattr(mod, clsname)
    if variant in ("v1", "v2"):
        # nn.TransformerDecoder/Encoder(custom_layer) breaks on torch >= 2: construct with stock containers
        # patched to the torch-1.8 loop
        orig_dec, orig_enc = nn.TransformerDecoder, nn.TransformerEncoder

        class Dec(nn.Module):
            def __init__(self, layer, n):
                super().__init__()
                import copy
                self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(n)])

            def forward(self, tgt, memory, tgt_mask=None):
                for l in self.layers:
                    tgt = l(tgt, memory, tgt_mask=tgt_mask)
                return tgt

        class Enc(nn.Module):
            def __init__(self, layer, num_layers):
                super().__init__()
                import copy
                self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(num_layers)])

            def forward(self, src, mask=None):
                for l in self.layers:
                    src = l(src, src_mask=mask)
                return src

        nn.TransformerDecoder = Dec
        if variant == "v2":
            nn.TransformerEncoder = Enc
        try:
            m = ctor(E, H, out_len, heads, 1, 1, 1, cat_d, col_d, fab_d, synth.STORE_N, 52, 3, 0,
                     use_encoder_mask=1, autoregressive=autoregressive)
        finally:
            nn.TransformerDecoder, nn.TransformerEncoder = orig_dec, orig_enc
    else:
        m = ctor(E, H, out_len, heads, 1, 1, 1, cat_d, col_d, fab_d, synth.STORE_N, 52, 3, 0,
                 use_encoder_mask=1, autoregressive=autoregressive, **kw)
    return refshim.strip_backbone(m)


def case_gtm(name, variant, B, E, H, out_len, heads, hw, mode, seed, demand, autoregressive=False,
             query_modality="image"):
    import visuelle2_multimodal_fusion_b200.synth as synth
    torch.manual_seed(seed)
    m = build_reference(variant, E, H, out_len, heads, autoregressive, query_modality)
    # make BatchNorm running stats non-trivial so that eval-mode parity means something
    for mod in m.modules():
        if isinstance(mod, nn.BatchNorm1d):
            mod.running_mean.uniform_(-0.2, 0.2)
            mod.running_var.uniform_(0.5, 1.5)
    if mode == "eval":
        m.eval()
    else:
        refshim.zero_dropout(m).train()
    data, feat = synth.make_batch(B, out_len=out_len if not demand else 10, demand=demand, seed=seed, feat_hw=hw)
    feat.requires_grad_(True)
    if demand:
        ts, cat, col, fab, store, temporal, gt = data
        y = ts[:, :out_len].contiguous()
        item_sales = torch.zeros(B, 1, 2)                    # GTM_Visuelle2.py:273-275
    else:
        item_sales, y, cat, col, fab, store, temporal, gt = data
    state = {k: v.detach().clone() for k, v in m.state_dict().items()}   # before BN momentum update
    out, _ = m(item_sales, cat, col, fab, store, temporal, gt, feat)
    loss = F.mse_loss(y.reshape(-1), out.reshape(-1))
    loss.backward()
    grads = {k: (p.grad.clone() if p.grad is not None else None) for k, p in m.named_parameters()}
    return dict(model="GTM:" + variant, cfg=dict(E=E, H=H, T=out_len, B=B, heads=heads, mode=mode, seed=seed,
                                                 demand=demand, autoregressive=autoregressive,
                                                 query_modality=query_modality),
                state=state, inputs=dict(item_sales=item_sales, y=y, cat=cat, col=col, fab=fab, store=store,
                                         temporal=temporal, gtrends=gt, feat=feat.detach()),
                tf_mask=None, out=out.detach(), loss=loss.detach(), grads=grads, grad_feat=feat.grad.clone())


def _mk(variant, **kw):
    base = dict(variant=variant, B=4, E=8, H=16, out_len=12, heads=4, hw=2, mode="eval", seed=31, demand=True)
    base.update(kw)
    return (case_gtm, base)


CASES = {
    "gtm_demand_eval": _mk("gtm"),
    "gtm_demand_train": _mk("gtm", mode="train_nodrop", seed=32),
    "gtm_sofore1_train": _mk("gtm", mode="train_nodrop", seed=33, demand=False, out_len=1, B=3),
    "gtm_ar_eval": _mk("gtm", seed=34, autoregressive=True),
    "v4_demand_train": _mk("v4", mode="train_nodrop", seed=35, E=16, H=32),
    "v4_sofore10_eval": _mk("v4", seed=36, demand=False, out_len=10),
    "v3_demand_train": _mk("v3", mode="train_nodrop", seed=37, query_modality="text"),
    "v1_demand_train": _mk("v1", mode="train_nodrop", seed=39),
    "v2_demand_train": _mk("v2", mode="train_nodrop", seed=40),
    "m4ft_demand_train": _mk("m4ft", mode="train_nodrop", seed=41),
    "m4ft_sofore10_eval": _mk("m4ft", seed=42, demand=False, out_len=10),
}
