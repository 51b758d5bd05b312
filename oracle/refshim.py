"""Import the UNMODIFIED reference modules from /root/reference.  Build container only.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  /root/reference does not exist on the GPU
box; everything that needs it is guarded by ``reference_available()``.

Three shims, none of which touches the arithmetic of the hot path:
  1. ``pytorch_lightning`` is absent -> stub whose LightningModule is an nn.Module with
     ``save_hyperparameters`` / ``log`` / ``device`` and a ``seed_everything``.
  2. ``fairseq`` is absent -> ``fairseq.optim.adafactor.Adafactor`` re-exports
     ``transformers.optimization.Adafactor`` (same algorithm and defaults).
  3. ``torchvision.models.resnet101(pretrained=True)`` needs the network -> build with
     ``weights=None``.
Proposed_model.py (v1) and Proposed_model_v2.py put custom layers inside
nn.TransformerDecoder / nn.TransformerEncoder, which torch >= 2.0 rejects; ``torch18_stack``
reproduces the torch-1.8 container loop (layers applied in order, no final norm).
"""
import importlib
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("V2F_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "CrossAttnRNN210.py"))


def _install_stubs():
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            def save_hyperparameters(self, *a, **k):
                pass

            def log(self, *a, **k):
                pass

            @property
            def device(self):
                try:
                    return next(self.parameters()).device
                except StopIteration:
                    return torch.device("cpu")

            def optimizers(self):
                return self._optimizers[0] if hasattr(self, "_optimizers") else None

        def seed_everything(seed):
            torch.manual_seed(seed)
            return seed

        pl.LightningModule = LightningModule
        pl.seed_everything = seed_everything
        sys.modules["pytorch_lightning"] = pl
    if "fairseq" not in sys.modules:
        from transformers.optimization import Adafactor

        fs = types.ModuleType("fairseq")
        fso = types.ModuleType("fairseq.optim")
        fsa = types.ModuleType("fairseq.optim.adafactor")
        fsa.Adafactor = Adafactor
        fs.optim = fso
        fso.adafactor = fsa
        sys.modules["fairseq"] = fs
        sys.modules["fairseq.optim"] = fso
        sys.modules["fairseq.optim.adafactor"] = fsa
    import torchvision.models as tvm

    if not getattr(tvm.resnet101, "_v2f_patched", False):
        orig = tvm.resnet101

        def resnet101(pretrained=False, **kw):
            kw.pop("weights", None)
            return orig(weights=None, **kw)

        resnet101._v2f_patched = True
        tvm.resnet101 = resnet101


def load_reference_module(name):
    """``name`` e.g. 'CrossAttnRNN210' -> the module object ``models.CrossAttnRNN210``."""
    if not reference_available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the product package also has a sub-package called ``models``; make sure the reference's wins
    mod = sys.modules.get("models")
    if mod is not None and REFERENCE_ROOT not in "".join(list(getattr(mod, "__path__", []))):
        del sys.modules["models"]
    return importlib.import_module("models." + name)


class torch18_stack(nn.Module):
    """torch-1.8 style TransformerEncoder/Decoder container: apply the layers in order."""

    def __init__(self, layers, decoder):
        super().__init__()
        self.layers = layers
        self.decoder = decoder

    def forward(self, x, *args, **kw):
        if self.decoder:
            memory = args[0]
            tgt_mask = args[1] if len(args) > 1 else kw.get("tgt_mask")
            for layer in self.layers:
                x = layer(x, memory, tgt_mask=tgt_mask, memory_mask=kw.get("memory_mask"),
                          tgt_key_padding_mask=None, memory_key_padding_mask=None)
            return x
        mask = args[0] if len(args) > 0 else kw.get("mask")
        for layer in self.layers:
            x = layer(x, src_mask=mask)
        return x


class _IdentityBackbone(nn.Module):
    """Stands in for the ResNet so that ``images`` can be a small feature map [B,2048,h,w]."""

    def forward(self, x):
        return x


def strip_backbone(model):
    """Replace ``image_encoder.cnn`` (torchvision, not restated by the oracle) with identity."""
    model.image_encoder.cnn = _IdentityBackbone()
    return model


def zero_dropout(model):
    """train()-mode gradient parity needs every dropout off (SURVEY.md section 8c hygiene)."""
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = 0.0
    return model
