"""Shared host-side pieces of the drop-in modules: Lightning compatibility, optimizer, backbone."""
import warnings

import torch
import torch.nn as nn

try:  # the reference derives from pytorch_lightning.LightningModule (absent in this image)
    import pytorch_lightning as _pl
    LightningBase = _pl.LightningModule
except Exception:  # pragma: no cover - depends on the image
    class LightningBase(nn.Module):
        """Minimal stand-in offering what the reference's hooks touch."""

        def __init__(self):
            super().__init__()
            self.hparams = {}
            self.logged = {}
            self._v2f_optimizers = None

        def save_hyperparameters(self, *args, **kwargs):
            import inspect
            frame = inspect.currentframe().f_back
            names = frame.f_code.co_varnames[1:frame.f_code.co_argcount]
            self.hparams = {k: frame.f_locals[k] for k in names if k in frame.f_locals}

        def log(self, name, value, *args, **kwargs):
            self.logged[name] = value.detach() if torch.is_tensor(value) else value

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def optimizers(self):
            if self._v2f_optimizers is None:
                self._v2f_optimizers = self.configure_optimizers()
            return self._v2f_optimizers[0]


def make_adafactor(params):
    """``Adafactor(scale_parameter=True, relative_step=True, warmup_init=True, lr=None)``
    (models/CrossAttnRNN210.py:229-230) as the multi-tensor CUDA step of this package (optim.Adafactor: same
    algorithm, same state keys as fairseq's / transformers')."""
    from ..optim import Adafactor
    return Adafactor(params, scale_parameter=True, relative_step=True, warmup_init=True, lr=None)


def resnet101_trunk():
    """torchvision ResNet-101 minus avgpool/fc; layer3/layer4 trainable, the rest frozen
    (models/CrossAttnRNN210.py:61-65).  Not replaced by this package."""
    import torchvision.models as tvm
    net = None
    try:
        import os
        from torchvision.models import ResNet101_Weights
        w = ResNet101_Weights.IMAGENET1K_V1
        cached = os.path.join(torch.hub.get_dir(), "checkpoints", os.path.basename(w.url))
        if os.path.isfile(cached):
            net = tvm.resnet101(weights=w)
    except Exception:
        net = None
    if net is None:
        warnings.warn("ImageNet weights for resnet101 are not cached and there is no network: "
                      "backbone starts from random init (load a checkpoint to restore it)")
        net = tvm.resnet101(weights=None)
    cnn = nn.Sequential(*list(net.children())[:-2])
    for p in cnn.parameters():
        p.requires_grad = False
    for c in list(cnn.children())[6:]:
        for p in c.parameters():
            p.requires_grad = True
    return cnn


def wape_mae(gt, pred, abs_den, norm_scalar=53.0):
    mae = torch.nn.functional.l1_loss(gt * norm_scalar, pred * norm_scalar)
    den = torch.sum(torch.abs(gt * norm_scalar)) if abs_den else torch.sum(gt * norm_scalar)
    wape = 100 * torch.sum(torch.abs((gt - pred) * norm_scalar)) / den
    return mae, wape


def current_lr(module):
    lr = module.optimizers().param_groups[0]["lr"]
    if lr is None:
        return 0.0
    return lr.item() if torch.is_tensor(lr) else lr
