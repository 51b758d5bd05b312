"""bf16 channels_last execution of the torchvision ResNet-101 trunk with the BatchNorm / residual add /
ReLU between the (cuDNN) convolutions fused into the sweeps of csrc/bn_act.cu.

The trunk's modules, parameters, buffers and state_dict keys are untouched (``image_encoder.cnn.*`` of
/root/reference/models/CrossAttnRNN210.py:58-72); only the order of evaluation of torchvision's
``Bottleneck.forward`` is restated here so that ``bn -> relu`` and ``bn -> (+identity) -> relu`` each
become one statistics pass + one apply pass (and two passes backward) over the activation instead of
torch's separate batch_norm / add / relu kernels, which were 60 % of the end-to-end step.
"""
import os

import torch
import torch.nn as nn
from torch.utils.weak import WeakTensorKeyDictionary

from . import _lib
from ._lib import check, ptr, stream

_CL = torch.channels_last


def _f32(*shape, device):
    return torch.empty(*shape, device=device, dtype=torch.float32)


_pending_counters = None     # inside forward(): num_batches_tracked buffers to bump with ONE multi-tensor add


def _count_batch(bn, use_batch, mom):
    """nn.BatchNorm bookkeeping: num_batches_tracked += 1 in train mode (deferred to one _foreach_add_ per trunk
    forward); momentum=None means a cumulative moving average."""
    if use_batch and bn.running_mean is not None and bn.num_batches_tracked is not None:
        if mom is None or _pending_counters is None:
            bn.num_batches_tracked.add_(1)
            if mom is None:
                mom = 1.0 / float(bn.num_batches_tracked)
        else:
            _pending_counters.append(bn.num_batches_tracked)
    return mom


def _stem_fusable(conv, bn, pool, x):
    ks = pool.kernel_size if isinstance(pool.kernel_size, tuple) else (pool.kernel_size,) * 2
    st = pool.stride if isinstance(pool.stride, tuple) else (pool.stride,) * 2
    pd = pool.padding if isinstance(pool.padding, tuple) else (pool.padding,) * 2
    dl = pool.dilation if isinstance(pool.dilation, tuple) else (pool.dilation,) * 2
    frozen = not (x.requires_grad or bn.weight.requires_grad or bn.bias.requires_grad or
                  any(p.requires_grad for p in conv.parameters()))
    return (frozen or not torch.is_grad_enabled()) and ks == (3, 3) and st == (2, 2) and pd == (1, 1) and \
        dl == (1, 1) and not pool.ceil_mode and not pool.return_indices


STEM_CONV_TC = os.environ.get("V2F_STEM_CONV", "1") != "0"      # A/B switch: 0 keeps conv1 on the library convolution
if os.environ.get("V2F_BN_PDL", "1") == "0":                    # A/B switch: BatchNorm chains without programmatic dependent launch
    _lib.lib().v2f_bn2d_pdl_enable(0)
_stem_pack_cache = WeakTensorKeyDictionary()                       # conv1.weight -> (version, data_ptr, packed bf16 [64,192])


def _stem_packed_weight(conv):
    """conv1.weight [64,3,7,7] as the K-major operand of csrc/stem_conv.cu: wpk[o, kh*24 + kw*3 + c], zero-padded to
    [64,192] bf16.  Cached while the (frozen) weight is unchanged."""
    w = conv.weight
    # like the bf16 weight copies: only a FROZEN weight is cached (optim.Adafactor updates trainable parameters through
    # raw pointers without bumping Tensor._version, so a cached pack of a trainable conv1 would go stale unnoticed)
    frozen = not w.requires_grad
    hit = _stem_pack_cache.get(w) if frozen else None
    if hit is not None and hit[0] == w._version and hit[1] == w.data_ptr():
        return hit[2]
    with torch.no_grad():
        k = w.detach().to(torch.bfloat16).permute(0, 2, 3, 1).reshape(64, 7, 21)      # [o, kh, kw*3 + c]
        pk = torch.zeros(64, 192, device=w.device, dtype=torch.bfloat16)
        pk[:, :168].view(64, 7, 24)[:, :, :21] = k
    if frozen:
        _stem_pack_cache[w] = (w._version, w.data_ptr(), pk)
    return pk


def _stem_conv_blocks(conv, x):
    """(CTAs, nchw flag) of the tcgen05 stem convolution for this input; 0 CTAs when conv1 / the input are not what it
    covers.  The kernel reads the images as they are: fp32 or bf16, channels_last or plain NCHW."""
    if not (STEM_CONV_TC and x.is_cuda and x.dim() == 4 and x.shape[1] == 3 and conv.bias is None):
        return 0, 0
    if (conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding, conv.dilation, conv.groups,
            conv.padding_mode) != (3, 64, (7, 7), (2, 2), (3, 3), (1, 1), 1, "zeros"):
        return 0, 0
    if x.dtype not in (torch.float32, torch.bfloat16):
        return 0, 0
    if x.is_contiguous(memory_format=_CL):
        nchw = 0
    elif x.is_contiguous():
        nchw = 1
    else:
        return 0, 0
    N, _, H, W = x.shape
    return _lib.lib().v2f_stem_conv_blocks(N, H, W, 1 if x.dtype == torch.bfloat16 else 0, nchw), nchw


def stem(conv, bn, pool, x):
    """maxpool(relu(bn(conv(x)))) with BN + ReLU + max-pool in one sweep when no gradient flows through the
    stem (it is frozen in the reference); otherwise BN+ReLU fused and torch's max-pool.  At the reference's image
    size the frozen conv1 itself runs on csrc/stem_conv.cu (tcgen05 implicit GEMM reading the fp32 images directly,
    batch statistics from its epilogue): no image cast, no statistics sweep."""
    bf16_mode = _w16 is not None and conv in _w16
    fusable = _stem_fusable(conv, bn, pool, x)
    nblk, nchw = _stem_conv_blocks(conv, x) if (bf16_mode and fusable) else (0, 0)
    use_batch = bn.training or bn.running_mean is None
    part = None
    if nblk > 0:
        N, _, H, W = x.shape
        OHc, OWc = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        c = torch.empty((N, 64, OHc, OWc), device=x.device, dtype=torch.bfloat16, memory_format=_CL)
        part = _f32(nblk * 2 * 64, device=x.device) if use_batch else None
        xd = x.detach()
        check(_lib.lib().v2f_stem_conv_fwd(N, H, W, xd.data_ptr(), 1 if xd.dtype == torch.bfloat16 else 0, nchw,
                                           _stem_packed_weight(conv).data_ptr(), c.data_ptr(),
                                           ptr(part) if part is not None else None, stream()), "v2f_stem_conv_fwd")
    else:
        if not x.is_contiguous(memory_format=_CL):
            x = x.contiguous(memory_format=_CL)
        c = _conv(conv, x.to(torch.bfloat16) if bf16_mode else x)
        if not fusable or c.dtype != torch.bfloat16:
            return pool(bn_act(c, bn, relu=True))
        c = c.detach()
        if not c.is_contiguous(memory_format=_CL):
            c = c.contiguous(memory_format=_CL)
    N, C, H, W = c.shape
    OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    dev = c.device
    y = torch.empty((N, C, OH, OW), device=dev, dtype=torch.bfloat16, memory_format=_CL)
    mean, rstd, ss = _f32(C, device=dev), _f32(C, device=dev), _f32(2, C, device=dev)
    mom = _count_batch(bn, use_batch, bn.momentum)
    if part is not None:
        check(_lib.lib().v2f_bn2d_relu_maxpool_fwd_parts(N, H, W, C, c.data_ptr(), ptr(bn.weight), ptr(bn.bias),
                                                         ptr(bn.running_mean, allow_none=True),
                                                         ptr(bn.running_var, allow_none=True),
                                                         float(mom if mom is not None else 0.0), float(bn.eps),
                                                         y.data_ptr(), ptr(mean), ptr(rstd), ptr(ss), ptr(part), nblk,
                                                         stream()), "v2f_bn2d_relu_maxpool_fwd_parts")
        return y
    part = _f32(max(_lib.lib().v2f_bn2d_blocks(N * H * W, C), 1) * 2 * C, device=dev)
    check(_lib.lib().v2f_bn2d_relu_maxpool_fwd(N, H, W, C, c.data_ptr(), ptr(bn.weight), ptr(bn.bias),
                                               ptr(bn.running_mean, allow_none=True),
                                               ptr(bn.running_var, allow_none=True), 1 if use_batch else 0,
                                               float(mom if mom is not None else 0.0), float(bn.eps), y.data_ptr(),
                                               ptr(mean), ptr(rstd), ptr(ss), ptr(part), stream()),
          "v2f_bn2d_relu_maxpool_fwd")
    return y


def _bn_forward(ctx, x, res, gamma, beta, bn, relu):
    if x.dtype != torch.bfloat16 or not x.is_cuda:
        raise RuntimeError("fused trunk expects bf16 CUDA activations (no CPU / fp32 fallback here)")
    if not x.is_contiguous(memory_format=_CL):
        x = x.contiguous(memory_format=_CL)
    if res is not None and not res.is_contiguous(memory_format=_CL):
        res = res.contiguous(memory_format=_CL)
    N, C, H, W = x.shape
    R = N * H * W
    dev = x.device
    use_batch = bn.training or bn.running_mean is None
    y = torch.empty_like(x)
    mean, rstd = _f32(C, device=dev), _f32(C, device=dev)
    ss = _f32(2, C, device=dev)
    nblk = _lib.lib().v2f_bn2d_blocks(R, C)
    part = _f32(max(nblk, 1) * 2 * C, device=dev)
    mom = _count_batch(bn, use_batch, bn.momentum)
    check(_lib.lib().v2f_bn2d_act_fwd(R, C, x.data_ptr(), res.data_ptr() if res is not None else None,
                                      ptr(gamma), ptr(beta),
                                      ptr(bn.running_mean, allow_none=True), ptr(bn.running_var, allow_none=True),
                                      1 if use_batch else 0, float(mom if mom is not None else 0.0), float(bn.eps),
                                      1 if relu else 0, y.data_ptr(), ptr(mean), ptr(rstd), ptr(ss), ptr(part),
                                      stream()), "v2f_bn2d_act_fwd")
    ctx.save_for_backward(x, y if relu else None, gamma, mean, rstd)
    ctx.cfg = (R, C, use_batch, relu, res is not None)
    return y


def _as_grad(t):
    if t.dtype != torch.bfloat16:
        t = t.to(torch.bfloat16)
    return t if t.is_contiguous(memory_format=_CL) else t.contiguous(memory_format=_CL)


def _bn_backward(ctx, dy, dy2):
    x, y, gamma, mean, rstd = ctx.saved_tensors
    R, C, use_batch, relu, has_res = ctx.cfg
    if dy is None:                        # only the second consumer produced a gradient
        dy, dy2 = dy2, None
    dy = _as_grad(dy)
    dy2 = _as_grad(dy2) if dy2 is not None else None
    dev = dy.device
    need_res = has_res and ctx.needs_input_grad[1]
    need_dz = dy2 is not None or (need_res and relu)
    dz = torch.empty_like(x) if need_dz else None
    dx = torch.empty_like(x)
    dgamma, dbeta = _f32(C, device=dev), _f32(C, device=dev)
    coef = _f32(3, C, device=dev)
    nblk = _lib.lib().v2f_bn2d_blocks(R, C)
    part = _f32(nblk * 2 * C, device=dev)
    check(_lib.lib().v2f_bn2d_act_bwd(R, C, dy.data_ptr(), dy2.data_ptr() if dy2 is not None else None, x.data_ptr(),
                                      y.data_ptr() if relu else None, ptr(gamma), ptr(mean), ptr(rstd),
                                      1 if use_batch else 0, 1 if relu else 0,
                                      dz.data_ptr() if need_dz else None, dx.data_ptr(), ptr(dgamma), ptr(dbeta),
                                      ptr(coef), ptr(part), stream()), "v2f_bn2d_act_bwd")
    dres = None
    if need_res:
        dres = dz if need_dz else dy
    return (dx if ctx.needs_input_grad[0] else None, dres,
            dgamma if ctx.needs_input_grad[2] else None, dbeta if ctx.needs_input_grad[3] else None, None, None)


class _BnAct(torch.autograd.Function):
    """y = act(BatchNorm2d(x) (+ res)) on bf16 channels_last tensors; ``bn`` is the nn.BatchNorm2d whose
    parameters / running statistics are used (and updated in train mode)."""

    @staticmethod
    def forward(ctx, x, res, gamma, beta, bn, relu):
        return _bn_forward(ctx, x, res, gamma, beta, bn, relu)

    @staticmethod
    def backward(ctx, dy):
        return _bn_backward(ctx, dy, None)


class _BnActFork(torch.autograd.Function):
    """Same, returning the result twice (two tensors on one storage) for an output with two consumers -- the next
    block's conv1 and its residual add.  Autograd then hands the two gradients over separately and the backward
    sweep sums them on the fly, instead of a separate elementwise add over the whole activation."""

    @staticmethod
    def forward(ctx, x, res, gamma, beta, bn, relu):
        y = _bn_forward(ctx, x, res, gamma, beta, bn, relu)
        ctx.set_materialize_grads(False)
        return y, y.detach()

    @staticmethod
    def backward(ctx, dy, dy2):
        if dy is None and dy2 is None:
            return None, None, None, None, None, None
        return _bn_backward(ctx, dy, dy2)


def bn_act(x, bn, relu=True, res=None):
    return _BnAct.apply(x, res, bn.weight, bn.bias, bn, relu)


def bn_act_fork(x, bn, relu=True, res=None):
    """(y, y): one storage, two autograd edges (see _BnActFork)."""
    return _BnActFork.apply(x, res, bn.weight, bn.bias, bn, relu)


# ---------------------------------------------------------------------------------------------- bf16 weights
# Under autocast every convolution casts its fp32 weight to bf16 (one elementwise launch each, 105 per forward) and
# every trainable one casts the bf16 weight gradient back (84 launches per backward): 0.9 ms of a 28.6 ms step in
# ~190 tiny kernels.  Here the trainable weights are cast by ONE multi-tensor copy before the first convolution and
# their gradients come back through ONE multi-tensor copy after the last weight gradient; frozen weights are cast
# once and cached until they change.
class _CastWeights(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *ws):
        ctx.set_materialize_grads(False)
        outs = [torch.empty_like(w, dtype=torch.bfloat16) for w in ws]       # same strides (channels_last stays)
        torch._foreach_copy_(outs, [w.detach() for w in ws])
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        idx = [i for i, g in enumerate(gs) if g is not None]
        outs = [torch.empty_like(gs[i], dtype=torch.float32) for i in idx]
        if idx:
            torch._foreach_copy_(outs, [gs[i] for i in idx])
        res = [None] * len(gs)
        for i, o in zip(idx, outs):
            res[i] = o
        return tuple(res)


BATCHED_WEIGHT_CASTS = os.environ.get("V2F_BATCHED_CASTS", "1") != "0"      # A/B switch
CAST_GROUP_ELEMS = int(os.environ.get("V2F_CAST_GROUP_ELEMS", str(6 << 20)))   # ~25 MB of fp32 gradient per group
# weight Parameter -> (version, data_ptr, bf16 copy); weak keys: an entry dies with its parameter (an id()-keyed dict
# handed the copy of a freed model's weight to the next model whose parameter reused the id)
_frozen_cache = WeakTensorKeyDictionary()


class _LazyCasts(dict):
    """{conv: bf16 weight} whose trainable entries are produced group by group at FIRST USE.  The autograd engine runs
    ready nodes in reverse order of creation: a cast node created before the whole forward has the lowest priority of
    the graph and its backward -- the fp32 weight gradients -- would only run after the last trunk kernel, so every
    gradient bucket of the trunk would be all-reduced in a tail behind the backward (1.3 ms at 2 GPUs).  Created right
    before the group's first convolution, the node's backward runs as soon as the group's earliest layer has produced
    its weight gradient, and ddp.GradReducer's buckets overlap the rest of the backward."""

    def __init__(self, groups):
        super().__init__()
        self.group_of = {c: g for g in groups for c in g}

    def __missing__(self, conv):
        grp = self.group_of.get(conv)
        if grp is None:
            raise KeyError(conv)
        cast = _CastWeights.apply(*[c.weight for c in grp])
        for c, w in zip(grp, cast):
            self[c] = w
        return self[conv]

    def get(self, conv, default=None):
        try:
            return self[conv]
        except KeyError:
            return default

    def __contains__(self, conv):
        return dict.__contains__(self, conv) or conv in self.group_of


def _bf16_weights(convs):
    """{conv: bf16 weight} for the convolutions of one trunk forward (``convs`` in forward order)."""
    # trainable weights are re-cast on EVERY forward, with or without autograd: optim.Adafactor updates them from a
    # CUDA kernel through raw pointers, which does not bump Tensor._version, so a cached copy keyed on the version went
    # stale after the first no_grad pass (validation / forward used the weights of that first pass for ever).  Only
    # frozen weights (requires_grad=False: layer1/layer2 and the stem in the reference) are cached.
    train = [c for c in convs if c.weight.requires_grad]
    out = {}
    if train:
        if torch.is_grad_enabled():
            # one multi-tensor cast per group of consecutive layers (~CAST_GROUP_ELEMS weights), issued lazily (see
            # _LazyCasts): a group's fp32 gradients appear as soon as ITS earliest layer's weight gradient exists
            groups, grp, n = [], [], 0
            for c in train:
                grp.append(c)
                n += c.weight.numel()
                if n >= CAST_GROUP_ELEMS:
                    groups.append(grp)
                    grp, n = [], 0
            if grp:
                groups.append(grp)
            out = _LazyCasts(groups)
        else:
            cast = [torch.empty_like(c.weight, dtype=torch.bfloat16) for c in train]
            torch._foreach_copy_(cast, [c.weight.detach() for c in train])
            for c, w in zip(train, cast):
                out[c] = w
    stale = []
    for c in convs:
        if c in out:
            continue
        hit = _frozen_cache.get(c.weight)
        if hit is not None and hit[0] == c.weight._version and hit[1] == c.weight.data_ptr() and \
                hit[2].shape == c.weight.shape:
            out[c] = hit[2]
        else:
            stale.append(c)
    if stale:
        with torch.no_grad():
            ws = [torch.empty_like(c.weight, dtype=torch.bfloat16) for c in stale]
            torch._foreach_copy_(ws, [c.weight.detach() for c in stale])
        for c, w in zip(stale, ws):
            _frozen_cache[c.weight] = (c.weight._version, c.weight.data_ptr(), w)
            out[c] = w
    return out


_w16 = None             # inside forward(): {conv: bf16 weight}


def _conv(conv, x):
    w = _w16.get(conv) if _w16 is not None else None
    if w is None or conv.bias is not None or conv.padding_mode != "zeros":
        return conv(x)
    return torch.nn.functional.conv2d(x, w, None, conv.stride, conv.padding, conv.dilation, conv.groups)


def _bottleneck(blk, x, x_res=None, fork=False):
    """torchvision.models.resnet.Bottleneck.forward with the fused normalisation sweeps.  ``x`` feeds conv1,
    ``x_res`` (same values; defaults to ``x``) the identity / downsample branch; ``fork``: return the output as a
    pair for the next block."""
    x_res = x if x_res is None else x_res
    out = bn_act(_conv(blk.conv1, x), blk.bn1, relu=True)
    out = bn_act(_conv(blk.conv2, out), blk.bn2, relu=True)
    identity = x_res
    if blk.downsample is not None:
        identity = bn_act(_conv(blk.downsample[0], x_res), blk.downsample[1], relu=False)
    if fork:
        return bn_act_fork(_conv(blk.conv3, out), blk.bn3, relu=True, res=identity)
    return bn_act(_conv(blk.conv3, out), blk.bn3, relu=True, res=identity)


def supported(cnn):
    """True when ``cnn`` is the torchvision ResNet trunk this module knows how to walk."""
    try:
        from torchvision.models.resnet import Bottleneck
    except Exception:
        return False
    mods = list(cnn.children())
    if len(mods) < 5 or not isinstance(mods[0], nn.Conv2d) or not isinstance(mods[1], nn.BatchNorm2d):
        return False
    if not isinstance(mods[2], nn.ReLU) or not isinstance(mods[3], nn.MaxPool2d):
        return False
    for layer in mods[4:]:
        if not isinstance(layer, nn.Sequential) or not all(isinstance(b, Bottleneck) for b in layer):
            return False
        for b in layer:
            if b.downsample is not None and not (len(b.downsample) == 2 and isinstance(b.downsample[0], nn.Conv2d)
                                                 and isinstance(b.downsample[1], nn.BatchNorm2d)):
                return False
    return True


def forward(cnn, images):
    """images [B,3,H,W] fp32 (any memory format) -> feature map [B,2048,h,w] bf16 channels_last."""
    global _pending_counters, _w16
    mods = list(cnn.children())
    x = images                   # the stem reads the images as collated (NCHW or channels_last, fp32 or bf16)
    _pending_counters = []
    blocks = [blk for layer in mods[4:] for blk in layer]
    try:
        convs = [mods[0]] + [m for blk in blocks for m in blk.modules() if isinstance(m, nn.Conv2d)]
        _w16 = _bf16_weights(convs) if BATCHED_WEIGHT_CASTS else None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            x = stem(mods[0], mods[1], mods[3], x)
            x_res = None
            for i, blk in enumerate(blocks):
                if i + 1 < len(blocks):
                    x, x_res = _bottleneck(blk, x, x_res, fork=True)     # two consumers downstream
                else:
                    x = _bottleneck(blk, x, x_res)
        if _pending_counters:
            torch._foreach_add_(_pending_counters, 1)
    finally:
        _pending_counters = None
        _w16 = None
    return x
