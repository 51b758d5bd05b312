"""Host -> device input pipeline piece: overlap the copy of the next batch with the current step.

The reference feeds ``pl.Trainer`` from a ``DataLoader`` (train_dl.py:60-76) whose batches are host
tensors in the layout of dataset_fusion.py:200-203; Lightning moves each batch to the GPU right before
``training_step``.  At B200 step times the 137 MB image tensor of a 128-item batch is ~2.7 ms of PCIe
time per step, so the drop-in trainer loop wraps its loader in ``DevicePrefetcher``: batch i+1 is copied
on a side stream while step i computes.
"""
import torch


def _to_device(obj, device, pin):
    if torch.is_tensor(obj):
        if pin and not obj.is_pinned() and obj.device.type == "cpu":
            obj = obj.pin_memory()
        return obj.to(device, non_blocking=True)
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_device(o, device, pin) for o in obj)
    if isinstance(obj, dict):
        return {k: _to_device(v, device, pin) for k, v in obj.items()}
    return obj


def _record(obj, stream):
    if torch.is_tensor(obj):
        obj.record_stream(stream)
    elif isinstance(obj, (list, tuple)):
        for o in obj:
            _record(o, stream)
    elif isinstance(obj, dict):
        for o in obj.values():
            _record(o, stream)


_COPY_STREAMS = {}


def _copy_stream(device):
    """One copy stream per device for the life of the process: the caching allocator keeps a block pool per
    stream, so a fresh stream per epoch would pay cudaMalloc (and its implicit synchronisation) again."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=key)
    return _COPY_STREAMS[key]


class DevicePrefetcher:
    """Iterate ``loader`` yielding batches already on ``device``; the copy of the next batch runs on a side
    stream concurrently with the consumer's work on the current stream."""

    def __init__(self, loader, device, pin=True):
        self.loader, self.device, self.pin = loader, torch.device(device), pin
        self.copy_stream = _copy_stream(self.device)

    def _stage(self, it):
        try:
            host = next(it)
        except StopIteration:
            return None
        with torch.cuda.stream(self.copy_stream):
            dev = _to_device(host, self.device, self.pin)
        ev = torch.cuda.Event()
        ev.record(self.copy_stream)
        return dev, ev

    def __iter__(self):
        it = iter(self.loader)
        nxt = self._stage(it)
        while nxt is not None:
            batch, ev = nxt
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            _record(batch, cur)               # the caching allocator must not recycle it under the consumer
            nxt = self._stage(it)             # next copy overlaps the consumer's step
            yield batch

    def __len__(self):
        return len(self.loader)


IMAGENET_MEAN = (0.485, 0.456, 0.406)      # dataset_fusion.py:55
IMAGENET_STD = (0.229, 0.224, 0.225)
_norm_consts = {}


def normalize_uint8_images(u8, dtype=torch.bfloat16, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """Device side of the reference's per-item transform (dataset_fusion.py:50-65): ``u8`` is a CUDA uint8 tensor
    [B,H,W,3] of decoded, already resized RGB pixels (what ``np.asarray(PIL image)`` gives); returns
    ``Normalize(mean, std)(ToTensor(img))`` for the whole batch as a [B,3,H,W] tensor in channels_last memory
    format (bf16 for the fused trunk, or fp32), computed by csrc/image_prep.cu.  The batch crosses PCIe as uint8:
    a quarter of the bytes of the fp32 batch the reference's DataLoader produces."""
    from . import _lib
    if not (u8.is_cuda and u8.dtype == torch.uint8 and u8.dim() == 4 and u8.is_contiguous()):
        raise RuntimeError("normalize_uint8_images expects a contiguous CUDA uint8 tensor [B,H,W,C] (no CPU fallback)")
    if dtype not in (torch.bfloat16, torch.float32):
        raise TypeError("dtype must be torch.bfloat16 or torch.float32")
    B, H, W, C = u8.shape
    key = (u8.device, tuple(mean), tuple(std))
    if key not in _norm_consts:
        _norm_consts[key] = (torch.tensor(mean, dtype=torch.float32, device=u8.device),
                             torch.tensor(std, dtype=torch.float32, device=u8.device))
    m, s = _norm_consts[key]
    if len(mean) != C or len(std) != C:
        raise ValueError("mean/std must have one entry per channel")
    out = torch.empty((B, H, W, C), dtype=dtype, device=u8.device)
    _lib.check(_lib.lib().v2f_image_normalize_u8(B * H * W, C, u8.data_ptr(), _lib.ptr(m), _lib.ptr(s),
                                                 1 if dtype == torch.bfloat16 else 0, out.data_ptr(), _lib.stream()),
               "v2f_image_normalize_u8")
    return out.permute(0, 3, 1, 2)          # [B,C,H,W] view = channels_last tensor
