"""Whole-step CUDA graph: forward + loss + backward of a drop-in module's ``training_step`` captured once
and replayed per batch.

The step is ~2 900 kernel launches (torchvision/cuDNN trunk + libv2f_b200) whose host-side issue time
(~27 ms) is close to their 33 ms of GPU time, so any hiccup of the launching thread stalls the GPU.
Captured into one graph the host cost per step is a few copies and one ``cudaGraphLaunch``.

What makes the step capturable:
  * every kernel of libv2f_b200 is launched on the caller's current stream with caller-owned buffers
    (include/v2f.h), so torch's capture sees them like its own kernels; the cooperative persistent-GRU
    launch is a capturable kernel node;
  * dropout masks come from torch's CUDA generator, which is graph-safe (offset advanced per replay);
  * the teacher-forcing decision of CrossAttnRNN210/Demand is drawn on the HOST in the reference
    (models/CrossAttnRNN210.py:216-217); here the T bits live in a device word (``tf_mask_dev`` of
    ``v2f_decode_params``) that is refreshed before each replay with the same host draws in the same order.
Gradients are left in ``p.grad`` (static storage): run the optimizer / ``ddp.GradReducer.reduce_now``
after the call.
"""
import contextlib

import torch


class _TfWord:
    """The teacher-forcing bits of one replay: a device word the captured kernels read, refreshed from a small ring
    of pinned host words guarded by events (the host may run several steps ahead of the GPU: a single pinned word
    could be overwritten before its copy has executed, and step i would train with step i+1's bits)."""

    def __init__(self, dev, depth=4):
        self.dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.host = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in range(depth)]
        self.ev = [None] * depth
        self.turn = 0

    def set(self, bits):
        t = self.turn = (self.turn + 1) % len(self.host)
        if self.ev[t] is not None:
            self.ev[t].synchronize()
        self.host[t][0] = int(bits) & 0x7FFFFFFF
        self.dev.copy_(self.host[t], non_blocking=True)
        self.ev[t] = torch.cuda.Event()
        self.ev[t].record()


@contextlib.contextmanager
def _tf_installed(model, word):
    """The module reads ``model._tf_mask_dev`` instead of drawing on the host ONLY while a Graphed* object warms up /
    captures; eager calls (validation_step between graphed epochs, a forecast after training) never see it."""
    if word is None:
        yield
        return
    prev = model.__dict__.get("_tf_mask_dev")
    model._tf_mask_dev = word.dev
    try:
        yield
    finally:
        if prev is None:
            model.__dict__.pop("_tf_mask_dev", None)
        else:
            model._tf_mask_dev = prev


def _tree_map(fn, obj):
    if torch.is_tensor(obj):
        return fn(obj)
    if isinstance(obj, (list, tuple)):
        return type(obj)(_tree_map(fn, o) for o in obj)
    if isinstance(obj, dict):
        return {k: _tree_map(fn, v) for k, v in obj.items()}
    return obj


def _tree_copy(dst, src):
    if torch.is_tensor(dst):
        dst.copy_(src, non_blocking=True)
    elif isinstance(dst, (list, tuple)):
        for d, s in zip(dst, src):
            _tree_copy(d, s)
    elif isinstance(dst, dict):
        for k in dst:
            _tree_copy(dst[k], src[k])


class GraphedTrainStep:
    """``step = GraphedTrainStep(model, example_batch); loss = step(batch)`` -- same arithmetic as
    ``loss = model.training_step(batch, i); loss.backward()`` with gradients in ``p.grad``."""

    def __init__(self, model, example_batch, warmup=3, reducer=None, image_grad=False):
        """``reducer``: a ``ddp.GradReducer`` built with hooks: its bucketed NCCL all-reduces are captured INSIDE
        the graph, on its side stream, so they overlap the rest of the backward exactly as in the eager loop
        (gradients in ``p.grad`` are already averaged when the replay returns).
        ``image_grad``: also produce the gradient w.r.t. the batch's image tensor (``self.static_batch[1].grad``) --
        for callers that feed precomputed feature maps and run the trunk's backward themselves.

        The capture runs on a private stream; the parameters' AccumulateGrad nodes must be created there too, so no
        autograd graph of an earlier eager step may still be alive (``del loss`` before constructing this)."""
        self.model = model
        self.reducer = reducer
        dev = next(model.parameters()).device
        self.device = dev
        self.static_batch = _tree_map(lambda t: t.to(dev, copy=True), example_batch)
        if image_grad:
            self.static_batch[1].requires_grad_(True)
        self.has_tf = hasattr(model, "draw_tf_mask")
        self._tf = _TfWord(dev) if self.has_tf else None      # owned here, not by the module
        self.params = [p for p in model.parameters() if p.requires_grad]
        # autograd graphs of earlier eager steps that are only reachable through reference cycles (custom Function
        # contexts) keep their AccumulateGrad nodes -- bound to the stream of that eager step -- alive until the cyclic
        # collector runs; a capture that meets one of them fails with "legacy stream depends on a capturing stream"
        import gc
        gc.collect()
        side = self.side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), _tf_installed(model, self._tf):
            for i in range(warmup):                       # lazy initialisations happen outside the capture
                self._refresh_tf()
                loss = model.training_step(self.static_batch, i)
                loss.backward()
                if reducer is not None:
                    reducer.finish()
                for p in self.params:
                    p.grad = None
                if image_grad:
                    self.static_batch[1].grad = None
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from . import _lib
        l0 = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        # same stream as the warm-up: no AccumulateGrad stream mismatch
        with _tf_installed(model, self._tf), torch.cuda.graph(self.graph, stream=side):
            self.static_loss = model.training_step(self.static_batch, 0)
            self.static_loss.backward()
            if reducer is not None:
                reducer.finish()
        self.launches_per_replay = _lib.launch_count() - l0   # libv2f_b200 kernels inside one replay

    def _refresh_tf(self):
        if self.has_tf:
            # the reference's host draws of this step, in its order; 0 when teacher forcing is off
            self._tf.set(self.model.draw_tf_mask(True))

    def __call__(self, batch):
        with torch.no_grad():
            _tree_copy(self.static_batch, batch)          # device->device (or pinned host->device) into the graph's inputs
        self._refresh_tf()
        self.graph.replay()
        return self.static_loss

    def close(self):
        """Drop the captured graph (and its memory pool).  With a reducer the graph contains the communicator's NCCL
        kernels: close every graphed step BEFORE ``torch.distributed.destroy_process_group()`` -- destroying the
        communicator while a graph that captured its kernels is alive does not return."""
        torch.cuda.synchronize(self.device)
        self.graph = None
        self.static_loss = None
        for p in self.params:
            p.grad = None

    def release(self):
        """Kept for callers of the first version: the module is never left in graph mode any more (the device word
        is installed only around warm-up and capture), so there is nothing to undo."""


class GraphedForecast:
    """The no-grad forecast loop of the reference's drivers (forecast_dl.py:123-171, forecast_Gated_v4.py:89-124:
    ``model.eval(); with torch.no_grad(): y_hat = model(*inputs)`` per batch) with the forward captured once into a
    CUDA graph and replayed per batch: ``fc = GraphedForecast(model, example_inputs); y_hat = fc(inputs)``.

    ``inputs`` is the positional tuple the module's ``forward`` takes (dataset_fusion.py batch layout: the data
    tuple followed by the images).  Batches must keep the example's shapes (the last, smaller batch of a loader runs
    eagerly through ``model`` -- ``__call__`` falls back to it when shapes differ).  The host random draws the
    reference makes per forward (CrossAttnRNNDemand draws 12 numbers even in eval, models/CrossAttnRNNDemand.py:343-345)
    are still made per call, so the host RNG stream stays the reference's."""

    def __init__(self, model, example_inputs, warmup=2):
        self.model = model.eval()
        dev = next(model.parameters()).device
        self.static_in = _tree_map(lambda t: t.to(dev, copy=True), tuple(example_inputs))
        self.has_tf = hasattr(model, "draw_tf_mask")
        self._tf = _TfWord(dev) if self.has_tf else None
        side = self.side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad(), _tf_installed(model, self._tf):
            for _ in range(warmup):
                self._refresh_tf(self.static_in)
                model(*self.static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), _tf_installed(model, self._tf), torch.cuda.graph(self.graph, stream=side):
            self.static_out = model(*self.static_in)

    def _refresh_tf(self, inputs):
        """Consume the host draws the reference makes in this forward and hand the bits to the captured kernels:
        0 unless the module has teacher forcing on AND targets are among the inputs (a forecast normally has it off;
        CrossAttnRNNDemand draws 12 numbers even then, models/CrossAttnRNNDemand.py:343-345)."""
        if self.has_tf:
            self._tf.set(self.model.draw_tf_mask(self.model.tf_targets_given(inputs)))

    def _same_shapes(self, inputs):
        flat_a, flat_b = [], []
        _tree_map(lambda t: flat_a.append(tuple(t.shape)), self.static_in)
        _tree_map(lambda t: flat_b.append(tuple(t.shape)), tuple(inputs))
        return flat_a == flat_b

    def __call__(self, inputs):
        if not self._same_shapes(inputs):
            with torch.no_grad():
                return self.model(*inputs)
        self._refresh_tf(inputs)
        _tree_copy(self.static_in, tuple(inputs))
        self.graph.replay()
        return self.static_out
