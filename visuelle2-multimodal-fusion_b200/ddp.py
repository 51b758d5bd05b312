"""Batch-sharded data parallelism: one process per GPU, bucketed gradient all-reduce overlapped
with backward.

The reference is single-process / single-GPU (``pl.Trainer(gpus=[n])``, train_dl.py:164-170); this
layer is new.  Every item is independent in forward and backward, so the only
exchange step is the gradient sum: parameters are grouped into ~25 MB buckets in reverse
registration order (roughly the order their gradients become ready: the fused head first, the
ResNet layers after it), each bucket is all-reduced on a side stream as soon as its last
gradient has been accumulated, and ``finish()`` joins the side stream before the optimizer runs.
``trend_linear.weight`` (54.5 MB fp32) exceeds the bucket size and travels alone.
Parameters that receive no gradient (e.g. Demand's ``gate.fc.weight``) are skipped: the set is
structural, hence identical on every rank, and their ``.grad`` stays ``None`` like in the reference.
"""
import torch
import torch.distributed as dist


def shard_batch(batch, rank, world):
    """Contiguous shard of the leading (item) dimension of every tensor of a dataset_fusion.py batch
    ``((t0, t1, ...), images)``; DistributedSampler-style without the sampler."""
    data, images = batch

    def cut(t):
        n = t.shape[0]
        if n % world:
            # a silently dropped remainder would also make the mean of the per-rank mean losses differ from the
            # global mean; the caller pads or drops the last incomplete batch (DataLoader(drop_last=True))
            raise ValueError(f"batch of {n} items does not split evenly over {world} ranks")
        per = n // world
        return t[rank * per:(rank + 1) * per]

    return tuple(cut(t) for t in data), cut(images)


def sync_batchnorm1d(module, group=None):
    """Mark every nn.BatchNorm1d of ``module`` (GTM_Visuelle2 / Proposed_model_v3 / M4FT fusion networks) so that its
    train-mode statistics are those of the GLOBAL batch: (sum, sum of squares) and the two backward sums are all-reduced
    across ``group`` (functional_gtm._BatchNorm1dSync).  With it, N ranks x B/N items reproduce the reference's
    single-process batch of B items; without it each rank normalises with its own shard's statistics (what
    DistributedDataParallel does by default).  The ResNet trunk's BatchNorm2d stays per-rank.  Returns the count."""
    n = 0
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.v2f_sync_group = "world" if group is None else group
            n += 1
    return n


class GradReducer:
    def __init__(self, module, bucket_bytes=25 << 20, group=None, hooks=True, broadcast=True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        params = [p for p in module.parameters() if p.requires_grad]
        self.params = params[::-1]
        self.buckets = []          # list of lists of params
        cur, size = [], 0
        for p in self.params:
            nbytes = p.numel() * p.element_size()
            if cur and size + nbytes > bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += nbytes
        if cur:
            self.buckets.append(cur)
        self.bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self.cuda = any(p.is_cuda for p in params)
        # NCCL averages inside the collective; gloo (CPU tests) only sums
        self.avg_in_collective = self.cuda and dist.is_initialized() and dist.get_backend(group) == "nccl"
        self.stream = torch.cuda.Stream() if self.cuda else None
        self._reset()
        if broadcast and self.world > 1:
            # replicas start from rank 0's parameters AND buffers (BatchNorm running statistics, counters), like
            # DistributedDataParallel does; afterwards the buffers evolve per rank (per-rank batch statistics)
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t.data, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        # hooks=False: no overlap with backward; call reduce_now() after the step (CUDA-graph replay, where the
        # backward is one opaque launch and the hooks would only fire at capture time)
        self.hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params] if hooks else []

    def _reset(self):
        self.pending = [len(b) for b in self.buckets]
        self.launched = [False] * len(self.buckets)
        self.work = []

    def _on_grad(self, p):
        if self.world == 1:
            return
        i = self.bucket_of[id(p)]
        self.pending[i] -= 1
        if self.pending[i] == 0:
            self._launch(i)

    def _launch(self, i):
        self.launched[i] = True
        ps = [p for p in self.buckets[i] if p.grad is not None]
        if not ps:
            return
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream())
            ctx = torch.cuda.stream(self.stream)
        else:
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            flat = torch.cat([p.grad.reshape(-1) for p in ps])
            if self.cuda:
                for p in ps:
                    p.grad.record_stream(self.stream)
            op = dist.ReduceOp.AVG if self.avg_in_collective else dist.ReduceOp.SUM
            h = dist.all_reduce(flat, op=op, group=self.group, async_op=True)
            self.work.append((h, flat, ps))

    def finish(self, copy_back=False):
        """Join: launch buckets that never filled (unused parameters), wait, hand the mean to ``p.grad``.
        ``copy_back``: write the mean into the existing ``p.grad`` storage instead of re-pointing ``p.grad`` at the
        reduced bucket -- needed when that storage is static and owned by someone else (gradients left behind by a
        CUDA-graph replay: the next replay writes there again)."""
        if self.world > 1:
            for i in range(len(self.buckets)):
                if not self.launched[i]:
                    self._launch(i)
            import contextlib
            main = torch.cuda.current_stream() if self.cuda else None
            with (torch.cuda.stream(self.stream) if self.cuda else contextlib.nullcontext()):
                for h, flat, ps in self.work:
                    h.wait()                      # side stream waits for the collective
                    if not self.avg_in_collective:
                        flat.div_(self.world)
                    # the averaged gradients are used where they are: p.grad is re-pointed at its slice of the
                    # reduced bucket (no copy back; inside a captured graph the bucket is static storage of the graph)
                    views, off = [], 0
                    for p in ps:
                        n = p.numel()
                        views.append(flat[off:off + n].view_as(p))
                        off += n
                    if copy_back:
                        torch._foreach_copy_([p.grad for p in ps], views)      # one multi-tensor launch per bucket
                        continue
                    if self.cuda and not torch.cuda.is_current_stream_capturing():
                        flat.record_stream(main)          # allocated on the side stream, consumed on the caller's
                    for p, v in zip(ps, views):
                        p.grad = v
            if self.cuda:
                torch.cuda.current_stream().wait_stream(self.stream)
        self._reset()

    def reduce_now(self):
        """All-reduce every bucket now (gradients already complete), then write the mean back."""
        if self.world > 1:
            for i in range(len(self.buckets)):
                if not self.launched[i]:
                    self._launch(i)
        self.finish(copy_back=True)

    def remove(self):
        for h in self.hooks:
            h.remove()
