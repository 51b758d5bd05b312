"""Headline benchmark: CrossAttnRNN210 (SO-fore2-10) training step, fwd+bwd samples/s on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on host cores

One JSON line on stdout (rank 0).  A step = forward + mse_loss + backward + zero_grad of the
whole model (torchvision ResNet-101 trunk + the fused head) on one synthetic VISUELLE2-shaped
batch of 128 items per GPU (BASELINE.json configs[1]; optimizer excluded, as the metric says).
``value``: inputs resident in HBM.  ``e2e``: the same step through the public LightningModule call
(``training_step``) from pinned HOST buffers, host->device copies and the loss read-back inside
the timed region.  ``roofline``: the fused recurrent-attention kernel timed with CUDA events on
its launching stream in a separate pass (so the event records do not perturb ``value``).
``cpu_baseline`` / ``--impl reference``: the oracle port (oracle/rnn.py, as-written algorithm) +
torchvision trunk on the box's host cores.

The same line carries the other BASELINE.json configs, each measured the same way (value, e2e,
head-only, bounded CPU sample) and under torchrun at N > 1 with the gradient all-reduce inside the step:
``other_configs.{rnn21,demand,gtm,v4}`` (configs[0], [2], [3], [4]), ``fp32_mode`` (the headline
workload in the 1e-5 parity mode) and ``v4_forecast`` (the no-grad forecast loop of
forecast_Gated_v4.py:89-124).  ``--only-headline`` skips them.
"""
import argparse
import gc
import hashlib
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

E = A = H = 512          # train_dl.py:197-199
LI, LT = 100, 52
GTM_E, GTM_H, GTM_OUT = 32, 64, 12      # train_GTM_visuelle2.py:162-176

# name -> (BASELINE.json configs index, description, out_len, synth.make_batch kwargs)
WORKLOADS = {
    "rnn210": (1, "CrossAttnRNN210 SO-fore2-10 train step (BASELINE.json configs[1]): fwd + mse_loss + bwd + zero_grad, "
                  "full model incl. ResNet-101 trunk", 10, dict(out_len=10)),
    "rnn21": (0, "CrossAttnRNN21 SO-fore2-1 train step (BASELINE.json configs[0]): 10 windows per item, fwd + mse_loss + "
                 "bwd + zero_grad, full model incl. ResNet-101 trunk", 1, dict(out_len=1)),
    "demand": (2, "CrossAttnRNNDemand new-product 12-week train step (BASELINE.json configs[2]): fwd + mse_loss + bwd + "
                  "zero_grad, full model incl. ResNet-101 trunk", 12, dict(out_len=10, demand=True)),
    "gtm": (3, "GTM_Visuelle2 demand train step (BASELINE.json configs[3]; E=32, H=64, 4 heads, FFN 2048, 12 weeks): "
               "fwd + mse_loss + bwd + zero_grad, full model incl. ResNet-101 trunk", 12, dict(out_len=10, demand=True)),
    "v4": (4, "Proposed_model_v4 gated fusion demand train step (BASELINE.json configs[4]; E=32, H=64, 12 weeks): "
              "fwd + mse_loss + bwd + zero_grad, full model incl. ResNet-101 trunk", 12, dict(out_len=10, demand=True)),
}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernels timed inside a step)"
    return 1384.8, "fallback (B200_PROFILING.md)"


def _gemm_tc_entry(nsteps):
    """The tcgen05 GEMM launches (csrc/gemm_tc.cu) recorded since the last read: FLOPs summed by the library
    (2 M N K per problem) over the summed launch durations, against the measured dense bf16 throughput."""
    from visuelle2_multimodal_fusion_b200 import _lib
    tot, n, flops = _lib.prof_read_bytes(_lib.K_GEMM_TC)
    if n == 0:
        return None
    peak, src = _tensor_peak()
    ach = flops / (tot * 1e-3) / 1e12
    return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
            "kernel": "gemm_tc_kernel", "launches_per_step": n / nsteps, "avg_launch_us": tot / n * 1e3,
            "gflop_per_step": flops / nsteps / 1e9, "ms_per_step": tot / nsteps, "peak_source": src,
            "note": "bf16 operands on the backbone features, tf32 on fp32 data (half the bf16 rate); M = 128 .. 12800 rows: "
                    "launch- and tail-bound, not tensor-pipe bound (SURVEY 8d expects exactly that)",
            "timing": "CUDA events on the launching stream around each launch, eager head-only pass"}


def _build_model(name, device, precision):
    import visuelle2_multimodal_fusion_b200.synth as synth
    from visuelle2_multimodal_fusion_b200.models import CrossAttnRNN21, CrossAttnRNN210, CrossAttnRNNDemand
    cat_d, col_d, fab_d = synth.label_dicts()
    torch.manual_seed(21)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if name == "rnn210":
            m = CrossAttnRNN210.CrossAttnRNN(A, E, H, cat_d, col_d, fab_d, synth.STORE_N, 3, use_img=True, out_len=10,
                                             use_teacher_forcing=True, teacher_forcing_ratio=0.5)
        elif name == "rnn21":
            m = CrossAttnRNN21.CrossAttnRNN(A, E, H, cat_d, col_d, fab_d, synth.STORE_N, 3)
        elif name == "demand":
            m = CrossAttnRNNDemand.CrossAttnRNN(A, E, 3, H, cat_d, col_d, fab_d, synth.STORE_N, True, True, True, True,
                                                out_len=12, use_teacher_forcing=True)
        elif name in ("gtm", "v4"):
            from visuelle2_multimodal_fusion_b200.models.GTM_Visuelle2 import GTM_Visuelle2
            from visuelle2_multimodal_fusion_b200.models.Proposed_model_v4 import GatedMultimodal_Visuelle2 as V4
            cls = GTM_Visuelle2 if name == "gtm" else V4
            m = cls(GTM_E, GTM_H, GTM_OUT, 4, 1, 1, 1, cat_d, col_d, fab_d, synth.STORE_N, 52, 3, 0, use_encoder_mask=1)
        else:
            raise KeyError(name)
    m = m.to(device).train()
    if hasattr(m, "on_train_epoch_start"):
        m.on_train_epoch_start()
    if precision == "bf16" and device != "cpu":
        m.image_encoder.use_bf16_backbone(True)
        m.precision = "bf16"
    return m


def _batch(name, batch, seed, device=None, pin=False):
    import visuelle2_multimodal_fusion_b200.synth as synth
    data, images = synth.make_batch(batch, seed=seed, **WORKLOADS[name][3])
    if name in ("gtm", "v4"):
        data = (data[0][:, :GTM_OUT].contiguous(),) + data[1:]      # demand tuple: (y[B,12], cat, ...)
    if pin:
        data = tuple(t.pin_memory() for t in data)
        images = images.pin_memory()
    if device is not None:
        data = tuple(t.to(device) for t in data)
        images = images.to(device)
    return data, images


def _nbytes(batch):
    data, images = batch
    return sum(t.numel() * t.element_size() for t in data) + images.numel() * images.element_size()


def _config(name, B, world):
    """The workload, identical for both arms (the reference arm runs `your arm's config`)."""
    idx, desc, out_len, _ = WORKLOADS[name]
    gtm = name in ("gtm", "v4")
    return {"workload": desc, "baseline_config_index": idx, "per_gpu_batch": B, "global_batch": B * world,
            "E": GTM_E if gtm else E, "A": None if gtm else A, "H": GTM_H if gtm else H, "out_len": out_len,
            "windows_per_item": 10 if name == "rnn21" else 1, "image": 299, "parallelism": f"dp{world}",
            "l2": "inputs larger than L2: two alternating batches, 137 MB of images each"}


class _Clocks:
    """SM clock and throttle reasons sampled every 250 ms DURING the timed region through NVML (the
    source nvidia-smi reads; polling nvidia-smi itself at 200 ms perturbed the step time by 2x)."""

    def __init__(self, gpu_index):
        import threading
        self.idx = gpu_index
        self.sm, self.mx, self.reasons = [], None, set()
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, b in bits.items():
                    if r & b:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.25)

    def start(self):
        if self.nv is None:
            return
        import threading
        try:
            self.mx = self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM)
        except Exception:
            self.mx = None
        self._thr = threading.Thread(target=self._poll, daemon=True)
        self._thr.start()

    def stop(self):
        if self._thr is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self._stop.set()
        self._thr.join()
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                "samples": len(sm), "source": "NVML (pynvml), 250 ms"}


# ------------------------------------------------------------------------------------------ CPU arm (oracle port)
def _cpu_reference_step_fn(name, batch_items, threads, forecast=False):
    """The reference's CPU implementation of the path: torchvision trunk + oracle head (as written),
    train mode (dropout on, host teacher-forcing draws), fwd + mse + bwd + zero_grad.  ``forecast``: the eval /
    no_grad forward of the forecast drivers instead."""
    from oracle import gtm as ogtm
    from oracle import rnn as orc
    torch.set_num_threads(threads)
    m = _build_model(name, "cpu", "fp32")
    if forecast:
        m.eval()
    cnn = m.image_encoder.cnn
    P = {k: v for k, v in m.named_parameters() if not k.startswith("image_encoder.cnn")}
    P.update({k: v for k, v in m.named_buffers() if not k.startswith("image_encoder.cnn")})
    data, images = _batch(name, batch_items, seed=21)
    params = [p for p in m.parameters() if p.requires_grad]
    mse = torch.nn.functional.mse_loss
    train = not forecast

    def fwd():
        feat = cnn(images)
        if name == "rnn210":
            X, y, cat, col, fab, store, temporal, gt = data
            out, _ = orc.rnn210_forward(P, X, y, cat, col, fab, store, temporal, gt, feat, out_len=10,
                                        use_teacher_forcing=train, teacher_forcing_ratio=0.5, training=train)
            return mse(y.reshape(out.shape), out)
        if name == "rnn21":
            X, y, cat, col, fab, store, temporal, gt = data
            out, _ = orc.rnn21_forward(P, X, y, cat, col, fab, store, temporal, gt, feat, training=train)
            return mse(y, out)
        if name == "demand":
            ts, cat, col, fab, store, temporal, gt = data
            out, _, _ = orc.demand_forward(P, ts, cat, col, fab, store, temporal, gt, feat, out_len=12,
                                           use_teacher_forcing=train, teacher_forcing_ratio=0.5, training=train)
            return mse(ts, out.squeeze())
        y, cat, col, fab, store, temporal, gt = data
        sales = torch.zeros(y.shape[0], 1, 2)                      # GTM_Visuelle2.py:275
        out, _ = ogtm.gtm_family_forward(name, P, sales, cat, col, fab, store, temporal, gt, feat,
                                         output_len=GTM_OUT, heads=4, training=train)
        return mse(y.reshape(-1), out.reshape(-1))

    def step():
        if forecast:
            with torch.no_grad():
                return float(fwd())
        loss = fwd()
        loss.backward()
        for p in params:
            p.grad = None
        return float(loss.detach())

    return step


def _cpu_sample(name, items, threads, n_it=2, forecast=False):
    step = _cpu_reference_step_fn(name, items, threads, forecast=forecast)
    step()
    t0 = time.perf_counter()
    for _ in range(n_it):
        step()
    dt = (time.perf_counter() - t0) / n_it
    what = "eval no_grad forward (forecast loop)" if forecast else "train step"
    return {"value": items / dt, "unit": "samples/s", "cores": threads, "kind": "port",
            "sample": f"{items} items/step of the same workload ({what}, full model incl. ResNet-101, fp32, oracle "
                      f"port + torchvision trunk), 1 warm-up + {n_it} timed steps"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = args.ref_batch
    step = _cpu_reference_step_fn("rnn210", sample, threads)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    desc = (f"{sample} items/step (the whole {args.batch}-item batch of the config)" if sample == args.batch else
            f"{sample} items/step (of the {args.batch}-item batch)") + \
        f", full model incl. ResNet-101, fp32, {threads} threads, oracle port + torchvision trunk"
    print(json.dumps({
        "impl": "reference", "metric": "train samples/sec (fwd+bwd) CrossAttnRNN210", "value": val,
        "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": _config("rnn210", args.batch, max(1, args.gpus)),
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ------------------------------------------------------------------------------------------ product arm
class _Env:
    pass


class _Runner:
    """One workload on this rank: model, two alternating batches (host pinned + resident), the whole-step CUDA graph
    (gradient all-reduces captured inside it at N > 1) and the timed loops."""

    def __init__(self, env, name, precision, B, use_graph=True, graph_nccl=True):
        from visuelle2_multimodal_fusion_b200.ddp import GradReducer
        self.env, self.name, self.B = env, name, B
        dev = env.dev
        self.model = _build_model(name, dev, precision)
        self.use_graph = use_graph
        self.graph_nccl = use_graph and graph_nccl
        self.reducer = GradReducer(self.model, hooks=(not use_graph) or self.graph_nccl) if env.world > 1 else None
        # two distinct batches per rank, alternated: 137 MB of images each, larger than the 126 MB L2
        self.host = [_batch(name, B, seed=21 + 1000 * env.rank + i, pin=True) for i in range(2)]
        self.resident = [(tuple(t.to(dev) for t in d), im.to(dev)) for d, im in self.host]
        self.h2d = _nbytes(self.host[0])
        self.params = [p for p in self.model.parameters() if p.requires_grad]
        self.step_ms = {}
        self.graphed = None
        if use_graph:
            from visuelle2_multimodal_fusion_b200.graphs import GraphedTrainStep
            torch.manual_seed(99)
            self.graphed = GraphedTrainStep(self.model, self.resident[0],
                                            reducer=self.reducer if self.graph_nccl else None)

    def zero(self):
        for p in self.params:
            p.grad = None

    def step_eager(self, i, batch=None):
        torch.manual_seed(1234 + i)          # same teacher-forcing draws on every rank
        loss = self.model.training_step(batch if batch is not None else self.resident[i & 1], i)
        loss.backward()
        if self.reducer:
            self.reducer.finish()
        self.zero()
        return loss

    def step(self, i, batch=None):
        if self.graphed is None:
            return self.step_eager(i, batch)
        torch.manual_seed(1234 + i)
        loss = self.graphed(batch if batch is not None else self.resident[i & 1])
        if self.reducer and not self.graph_nccl:
            self.reducer.reduce_now()
        return loss

    def _max_ms(self, ms):
        if self.env.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=self.env.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    def timed(self, fn, steps, tag=None):
        env = self.env
        env.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record()
        for i in range(steps):
            fn(i)
            ev[i + 1].record()
            # keep the launching thread at most one step ahead of the GPU (what reading the loss every step does
            # in a trainer): with an unbounded run-ahead the driver's launch queue fills and its back-off when
            # the thread is finally let through showed up as isolated 100-300 ms steps
            ev[i].synchronize()
        env.barrier()
        ms = ev[0].elapsed_time(ev[steps])
        if tag:
            self.step_ms[tag] = [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(steps)]
        return self._max_ms(ms)

    def run_e2e(self, steps, tag=None):
        """End to end through the public pipeline: pinned host batches -> DevicePrefetcher (the copy of batch i+1
        overlaps step i on a side stream) -> training_step -> backward -> loss read-back.  Every timed step's
        host->device copy is issued inside the timed region."""
        from visuelle2_multimodal_fusion_b200.data import DevicePrefetcher
        env, host = self.env, self.host

        class _HostBatches:
            def __iter__(self):
                return ((host[i & 1][0], host[i & 1][1]) for i in range(steps))

            def __len__(self):
                return steps

        env.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record()
        for i, batch in enumerate(DevicePrefetcher(_HostBatches(), env.dev)):
            loss = self.step(i, batch)
            float(loss.detach())                 # device->host read of the step's result
            ev[i + 1].record()
        env.barrier()
        ms = ev[0].elapsed_time(ev[steps])
        if tag:
            self.step_ms[tag] = [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(steps)]
        return self._max_ms(ms)

    def measure(self, steps, warmup):
        """(ms resident, ms e2e) for ``steps`` steps each."""
        for i in range(max(warmup, 3)):
            self.step(i)
        gc.collect()
        ms = self.timed(self.step, steps, "resident")
        gc.collect()
        # warm-up of the end-to-end loop: long enough for the caching allocator to reach the steady state of the
        # prefetch pipeline (three staged image batches alive at once: a 2-step warm-up left the third 137 MB block
        # to be cudaMalloc'ed -- an implicit device synchronisation, one 100+ ms step -- inside the timed region)
        self.run_e2e(max(warmup, 3) + 3)
        gc.collect()
        ms_e2e = self.run_e2e(steps, "e2e")
        gc.collect()
        return ms, ms_e2e

    def to_eager(self):
        """The explanatory passes run eagerly (per-launch CUDA events), on the stream the capture warm-up used: the
        parameters' AccumulateGrad nodes live there, any other stream would add a sync per gradient."""
        if self.graphed is not None:
            torch.cuda.set_stream(self.graphed.side)
            self.zero()
            self.graphed = None

    def head_setup(self):
        """Precomputed feature maps in, trunk replaced by identity: everything libv2f_b200 covers."""
        m = self.model
        feats = []
        bf16 = m.image_encoder.backbone_dtype is not None
        with torch.no_grad():
            for d, im in self.resident:
                f = m.image_encoder.cnn(im.contiguous(memory_format=torch.channels_last) if bf16 else im)
                feats.append(f.float().detach())
        self._cnn, self._bdt = m.image_encoder.cnn, m.image_encoder.backbone_dtype
        m.image_encoder.cnn = torch.nn.Identity()
        m.image_encoder.backbone_dtype = None
        self.feats = feats

    def head_restore(self):
        self.model.image_encoder.cnn = self._cnn
        self.model.image_encoder.backbone_dtype = self._bdt

    def step_head(self, i, lag=0):
        torch.manual_seed(1234 + i)
        d, _ = self.resident[i & 1]
        f = self.feats[i & 1].clone().requires_grad_(True)
        if lag:
            torch.cuda._sleep(lag)
        loss = self.model.training_step((d, f), i)
        if lag:
            torch.cuda._sleep(lag)
        loss.backward()
        if self.reducer:
            self.reducer.finish()
        self.zero()

    def head_ms(self, steps):
        """Head-only step (feature maps in, gradient w.r.t. them out) replayed from a CUDA graph like the full step:
        the eager pass is ~700 launches issued by one host thread, and its figure followed the box's host jitter
        (4.7 ... 60 ms from box to box for the same kernels); the eager median is kept beside it."""
        from visuelle2_multimodal_fusion_b200.graphs import GraphedTrainStep
        for i in range(3):
            self.step_head(i)
        self.timed(self.step_head, steps, "head_only_eager")
        hs = sorted(self.step_ms["head_only_eager"])
        self.head_eager_ms = hs[len(hs) // 2]
        if not self.use_graph:
            self.step_ms["head_only"] = self.step_ms["head_only_eager"]
            return self.head_eager_ms
        torch.manual_seed(98)
        batches = [(d, f) for (d, _), f in zip(self.resident, self.feats)]
        g = GraphedTrainStep(self.model, batches[0], reducer=self.reducer if self.graph_nccl else None, image_grad=True)

        def step(i):
            torch.manual_seed(1234 + i)
            g(batches[i & 1])
            if self.reducer and not self.graph_nccl:
                self.reducer.reduce_now()

        for i in range(3):
            step(i)
        ms = self.timed(step, steps, "head_only")
        del g
        self.zero()
        return ms / steps

    def close(self):
        if self.reducer is not None:
            self.reducer.remove()
        self.graphed = None
        self.model = self.reducer = self.resident = self.host = self.feats = None
        gc.collect()
        torch.cuda.empty_cache()


def _traffic(kernel):
    """DRAM bytes per launch from the committed `ncu --set full` capture -- only while the kernel source it was taken
    from is unchanged (sha of the .cu file recorded beside the number); otherwise null: a profile of other code."""
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.isfile(tpath):
        return None, "no capture committed"
    tj = json.load(open(tpath)).get(kernel)
    if not isinstance(tj, dict):
        return None, "no capture for this kernel"
    src = os.path.join(ROOT, "visuelle2-multimodal-fusion_b200", "csrc", tj.get("source", ""))
    sha = hashlib.sha256(open(src, "rb").read()).hexdigest()[:16] if os.path.isfile(src) else None
    if sha != tj.get("source_sha16"):
        return None, f"stale: {tj.get('profile')} was captured from another revision of {tj.get('source')}"
    return tj.get("dram_bytes_per_launch"), f"{tj.get('profile')} (ncu --set full, one launch; source sha16 {sha} matches)"


def _roofline_passes(r, args, nsteps):
    """Per-launch CUDA events around the attention / decode kernels (head-only eager passes with the host running
    ahead of the GPU), phase stamps of the persistent decoder, and the BatchNorm sweeps of the trunk."""
    from visuelle2_multimodal_fusion_b200 import _lib
    import visuelle2_multimodal_fusion_b200.functional as Fv
    roof = {}
    lag = int(0.03 * 1.9e9)                  # ~30 ms of GPU spin (cycles)
    _lib.prof_enable(True)
    nprof = min(nsteps, 5)
    for i in range(nprof):
        r.step_head(i, lag)
    torch.cuda.synchronize()
    _lib.prof_enable(False)
    phases = None
    if Fv.PERSISTENT_DECODE and hasattr(Fv, "persist_phase_times"):
        _lib.lib().v2f_decode_persist_stamps_enable(1)
        _lib.lib().v2f_decode_team_stamps_enable(1)
        Fv.KEEP_LAST_PERSIST_WS = True
        try:
            r.step_head(0, lag)
            phases = Fv.persist_phase_times()
        finally:
            Fv.KEEP_LAST_PERSIST_WS = False
            _lib.lib().v2f_decode_persist_stamps_enable(0)
            _lib.lib().v2f_decode_team_stamps_enable(0)
    peak, peak_src = _peaks()
    N, T = r.B, 10
    team = bool(phases) and len([k for k in phases if k.startswith("P") and "." not in k]) == 5
    # SURVEY 8(d) per-step byte model: every row streams its image and trend tiles (H and V: 2 x (100 + 52) positions x E)
    # once per step, plus the row's small vectors.  The row-team kernel streams bf16 copies of the tiles (2 B / element);
    # the fp32-tile figure of the survey's formula is reported beside it so that rounds stay comparable.
    tile_bytes_fp32 = N * 4 * (2 * LI + 2 * LT) * E
    tile_bytes = tile_bytes_fp32 // 2 if team else tile_bytes_fp32
    small = N * 4 * (5 * H + LI + LT + 4)
    timing = ("CUDA events on the launching stream around each launch, separate eager pass with the host running ahead "
              "of the GPU (spin kernel before forward / backward)")
    for name, kid, bytes_per_launch in (("decode_persist_fwd_kernel", _lib.K_DECODE_PERSIST_FWD, T * (tile_bytes + small)),
                                        ("decode_persist_bwd_kernel", _lib.K_DECODE_PERSIST_BWD, T * (tile_bytes + small)),
                                        ("attn_fwd_kernel", _lib.K_ATTN_FWD, tile_bytes + small),
                                        ("attn_bwd_kernel", _lib.K_ATTN_BWD, tile_bytes + small),
                                        ("tilegrad_kernel", _lib.K_TILEGRAD, None)):
        tot, n = _lib.prof_read(kid)
        if n == 0:
            continue
        avg_ms = tot / n
        if name == "tilegrad_kernel":
            # reads H once, writes dH and dV once; two launches (img, trend) per backward
            bytes_per_launch = N * 4 * 3 * (LI + LT) * E / 2
        ach = bytes_per_launch / (avg_ms * 1e-3) / 1e9
        traffic, tsrc = _traffic(name)
        roof[name] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                      "traffic": traffic, "traffic_source": tsrc, "kernel": name, "avg_launch_us": avg_ms * 1e3,
                      "launches_per_step": n / nprof, "algorithmic_bytes_per_launch": bytes_per_launch,
                      "peak_source": peak_src, "timing": timing}
    g = _gemm_tc_entry(nprof)
    if g is not None:
        roof["gemm_tc_kernel"] = g
    for kname in ("decode_persist_fwd_kernel", "decode_persist_bwd_kernel"):
        if kname in roof and team:
            rr = roof[kname]
            rr["tile_storage"] = "bf16 copies of the tiles (csrc/decode_team.cu)"
            fp32_bytes = T * (tile_bytes_fp32 + small)
            rr["algorithmic_bytes_per_launch_fp32_tiles"] = fp32_bytes
            rr["frac_fp32_tile_bytes"] = fp32_bytes / (rr["avg_launch_us"] * 1e-6) / 1e9 / peak
            rr["note_bound"] = ("the bf16 tiles of all rows (40 MB) stay resident in the 126 MB L2 for the whole launch: DRAM traffic "
                                "(`traffic`) is a fraction of the algorithmic bytes, so HBM is not what bounds this kernel -- "
                                "the per-step chain of five team barriers, tcgen05 issue (~85 cycles per instruction in the "
                                "issuing thread) and L2 -> SM tile delivery is (profiles/r02_summary.md)")
    if "decode_persist_fwd_kernel" in roof and phases:
        rr = roof["decode_persist_fwd_kernel"]
        rr["phases_us_per_step"] = {k: {"work": round(w, 2), "barrier_wait": round(b, 2)} for k, (w, b) in phases.items()}
        rr["decoder"] = "row-team tcgen05 kernel (csrc/decode_team.cu)" if team else "column-split kernel (csrc/decode_persist.cu)"
        if "P2 attention sweep" in phases:
            w, b = phases["P2 attention sweep"]
            pa = (tile_bytes + small) / ((w + b) * 1e-6) / 1e9
            rr["attention_phase"] = {"algorithmic_bytes_per_step": tile_bytes + small, "us_per_step": round(w + b, 2),
                                     "achieved": pa, "frac": pa / peak, "unit": "GB/s",
                                     "timing": "%globaltimer stamps of CTA 0 around the phase including its closing "
                                               "barrier, mean over the steps of one launch"}
    return roof


def _bn_passes(r, roof):
    """The BatchNorm / add / ReLU sweeps of the trunk (csrc/bn_act.cu): full-model steps, CUDA events around every
    launch, algorithmic bytes summed by the library (they differ per layer)."""
    from visuelle2_multimodal_fusion_b200 import _lib
    if not getattr(r.model.image_encoder, "fused_trunk", False):
        return
    peak, peak_src = _peaks()
    lag = int(0.06 * 1.9e9)
    _lib.prof_enable(True)
    for i in range(2):
        torch.manual_seed(1234 + i)
        torch.cuda._sleep(lag)                    # host runs ahead: launches execute back to back
        loss = r.model.training_step(r.resident[i & 1], i)
        torch.cuda._sleep(lag)
        loss.backward()
        if r.reducer:
            r.reducer.finish()
        r.zero()
    torch.cuda.synchronize()
    _lib.prof_enable(False)
    for name, kid in (("bn_stats_kernel", _lib.K_BN_STATS), ("bn_apply_kernel", _lib.K_BN_APPLY),
                      ("bn_bwd_reduce_kernel", _lib.K_BN_BWD_REDUCE), ("bn_bwd_elemt_kernel", _lib.K_BN_BWD_ELEMT),
                      ("stem_conv_kernel", _lib.K_STEM_CONV)):
        tot, n, nbytes = _lib.prof_read_bytes(kid)
        if n == 0:
            continue
        ach = nbytes / (tot * 1e-3) / 1e9
        traffic, tsrc = _traffic(name) if name == "stem_conv_kernel" else (None, None)
        roof[name] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                      "traffic": traffic, "traffic_source": tsrc, "kernel": name, "avg_launch_us": tot / n * 1e3,
                      "launches_per_step": n / 2, "algorithmic_bytes_per_step": nbytes / 2,
                      "ms_per_step": tot / 2, "peak_source": peak_src,
                      "timing": "CUDA events on the launching stream around each launch, full-model pass"}
    for kid in (_lib.K_ATTN_FWD, _lib.K_ATTN_BWD, _lib.K_TILEGRAD, _lib.K_DECODE_PERSIST_FWD, _lib.K_DECODE_PERSIST_BWD):
        _lib.prof_read(kid)          # drop the head spans of this pass
    _lib.prof_read_bytes(_lib.K_GEMM_TC)


def _forecast_pass(env, name, B, steps, cpu):
    """The no-grad forecast loop of the reference's drivers (forecast_Gated_v4.py:89-124: eval, no_grad, forward per
    batch, predictions moved to the host) through graphs.GraphedForecast; each rank forecasts its own shard."""
    from visuelle2_multimodal_fusion_b200.graphs import GraphedForecast
    dev = env.dev
    m = _build_model(name, dev, "bf16").eval()
    host = [_batch(name, B, seed=77 + 1000 * env.rank + i, pin=True) for i in range(2)]

    def inputs(b, device=None):
        (y, cat, col, fab, store, temporal, gt), im = b
        t = (torch.zeros(y.shape[0], 1, 2).pin_memory() if device is None else torch.zeros(y.shape[0], 1, 2, device=device),
             cat, col, fab, store, temporal, gt, im)
        return t if device is None else tuple(x.to(device) for x in t)

    res = [inputs(b, dev) for b in host]
    hin = [inputs(b) for b in host]
    fc = GraphedForecast(m, res[0])
    for i in range(3):
        fc(res[i & 1])
    out_host = torch.empty(B, GTM_OUT).pin_memory()

    def loop(src, steps, e2e):
        env.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(steps):
            x = src[i & 1]
            if e2e:
                x = tuple(t.to(dev, non_blocking=True) for t in x)
            out, _ = fc(x)
            if e2e:
                out_host.copy_(out, non_blocking=False)          # `.cpu()` of the predictions
        ev1.record()
        env.barrier()
        ms = ev0.elapsed_time(ev1)
        if env.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    loop(hin, 3, True)
    ms = loop(res, steps, False)
    ms_e2e = loop(hin, steps, True)
    tot = B * env.world * steps
    out = {"workload": "Proposed_model_v4 demand forecast loop (forecast_Gated_v4.py:89-124): eval, no_grad, one forward "
                       "per 128-item batch replayed from a CUDA graph (graphs.GraphedForecast), full model incl. ResNet-101",
           "value": tot / (ms * 1e-3), "unit": "samples/s", "ms_per_batch": ms / steps, "steps": steps,
           "e2e": {"value": tot / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_batch": ms_e2e / steps,
                   "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in hin[0]),
                   "d2h_bytes_per_step": out_host.numel() * 4},
           "cpu_baseline": _cpu_sample(name, 16, os.cpu_count() or 1, n_it=2, forecast=True) if cpu else None}
    del fc, m, res
    gc.collect()
    torch.cuda.empty_cache()
    return out


def run_product(args):
    import torch.distributed as dist
    from visuelle2_multimodal_fusion_b200 import _lib
    env = _Env()
    env.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    env.rank = rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    env.dev = dev = f"cuda:{local}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    env.barrier = barrier
    B = args.batch
    if os.environ.get("V2F_NO_PERSISTENT_GRU"):          # A/B knob for experiments
        _lib.lib().v2f_gru_persistent_enable(0)
    use_graph = not args.no_graph
    graph_nccl = not args.no_graph_nccl
    default_stream = torch.cuda.current_stream()
    threads = os.cpu_count() or 1
    want_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline

    # ---- headline: CrossAttnRNN210.  Default: the whole step (forward + loss + backward, and at N > 1 the bucketed
    # gradient all-reduces on a side stream, overlapping the rest of the backward) replayed from ONE CUDA graph
    # (graphs.GraphedTrainStep); gradients stay in p.grad (static storage), so there is no zero_grad.
    r = _Runner(env, "rnn210", args.precision, B, use_graph, graph_nccl)
    for i in range(max(args.warmup, 3)):
        r.step(i)
    # a generation-2 pass of Python's cyclic GC over the (large, import-heavy) heap stalls the launching
    # thread for ~200 ms and shows up as an isolated 3-4x step: park the start-up heap in the permanent
    # generation, as a long-running trainer would; collections happen between the timed regions
    gc.collect()
    gc.freeze()
    gc.disable()
    clocks = _Clocks(local)
    if not os.environ.get("V2F_BENCH_NO_CLOCKS"):        # experiment knob: is the NVML poll what perturbs?
        clocks.start()
    l0 = _lib.launch_count()
    ms = r.timed(r.step, args.steps, "resident")
    launches = _lib.launch_count() - l0
    if r.graphed is not None:
        launches = r.graphed.launches_per_replay * args.steps   # replays do not pass through the host-side counter
    clk = clocks.stop()
    gc.collect()
    r.run_e2e(max(args.warmup, 3) + 3)
    gc.collect()
    ms_e2e = r.run_e2e(args.steps, "e2e")
    gc.collect()
    step_ms = dict(r.step_ms)

    # ---- head-only figure (precomputed feature maps in) and the roofline passes
    r.to_eager()
    r.head_setup()
    head_ms = r.head_ms(args.steps)
    head_eager_ms = r.head_eager_ms
    step_ms["head_only"] = r.step_ms["head_only"]
    step_ms["head_only_eager"] = r.step_ms["head_only_eager"]
    roof = _roofline_passes(r, args, args.steps)
    r.head_restore()
    _bn_passes(r, roof)
    torch.cuda.set_stream(default_stream)
    h2d = r.h2d
    r.close()
    roof_main = next((k for k in ("decode_persist_fwd_kernel", "attn_fwd_kernel") if k in roof), None)

    cpu = _cpu_sample("rnn210", args.ref_batch, threads) if want_cpu else None

    # ---- the other BASELINE.json configs, the fp32 parity mode and the forecast loop
    other, fp32_mode, forecast = {}, None, None
    if not args.only_headline:
        k_other = max(3, min(args.steps, args.other_steps))
        for name in ("rnn21", "demand", "gtm", "v4"):
            ro = _Runner(env, name, args.precision, B, use_graph, graph_nccl)
            m1, m2 = ro.measure(k_other, 3)
            lp = ro.graphed.launches_per_replay if ro.graphed is not None else None
            ro.to_eager()
            ro.head_setup()
            hm = ro.head_ms(k_other)
            _lib.prof_enable(True)                      # two eager head-only steps: the tcgen05 GEMMs of this head
            for i in range(2):
                ro.step_head(i)
            torch.cuda.synchronize()
            _lib.prof_enable(False)
            head_gemm = _gemm_tc_entry(2)
            for kid in range(_lib.K_GEMM_TC):
                _lib.prof_read(kid)
            ro.head_restore()
            torch.cuda.set_stream(default_stream)
            tot = B * world * k_other
            other[name] = {"config": _config(name, B, world), "value": tot / (m1 * 1e-3), "unit": "samples/s",
                           "ms_per_step": m1 / k_other, "steps": k_other, "n_gpus": world, "dtype": args.precision,
                           "e2e": {"value": tot / (m2 * 1e-3), "unit": "samples/s", "ms_per_step": m2 / k_other,
                                   "h2d_bytes_per_step": ro.h2d, "d2h_bytes_per_step": 4},
                           "head_only": {"ms_per_step": hm, "value": B * world / (hm * 1e-3), "unit": "samples/s",
                                         "eager_median_ms": ro.head_eager_ms,
                                         "note": "feature maps in, gradient w.r.t. them out; CUDA-graph replay",
                                         "gemm_tc": head_gemm},
                           "gpu_launches_per_step": lp,
                           "cpu_baseline": _cpu_sample(name, 16, threads, n_it=1) if want_cpu else None}
            ro.close()
        k32 = max(3, min(args.steps, 6))
        tf32_prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False     # true fp32 convolutions
        r32 = _Runner(env, "rnn210", "fp32", B, use_graph, graph_nccl)
        m1, m2 = r32.measure(k32, 3)
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32_prev
        tot = B * world * k32
        fp32_mode = {"config": _config("rnn210", B, world), "value": tot / (m1 * 1e-3), "unit": "samples/s",
                     "ms_per_step": m1 / k32, "steps": k32, "dtype": "fp32",
                     "e2e": {"value": tot / (m2 * 1e-3), "unit": "samples/s", "ms_per_step": m2 / k32,
                             "h2d_bytes_per_step": r32.h2d, "d2h_bytes_per_step": 4},
                     "note": "the headline workload in the 1e-5 parity mode: fp32 torchvision/cuDNN trunk (TF32 off), every "
                             "product of the head on exact fp32 CUDA-core kernels"}
        torch.cuda.set_stream(default_stream)
        r32.close()
        forecast = _forecast_pass(env, "v4", B, k_other, want_cpu)

    if rank == 0:
        total = B * world * args.steps
        out = {
            "metric": "train samples/sec (fwd+bwd) CrossAttnRNN210", "value": total / (ms * 1e-3),
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "fp32", "data": "synthetic",
            "config": _config("rnn210", B, world),
            "execution": {"precision": {"backbone": "bf16 channels_last: torchvision modules, cuDNN convolutions except the stem (conv1 as a "
                                                    "tcgen05 implicit GEMM reading the fp32 NCHW images in place, BatchNorm statistics from its "
                                                    "epilogue: csrc/stem_conv.cu); BatchNorm + residual add + ReLU (+ stem max-pool) as fused HBM "
                                                    "sweeps of libv2f_b200.so (csrc/bn_act.cu)"
                                        if args.precision == "bf16" else "fp32 torchvision/cuDNN, untouched",
                                        "head": ("decode loop and its BPTT as one persistent cooperative launch each with the products on "
                                                 "tcgen05 and bf16 attention tiles (csrc/decode_team.cu); tcgen05 GEMMs: bf16 on backbone "
                                                 "features, tf32 elsewhere; fp32 state, softmax and gates"
                                                 if args.precision == "bf16" else "fp32 CUDA-core kernels") + " (libv2f_b200.so)"},
                          "step": ("whole step (fwd + loss + bwd" + (" + bucketed NCCL gradient all-reduce on a side stream"
                                                                      if (world > 1 and graph_nccl) else "") +
                                   ") replayed from one CUDA graph, graphs.GraphedTrainStep" if use_graph else "eager launches"),
                          "allreduce": (None if world == 1 else
                                        ("captured inside the graph, overlapping the backward" if (use_graph and graph_nccl)
                                         else "after the replay, not overlapped" if use_graph else "autograd hooks, side stream"))},
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "gpu_launches_per_step": launches / args.steps,
            "clocks": clk, "step_ms": step_ms,
            "head_only": {"value": B * world / (head_ms * 1e-3), "unit": "samples/s", "ms_per_step": head_ms,
                          "eager_median_ms": head_eager_ms,
                          "note": "feature maps [B,2048,10,10] in, gradient w.r.t. them out; everything libv2f_b200 covers; "
                                  "CUDA-graph replay like the full step (eager_median_ms: the same step issued eagerly, "
                                  "~700 launches from one host thread -- follows the box's host jitter)"},
            "roofline": roof.get(roof_main),
            "roofline_other": {k: v for k, v in roof.items() if k != roof_main},
            "cpu_baseline": cpu,
            "other_configs": other, "fp32_mode": fp32_mode, "v4_forecast": forecast,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--ref-batch", type=int, default=128,
                    help="items per step of the CPU arm (128 = the config's whole batch)")
    ap.add_argument("--other-steps", type=int, default=10, help="timed steps per secondary config")
    ap.add_argument("--only-headline", action="store_true", help="skip other_configs / fp32_mode / v4_forecast")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph-nccl", action="store_true",
                    help="reduce the gradients after the graph replay instead of inside it (not overlapped)")
    ap.add_argument("--no-graph", action="store_true", help="time the eager loop instead of the CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        run_product(args)


if __name__ == "__main__":
    main()
