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
torchvision trunk on the box's host cores, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

E = A = H = 512          # train_dl.py:197-199
OUT_LEN = 10
LI, LT = 100, 52


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _build_model(device, precision):
    import visuelle2_multimodal_fusion_b200.synth as synth
    from visuelle2_multimodal_fusion_b200.models.CrossAttnRNN210 import CrossAttnRNN
    cat_d, col_d, fab_d = synth.label_dicts()
    torch.manual_seed(21)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = CrossAttnRNN(A, E, H, cat_d, col_d, fab_d, synth.STORE_N, 3, use_img=True, out_len=OUT_LEN,
                         use_teacher_forcing=True, teacher_forcing_ratio=0.5)
    m = m.to(device).train()
    m.on_train_epoch_start()
    if precision == "bf16" and device != "cpu":
        m.image_encoder.use_bf16_backbone(True)
        m.precision = "bf16"
    return m


def _batch(batch, seed, device=None, pin=False):
    import visuelle2_multimodal_fusion_b200.synth as synth
    data, images = synth.make_batch(batch, out_len=OUT_LEN, seed=seed)
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


def _cpu_reference_step_fn(batch_items, threads):
    """The reference's CPU implementation of the path: torchvision trunk + oracle head (as written),
    train mode (dropout on, host teacher-forcing draws), fwd + mse + bwd + zero_grad."""
    from oracle import rnn as orc
    torch.set_num_threads(threads)
    m = _build_model("cpu", "fp32")
    cnn = m.image_encoder.cnn
    P = {k: v for k, v in m.named_parameters() if not k.startswith("image_encoder.cnn")}
    (X, y, cat, col, fab, store, temporal, gt), images = _batch(batch_items, seed=21)
    params = [p for p in m.parameters() if p.requires_grad]

    def step():
        feat = cnn(images)
        out, _ = orc.rnn210_forward(P, X, y, cat, col, fab, store, temporal, gt, feat, out_len=OUT_LEN,
                                    use_teacher_forcing=True, teacher_forcing_ratio=0.5, training=True)
        loss = torch.nn.functional.mse_loss(y.reshape(out.shape), out)
        loss.backward()
        for p in params:
            p.grad = None
        return float(loss.detach())

    return step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = args.ref_batch
    step = _cpu_reference_step_fn(sample, threads)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    desc = f"{sample} items/step (of the 128-item batch), full model incl. ResNet-101, fp32, {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": "train samples/sec (fwd+bwd) CrossAttnRNN210", "value": val,
        "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "CrossAttnRNN210 SO-fore2-10 train step (BASELINE.json configs[1])",
                   "per_gpu_batch": 128, "E": E, "A": A, "H": H, "out_len": OUT_LEN, "image": 299},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def run_product(args):
    import torch.distributed as dist
    from visuelle2_multimodal_fusion_b200 import _lib
    import visuelle2_multimodal_fusion_b200.functional as Fv
    from visuelle2_multimodal_fusion_b200.ddp import GradReducer
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))
    B = args.batch
    if os.environ.get("V2F_NO_PERSISTENT_GRU"):          # A/B knob for experiments
        _lib.lib().v2f_gru_persistent_enable(0)
    model = _build_model(dev, args.precision)
    use_graph = not args.no_graph
    # --graph-nccl: capture the bucketed all-reduces inside the graph (overlap with backward); default: reduce after
    graph_nccl = use_graph and args.graph_nccl
    reducer = GradReducer(model, hooks=(not use_graph) or graph_nccl) if world > 1 else None
    # two distinct batches per rank, alternated: 137 MB of images each, larger than the 126 MB L2
    host = [_batch(B, seed=21 + 1000 * rank + i, pin=True) for i in range(2)]
    resident = [(tuple(t.to(dev) for t in d), im.to(dev)) for d, im in host]
    h2d = _nbytes(host[0])
    params = [p for p in model.parameters() if p.requires_grad]

    def zero():
        for p in params:
            p.grad = None

    def step_resident(i):
        torch.manual_seed(1234 + i)          # same teacher-forcing draws on every rank
        loss = model.training_step(resident[i & 1], i)
        loss.backward()
        if reducer:
            reducer.finish()
        zero()
        return loss

    def step_e2e(i):
        torch.manual_seed(1234 + i)
        d, im = host[i & 1]
        batch = (tuple(t.to(dev, non_blocking=True) for t in d), im.to(dev, non_blocking=True))
        loss = model.training_step(batch, i)
        loss.backward()
        if reducer:
            reducer.finish()
        zero()
        return float(loss.detach())          # device->host read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_ms = {}

    def timed(fn, steps, tag=None):
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record()
        for i in range(steps):
            fn(i)
            ev[i + 1].record()
            # keep the launching thread at most one step ahead of the GPU (what reading the loss every step does
            # in a trainer): with an unbounded run-ahead the driver's launch queue fills and its back-off when
            # the thread is finally let through showed up as isolated 100-300 ms steps
            ev[i].synchronize()
        barrier()
        ms = ev[0].elapsed_time(ev[steps])
        if tag:
            step_ms[tag] = [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(steps)]
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    # ---- default: the whole step (forward + loss + backward) replayed from ONE CUDA graph
    # (graphs.GraphedTrainStep); gradients stay in p.grad (static storage), so there is no zero_grad, and the
    # DDP all-reduce runs right after the replay.  --no-graph times the eager loop instead.
    step_eager = step_resident
    graphed = None
    if use_graph:
        from visuelle2_multimodal_fusion_b200.graphs import GraphedTrainStep
        torch.manual_seed(99)
        graphed = GraphedTrainStep(model, resident[0], reducer=reducer if graph_nccl else None)

        def step_resident(i):                                # noqa: F811
            torch.manual_seed(1234 + i)
            loss = graphed(resident[i & 1])
            if reducer and not graph_nccl:
                reducer.reduce_now()
            return loss

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    # a generation-2 pass of Python's cyclic GC over the (large, import-heavy) heap stalls the launching
    # thread for ~200 ms and shows up as an isolated 3-4x step: park the start-up heap in the permanent
    # generation, as a long-running trainer would
    import gc
    gc.collect()
    gc.freeze()
    gc.disable()          # collections happen between the timed regions (gc.collect() below), not inside them
    clocks = _Clocks(local)
    if not os.environ.get("V2F_BENCH_NO_CLOCKS"):        # experiment knob: is the NVML poll what perturbs?
        clocks.start()
    l0 = _lib.launch_count()
    ms = timed(step_resident, args.steps, "resident")
    launches = _lib.launch_count() - l0
    if graphed is not None:
        launches = graphed.launches_per_replay * args.steps   # replays do not pass through the host-side counter
    clk = clocks.stop()
    gc.collect()
    if graphed is None:
        for i in range(2):
            step_e2e(i)
    # end to end through the public pipeline: pinned host batches -> DevicePrefetcher (the copy of batch i+1
    # overlaps step i on a side stream) -> training_step -> backward -> loss read-back.  Every timed step's
    # host->device copy is issued inside the timed region.
    from visuelle2_multimodal_fusion_b200.data import DevicePrefetcher

    class _HostBatches:
        def __init__(self, n):
            self.n = n

        def __iter__(self):
            return ((host[i & 1][0], host[i & 1][1]) for i in range(self.n))

        def __len__(self):
            return self.n

    def run_e2e(steps, tag=None):
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record()
        for i, batch in enumerate(DevicePrefetcher(_HostBatches(steps), dev)):
            torch.manual_seed(1234 + i)
            if graphed is not None:
                loss = graphed(batch)            # staged device batch -> the graph's input buffers -> replay
                if reducer and not graph_nccl:
                    reducer.reduce_now()
            else:
                loss = model.training_step(batch, i)
                loss.backward()
                if reducer:
                    reducer.finish()
                zero()
            float(loss.detach())                 # device->host read of the step's result
            ev[i + 1].record()
        barrier()
        ms = ev[0].elapsed_time(ev[steps])
        if tag:
            step_ms[tag] = [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(steps)]
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    # warm-up of the end-to-end loop: long enough for the caching allocator to reach the steady state of the
    # prefetch pipeline (three staged image batches alive at once: a 2-step warm-up left the third 137 MB block to be
    # cudaMalloc'ed -- an implicit device synchronisation, one 100+ ms step -- inside the timed region)
    run_e2e(max(args.warmup, 3) + 3)
    gc.collect()
    ms_e2e = run_e2e(args.steps, "e2e")
    gc.collect()

    # ---- head-only figure (precomputed feature maps in), explains the roofline numbers
    head_ms = None
    roof = {}
    default_stream = torch.cuda.current_stream()
    if graphed is not None:
        # the explanatory passes below run eagerly (per-launch CUDA events), on the stream the capture warm-up
        # used: the parameters' AccumulateGrad nodes live there, any other stream would add a sync per gradient
        graphed.release()
        torch.cuda.set_stream(graphed.side)
        for p in params:
            p.grad = None
        step_resident = step_eager
    if rank == 0 or world > 1:
        feats = []
        with torch.no_grad():
            for d, im in resident:
                f = model.image_encoder.cnn(im.contiguous(memory_format=torch.channels_last)
                                            if args.precision == "bf16" else im)
                feats.append(f.float().detach())
        cnn = model.image_encoder.cnn
        model.image_encoder.cnn = torch.nn.Identity()
        saved_dtype = model.image_encoder.backbone_dtype
        model.image_encoder.backbone_dtype = None

        def step_head(i):
            torch.manual_seed(1234 + i)
            d, _ = resident[i & 1]
            f = feats[i & 1].clone().requires_grad_(True)
            loss = model.training_step((d, f), i)
            loss.backward()
            zero()

        for i in range(3):
            step_head(i)
        timed(step_head, args.steps, "head_only")
        hs = sorted(step_ms["head_only"])
        head_ms = hs[len(hs) // 2] * args.steps        # median step: this eager, launch-bound pass is jitter-prone
        # ---- roofline pass: CUDA events around every launch of the attention kernels.  The eager loop is
        # launch-bound (the GPU waits for the host between kernels), so an event pair around one launch would
        # also time the host's issue latency (~5 us on a 16 us kernel).  A spin kernel ahead of the forward and of
        # the backward lets the host run ahead; the bracketed launches then execute back to back from the queue.
        lag = int(0.03 * 1.9e9)                  # ~30 ms of GPU spin (cycles)

        def step_head_prof(i):
            torch.manual_seed(1234 + i)
            d, _ = resident[i & 1]
            f = feats[i & 1].clone().requires_grad_(True)
            torch.cuda._sleep(lag)
            loss = model.training_step((d, f), i)
            torch.cuda._sleep(lag)
            loss.backward()
            zero()

        _lib.prof_enable(True)
        nprof = min(args.steps, 5)
        for i in range(nprof):
            step_head_prof(i)
        torch.cuda.synchronize()
        _lib.prof_enable(False)
        # phase times inside the persistent decoder (a phase of a persistent kernel has no launch to bracket with CUDA
        # events): CTA 0 stamps %globaltimer at its phase boundaries during one extra forward
        phases = None
        if Fv.PERSISTENT_DECODE:
            _lib.lib().v2f_decode_persist_stamps_enable(1)
            Fv.KEEP_LAST_PERSIST_WS = True
            try:
                step_head_prof(0)
                phases = Fv.persist_phase_times()
            finally:
                Fv.KEEP_LAST_PERSIST_WS = False
                _lib.lib().v2f_decode_persist_stamps_enable(0)
        peak, peak_src = _peaks()
        N = B
        tile_bytes = N * 4 * (2 * LI + 2 * LT) * E
        small = N * 4 * (5 * H + LI + LT + 4)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        tj = json.load(open(tpath)) if os.path.isfile(tpath) else {}
        for name, kid, bytes_per_launch in (("decode_persist_fwd_kernel", _lib.K_DECODE_PERSIST_FWD,
                                             OUT_LEN * (tile_bytes + small)),
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
            roof[name] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                          "traffic": tj.get(name), "kernel": name, "avg_launch_us": avg_ms * 1e3,
                          "launches_per_step": n / nprof, "algorithmic_bytes_per_launch": bytes_per_launch,
                          "peak_source": peak_src,
                          "timing": "CUDA events on the launching stream around each launch, separate eager pass with the "
                                    "host running ahead of the GPU (spin kernel before forward / backward)"}
        if "decode_persist_fwd_kernel" in roof and phases:
            r = roof["decode_persist_fwd_kernel"]
            r["note"] = ("one cooperative launch = all %d decode steps: per step six phases (three weight-stationary "
                         "products out of shared memory, the HBM/L2-bound attention sweep, two row-local phases) "
                         "separated by grid barriers; achieved/frac are the by-the-book whole-launch figures "
                         "(algorithmic tile bytes of all steps / launch duration), attention_phase is the sweep alone" % OUT_LEN)
            r["phases_us_per_step"] = {k: {"work": round(w, 2), "barrier_wait": round(b, 2)} for k, (w, b) in phases.items()}
            w, b = phases["P2 attention sweep"]
            pa = (tile_bytes + small) / ((w + b) * 1e-6) / 1e9
            r["attention_phase"] = {"algorithmic_bytes_per_step": tile_bytes + small, "us_per_step": round(w + b, 2),
                                    "achieved": pa, "frac": pa / peak, "unit": "GB/s",
                                    "timing": "%globaltimer stamps of CTA 0 around the phase including its closing grid "
                                              "barrier, mean over the steps of one launch"}
        model.image_encoder.cnn = cnn
        model.image_encoder.backbone_dtype = saved_dtype
        # ---- the BatchNorm / add / ReLU sweeps of the trunk (csrc/bn_act.cu): full-model steps, CUDA events
        # around every launch, algorithmic bytes summed by the library (they differ per layer)
        if getattr(model.image_encoder, "fused_trunk", False):
            _lib.prof_enable(True)
            for i in range(2):
                torch.manual_seed(1234 + i)
                torch.cuda._sleep(2 * lag)                    # host runs ahead: launches execute back to back
                loss = model.training_step(resident[i & 1], i)
                torch.cuda._sleep(2 * lag)
                loss.backward()
                zero()
            torch.cuda.synchronize()
            _lib.prof_enable(False)
            for name, kid in (("bn_stats_kernel", _lib.K_BN_STATS), ("bn_apply_kernel", _lib.K_BN_APPLY),
                              ("bn_bwd_reduce_kernel", _lib.K_BN_BWD_REDUCE),
                              ("bn_bwd_elemt_kernel", _lib.K_BN_BWD_ELEMT)):
                tot, n, nbytes = _lib.prof_read_bytes(kid)
                if n == 0:
                    continue
                ach = nbytes / (tot * 1e-3) / 1e9
                roof[name] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                              "traffic": None, "kernel": name, "avg_launch_us": tot / n * 1e3,
                              "launches_per_step": n / 2, "algorithmic_bytes_per_step": nbytes / 2,
                              "ms_per_step": tot / 2, "peak_source": peak_src,
                              "timing": "CUDA events on the launching stream around each launch, full-model pass"}
            for kid in (_lib.K_ATTN_FWD, _lib.K_ATTN_BWD, _lib.K_TILEGRAD):
                _lib.prof_read(kid)          # drop the head spans of this pass

    torch.cuda.set_stream(default_stream)
    roof_main = "decode_persist_fwd_kernel" if "decode_persist_fwd_kernel" in roof else "attn_fwd_kernel"
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample = args.ref_batch
        step = _cpu_reference_step_fn(sample, threads)
        step()
        t0 = time.perf_counter()
        n_it = 2
        for _ in range(n_it):
            step()
        dt = (time.perf_counter() - t0) / n_it
        cpu = {"value": sample / dt, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": f"{sample} items/step of the same workload (full model incl. ResNet-101, fp32), "
                         f"1 warm-up + {n_it} timed steps"}
    if rank == 0:
        total = B * world * args.steps
        out = {
            "metric": "train samples/sec (fwd+bwd) CrossAttnRNN210", "value": total / (ms * 1e-3),
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "fp32", "data": "synthetic",
            "config": {"workload": "CrossAttnRNN210 SO-fore2-10 train step (BASELINE.json configs[1]): "
                                   "fwd + mse_loss + bwd + zero_grad, full model incl. ResNet-101 trunk",
                       "per_gpu_batch": B, "global_batch": B * world, "E": E, "A": A, "H": H, "out_len": OUT_LEN,
                       "image": 299, "parallelism": f"dp{world}",
                       "precision": {"backbone": "bf16 channels_last: torchvision modules, cuDNN convolutions; BatchNorm + residual add + "
                                                 "ReLU (+ stem max-pool) as fused HBM sweeps of libv2f_b200.so (csrc/bn_act.cu)"
                                     if args.precision == "bf16" else "fp32 torchvision/cuDNN, untouched",
                                     "head": ("tcgen05 GEMMs: bf16 on backbone features, tf32 elsewhere; fp32 state, softmax and gates"
                                              if args.precision == "bf16" else "fp32 CUDA-core kernels") + " (libv2f_b200.so)"},
                       "l2": "inputs larger than L2: two alternating batches, 137 MB of images each",
                       "execution": ("whole step (fwd + loss + bwd) replayed from one CUDA graph, graphs.GraphedTrainStep"
                                     if use_graph else "eager launches")},
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "gpu_launches_per_step": launches / args.steps,
            "clocks": clk, "step_ms": step_ms,
            "head_only": {"value": (B * world * args.steps / (head_ms * 1e-3)) if head_ms else None,
                          "unit": "samples/s", "ms_per_step": head_ms / args.steps if head_ms else None,
                          "note": "feature maps [B,2048,10,10] in; everything libv2f_b200 covers"},
            "roofline": roof.get(roof_main),
            "roofline_other": {k: v for k, v in roof.items() if k != roof_main},
            "cpu_baseline": cpu,
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
    ap.add_argument("--ref-batch", type=int, default=8, help="items per step of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph-nccl", action="store_true", help="capture the gradient all-reduces inside the CUDA graph")
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
