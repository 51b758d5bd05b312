"""Throughput of the other BASELINE.json configs (not the headline; bench.py carries that): CrossAttnRNN21 (W=10),
CrossAttnRNNDemand, GTM_Visuelle2 and Proposed_model_v4, full model incl. the ResNet-101 trunk, B=128 per GPU, bf16 mode,
whole step replayed from one CUDA graph, CUDA events, 1 GPU.  One JSON line per config.
    python tools/bench_configs.py [--steps 20] [rnn21 demand gtm v4]"""
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 20
names = [a for a in sys.argv[1:] if a in ("rnn21", "demand", "gtm", "v4", "rnn210")] or ["rnn21", "demand", "gtm", "v4"]
import visuelle2_multimodal_fusion_b200.synth as synth  # noqa: E402
from visuelle2_multimodal_fusion_b200 import _lib  # noqa: E402
from visuelle2_multimodal_fusion_b200.graphs import GraphedTrainStep  # noqa: E402

cat_d, col_d, fab_d = synth.label_dicts()
B, dev = 128, "cuda:0"


def build(name):
    from visuelle2_multimodal_fusion_b200.models import CrossAttnRNN21, CrossAttnRNN210, CrossAttnRNNDemand
    from visuelle2_multimodal_fusion_b200.models.GTM_Visuelle2 import GTM_Visuelle2
    from visuelle2_multimodal_fusion_b200.models.Proposed_model_v4 import GatedMultimodal_Visuelle2 as V4
    torch.manual_seed(21)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if name == "rnn21":
            m, kw = CrossAttnRNN21.CrossAttnRNN(512, 512, 512, cat_d, col_d, fab_d, synth.STORE_N, 3), dict(out_len=1)
        elif name == "rnn210":
            m, kw = CrossAttnRNN210.CrossAttnRNN(512, 512, 512, cat_d, col_d, fab_d, synth.STORE_N, 3), dict(out_len=10)
        elif name == "demand":
            m = CrossAttnRNNDemand.CrossAttnRNN(512, 512, 3, 512, cat_d, col_d, fab_d, synth.STORE_N, True, True, True, True,
                                                out_len=12, use_teacher_forcing=True)
            kw = dict(out_len=10, demand=True)
        else:
            cls = GTM_Visuelle2 if name == "gtm" else V4
            m = cls(32, 64, 12, 4, 1, 1, 1, cat_d, col_d, fab_d, synth.STORE_N, 52, 3, 0, use_encoder_mask=1)
            kw = dict(out_len=10, demand=True)
    m = m.to(dev).train()
    if hasattr(m, "on_train_epoch_start"):
        m.on_train_epoch_start()
    m.image_encoder.use_bf16_backbone(True)
    m.precision = "bf16"
    return m, kw


for name in names:
    m, kw = build(name)
    batches = []
    for s in (21, 22):
        data, im = synth.make_batch(B, seed=s, **kw)
        if name in ("gtm", "v4"):
            data = (data[0][:, :12],) + data[1:]            # demand tuple: (y[B,12], cat, ...)
        batches.append((tuple(t.to(dev) for t in data), im.to(dev)))
    step = GraphedTrainStep(m, batches[0])
    for i in range(3):
        step(batches[i & 1])
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        torch.manual_seed(1234 + i)
        step(batches[i & 1])
        ev[i + 1].record()
        ev[i].synchronize()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[steps]) / steps
    rows = batches[0][0][0].shape[0] * (batches[0][0][0].shape[1] if name == "rnn21" else 1)
    print(json.dumps({"workload": name, "model": type(m).__name__, "per_gpu_batch": B, "decoder_rows": rows,
                      "ms_per_step": round(ms, 3), "samples_per_s": round(B / ms * 1e3, 1),
                      "libv2f_launches_per_step": step.launches_per_replay,
                      "loss": float(step.static_loss), "execution": "CUDA graph replay, bf16 mode, 1 GPU"}), flush=True)
    step.release()
    del step, m, batches
    torch.cuda.empty_cache()
