"""Forecast (inference) loop of the reference's drivers (forecast_dl.py:123-171): eval + no_grad forward per batch,
eager launches vs graphs.GraphedForecast, full CrossAttnRNN210 incl. the ResNet-101 trunk (bf16 mode).
    python tools/forecast_bench.py [batch ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from visuelle2_multimodal_fusion_b200.graphs import GraphedForecast
    model = bench._build_model("rnn210", "cuda:0", "bf16").eval()
    model.on_validation_epoch_start()
    for B in [int(a) for a in sys.argv[1:]] or [1, 16, 128]:
        d, im = bench._batch("rnn210", B, seed=3)
        inputs = tuple(t.cuda() for t in d) + (im.cuda(),)
        fc = GraphedForecast(model, inputs)
        res = {}
        for name, fn in (("eager", lambda: model(*inputs)), ("graph", lambda: fc(inputs))):
            ts = []
            with torch.no_grad():
                for it in range(23):
                    torch.cuda.synchronize()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    fn()
                    b.record()
                    torch.cuda.synchronize()
                    if it >= 3:
                        ts.append(a.elapsed_time(b))
            ts.sort()
            res[name] = ts[len(ts) // 2]
        print(f"B={B:4d}: eager {res['eager']:.3f} ms ({B / res['eager'] * 1e3:.0f} items/s)   "
              f"graph {res['graph']:.3f} ms ({B / res['graph'] * 1e3:.0f} items/s)")


if __name__ == "__main__":
    main()
