"""A complete training loop on synthetic VISUELLE2-shaped data, the way a user of the drop-in would run it without
Lightning (train_dl.py:164-186 hands the same calls to pl.Trainer): pinned host batches -> data.DevicePrefetcher ->
graphs.GraphedTrainStep (forward + loss + backward replayed from one CUDA graph) -> optimizer step with the module's
own configure_optimizers() (optim.Adafactor, one multi-tensor CUDA step) -> loss read-back.  Prints the loss curve and
the throughput INCLUDING the optimizer (bench.py's metric excludes it, as BASELINE.json's does).
    python tools/train_synthetic.py [--steps 40] [--batch 128]"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--batch", type=int, default=128)
    args = ap.parse_args()
    import bench
    from visuelle2_multimodal_fusion_b200.data import DevicePrefetcher
    from visuelle2_multimodal_fusion_b200.graphs import GraphedTrainStep
    dev = "cuda:0"
    model = bench._build_model("rnn210", dev, "bf16")
    model.train()
    model.on_train_epoch_start()
    opt = model.configure_optimizers()[0]
    host = [bench._batch("rnn210", args.batch, seed=100 + i, pin=True) for i in range(4)]
    example = (tuple(t.to(dev) for t in host[0][0]), host[0][1].to(dev))
    step = GraphedTrainStep(model, example)

    class Loader:
        def __iter__(self):
            return ((host[i % 4][0], host[i % 4][1]) for i in range(args.steps))

        def __len__(self):
            return args.steps

    losses, t_all = [], []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i, batch in enumerate(DevicePrefetcher(Loader(), dev)):
        torch.manual_seed(1000 + i)
        loss = step(batch)
        opt.step()                             # no zero_grad: every replay overwrites the static gradient storage
        losses.append(float(loss.detach()))    # device -> host read-back, as Lightning's progress bar does
        t_all.append(time.perf_counter())
    torch.cuda.synchronize()
    warm = min(5, args.steps // 2)
    dt = (t_all[-1] - t_all[warm - 1]) / (args.steps - warm)
    print("loss:", " ".join(f"{v:.4f}" for v in losses[:3]), "...", " ".join(f"{v:.4f}" for v in losses[-3:]))
    print(f"{args.batch / dt:.0f} samples/s including the optimizer step ({1e3 * dt:.2f} ms per step, "
          f"{args.steps - warm} steps after {warm} warm-up steps, 1 GPU)")


if __name__ == "__main__":
    main()
