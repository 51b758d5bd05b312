"""Head-only CrossAttnRNN210 training steps (feature maps in) for ncu / compute-sanitizer runs.
Usage: python tools/head_step.py [--steps N] [--batch B] [--model 210|21|demand]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--eval", action="store_true")
    ap.add_argument("--precision", default="bf16")
    a = ap.parse_args()
    import visuelle2_multimodal_fusion_b200.synth as synth
    import visuelle2_multimodal_fusion_b200.models.modules as mods
    from visuelle2_multimodal_fusion_b200 import _lib
    from visuelle2_multimodal_fusion_b200.models.CrossAttnRNN210 import CrossAttnRNN
    mods.resnet101_trunk = lambda: nn.Identity()
    cat_d, col_d, fab_d = synth.label_dicts()
    torch.manual_seed(21)
    m = CrossAttnRNN(a.dim, a.dim, a.dim, cat_d, col_d, fab_d, synth.STORE_N, 3).cuda()
    m = m.eval() if a.eval else m.train()
    m.precision = a.precision
    data, feat = synth.make_batch(a.batch, out_len=10, seed=21, feat_hw=10)
    data = tuple(t.cuda() for t in data)
    feat = feat.cuda()
    import time
    for i in range(a.steps):
        if i == a.steps - 1:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        torch.manual_seed(i)
        f = feat.clone().requires_grad_(True)
        loss = m.training_step((data, f), i)
        loss.backward()
        for p in m.parameters():
            p.grad = None
    torch.cuda.synchronize()
    print("last step wall ms", 1e3 * (time.perf_counter() - t0))
    print("ok", float(loss.detach()), "launches", _lib.launch_count())


if __name__ == "__main__":
    main()
