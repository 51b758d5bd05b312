"""Optimizer step on the full CrossAttnRNN210 parameter set (ResNet-101 layer3/4 + head, 61.9 M trainable scalars):
the multi-tensor CUDA Adafactor of this package (csrc/adafactor.cu) vs transformers.optimization.Adafactor on the
same GPU tensors.  CUDA events, median of 10 steps, launch counts.
    python tools/adafactor_bench.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from transformers.optimization import Adafactor as Ref
    from visuelle2_multimodal_fusion_b200 import _lib
    from visuelle2_multimodal_fusion_b200.optim import Adafactor
    model = bench._build_model("rnn210", "cuda:0", "bf16")
    params = [p for p in model.parameters() if p.requires_grad]
    g = torch.Generator(device="cuda").manual_seed(0)
    print(f"{len(params)} trainable tensors, {sum(p.numel() for p in params) / 1e6:.1f} M scalars")
    kw = dict(scale_parameter=True, relative_step=True, warmup_init=True, lr=None)
    for name, cls in (("fused (libv2f_b200)", Adafactor), ("transformers (torch ops)", Ref)):
        opt = cls(params, **kw)
        ts = []
        for it in range(13):
            for p in params:
                p.grad = torch.randn(p.shape, device=p.device, generator=g) * 1e-3
            torch.cuda.synchronize()
            n0 = _lib.launch_count()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            opt.step()
            b.record()
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(a.elapsed_time(b))
            opt.zero_grad(set_to_none=True)
        ts.sort()
        print(f"{name}: median {ts[len(ts) // 2]:.3f} ms / step, min {ts[0]:.3f} ms, "
              f"libv2f launches in the last step: {_lib.launch_count() - n0}")
        del opt


if __name__ == "__main__":
    main()
