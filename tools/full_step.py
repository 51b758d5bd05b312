"""Full CrossAttnRNN210 training steps exactly as bench.py runs them (ResNet-101 trunk + head, B=128, bf16 mode),
for ncu launch lists / full captures.  Usage: python tools/full_step.py [--steps N]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 2
dev = "cuda:0"
model = bench._build_model("rnn210", dev, "bf16")
d, im = bench._batch("rnn210", 128, seed=21)
batch = (tuple(t.to(dev) for t in d), im.to(dev))
params = [p for p in model.parameters() if p.requires_grad]
for i in range(steps):
    torch.manual_seed(1234 + i)
    loss = model.training_step(batch, i)
    loss.backward()
    for p in params:
        p.grad = None
torch.cuda.synchronize()
print("loss", float(loss))
