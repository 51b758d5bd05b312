"""Not product code: kernel-level time table (torch.profiler, CUDA activities) of the bench step.
python tools/step_profile.py [--rows 45]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

rows = int(sys.argv[sys.argv.index("--rows") + 1]) if "--rows" in sys.argv else 45
dev = "cuda:0"
model = bench._build_model(dev, "bf16")
d, im = bench._batch(128, seed=21)
batch = (tuple(t.to(dev) for t in d), im.to(dev))
params = [p for p in model.parameters() if p.requires_grad]

def step(i):
    torch.manual_seed(1234 + i)
    loss = model.training_step(batch, i)
    loss.backward()
    for p in params:
        p.grad = None

for i in range(3):
    step(i)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
N = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(N):
        step(i)
    torch.cuda.synchronize()
ka = prof.key_averages()
tot = sum(k.self_device_time_total for k in ka)
print(f"total device time per step: {tot / N / 1e3:.2f} ms over {N} steps")
for k in sorted(ka, key=lambda k: -k.self_device_time_total)[:rows]:
    print(f"{k.self_device_time_total / N / 1e3:8.3f} ms {k.count // N:5d}x  {k.key[:110]}")
