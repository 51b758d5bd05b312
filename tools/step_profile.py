"""Not product code: kernel-level time table (torch.profiler, CUDA activities) of the bench step.
python tools/step_profile.py [--rows 45]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

rows = int(sys.argv[sys.argv.index("--rows") + 1]) if "--rows" in sys.argv else 45
dev = "cuda:0"
model = bench._build_model("rnn210", dev, "bf16")
d, im = bench._batch("rnn210", 128, seed=21)
batch = (tuple(t.to(dev) for t in d), im.to(dev))
params = [p for p in model.parameters() if p.requires_grad]

HEAD = "--head" in sys.argv
if HEAD:
    with torch.no_grad():
        feat = model.image_encoder.cnn(batch[1].contiguous(memory_format=torch.channels_last)).float()
    model.image_encoder.cnn = torch.nn.Identity()
    model.image_encoder.backbone_dtype = None


def step(i):
    torch.manual_seed(1234 + i)
    b = (batch[0], feat.clone().requires_grad_(True)) if HEAD else batch
    loss = model.training_step(b, i)
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
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(5):
    step(i)
e1.record()
torch.cuda.synchronize()
print(f"total device time per step: {tot / N / 1e3:.2f} ms over {N} steps; wall (events, unprofiled) {e0.elapsed_time(e1) / 5:.2f} ms/step; "
      f"kernels per step {sum(k.count for k in ka) // N}")
for k in sorted(ka, key=lambda k: -k.self_device_time_total)[:rows]:
    print(f"{k.self_device_time_total / N / 1e3:8.3f} ms {k.count // N:5d}x  {k.key[:110]}")
