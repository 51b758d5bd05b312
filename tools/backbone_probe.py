"""Not product code: times the (unreplaced) torchvision ResNet-101 trunk fwd+bwd in the bench configuration
to see what the end-to-end step is made of and which library knobs matter.  python tools/backbone_probe.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from visuelle2_multimodal_fusion_b200.models._base import resnet101_trunk

def run(tag, cnn, x, autocast, steps=6):
    def step():
        if autocast:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                f = cnn(x)
        else:
            f = cnn(x)
        f.float().square().mean().backward()
        for p in cnn.parameters():
            p.grad = None
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f"{tag:50s} {e0.elapsed_time(e1) / steps:8.2f} ms/step", flush=True)

import warnings
warnings.simplefilter("ignore")
B = 128
x = torch.randn(B, 3, 299, 299, device="cuda").contiguous(memory_format=torch.channels_last)
torch.manual_seed(0)
cnn = resnet101_trunk().cuda().train().to(memory_format=torch.channels_last)
run("autocast bf16 channels_last (bench config)", cnn, x, True)
torch.backends.cudnn.benchmark = True
run("+ cudnn.benchmark", cnn, x, True)
cnn_b = resnet101_trunk().cuda().train().to(memory_format=torch.channels_last).to(torch.bfloat16)
run("pure bf16 weights, no autocast, cudnn.benchmark", cnn_b, x.bfloat16(), False)
# forward only
with torch.no_grad():
    for _ in range(2):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            cnn(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            cnn(x)
    torch.cuda.synchronize()
    print(f"{'forward only (no_grad) autocast':50s} {(time.perf_counter() - t0) / 5 * 1e3:8.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            f = cnn(x)
        f.float().square().mean().backward()
        for p in cnn.parameters():
            p.grad = None
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=90))
