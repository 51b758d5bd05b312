"""Not product code: find what stalls isolated steps of the resident loop (CPU wall per phase vs GPU events)."""
import os, sys, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = "cuda:0"
model = bench._build_model("rnn210", dev, "bf16")
res = []
for s in (21, 22):
    d, im = bench._batch("rnn210", 128, seed=s)
    res.append((tuple(t.to(dev) for t in d), im.to(dev)))
params = [p for p in model.parameters() if p.requires_grad]
mode = sys.argv[1] if len(sys.argv) > 1 else "lag1"
N = 40
gc.collect(); gc.freeze(); gc.disable()
def step(i):
    t0 = time.perf_counter()
    torch.manual_seed(1234 + i)
    loss = model.training_step(res[i & 1], i)
    t1 = time.perf_counter()
    loss.backward()
    t2 = time.perf_counter()
    for p in params:
        p.grad = None
    t3 = time.perf_counter()
    return loss, (t1 - t0, t2 - t1, t3 - t2)
for i in range(4):
    step(i)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
cpu = []
ev[0].record()
side = torch.cuda.Stream()
pin = torch.empty(int(os.environ.get("COPY_MB", "0")) << 20 or 1, dtype=torch.uint8).pin_memory()
dst = torch.empty_like(pin, device=dev)
for i in range(N):
    if os.environ.get("COPY_MB"):
        with torch.cuda.stream(side):
            dst.copy_(pin, non_blocking=True)
    loss, ts = step(i)
    ev[i + 1].record()
    t0 = time.perf_counter()
    if mode == "sync":
        float(loss)
    elif mode == "lag1":
        ev[i].synchronize()
    cpu.append(ts + (time.perf_counter() - t0,))
torch.cuda.synchronize()
for i in range(N):
    g = ev[i].elapsed_time(ev[i + 1])
    flag = " <<<" if g > 37 else ""
    print(f"{i:3d} gpu {g:7.2f} ms | cpu fwd {cpu[i][0]*1e3:6.2f} bwd {cpu[i][1]*1e3:6.2f} zero {cpu[i][2]*1e3:5.2f} wait {cpu[i][3]*1e3:6.2f}{flag}")
