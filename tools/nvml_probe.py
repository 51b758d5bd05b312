import time, torch, pynvml, threading
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
x = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
stop = False
def load():
    while not stop:
        for _ in range(50): (x @ x)
        torch.cuda.synchronize()
th = threading.Thread(target=load); th.start(); time.sleep(0.5)
for name, fn in [("clock_sm", lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                 ("reasons", lambda: pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)),
                 ("power", lambda: pynvml.nvmlDeviceGetPowerUsage(h)),
                 ("max_clock", lambda: pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))]:
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); v = fn(); ts.append((time.perf_counter() - t0) * 1e3); time.sleep(0.05)
    print(name, v, "ms:", [round(t, 2) for t in ts])
stop = True; th.join()
# effect on a launch-heavy loop: many tiny kernels, with and without polling
y = torch.zeros(1024, device="cuda")
def tiny_loop(n=20000):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): y.add_(1.0)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
print("tiny loop no polling ms", round(tiny_loop(), 1), round(tiny_loop(), 1))
for which in ("clock_sm", "reasons"):
    stop = False
    def poll():
        while not stop:
            if which == "clock_sm": pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            else: pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            time.sleep(0.1)
    th = threading.Thread(target=poll); th.start()
    print("tiny loop polling", which, "ms", round(tiny_loop(), 1), round(tiny_loop(), 1))
    stop = True; th.join()
