"""Time the tcgen05 stem convolution against the library path it replaces (cast + cuDNN conv + statistics sweep)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visuelle2_multimodal_fusion_b200 import _lib, trunk
import torch.nn as nn

def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

N, H, W = 128, 299, 299
conv = nn.Conv2d(3, 64, 7, 2, 3, bias=False).cuda()
pk = trunk._stem_packed_weight(conv)
w16 = conv.weight.detach().bfloat16().contiguous(memory_format=torch.channels_last)
L = _lib.lib()
import ctypes
r, sm, ps = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
L.v2f_stem_conv_occupancy.argtypes = [ctypes.c_int] + [ctypes.POINTER(ctypes.c_int)] * 3
L.v2f_stem_conv_occupancy(W, ctypes.byref(r), ctypes.byref(sm), ctypes.byref(ps))
print("occupancy: regs", r.value, "smem", sm.value, "CTAs/SM", ps.value)
for nchw in (0, 1):
    for bf in (0, 1):
        x = torch.randn(N, 3, H, W, device="cuda")
        if not nchw:
            x = x.contiguous(memory_format=torch.channels_last)
        if bf:
            x = x.bfloat16()
        nblk = L.v2f_stem_conv_blocks(N, H, W, bf, nchw)
        y = torch.empty((N, 64, 150, 150), device="cuda", dtype=torch.bfloat16, memory_format=torch.channels_last)
        part = torch.empty(nblk, 2, 64, device="cuda")
        for use_part in (1, 0):
            us = timeit(lambda: _lib.check(L.v2f_stem_conv_fwd(N, H, W, x.data_ptr(), bf, nchw, pk.data_ptr(), y.data_ptr(),
                                                               part.data_ptr() if use_part else None, _lib.stream()), "stem"))
            print(f"stem_conv nchw={nchw} bf16_in={bf} stats={use_part} grid={nblk}: {us:.1f} us")
x = torch.randn(N, 3, H, W, device="cuda")
xcl = x.contiguous(memory_format=torch.channels_last)
print("transpose NCHW->NHWC fp32: %.1f us" % timeit(lambda: x.contiguous(memory_format=torch.channels_last)))
print("cast fp32->bf16: %.1f us" % timeit(lambda: xcl.bfloat16()))
xb = xcl.bfloat16()
print("cudnn conv bf16 CL: %.1f us" % timeit(lambda: torch.nn.functional.conv2d(xb, w16, None, 2, 3)))
