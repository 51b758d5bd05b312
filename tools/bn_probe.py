"""Per-shape timing of the fused BatchNorm sweeps (csrc/bn_act.cu) at the ResNet-101 activation shapes of a 128-image
batch: CUDA events, L2-cold (rotating over > 126 MB of buffers) and L2-warm (same buffer, as right after the
producing convolution).    python tools/bn_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from visuelle2_multimodal_fusion_b200 import _lib
    lib = _lib.lib()
    shapes = [("stem", 128 * 150 * 150, 64), ("l1.c64", 128 * 75 * 75, 64), ("l1.c256", 128 * 75 * 75, 256),
              ("l2.c128", 128 * 38 * 38, 128), ("l2.c512", 128 * 38 * 38, 512), ("l3.c256", 128 * 19 * 19, 256),
              ("l3.c1024", 128 * 19 * 19, 1024), ("l4.c512", 128 * 10 * 10, 512), ("l4.c2048", 128 * 10 * 10, 2048)]
    peak = 6545.9
    st = torch.cuda.current_stream().cuda_stream
    for name, R, C in shapes:
        nbytes = R * C * 2
        nbuf = max(2, int(300e6 // nbytes) + 1)
        xs = [torch.randn(R, C, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
        y = torch.empty_like(xs[0])
        g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
        mean, rstd, ss = torch.empty(C, device="cuda"), torch.empty(C, device="cuda"), torch.empty(2, C, device="cuda")
        part = torch.empty(lib.v2f_bn2d_blocks(R, C) * 2 * C, device="cuda")
        res = {}
        for mode in ("cold", "warm"):
            _lib.prof_enable(True)
            for i in range(12):
                x = xs[i % nbuf] if mode == "cold" else xs[0]
                _lib.check(lib.v2f_bn2d_act_fwd(R, C, x.data_ptr(), None, _lib.ptr(g), _lib.ptr(b), _lib.ptr(rm),
                                                _lib.ptr(rv), 1, 0.1, 1e-5, 1, y.data_ptr(), _lib.ptr(mean),
                                                _lib.ptr(rstd), _lib.ptr(ss), _lib.ptr(part), st), "fwd")
            torch.cuda.synchronize()
            _lib.prof_enable(False)
            for kname, kid, passes in (("stats", _lib.K_BN_STATS, 1), ("apply", _lib.K_BN_APPLY, 2)):
                ms, n = _lib.prof_read(kid)
                us = ms / n * 1e3
                res[(mode, kname)] = (us, nbytes * passes / (us * 1e-6) / 1e9 / peak)
        print(f"{name:9s} R={R:8d} C={C:5d} {nbytes / 1e6:7.1f} MB | " +
              " | ".join(f"{m} {k}: {res[(m, k)][0]:6.1f} us {res[(m, k)][1]:.2f}" for m in ("cold", "warm")
                         for k in ("stats", "apply")))


if __name__ == "__main__":
    main()
