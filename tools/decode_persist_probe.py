"""A/B probe of the persistent decoder (csrc/decode_persist.cu) on one B200: head-only CrossAttnRNN210 forward at
the bench dims, persistent vs step-per-launch, CUDA-event times and the in-kernel phase stamps of CTA 0.
    python tools/decode_persist_probe.py [B] [precision]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    import visuelle2_multimodal_fusion_b200.functional as Fv
    import visuelle2_multimodal_fusion_b200.synth as synth
    from visuelle2_multimodal_fusion_b200 import _lib
    from test_gpu_rnn import _head_model
    m = _head_model("CrossAttnRNN210", 512)
    m.precision = precision
    m.use_teacher_forcing = True
    data, feat = synth.make_batch(B, out_len=10, seed=5, feat_hw=10)
    data = tuple(t.cuda() for t in data)
    feat = feat.cuda()
    lib = _lib.lib()
    dbg = int(os.environ.get("DP_DBG", "0"))
    if dbg:
        lib.v2f_decode_persist_debug(dbg)
    # capture the persist workspace of the last forward to read the stamps
    holder = {}
    orig = Fv._f32

    def spy(*shape, **kw):
        t = orig(*shape, **kw)
        holder["last"] = holder.get("last", []) + [t]
        return t

    for flag in (True, False):
        Fv.PERSISTENT_DECODE = flag
        for fwd_only in (True, False):
            ts = []
            for it in range(8):
                torch.manual_seed(9)
                f = feat.clone().requires_grad_(not fwd_only)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                with torch.set_grad_enabled(not fwd_only):
                    out = m(*data, f)[0]
                    if not fwd_only:
                        out.square().mean().backward()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
                m.zero_grad(set_to_none=True)
            print(f"persistent={flag} {'fwd' if fwd_only else 'fwd+bwd'} head ms: min {min(ts):.3f} med {sorted(ts)[len(ts)//2]:.3f}")
    Fv.PERSISTENT_DECODE = True
    # kernel time of the persistent launch alone
    _lib.prof_enable(True)
    for _ in range(5):
        with torch.no_grad():
            m(*data, feat)
    ms, n = _lib.prof_read(_lib.K_DECODE_PERSIST_FWD)
    _lib.prof_enable(False)
    print(f"decode_persist_fwd_kernel: {n} launches, {1e3 * ms / max(n, 1):.1f} us each")
    # phase stamps
    lib.v2f_decode_persist_stamps_enable(1)
    Fv._f32 = spy
    try:
        with torch.no_grad():
            m(*data, feat)
        torch.cuda.synchronize()
    finally:
        Fv._f32 = orig
        lib.v2f_decode_persist_stamps_enable(0)
    N, E, H, T = B, 512, 512, 10
    want = lib.v2f_decode_persist_ws_floats(N, E, H, T)
    ws = [t for t in holder["last"] if t.numel() == want][-1]
    off = lib.v2f_decode_persist_stamps_offset(N, E, H) // 4
    st = ws[off:off + 2 * T * 16].cpu().view(torch.int64).view(T, 16)
    names = ["P1 S-product", "P2 attention", "P2b combine", "P3 HC-product", "P4 mm-attn", "P5/6 embed+gates"]
    work, wait = [0.0] * 6, [0.0] * 6
    for t in range(T):
        w = [(int(st[t, 2 * k + 1]) - int(st[t, 2 * k])) / 1e3 for k in range(6)]
        b = [(int(st[t, 2 * k + 2]) - int(st[t, 2 * k + 1])) / 1e3 for k in range(6)]
        work = [x + y for x, y in zip(work, w)]
        wait = [x + y for x, y in zip(wait, b)]
        if t in (0, T - 1):
            print(f"step {t}: " + "  ".join(f"{n}={x:.1f}+{y:.1f}us" for n, x, y in zip(names, w, b)))
    tot = [x + y for x, y in zip(work, wait)]
    print("mean per step (CTA 0: work + barrier wait): " +
          "  ".join(f"{n}={x / T:.1f}+{y / T:.1f}us" for n, x, y in zip(names, work, wait)),
          f" | step total {sum(tot) / T:.1f} us")
    bytes_step = N * 4 * (2 * 100 + 2 * 52) * 512 + N * 4 * (5 * 512 + 100 + 52 + 4)
    print(f"attention phase: {bytes_step / 1e6:.2f} MB per step -> {bytes_step / (tot[1] / T * 1e-6) / 1e9:.0f} GB/s")


if __name__ == "__main__":
    main()
