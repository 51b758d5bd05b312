"""Not a test: prints a full parity table (every tensor, every golden) for debugging on the GPU box.
Usage: python tests/gpu_report.py > gpurun_out/parity_report.txt"""
import sys
import traceback

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
sys.path.insert(0, __file__.rsplit("/", 1)[0])
from helpers import compare_blob, load_golden  # noqa: E402

if __name__ == "__main__":
    prec = "bf16" if "--bf16" in sys.argv else "fp32"
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or \
        ["rnn210_notf", "rnn210_small", "rnn21_small", "demand_notf", "demand_small"]
    for nm in names:
        print("=" * 20, nm, prec)
        try:
            for what, e, s, ok in compare_blob(load_golden(nm), 1e-5 if prec == "fp32" else 2e-2, precision=prec):
                print(f"{'ok ' if ok else 'BAD'} {what:60s} rel={e:.3e} scale={s:.3e}")
        except Exception:
            traceback.print_exc(file=sys.stdout)
        torch.cuda.synchronize()
