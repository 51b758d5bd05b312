"""Timeline evidence that the bucketed NCCL gradient all-reduces run CONCURRENTLY with the backward kernels of the
graph-replayed step (nsys is not in the image: torch.profiler / CUPTI kernel records instead).

  torchrun --nproc-per-node 2 tools/overlap_trace.py > profiles/r02_nccl_overlap.md      (rank 0 prints)

For 3 replays of the CrossAttnRNN210 step at B = 128 per rank: every kernel's [start, end) on the device, split into NCCL
kernels and the step's own kernels; reported: per replay the step span, the union of NCCL-kernel time, how much of it
lies inside intervals where a non-NCCL kernel of the step is running, and the un-hidden tail after the last compute kernel."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402


def union(iv):
    iv = sorted(iv)
    out = []
    for a, b in iv:
        if out and a <= out[-1][1]:
            out[-1][1] = max(out[-1][1], b)
        else:
            out.append([a, b])
    return out


def overlap(u1, u2):
    i = j = 0
    tot = 0.0
    while i < len(u1) and j < len(u2):
        a, b = max(u1[i][0], u2[j][0]), min(u1[i][1], u2[j][1])
        if b > a:
            tot += b - a
        if u1[i][1] < u2[j][1]:
            i += 1
        else:
            j += 1
    return tot


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    env = bench._Env()
    env.rank, env.world, env.dev = rank, world, torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=env.dev)
    env.barrier = (lambda: (dist.barrier(), torch.cuda.synchronize())) if world > 1 else torch.cuda.synchronize
    r = bench._Runner(env, "rnn210", "bf16", 128, True, True)
    for i in range(5):
        r.step(i)
    torch.cuda.synchronize()
    steps = 3
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
        for i in range(steps):
            r.step(i)
            torch.cuda.synchronize()
    if rank == 0:
        path = os.path.join(tempfile.gettempdir(), "v2f_trace.json")
        prof.export_chrome_trace(path)
        ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
        ev.sort(key=lambda e: e["ts"])
        # split into replays: every replay launches the same kernel sequence
        per = len(ev) // steps
        groups = [ev[i * per:(i + 1) * per] for i in range(steps)] if per * steps == len(ev) else [ev]
        print(f"# NCCL all-reduce vs backward overlap inside the graph-replayed step ({world} GPUs, rank 0, torch.profiler/CUPTI)\n")
        print("| replay | kernels | step span ms | NCCL kernels | NCCL busy ms | of it under compute kernels | tail after last compute kernel ms |")
        print("|---:|---:|---:|---:|---:|---:|---:|")
        for gi, g in enumerate(groups):
            nccl = [(e["ts"], e["ts"] + e["dur"]) for e in g if "nccl" in e["name"].lower()]
            comp = [(e["ts"], e["ts"] + e["dur"]) for e in g if "nccl" not in e["name"].lower()]
            un, uc = union(nccl), union(comp)
            busy = sum(b - a for a, b in un)
            ov = overlap(un, uc)
            t0, t1 = g[0]["ts"], max(e["ts"] + e["dur"] for e in g)
            last_comp = max(b for _, b in comp)
            print(f"| {gi} | {len(g)} | {(t1 - t0) / 1e3:.2f} | {len(nccl)} | {busy / 1e3:.3f} | {100 * ov / max(busy, 1e-9):.1f} % | {(t1 - last_comp) / 1e3:.3f} |")
        g = groups[-1]
        t0 = g[0]["ts"]
        print("\nNCCL kernels of the last replay (start and end relative to the replay's first kernel, ms):\n")
        for e in g:
            if "nccl" in e["name"].lower():
                print(f"* `{e['name'][:60]}` {(e['ts'] - t0) / 1e3:.3f} -> {(e['ts'] + e['dur'] - t0) / 1e3:.3f}")
        last = max(g, key=lambda e: e["ts"] + e["dur"])
        print(f"\nlast kernel of the replay: `{last['name'][:70]}` ends at {(last['ts'] + last['dur'] - t0) / 1e3:.3f} ms")
    r.close()                        # the graph holds NCCL kernels: it must die before the process group
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
