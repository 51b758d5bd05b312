"""Summarise ncu outputs into small text files for profiles/.
  python tools/ncu_extract.py launches <launches.csv> <steps> > profiles/...md
  python tools/ncu_extract.py full <report.ncu-rep> > profiles/...csv
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "l1tex__t_bytes.sum", "sm__inst_executed_pipe_uniform.sum", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(path, steps):
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")[:90]
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list ({path.split('/')[-1]}): {sum(v[0] for v in agg.values())} launches over {steps} steps, "
          f"{tot / steps / 1e6:.2f} ms of kernel time per step (cold-cache, serialised replays)\n")
    print("| kernel | launches/step | ms/step | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"| `{k}` | {v[0] / steps:.0f} | {v[1] / steps / 1e6:.3f} | {100 * v[1] / tot:.1f} % |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    print("metric,unit," + ",".join(r[h.index("Kernel Name")][:60].replace(",", ";") for r in rows[2:]))
    for i, k in enumerate(h):
        if k in KEYS or k.startswith("smsp__average_warps_issue_stalled") or k.startswith("sm__pipe_tensor_cycles_active") \
                or ("utchmma" in k and "sparsity_off" in k and ("pct_of_peak_sustained_elapsed" in k or k.endswith(".sum"))
                    and rows[2][i] not in ("0", "0.0")):
            print(",".join([k, units[i]] + [r[i] for r in rows[2:]]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]))
    else:
        full(sys.argv[2])
