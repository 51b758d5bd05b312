"""Per-kernel census of the Blackwell-specific SASS in libv2f_b200.so (cuobjdump -sass; no GPU needed):
tcgen05 MMA (UTCHMMA / UTCQMMA ...), TMEM loads/stores (LDTM / STTM), TMA tensor loads (UTMALDG), bulk copies
(UBLKCP), mbarrier traffic (SYNCS), cluster barriers (UCGABAR), legacy tensor-core MMA (HMMA), plus registers per
thread from cuobjdump -res-usage.  Writes profiles/sass_census.md.
    python tools/sass_census.py [--out profiles/sass_census.md]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "visuelle2-multimodal-fusion_b200", "libv2f_b200.so")
OUT = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else os.path.join(ROOT, "profiles", "sass_census.md")
COLS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UCGABAR", "HMMA",
        "LDGSTS", "MUFU"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for c in COLS:
                if op.startswith(c):
                    counts[cur][c] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    fn = None
    for line in res.split("\n"):
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and fn:
            regs[fn] = (int(m.group(1)), int(m.group(2)))
            fn = None
    names = demangle(order)
    rows = []
    for f in order:
        c = counts[f]
        n = re.sub(r"\(.*$", "", names.get(f, f)).replace("void ", "")
        rows.append((n, c, regs.get(f, ("", ""))))
    rows.sort(key=lambda r: r[0])
    with open(OUT, "w") as fo:
        fo.write("# SASS census of libv2f_b200.so (sm_100a)\n\n")
        fo.write("`python tools/sass_census.py` = `cuobjdump -sass` + `-res-usage` of the shipped library, instruction counts "
                 "per kernel.  UTCHMMA/UTCQMMA = `tcgen05.mma`, LDTM/STTM = `tcgen05.ld/st` (TMEM), UTMALDG/UTMASTG = TMA "
                 "tensor copies (`cp.async.bulk.tensor`), UBLKCP = `cp.async.bulk`, SYNCS = mbarrier operations, UCGABAR = "
                 "cluster barrier, HMMA = legacy `mma.sync`, LDGSTS = `cp.async`.\n\n")
        fo.write("| kernel | instr | regs | static smem | " + " | ".join(COLS) + " |\n")
        fo.write("|---|---:|---:|---:|" + "---:|" * len(COLS) + "\n")
        for n, c, (rg, sm) in rows:
            fo.write(f"| `{n}` | {c['_total']} | {rg} | {sm} | " + " | ".join(str(c[k]) if c[k] else "" for k in COLS) + " |\n")
        tot = collections.Counter()
        for _, c, _ in rows:
            tot.update(c)
        fo.write(f"| **all {len(rows)} kernels** | {tot['_total']} | | | " + " | ".join(str(tot[k]) for k in COLS) + " |\n")
    print("wrote", OUT, len(rows), "kernels")


if __name__ == "__main__":
    main()
