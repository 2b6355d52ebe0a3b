"""Per-source-line instruction counts of one kernel from an ncu report.

ncu's CSV export of the source page correlates metrics with SASS only; this joins the SASS rows (by offset from the kernel's
first instruction) with the line table of `nvdisasm -g` run on a cubin of the SAME source and prints, per source line, the warp
instructions executed, the average active lanes and the stall samples.

  python tools/ncu_by_line.py REPORT.ncu-rep CUBIN KERNEL_SUBSTRING [--top N] [--ranges a-b,c-d]
"""
from __future__ import annotations

import collections
import csv
import io
import re
import subprocess
import sys


def sass_rows(report: str):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hd = rows[1]
    col = {h: i for i, h in enumerate(hd)}
    base = None
    for r in rows[2:]:
        if len(r) < len(hd) or not r[0].startswith("0x"):
            continue
        addr = int(r[0], 16)
        if base is None:
            base = addr
        yield addr - base, r[1].strip(), int(r[col["Instructions Executed"]]), int(r[col["Thread Instructions Executed"]]), int(r[col["# Samples"]])


def line_table(cubin: str, kernel: str):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    table = {}
    inside = False
    cur = ("?", 0)
    for ln in txt.splitlines():
        if ln.startswith("//-----") and ".text." in ln:
            inside = kernel in ln
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            if "inlined at" not in ln:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
            else:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    report, cubin, kernel = sys.argv[1:4]
    top = 40
    if "--top" in sys.argv:
        top = int(sys.argv[sys.argv.index("--top") + 1])
    table = line_table(cubin, kernel)
    per = collections.defaultdict(lambda: [0, 0, 0, 0])
    total = 0
    missing = 0
    for off, text, inst, tinst, samples in sass_rows(report):
        total += inst
        if off not in table:
            missing += inst
            continue
        key = table[off][0]
        a = per[key]
        a[0] += inst
        a[1] += tinst
        a[2] += samples
        a[3] += 1
    print(f"total warp instructions {total:.4g}; unmapped {missing:.3g}")
    srcs = {}
    for (f, l), a in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{f}:{l:5d}  inst {a[0]:12d} ({100 * a[0] / total:5.2f} %)  lanes {a[1] / max(a[0], 1):5.1f}  samples {a[2]:7d}  sass {a[3]}")
    if "--files" in sys.argv:
        agg = collections.defaultdict(int)
        for (f, l), a in per.items():
            agg[f] += a[0]
        print(dict(agg))
    if "--ranges" in sys.argv:
        spec = sys.argv[sys.argv.index("--ranges") + 1]
        for part in spec.split(","):
            f, rng = part.split(":")
            lo, hi = map(int, rng.split("-"))
            s = [0, 0, 0]
            for (ff, l), a in per.items():
                if ff == f and lo <= l <= hi:
                    s[0] += a[0]
                    s[1] += a[1]
                    s[2] += a[2]
            print(f"{part}: inst {s[0]} ({100 * s[0] / total:.2f} %) lanes {s[1] / max(s[0], 1):.1f} samples {s[2]}")


if __name__ == "__main__":
    main()
