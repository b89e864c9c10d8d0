#!/usr/bin/env python
"""Aggregate an ncu source page (ncu -i x.ncu-rep --page source --csv --print-source cuda,sass -s N -c 1)
per CUDA source line: stall samples, warp instructions executed, shared wavefronts.
usage: python profiles/src_lines.py /tmp/src.csv [top_n]"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    lines = open(path).read().split("\n")
    # find the header row
    h = next(i for i, l in enumerate(lines) if l.startswith('"Line No"'))
    rows = list(csv.reader(lines[h:]))
    hdr = rows[0]
    idx = {n: i for i, n in enumerate(hdr)}
    # first "Source" = CUDA source text, second = SASS
    src_i = hdr.index("Source")
    sass_i = hdr.index("Source", src_i + 1)
    samp_i = idx["# Samples"]
    inst_i = idx["Instructions Executed"]
    tinst_i = idx["Thread Instructions Executed"]
    wf_i = idx["L1 Wavefronts Shared"]
    agg = defaultdict(lambda: [0, 0, 0, 0, ""])
    tot = [0, 0, 0, 0]
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        def f(i):
            try:
                return int(float(r[i].replace(",", "") or 0))
            except ValueError:
                return 0
        a = agg[ln]
        vals = (f(samp_i), f(inst_i), f(tinst_i), f(wf_i))
        for k in range(4):
            a[k] += vals[k]
            tot[k] += vals[k]
        a[4] = r[src_i]
    print(f"total: samples={tot[0]} warp_inst={tot[1]} thread_inst={tot[2]} smem_wavefronts={tot[3]}")
    print(f"{'line':>5} {'samp%':>6} {'inst%':>6} {'thr/inst':>8} {'smemWf%':>7}  source")
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
        print(f"{ln:5d} {100 * a[0] / max(tot[0], 1):6.2f} {100 * a[1] / max(tot[1], 1):6.2f} "
              f"{a[2] / max(a[1], 1):8.1f} {100 * a[3] / max(tot[3], 1):7.2f}  {a[4].strip()[:90]}")


if __name__ == "__main__":
    main()
