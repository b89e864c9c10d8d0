#!/usr/bin/env python
"""Summarise an ncu launch list (gpu__time_duration.sum per launch) for one multiply.
usage: python profiles/parse_launches.py gpurun_out/<tag>_launches.csv [multiply_index]"""
import csv
import re
import sys


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    idx = [i for i, x in enumerate(rows) if "k_analyze" in x["Kernel Name"]]
    s = idx[which]
    e = idx[which + 1] if len(idx) > which + 1 else len(rows)
    tot = 0.0
    out = []
    for x in rows[s:e]:
        v = float(x["Metric Value"].replace(",", ""))
        unit = x["Metric Unit"]
        ms = v / 1e6 if unit.startswith("n") else (v / 1e3 if unit.startswith("u") else v)
        name = x["Kernel Name"]
        m = re.match(r"(?:void )?(\w+)(<[^(]*>)?\(", name)
        short = (m.group(1) + (m.group(2) or "")) if m else name[:60]
        tot += ms
        out.append((ms, x["Grid Size"], x["Block Size"], short))
    print(f"# {path}: multiply #{which}, {e - s} launches, sum of kernel times {tot:.3f} ms (serialised, cold cache)")
    for ms, g, b, n in out:
        print(f"{ms:9.3f} ms {ms / tot * 100:5.1f}%  grid={g:<14} block={b:<13} {n}")


if __name__ == "__main__":
    main()
