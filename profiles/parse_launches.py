#!/usr/bin/env python
"""Summarise an ncu launch list (gpu__time_duration.sum and, when captured, dram__bytes_{read,write}.sum per
launch) for one multiply.   usage: python profiles/parse_launches.py gpurun_out/<tag>_launches.csv [multiply_index]"""
import csv
import re
import sys
from collections import OrderedDict


def to_ms(v, unit):
    return v / 1e6 if unit.startswith("n") else (v / 1e3 if unit.startswith("u") else (v * 1e3 if unit == "s" else v))


def to_gb(v, unit):
    u = unit.lower()
    return v / 1e9 if u == "byte" else (v / 1e6 if u.startswith("k") else (v / 1e3 if u.startswith("m") else v))


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    lines = [l for l in open(path) if not l.startswith("==")]
    launches = OrderedDict()
    for x in csv.DictReader(lines):
        k = launches.setdefault(x["ID"], {"name": x["Kernel Name"], "grid": x["Grid Size"], "block": x["Block Size"]})
        v = float(x["Metric Value"].replace(",", ""))
        m = x["Metric Name"]
        if m == "gpu__time_duration.sum":
            k["ms"] = to_ms(v, x["Metric Unit"])
        elif m == "dram__bytes_read.sum":
            k["rd"] = to_gb(v, x["Metric Unit"])
        elif m == "dram__bytes_write.sum":
            k["wr"] = to_gb(v, x["Metric Unit"])
        elif m == "smsp__inst_executed.sum":
            k["inst"] = v
        elif m == "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum":
            k["wf"] = v
    rows = list(launches.values())
    idx = [i for i, x in enumerate(rows) if "k_analyze" in x["name"]]
    s = idx[which]
    e = idx[which + 1] if len(idx) > which + 1 else len(rows)
    sel = rows[s:e]
    tot = sum(x["ms"] for x in sel)
    print(f"# {path}: multiply #{which}, {e - s} launches, sum of kernel times {tot:.3f} ms (serialised, cold cache)")
    phase = "analysis"
    agg = {}
    for x in sel:
        m = re.match(r"(?:void )?(?:sb::)?(\w+)(<[^(]*>)?\(", x["name"])
        short = (m.group(1) + (m.group(2) or "")) if m else x["name"][:60]
        if "k_scan<unsigned int>" in short:
            phase = "scan"
        elif phase == "analysis" and ("k_rank_rows" in short or "k_rank_flat" in short or "k_hash_count" in short or "k_sort_rows" in short or "k_dense_rows" in short):
            phase = "symbolic"
        elif phase == "scan" and "k_scan" not in short:
            phase = "numeric"
        a = agg.setdefault(phase, [0.0, 0.0, 0.0])
        a[0] += x["ms"]
        a[1] += x.get("rd", 0.0)
        a[2] += x.get("wr", 0.0)
        dram = f"  dram rd {x['rd']:6.3f} wr {x['wr']:6.3f} GB" if "rd" in x else ""
        if "inst" in x:
            dram += f"  inst {x['inst'] / 1e6:8.1f} M  smem wf {x.get('wf', 0.0) / 1e6:8.1f} M"
        print(f"{x['ms']:9.3f} ms {x['ms'] / tot * 100:5.1f}%  grid={x['grid']:<14} block={x['block']:<13} {short[:58]:58s}{dram}")
    for ph, a in agg.items():
        print(f"# phase {ph:9s}: {a[0]:7.3f} ms  dram read {a[1]:7.3f} GB  write {a[2]:7.3f} GB")


if __name__ == "__main__":
    main()
