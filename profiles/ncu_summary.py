#!/usr/bin/env python
"""Summarise a .ncu-rep (ncu --set full): per kernel duration, DRAM bytes, issue activity, occupancy,
stall mix, and -- with --regions -- the stall/instruction share of SASS regions split at given
instruction indices.  usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv
import io
import subprocess
import sys


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    path = sys.argv[1]
    hdr, units, rows = raw(path)
    col = {h: i for i, h in enumerate(hdr)}
    want = [("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "rdGB"), ("dram__bytes_write.sum", "wrGB"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
            ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "inst"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
            ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smemWf")]
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    for r in rows:
        name = r[col["Kernel Name"]]
        short = name[name.find("k_"):name.find("(")][:70]
        vals = []
        for k, lab in want:
            if k in col:
                v = r[col[k]]
                u = units[col[k]]
                try:
                    f = float(v.replace(",", ""))
                    if lab == "ms" and u.startswith("us"):
                        f /= 1e3
                    if lab in ("rdGB", "wrGB") and u.startswith("M"):
                        f /= 1e3
                    vals.append(f"{lab}={f:.4g}")
                except ValueError:
                    vals.append(f"{lab}={v}")
        st = sorted(((float(r[col[h]]), h[34:-23]) for h in stalls), reverse=True)[:6]
        print(short)
        print("   ", " ".join(vals))
        print("    stalls/issue:", " ".join(f"{n}={v:.2f}" for v, n in st))


if __name__ == "__main__":
    main()
