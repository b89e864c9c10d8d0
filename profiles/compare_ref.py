#!/usr/bin/env python
"""Ours vs the reference spECK (sm_100 build in oracle/_ref) on every workload of BASELINE.json,
same box, same run: mean/min ms of the complete multiply (the reference's timings.complete scope:
events around the whole call, C already allocated), GFLOPS = 2P/t, index parity vs the CPU oracle.
usage (GPU box): python profiles/compare_ref.py [--out gpurun_out/compare.jsonl] [workloads...]"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="*", default=["econ_like", "circuit_like", "webbase_like", "cant_like", "banded_like", "rmat20"])
    ap.add_argument("--out", default="gpurun_out/compare.jsonl")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (experiments)")
    args = ap.parse_args()
    from bench import load_workload
    from speck_b200 import api
    import oracle
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    ctx = api.Context(0)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    with open(args.out, "a") as fo:
        for w in args.workloads:
            A = load_workload(w, (24 if w == "rmat24" else 20) if w.startswith("rmat") else {"webbase_like": 3, "cant_like": 44, "banded_like": 41, "econ_like": 42, "circuit_like": 43}[w])
            seed = (24 if w == "rmat24" else 20) if w.startswith("rmat") else {"webbase_like": 3, "cant_like": 44, "banded_like": 41, "econ_like": 42, "circuit_like": 43}[w]
            dA = ctx.upload(A)
            dC = api.DeviceCSR(ctx)
            for _ in range(args.warmup):
                ctx.multiply(dA, dA, dC)
            tot, wall, stages = [], [], []
            for _ in range(args.iters):
                t0 = time.perf_counter()
                ctx.multiply(dA, dA, dC)
                wall.append((time.perf_counter() - t0) * 1e3)
                s = ctx.stats()
                tot.append(s["ms_total"])
                stages.append([s["ms_analysis"], s["ms_symbolic"], s["ms_scan"], s["ms_numeric"]])
            st = ctx.stats()
            C = ctx.download(dC)
            rp, ci, v = oracle.spgemm(A.row_offsets, A.col_ids, A.data, A.row_offsets, A.col_ids, A.data, A.cols)
            ok = bool(np.array_equal(rp, C.row_offsets) and np.array_equal(ci, C.col_ids))
            rel = float(np.max(np.abs(v - C.data) / np.maximum(np.abs(v), 1e-300))) if v.size else 0.0
            P = st["products"]
            sm = np.mean(np.array(stages), axis=0)
            line = {"workload": w, "rows": A.rows, "nnz_a": A.nnz, "products": P, "nnz_c": st["nnz_c"],
                    "ours": {"mean_ms": float(np.mean(tot)), "min_ms": float(np.min(tot)), "wall_mean_ms": float(np.mean(wall)),
                             "gflops_mean": 2.0 * P / (np.mean(wall) * 1e-3) / 1e9,
                             "stage_ms": {"analysis": float(sm[0]), "symbolic": float(sm[1]), "scan": float(sm[2]), "numeric": float(sm[3])},
                             "idx_bit_exact_vs_oracle": ok, "max_rel_err_vs_oracle": rel, "class_rows": st["class_rows"],
                             "numeric_alg_gbs": api.numeric_bytes(A.rows, A.nnz, P, st["nnz_c"]) / (float(sm[3]) * 1e-3) / 1e9}}
            dC.free()
            dA.free()
            if not args.no_ref:
                for variant in ("stock", "tuned"):
                    try:
                        out = subprocess.run([sys.executable, "-m", "oracle.ref_run", "--workload", w, "--seed", str(seed),
                                              "--variant", variant, "--warmup", str(args.warmup), "--iters", str(args.iters),
                                              "--check"], capture_output=True, text=True, timeout=900, cwd=ROOT)
                        js = [l for l in out.stdout.splitlines() if l.startswith("{")]
                        line[f"ref_{variant}"] = json.loads(js[-1]) if js else {"error": out.stderr[-300:]}
                    except Exception as e:  # noqa: BLE001
                        line[f"ref_{variant}"] = {"error": repr(e)}
                best = min((line[f"ref_{v}"].get("mean_ms", 1e30) for v in ("stock", "tuned")))
                line["speedup_vs_best_ref_mean"] = best / line["ours"]["wall_mean_ms"] if best < 1e29 else None
            print(json.dumps(line), flush=True)
            fo.write(json.dumps(line) + "\n")
            fo.flush()
    ctx.close()


if __name__ == "__main__":
    main()
