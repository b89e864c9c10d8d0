"""Run the compiled reference spECK (oracle/_ref) on one workload and print one JSON line.
Used by bench.py (in a subprocess, so a crash of the reference cannot take the bench down) and by
profiles/.  usage: python -m oracle.ref_run --workload rmat20 --variant stock --warmup 3 --iters 5"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="rmat20")
    ap.add_argument("--seed", type=int, default=20)
    ap.add_argument("--variant", default="stock")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--stages", action="store_true")
    ap.add_argument("--check", action="store_true", help="compare indices with the CPU oracle")
    args = ap.parse_args()
    from bench import load_workload
    from oracle import ref
    import oracle
    if not ref.available(args.variant):
        print(json.dumps({"unavailable": f"oracle/_ref/libspeck_ref_{args.variant}.so not built"}))
        return
    A = load_workload(args.workload, args.seed)
    _, _, P, _ = oracle.row_products(A.row_offsets, A.col_ids, A.row_offsets)
    out = ref.multiply(A, None, args.warmup, args.iters, args.variant, stages=args.stages, fetch=args.check)
    t = out["times_ms"]
    line = {"impl": f"reference spECK (sm_100 build, {args.variant} smem constants)", "workload": args.workload,
            "products": P, "nnz_c": out["nnz"], "mean_ms": float(t.mean()), "min_ms": float(t.min()),
            "gflops_mean": 2.0 * P / (float(t.mean()) * 1e-3) / 1e9, "gflops_best": 2.0 * P / (float(t.min()) * 1e-3) / 1e9,
            "warmup": args.warmup, "iters": args.iters, "stage_ms": out["stage_ms"]}
    if args.check:
        rp, ci, v = oracle.spgemm(A.row_offsets, A.col_ids, A.data, A.row_offsets, A.col_ids, A.data, A.cols)
        line["row_ptr_equal"] = bool(np.array_equal(rp, out["rp"]))
        line["col_idx_equal"] = bool(np.array_equal(ci, out["ci"]))
        line["max_rel_err"] = float(np.max(np.abs(v - out["v"]) / np.maximum(np.abs(v), 1e-300))) if v.size else 0.0
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
