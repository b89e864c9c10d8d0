#!/usr/bin/env python
"""A few multiplies of one workload and nothing else (for ncu launch lists / captures).
usage: python profiles/one_multiply.py [--workload rmat20] [--reps 3] [--opt key=value ...]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_workload  # noqa: E402
from speck_b200 import api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="rmat20")
ap.add_argument("--seed", type=int, default=None)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
A = load_workload(a.workload, a.seed)
with api.Context(0) as ctx:
    for kv in a.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    dA = ctx.upload(A)
    dC = api.DeviceCSR(ctx)
    for _ in range(a.reps):
        ctx.multiply(dA, dA, dC)
    st = ctx.stats()
    print({k: st[k] for k in ("products", "nnz_c", "kernel_launches", "ms_analysis", "ms_symbolic", "ms_numeric", "ms_total")})
    dC.free()
    dA.free()
