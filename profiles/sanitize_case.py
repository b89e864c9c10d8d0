#!/usr/bin/env python
"""Small multiplies that touch every kernel family, for compute-sanitizer (memcheck / racecheck / initcheck):
the smoke matrix, the class-boundary rows of tests/test_gpu_parity.py, a FEM-like matrix (bitmap rows, sequential-k
kernel in both staging variants), the deterministic mode.  Checks every result against the CPU oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from speck_b200 import api, matrices as M  # noqa: E402
from test_gpu_parity import _rows_with_products  # noqa: E402


def check(ctx, A, B=None, what=""):
    dA = ctx.upload(A)
    dB = dA if B is None else ctx.upload(B)
    C = ctx.download(ctx.multiply(dA, dB))
    Bh = A if B is None else B
    rp, ci, v = oracle.spgemm(A.row_offsets, A.col_ids, A.data, Bh.row_offsets, Bh.col_ids, Bh.data, Bh.cols)
    assert np.array_equal(C.row_offsets, rp) and np.array_equal(C.col_ids, ci), what
    assert np.allclose(C.data, v, rtol=1e-6, atol=0), what
    print("ok", what, "P =", ctx.stats()["products"], flush=True)


with api.Context(0) as ctx:
    check(ctx, M.rmat(12, 16, seed=7), what="smoke matrix")
    targets = []
    for c in range(0, 12):
        b = 4 << c
        targets += [b - 1, b, b + 1]
    targets += [1, 2, 3, 8255, 8256, 12345, 16383, 16384, 16385, 20000]
    A, B = _rows_with_products(targets, cols=1 << 18, nb=256)
    check(ctx, A, B, what="class boundaries (mapped)")
    ctx.set_option("rank_map", 0)
    check(ctx, A, B, what="class boundaries (self-contained numeric kernels)")
    ctx.set_option("rank_map", 1)
    ctx.set_option("flat_sym", 0)
    check(ctx, A, B, what="class boundaries (register-slot symbolic kernel)")
    ctx.set_option("flat_sym", 1)
    # lane-group shapes (several rows per warp is the default), capped-grid sort kernels, several CTAs per large row
    S = M.rmat(13, 4, seed=21)
    for v in (0, 1, 2):
        ctx.set_option("narrow_groups", v)
        check(ctx, S, what=f"small rows, narrow_groups={v}")
    ctx.set_option("sym_mix", 1)
    check(ctx, M.rmat(13, 16, seed=22), what="sym_mix=1 (k_sort_rows_loop)")
    ctx.set_option("sym_mix", 0)
    for v in (1, 2):
        ctx.set_option("big_split", v)
        check(ctx, A, B, what=f"class boundaries, big_split={v}")
    ctx.set_option("big_split", 0)
    ctx.set_option("flat_min_class", 6)
    check(ctx, A, B, what="class boundaries, flat_min_class=6")
    ctx.set_option("flat_min_class", 8)
    F = M.fem3d_like(7, 6, 5)
    for v in (1, 2, 0):
        ctx.set_option("dense_seq", v)
        check(ctx, F, what=f"FEM-like, dense_seq={v}")
    ctx.set_option("dense_seq", 1)
    ctx.set_option("deterministic", 1)
    check(ctx, M.rmat(13, 16, seed=13), what="deterministic mode")
    ctx.set_option("deterministic", 0)
    # hub rows of A (k_analyze_long) and the GPU-side COO -> CSR
    rng = np.random.default_rng(8)
    n = 4000
    r = list(rng.integers(0, n, 12000)) + [7] * 1500 + [3999] * 2500
    c = list(rng.integers(0, n, 12000)) + list(rng.choice(n, 1500, replace=False)) + list(rng.choice(n, 2500, replace=False))
    H = M.from_coo(n, n, np.array(r), np.array(c), seed=9)
    check(ctx, H, what="hub rows of A")
    d = ctx.coo_to_csr(n, n, np.array(r, np.uint32), np.array(c, np.uint32), rng.standard_normal(len(r)), sum_duplicates=True)
    h = ctx.download(d)
    assert h.nnz == H.nnz and np.array_equal(h.col_ids, H.col_ids), "coo_to_csr"
    d.free()
    print("ok GPU COO -> CSR", flush=True)
    eq, mm = ctx.compare_report(ctx.upload(H), ctx.upload(H), True)
    assert eq
print("all ok")
