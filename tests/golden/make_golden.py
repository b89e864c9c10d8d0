#!/usr/bin/env python
"""Generates the committed golden fixtures of tests/golden/ with scipy.sparse (NOT with the oracle
and not with the CUDA path), so that the oracle itself is pinned against an independent
implementation.  The reference ships no golden vectors of its own.

    python tests/golden/make_golden.py

Fixtures:
  tiny8.mtx              BASELINE config #1 input (general real coordinate, 1-based)
  tiny8_expected.npz     C = A.A for tiny8 (row_offsets, col_ids, data), scipy, sorted indices
  small_cases.npz        a few seeded small A (and B) with scipy's C: random, with empty rows,
                         rectangular, and one with exact cancellation (structural zero kept; hand-written, scipy drops it)
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from speck_b200 import matrices as M  # noqa: E402


def product(A, B):
    C = (A.to_scipy() @ B.to_scipy()).tocsr()
    C.sort_indices()
    return C.indptr.astype(np.uint32), C.indices.astype(np.uint32), C.data.astype(np.float64)


def main():
    T = M.tiny8()
    with open(os.path.join(HERE, "tiny8.mtx"), "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n% BASELINE config #1: tiny 8x8\n")
        f.write(f"{T.rows} {T.cols} {T.nnz}\n")
        for i in range(T.rows):
            for p in range(T.row_offsets[i], T.row_offsets[i + 1]):
                f.write(f"{i + 1} {T.col_ids[p] + 1} {float(T.data[p])!r}\n")
    rp, ci, v = product(T, T)
    np.savez(os.path.join(HERE, "tiny8_expected.npz"), a_rp=T.row_offsets, a_ci=T.col_ids, a_v=T.data,
             c_rp=rp, c_ci=ci, c_v=v)

    cases = {}

    def add(name, A, B=None):
        B = A if B is None else B
        rp, ci, v = product(A, B)
        for tag, m in (("a", A), ("b", B)):
            cases[f"{name}.{tag}_shape"] = np.array([m.rows, m.cols], np.int64)
            cases[f"{name}.{tag}_rp"], cases[f"{name}.{tag}_ci"], cases[f"{name}.{tag}_v"] = m.row_offsets, m.col_ids, m.data
        cases[f"{name}.c_rp"], cases[f"{name}.c_ci"], cases[f"{name}.c_v"] = rp, ci, v

    add("random64", M.uniform_random(64, 64, 5, seed=101))
    rng = np.random.default_rng(102)
    r, c = rng.integers(0, 90, 400), rng.integers(0, 90, 400)
    keep = r % 3 != 0
    add("empty_rows", M.from_coo(90, 90, r[keep], c[keep], seed=103))
    add("rect", M.uniform_random(40, 25, 4, seed=104), M.uniform_random(25, 70, 6, seed=105))
    add("rmat8", M.rmat(8, 8, seed=106))
    A = M.HostCSR(2, 2, np.array([0, 2, 3], np.uint32), np.array([0, 1, 1], np.uint32), np.array([1.0, 1.0, 5.0]))
    B = M.HostCSR(2, 2, np.array([0, 1, 2], np.uint32), np.array([0, 0], np.uint32), np.array([2.0, -2.0]))
    add("cancel", A, B)
    # scipy's csr_matmat drops entries whose sum is exactly 0; the reference never drops
    # numerically, so this one expectation is written by hand:
    # C[0,0] = 1*2 + 1*(-2) = 0 (kept), C[1,0] = 5*(-2) = -10
    cases["cancel.c_rp"] = np.array([0, 1, 2], np.uint32)
    cases["cancel.c_ci"] = np.array([0, 0], np.uint32)
    cases["cancel.c_v"] = np.array([0.0, -10.0])
    np.savez_compressed(os.path.join(HERE, "small_cases.npz"), **cases)
    print("golden fixtures written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
