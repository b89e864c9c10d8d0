"""Shared helpers of the parity tests: run the CUDA path through the C ABI, compare with the oracle."""
import numpy as np

import oracle
from speck_b200.matrices import HostCSR


def oracle_multiply(A: HostCSR, B: HostCSR) -> HostCSR:
    rp, ci, v = oracle.spgemm(A.row_offsets, A.col_ids, A.data, B.row_offsets, B.col_ids, B.data, B.cols)
    return HostCSR(A.rows, B.cols, rp, ci, v)


def assert_csr_equal(got: HostCSR, want: HostCSR, rtol=1e-6, what=""):
    """Indices bit-exact, values within rtol relative (BASELINE.json north_star: 1e-6 for fp64)."""
    assert got.rows == want.rows and got.cols == want.cols, f"{what}: shape"
    np.testing.assert_array_equal(got.row_offsets, want.row_offsets, err_msg=f"{what}: row_offsets")
    np.testing.assert_array_equal(got.col_ids, want.col_ids, err_msg=f"{what}: col_ids")
    if want.nnz:
        g, w = got.data.astype(np.float64), want.data.astype(np.float64)
        scale = max(float(np.abs(w).max()), 1e-300)
        # entries that cancel to ~0 are compared absolutely against the matrix scale
        np.testing.assert_allclose(g, w, rtol=rtol, atol=rtol * scale * 1e-3, err_msg=f"{what}: values")


def gpu_multiply(ctx, A: HostCSR, B: HostCSR = None):
    dA = ctx.upload(A)
    dB = dA if B is None else ctx.upload(B)
    dC = ctx.multiply(dA, dB)
    out = ctx.download(dC)
    st = ctx.stats()
    dC.free()
    dA.free()
    if B is not None:
        dB.free()
    return out, st


def check_case(ctx, A: HostCSR, B: HostCSR = None, rtol=1e-6, what=""):
    got, st = gpu_multiply(ctx, A, B)
    want = oracle_multiply(A, A if B is None else B)
    assert_csr_equal(got, want, rtol, what)
    return got, st
