"""CPU tests (no GPU): the oracle against the committed golden fixtures (made with scipy by
tests/golden/make_golden.py) and against scipy on seeded random inputs; edge cases of SURVEY
Appendix C that the oracle must get right before it may judge the CUDA path."""
import os

import numpy as np
import pytest

import oracle
from speck_b200 import matrices as M
from speck_b200.matrices import HostCSR

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def mul(A, B=None):
    B = A if B is None else B
    return oracle.spgemm(A.row_offsets, A.col_ids, A.data, B.row_offsets, B.col_ids, B.data, B.cols)


def scipy_mul(A, B=None):
    B = A if B is None else B
    C = (A.to_scipy() @ B.to_scipy()).tocsr()
    C.sort_indices()
    return C.indptr.astype(np.uint32), C.indices.astype(np.uint32), C.data


def test_tiny8_golden():
    z = np.load(os.path.join(GOLD, "tiny8_expected.npz"))
    T = M.tiny8()
    np.testing.assert_array_equal(T.row_offsets, z["a_rp"])
    np.testing.assert_array_equal(T.col_ids, z["a_ci"])
    rp, ci, v = mul(T)
    np.testing.assert_array_equal(rp, z["c_rp"])
    np.testing.assert_array_equal(ci, z["c_ci"])
    np.testing.assert_allclose(v, z["c_v"], rtol=1e-14)


def test_small_golden_cases():
    z = np.load(os.path.join(GOLD, "small_cases.npz"))
    names = sorted({k.split(".")[0] for k in z.files})
    assert {"random64", "empty_rows", "rect", "rmat8", "cancel"} <= set(names)
    for n in names:
        A = HostCSR(int(z[f"{n}.a_shape"][0]), int(z[f"{n}.a_shape"][1]), z[f"{n}.a_rp"], z[f"{n}.a_ci"], z[f"{n}.a_v"])
        B = HostCSR(int(z[f"{n}.b_shape"][0]), int(z[f"{n}.b_shape"][1]), z[f"{n}.b_rp"], z[f"{n}.b_ci"], z[f"{n}.b_v"])
        rp, ci, v = mul(A, B)
        np.testing.assert_array_equal(rp, z[f"{n}.c_rp"], err_msg=n)
        np.testing.assert_array_equal(ci, z[f"{n}.c_ci"], err_msg=n)
        np.testing.assert_allclose(v, z[f"{n}.c_v"], rtol=1e-13, atol=1e-13, err_msg=n)


def test_cancellation_keeps_entry():
    z = np.load(os.path.join(GOLD, "small_cases.npz"))
    assert z["cancel.c_v"][0] == 0.0 and z["cancel.c_rp"][-1] == 2   # structural zero stays (no numeric dropping)


@pytest.mark.parametrize("seed", range(6))
def test_random_vs_scipy(seed):
    rng = np.random.default_rng(seed)
    n, k, m = int(rng.integers(1, 400)), int(rng.integers(1, 400)), int(rng.integers(1, 400))
    A = M.uniform_random(n, k, float(rng.uniform(0.5, 12)), seed=seed)
    B = M.uniform_random(k, m, float(rng.uniform(0.5, 12)), seed=seed + 100)
    rp, ci, v = mul(A, B)
    srp, sci, sv = scipy_mul(A, B)
    np.testing.assert_array_equal(rp, srp)
    np.testing.assert_array_equal(ci, sci)
    np.testing.assert_allclose(v, sv, rtol=1e-12)


def test_rmat_vs_scipy_and_row_products():
    A = M.rmat(12, 16, seed=3)
    rp, ci, v = mul(A)
    srp, sci, sv = scipy_mul(A)
    np.testing.assert_array_equal(rp, srp)
    np.testing.assert_array_equal(ci, sci)
    np.testing.assert_allclose(v, sv, rtol=1e-12)
    ops, mx, P, gmax = oracle.row_products(A.row_offsets, A.col_ids, A.row_offsets)
    blen = np.diff(A.row_offsets.astype(np.int64))
    want = np.add.reduceat(np.concatenate([blen[A.col_ids], [0]]), A.row_offsets[:-1].astype(np.int64))
    want[np.diff(A.row_offsets.astype(np.int64)) == 0] = 0
    np.testing.assert_array_equal(ops, want.astype(np.uint32))
    assert P == int(want.sum()) and gmax == int(want.max())
    # symbolic alone agrees with the numeric row_ptr
    srp2, nnz = oracle.symbolic(A.row_offsets, A.col_ids, A.row_offsets, A.col_ids, A.cols)
    np.testing.assert_array_equal(srp2, rp)
    assert nnz == ci.size


def test_fp32_oracle():
    A = M.rmat(9, 8, seed=5, dtype=np.float32)
    rp, ci, v = mul(A)
    assert v.dtype == np.float32
    srp, sci, sv = scipy_mul(A.astype(np.float64))
    np.testing.assert_array_equal(ci, sci)
    np.testing.assert_allclose(v, sv, rtol=1e-5)


def test_empty_conventions():
    Z = HostCSR(5, 5, np.zeros(6, np.uint32), np.zeros(0, np.uint32), np.zeros(0))
    rp, ci, v = mul(Z)
    assert ci.size == 0 and not rp.any()
    # A references only empty rows of B -> P = 0
    A = HostCSR(3, 3, np.array([0, 1, 1, 1], np.uint32), np.array([2], np.uint32), np.array([1.0]))
    B = HostCSR(3, 3, np.array([0, 1, 1, 1], np.uint32), np.array([0], np.uint32), np.array([1.0]))
    rp, ci, v = mul(A, B)
    assert ci.size == 0
    _, _, P, _ = oracle.row_products(A.row_offsets, A.col_ids, B.row_offsets)
    assert P == 0


def test_compare_restatement():
    A = M.rmat(8, 8, seed=1)
    rp, ci, v = mul(A)
    assert oracle.compare(rp, ci, v, rp, ci, v)[0] == 0
    ci2 = ci.copy()
    ci2[7] ^= 1
    assert oracle.compare(rp, ci, v, rp, ci2, v)[0] == 2
    v2 = v.copy()
    v2[3] *= 1.0 + 1e-3
    assert oracle.compare(rp, ci, v, rp, ci, v2, rel_tol=1e-6)[0] == 3
    assert oracle.compare(rp, ci, v, rp, ci, v2, rel_tol=1e-2)[0] == 0
    rp2 = rp.copy()
    rp2[5] += 1
    assert oracle.compare(rp, ci, v, rp2, ci, v)[0] == 1


def test_generators_are_sorted_and_duplicate_free():
    for A in (M.tiny8(), M.rmat(10, 8, seed=2), M.banded_fem_like(n=500, band=100), M.econ_like(n=3000),
              M.circuit_like(n=3000), M.webbase_like(n=5000)):
        for i in range(0, A.rows, max(1, A.rows // 200)):
            row = A.col_ids[A.row_offsets[i]:A.row_offsets[i + 1]].astype(np.int64)
            assert np.all(np.diff(row) > 0)
        assert A.row_offsets[-1] == A.nnz and A.col_ids.max() < A.cols


def test_config2_counts_at_reduced_scale():
    """The R-MAT generator is the one SURVEY Appendix D pins (counts quoted there for scale 20 are
    checked on the GPU box by bench.py); at scale 12 the counts are pinned here."""
    A = M.rmat(12, 16, seed=20)
    _, _, P, _ = oracle.row_products(A.row_offsets, A.col_ids, A.row_offsets)
    assert (A.rows, A.nnz) == (4096, A.nnz) and A.nnz > 60000 and P > 1_000_000
