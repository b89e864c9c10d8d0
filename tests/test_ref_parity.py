"""Three-way parity on the GPU box: reference spECK (compiled from /root/reference into
oracle/_ref by oracle/build_ref.sh) vs the CPU oracle vs this repository's CUDA path.
This is what pins the oracle to the reference (the reference ships no golden vectors):
row_ptr / col_idx bit-exact, values within 1e-6 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

from speck_b200 import matrices as M
from speck_b200.matrices import HostCSR
from helpers import assert_csr_equal, gpu_multiply, oracle_multiply

pytestmark = pytest.mark.gpu

ref = pytest.importorskip("oracle.ref")


def _cases():
    yield "rmat13", M.rmat(13, 16, seed=13), None
    yield "rmat16", M.rmat(16, 16, seed=16), None
    yield "uniform", M.uniform_random(3000, 3000, 12, seed=2), None
    yield "banded", M.banded_fem_like(n=4000, per_row=64, clusters=8, band=300, seed=41), None
    yield "econ_like", M.econ_like(n=20000, seed=42), None
    yield "circuit_like", M.circuit_like(n=20000, seed=43), None
    yield "rect", M.uniform_random(900, 400, 6, seed=5), M.uniform_random(400, 2500, 9, seed=6)


@pytest.mark.parametrize("variant", ["stock", "tuned"])
def test_reference_vs_oracle_vs_ours(ctx, variant):
    if not ref.available(variant):
        pytest.skip(f"oracle/_ref/libspeck_ref_{variant}.so not built (needs /root/reference at build time)")
    for name, A, B in _cases():
        want = oracle_multiply(A, A if B is None else B)
        r = ref.multiply(A, B, warmup=0, iters=1, variant=variant)
        got_ref = HostCSR(A.rows, want.cols, r["rp"], r["ci"], r["v"])
        assert_csr_equal(got_ref, want, rtol=1e-6, what=f"reference[{variant}] vs oracle: {name}")
        ours, _ = gpu_multiply(ctx, A, B)
        assert_csr_equal(ours, got_ref, rtol=1e-6, what=f"ours vs reference[{variant}]: {name}")


def _disjoint_rows(targets, nb=128, seed=0):
    """A whose row i has exactly targets[i] products and -- the B rows having pairwise disjoint column sets --
    exactly targets[i] entries in C: row k of B holds k + 1 columns of its own."""
    br, bc, base = [], [], 0
    for k in range(nb):
        br += [k] * (k + 1)
        bc += list(range(base, base + k + 1))
        base += k + 1
    B = M.from_coo(nb, base, br, bc, seed=seed + 1)
    ar, ac = [], []
    for i, t in enumerate(targets):
        left = t
        for k in range(nb - 1, -1, -1):
            if left >= k + 1:
                ar.append(i), ac.append(k)
                left -= k + 1
        assert left == 0, (t, left)
    return M.from_coo(len(targets), nb, ar, ac, seed=seed + 2), B


@pytest.mark.parametrize("variant", ["stock", "tuned"])
def test_reference_class_edges(ctx, variant):
    """The reference's own branch boundaries (stock constants): 594/595 ... 9522/9523
    products (symbolic classes) and 167/168 ... 2687/2688 entries of C (numeric classes), three ways."""
    if not ref.available(variant):
        pytest.skip(f"oracle/_ref/libspeck_ref_{variant}.so not built (needs /root/reference at build time)")
    from test_gpu_parity import _rows_with_products
    sym = [594, 595, 1189, 1190, 2379, 2380, 4760, 4761, 9522, 9523, 9524]
    num = [167, 168, 335, 336, 671, 672, 1343, 1344, 2687, 2688, 2689]
    cases = [("products", _rows_with_products(sym * 3, cols=1 << 17, nb=256, seed=3)),
             ("nnz", _disjoint_rows(num * 3, nb=128, seed=4)),
             ("mixed", _disjoint_rows(sorted(sym + num), nb=192, seed=5))]
    for name, (A, B) in cases:
        want = oracle_multiply(A, B)
        r = ref.multiply(A, B, warmup=0, iters=1, variant=variant)
        got_ref = HostCSR(A.rows, want.cols, r["rp"], r["ci"], r["v"])
        assert_csr_equal(got_ref, want, rtol=1e-6, what=f"reference[{variant}] vs oracle: class edges ({name})")
        ours, st = gpu_multiply(ctx, A, B)
        assert_csr_equal(ours, got_ref, rtol=1e-6, what=f"ours vs reference[{variant}]: class edges ({name})")
        if name != "products":
            assert st["products"] == want.nnz   # disjoint B rows: nothing folds, every edge value is a C row length


@pytest.mark.parametrize("variant", ["stock", "tuned"])
def test_fp32_reference_vs_oracle_vs_ours(ctx, variant):
    """The float instantiation (reference source/GPU/Multiply.cu:1130): indices bit-exact, values to fp32 accuracy
    (1e-4 relative: sums of up to a few hundred products in single precision, order not fixed on either side)."""
    if not ref.available(variant):
        pytest.skip(f"oracle/_ref/libspeck_ref_{variant}.so not built (needs /root/reference at build time)")
    for name, A in (("rmat14", M.rmat(14, 16, seed=14, dtype=np.float32)),
                    ("banded", M.banded_fem_like(n=3000, per_row=48, clusters=6, band=200, seed=8, dtype=np.float32)),
                    ("econ_like", M.econ_like(n=20000, seed=42, dtype=np.float32))):
        want = oracle_multiply(A.astype(np.float64), A.astype(np.float64))
        r = ref.multiply(A, None, warmup=0, iters=1, variant=variant)
        assert r["v"].dtype == np.float32
        got_ref = HostCSR(A.rows, want.cols, r["rp"], r["ci"], r["v"])
        assert_csr_equal(got_ref, want, rtol=1e-4, what=f"fp32 reference[{variant}] vs oracle: {name}")
        ours, _ = gpu_multiply(ctx, A)
        assert ours.data.dtype == np.float32
        assert_csr_equal(ours, want, rtol=1e-4, what=f"fp32 ours vs oracle: {name}")
        np.testing.assert_array_equal(ours.col_ids, got_ref.col_ids)
