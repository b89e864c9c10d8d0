"""GPU, full size: the BASELINE.json configurations (config #2, the SuiteSparse-shaped generators of configs #3/#4,
the config-5 matrix on one GPU) three ways -- reference spECK (oracle/_ref, compiled from /root/reference) vs the
CPU oracle vs this repository's CUDA path through the C ABI.  row_ptr / col_idx bit-exact, values within 1e-6
relative (north_star).  Plus size-independent properties at full size (sortedness, row_ptr consistency, linearity)."""
import numpy as np
import pytest

from bench import load_workload
from speck_b200.matrices import HostCSR
from helpers import assert_csr_equal, gpu_multiply, oracle_multiply

pytestmark = pytest.mark.gpu

ref = pytest.importorskip("oracle.ref")

WORKLOADS = ["rmat20", "webbase_like", "cant_like", "banded_like", "econ_like", "circuit_like", "rmat24"]


@pytest.mark.parametrize("name", WORKLOADS)
def test_fullsize_reference_vs_oracle_vs_ours(ctx, name):
    A = load_workload(name)
    want = oracle_multiply(A, A)
    ours, st = gpu_multiply(ctx, A)
    assert_csr_equal(ours, want, rtol=1e-6, what=f"ours vs oracle: {name}")
    assert st["nnz_c"] == want.nnz
    # size-independent properties of the result itself
    rp = ours.row_offsets.astype(np.int64)
    assert rp[0] == 0 and rp[-1] == ours.nnz and np.all(np.diff(rp) >= 0)
    d = np.diff(ours.col_ids.astype(np.int64))
    starts = rp[1:-1][(rp[1:-1] > 0) & (rp[1:-1] < ours.nnz)]
    d[starts - 1] = 1                      # a row boundary may step down
    assert np.all(d > 0), "columns must ascend strictly inside every row"
    del ours
    ctx.set_option("release_workspace", 1)
    checked = 0
    for variant in ("tuned", "stock"):
        if not ref.available(variant):
            continue
        r = ref.multiply(A, None, warmup=0, iters=1, variant=variant)
        got_ref = HostCSR(A.rows, want.cols, r["rp"], r["ci"], r["v"])
        assert_csr_equal(got_ref, want, rtol=1e-6, what=f"reference[{variant}] vs oracle: {name}")
        checked += 1
        if A.nnz > 10_000_000:             # one variant is enough for the 7-8 GB results
            break
    if checked == 0:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time); ours vs oracle checked")


def test_fullsize_linearity_in_values(ctx):
    """(2A).A == 2 (A.A) exactly in fp64 (scaling by a power of two commutes with rounding); same indices."""
    A = load_workload("webbase_like")
    C1, _ = gpu_multiply(ctx, A)
    A2 = HostCSR(A.rows, A.cols, A.row_offsets, A.col_ids, A.data * 2.0)
    dA2, dA = ctx.upload(A2), ctx.upload(A)
    dC = ctx.multiply(dA2, dA)
    C2 = ctx.download(dC)
    dC.free(), dA2.free(), dA.free()
    np.testing.assert_array_equal(C1.row_offsets, C2.row_offsets)
    np.testing.assert_array_equal(C1.col_ids, C2.col_ids)
    np.testing.assert_allclose(C2.data, 2.0 * C1.data, rtol=1e-12)
