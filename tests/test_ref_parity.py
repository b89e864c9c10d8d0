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
