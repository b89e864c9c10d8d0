"""GPU: the row-sharded path of SURVEY 8e through the CUDA library.  On one device the slabs are
multiplied one after the other (on N devices each rank does one, see bench.py --gpus N); their
concatenation must be bit-identical in indices to the unsharded product and to the oracle."""
import numpy as np
import pytest

from speck_b200 import matrices as M
from speck_b200.matrices import HostCSR
from speck_b200.sharding import concat_slabs
from helpers import assert_csr_equal, gpu_multiply, oracle_multiply

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("parts", [2, 8])
def test_sharded_equals_unsharded(ctx, parts):
    A = M.rmat(14, 8, seed=24)
    full, st = gpu_multiply(ctx, A)
    cuts = M.product_balanced_cuts(A, A.row_offsets, parts)
    dB = ctx.upload(A)
    slabs, prods = [], []
    for g in range(parts):
        S = A.row_slice(int(cuts[g]), int(cuts[g + 1]))
        dS = ctx.upload(S)
        dC = ctx.multiply(dS, dB)
        C = ctx.download(dC)
        prods.append(ctx.stats()["products"])
        slabs.append((C.row_offsets, C.col_ids, C.data))
        dC.free(), dS.free()
    dB.free()
    rp, ci, v = concat_slabs(slabs)
    got = HostCSR(A.rows, A.cols, rp, ci, v)
    assert_csr_equal(got, full, what="sharded vs unsharded")
    assert_csr_equal(got, oracle_multiply(A, A), what="sharded vs oracle")
    assert sum(prods) == st["products"]
    assert max(prods) < 1.6 * st["products"] / parts   # product-balanced cuts
