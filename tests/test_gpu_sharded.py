"""GPU: the row-sharded path of SURVEY 8e through the CUDA library.  On one device the slabs are
multiplied one after the other (on N devices each rank does one, see bench.py --gpus N); their
concatenation must be bit-identical in indices to the unsharded product and to the oracle."""
import numpy as np
import pytest

from speck_b200 import api, matrices as M
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


@pytest.mark.parametrize("parts", [1, 3, 8])
def test_partition_rows_on_device_matches_host_cuts(ctx, parts):
    """speck_b200_partition_rows (analysis -> u64 scan -> search on the device) against the numpy definition."""
    A = M.rmat(13, 8, seed=5)
    dA = ctx.upload(A)
    cuts, prods = ctx.partition_rows(dA, dA, parts)
    dA.free()
    want = M.product_balanced_cuts(A, A.row_offsets, parts)
    np.testing.assert_array_equal(cuts.astype(np.int64), want)
    blen = np.diff(A.row_offsets.astype(np.int64))
    row_ops = np.add.reduceat(np.concatenate([blen[A.col_ids], [0]]), A.row_offsets[:-1].astype(np.int64))
    row_ops[np.diff(A.row_offsets.astype(np.int64)) == 0] = 0
    for g in range(parts):
        assert int(prods[g]) == int(row_ops[want[g]:want[g + 1]].sum())


def _plan_contexts(n):
    """n contexts: one per device when the box has n GPUs, else all on device 0 (the slab logic is the same)."""
    import torch
    from speck_b200 import api
    ndev = torch.cuda.device_count()
    return [api.Context(g if ndev >= n else 0) for g in range(n)], ndev >= n


@pytest.mark.parametrize("n", [2, 4])
def test_shard_plan_multiply_and_concat(n):
    """The C-level multi-device driver (speck_b200_sharded_*): host A/B in, slabs on the contexts' devices, B
    replicated by peer copies, concurrent slab multiplies, slabs concatenated on the first device."""
    from speck_b200 import api
    A = M.rmat(14, 8, seed=24)
    ctxs, real = _plan_contexts(n)
    try:
        plan = api.ShardPlan(ctxs, A)
        for _ in range(2):     # second call reuses the slabs of C
            info = plan.multiply()
        dC = plan.concat()
        got = ctxs[0].download(dC)
        want = oracle_multiply(A, A)
        assert_csr_equal(got, want, what=f"sharded plan x{n} ({'distinct devices' if real else 'one device'})")
        assert info["cuts"] == [int(x) for x in M.product_balanced_cuts(A, A.row_offsets, n)]
        assert sum(info["nnz_c"]) == want.nnz
        assert max(info["products"]) < 1.6 * sum(info["products"]) / n
        # slabs are visible one by one as well
        a0, c0 = plan.slab(0)
        assert a0.rows == info["cuts"][1] and c0.nnz == info["nnz_c"][0]
        dC.free()
        plan.close()
    finally:
        for c in ctxs:
            c.close()


def test_shard_plan_empty_slabs_and_rectangular():
    """More devices than non-empty rows: empty slabs must still concatenate to a well-formed CSR."""
    from speck_b200 import api
    A = M.uniform_random(5, 40, 3, seed=1)
    B = M.uniform_random(40, 60, 4, seed=2)
    ctxs, _ = _plan_contexts(8)
    try:
        plan = api.ShardPlan(ctxs, A, B)
        plan.multiply()
        got = ctxs[0].download(plan.concat())
        assert_csr_equal(got, oracle_multiply(A, B), what="8 slabs of a 5-row matrix")
        plan.close()
    finally:
        for c in ctxs:
            c.close()


def test_partition_rows_with_row_and_entry_costs(ctx):
    """partition_row_cost / partition_entry_cost: the cuts balance products + 2 * nnz(A row) + 8 per row (what
    bench.py --gpus N uses: tiny rows cost more than their products say)."""
    A = M.rmat(13, 4, seed=9)
    dA = ctx.upload(A)
    ctx.set_option("partition_row_cost", 8)
    ctx.set_option("partition_entry_cost", 2)
    try:
        cuts, cost = ctx.partition_rows(dA, dA, 4)
    finally:
        ctx.set_option("partition_row_cost", 0)
        ctx.set_option("partition_entry_cost", 0)
        dA.free()
    blen = np.diff(A.row_offsets.astype(np.int64))
    alen = np.diff(A.row_offsets.astype(np.int64))
    ops = np.zeros(A.rows, np.int64)
    nz = alen > 0
    ops[nz] = np.add.reduceat(blen[A.col_ids], A.row_offsets[:-1].astype(np.int64)[nz])
    c = ops + 2 * alen + 8
    prefix = np.concatenate([[0], np.cumsum(c)])
    want = np.concatenate([[0], np.searchsorted(prefix, (np.arange(1, 4) * prefix[-1]) // 4, side="left"), [A.rows]])
    np.testing.assert_array_equal(cuts.astype(np.int64), want)
    assert [int(x) for x in cost] == [int(prefix[want[g + 1]] - prefix[want[g]]) for g in range(4)]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_push_slab_concatenates_slabs(ctx, dtype):
    """speck_b200_push_slab_*: slabs of C pushed into the arrays of the concatenated C (here local memory; in the
    one-process-per-GPU bench the same kernel stores into IPC-opened peer memory).  Cuts chosen so that the nnz offsets
    are odd / even / multiples of 4 (16-byte store alignment of both arrays), two one-row slabs, one empty slab."""
    import ctypes
    A = M.rmat(12, 8, seed=31, dtype=dtype)
    C = oracle_multiply(A.astype(np.float64), A.astype(np.float64))
    C = HostCSR(C.rows, C.cols, C.row_offsets, C.col_ids, C.data.astype(dtype))
    rp = C.row_offsets.astype(np.int64)
    cuts = [0, 1, 2, 700, 700, 1501, 2222, C.rows]
    # make sure the offsets cover every residue mod 4 somewhere
    assert len({int(rp[c]) % 4 for c in cuts[1:-1]}) >= 3
    d_rp = ctx._alloc((C.rows + 1) * 4)
    d_ci = ctx._alloc(max(C.nnz, 1) * 4)
    d_v = ctx._alloc(max(C.nnz, 1) * np.dtype(dtype).itemsize)
    try:
        for g in range(len(cuts) - 1):
            r0, r1 = cuts[g], cuts[g + 1]
            slab = C.row_slice(r0, r1)
            last = g == len(cuts) - 2
            if slab.nnz == 0 and r1 == r0:   # empty slab: nothing was multiplied, device arrays are null
                d = api.DeviceCSR(ctx, dtype)
                d.s.rows = 0
                ms = ctx.push_slab(d, int(rp[r0]), r0, last, d_rp, d_ci, d_v)
            else:
                d = ctx.upload(slab)
                ms = ctx.push_slab(d, int(rp[r0]), r0, last, d_rp, d_ci, d_v)
                assert ms >= 0.0
                d.free()
        out = api.DeviceCSR.from_pointers(ctx, C.rows, C.cols, C.nnz, d_rp, d_ci, d_v, dtype)
        got = ctx.download(out)
        np.testing.assert_array_equal(got.row_offsets, C.row_offsets)
        np.testing.assert_array_equal(got.col_ids, C.col_ids)
        np.testing.assert_array_equal(got.data, C.data)
    finally:
        for p in (d_rp, d_ci, d_v):
            ctx.lib.speck_b200_free(ctx.h, p)
