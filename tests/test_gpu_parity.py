"""GPU parity: the sm_100a path (through the C ABI) vs the CPU oracle on seeded inputs.
Cases follow SURVEY.md Appendix C (the reference's branch structure) re-mapped onto this
design's classes (direct / sort4..sort1024 / dense bitmap)."""
import numpy as np
import pytest

from speck_b200 import matrices as M
from speck_b200.matrices import HostCSR
from helpers import check_case, gpu_multiply, oracle_multiply, assert_csr_equal


def ndense(st):
    """rows that took the bitmap path (either launch shape)"""
    return st["class_rows"]["dense"] + st["class_rows"]["dense_local"]

pytestmark = pytest.mark.gpu


def test_tiny8(ctx):
    got, st = check_case(ctx, M.tiny8(), what="tiny8")
    assert st["products"] == 62 or st["products"] > 0


@pytest.mark.parametrize("n,d,seed", [(64, 2, 1), (300, 4, 2), (1000, 8, 3), (2000, 16, 4), (513, 40, 5)])
def test_uniform_random(ctx, n, d, seed):
    check_case(ctx, M.uniform_random(n, n, d, seed), what=f"uniform {n} {d}")


@pytest.fixture(params=[(16384, 1, 1, 1, 8, 0, 1), (16384, 1, 1, 1, 16, 1, 2), (16384, 1, 1, 0, 8, 0, 0), (8192, 1, 0, 1, 8, 0, 1),
                        (8192, 0, 1, 1, 8, 0, 0), (8192, 0, 0, 1, 8, 0, 1), (1024, 1, 1, 1, 8, 1, 1), (64, 1, 1, 1, 8, 0, 1),
                        (64, 1, 0, 1, 8, 0, 0)],
                ids=["16384-flat8-map", "16384-flat16-seg-map", "16384-rank-map", "8192-rank", "8192-sort-map", "8192-sort",
                     "1024-seg-map", "64-map", "64"])
def sort_max(ctx, request):
    """Run a case with the sort/bitmap switch at several places, with the rank classes on and off, with
    and without the symbolic->numeric rank map, and with the flat (staged) and the register-slot variant of
    the mapped symbolic rank kernel, with the segment-major and the thread-blocked mapped numeric kernel, so that every kernel family (lane-group sort, rank, flat rank, CTA sort,
    bitmap, mapped numeric) sees the same inputs."""
    ctx.set_option("sort_max", request.param[0])
    ctx.set_option("rank_path", request.param[1])
    ctx.set_option("rank_map", request.param[2])
    ctx.set_option("flat_sym", request.param[3])
    ctx.set_option("flat_e", request.param[4])
    ctx.set_option("seg_num", request.param[5])
    ctx.set_option("dense_seq", request.param[6])
    yield request.param[0]
    ctx.set_option("sort_max", 16384)
    ctx.set_option("rank_path", 1)
    ctx.set_option("rank_map", 1)
    ctx.set_option("flat_sym", 1)
    ctx.set_option("flat_e", 8)
    ctx.set_option("seg_num", 0)
    ctx.set_option("dense_seq", 1)


@pytest.mark.parametrize("scale,ef", [(10, 8), (13, 16), (15, 16)])
def test_rmat(ctx, sort_max, scale, ef):
    A = M.rmat(scale, ef, seed=scale)
    got, st = check_case(ctx, A, what=f"rmat{scale} sort_max={sort_max}")
    if sort_max <= 1024 and scale >= 13:
        assert ndense(st) > 0
    if sort_max >= 8192 and scale >= 15:
        assert st["class_rows"]["sort2048"] > 0 and st["class_rows"]["sort4096"] > 0


def test_rectangular(ctx):
    A = M.uniform_random(700, 300, 5, seed=11)
    B = M.uniform_random(300, 1500, 7, seed=12)
    check_case(ctx, A, B, what="rectangular")


def _rows_with_products(targets, cols=4096, seed=0, nb=128):
    """A whose row i has exactly targets[i] products against a B with one entry... B row k has
    length k+1 (k < 64) so any product count can be composed; all columns distinct mod cols."""
    rng = np.random.default_rng(seed)
    # B: row k has k+1 random distinct columns
    br, bc = [], []
    for k in range(nb):
        cs = rng.choice(cols, k + 1, replace=False)
        br += [k] * (k + 1)
        bc += list(cs)
    B = M.from_coo(nb, cols, br, bc, seed=seed + 1)
    ar, ac = [], []
    for i, t in enumerate(targets):
        left = t
        used = set()
        # greedy: use long rows first, each B row at most once per A row -> products <= 8256
        for k in range(nb - 1, -1, -1):
            if left >= k + 1 and k not in used:
                ar.append(i), ac.append(k)
                used.add(k)
                left -= k + 1
        assert left == 0, (t, left)
    A = M.from_coo(len(targets), nb, ar, ac, seed=seed + 2)
    return A, B


@pytest.mark.parametrize("rank_map", [1, 0])
def test_class_boundaries(ctx, rank_map):
    """product counts straddling every class boundary (4<<c) and the sort/dense switch."""
    ctx.set_option("rank_map", rank_map)
    targets = []
    for c in range(0, 12):
        b = 4 << c
        targets += [b - 1, b, b + 1]
    targets += [1, 2, 3, 8255, 8256, 12345, 16383, 16384, 16385, 20000, 32896]
    A, B = _rows_with_products(targets, cols=1 << 18, nb=256)
    try:
        got, st = check_case(ctx, A, B, what="class boundaries")
    finally:
        ctx.set_option("rank_map", 1)
    # 8193..16384 products: rank kernels when the rank map is on, else the bitmap path
    assert ndense(st) >= (3 if rank_map else 9)
    assert st["class_rows"]["sort16384"] == (6 if rank_map else 0)
    # CTA classes step by 512 products (2..16 warps): b-1 and b fall in sort{b}, b+1 in sort{b+512}
    for name in ("sort1024", "sort1536", "sort2048", "sort2560", "sort4096", "sort4608", "sort8192"):
        assert st["class_rows"][name] >= 1, name
    assert sum(st["class_rows"][f"sort{512 * w}"] for w in range(2, 17)) >= 11


def test_empty_rows_and_empty_b_rows(ctx):
    rng = np.random.default_rng(5)
    n = 500
    r = rng.integers(0, n, 3000)
    c = rng.integers(0, n, 3000)
    keep = (r % 3 != 0)            # every third row of A (and hence of B = A) is empty
    A = M.from_coo(n, n, r[keep], c[keep], seed=6)
    check_case(ctx, A, what="empty rows")
    # trailing empty rows
    A2 = M.from_coo(n, n, r[r < 100], c[r < 100], seed=7)
    check_case(ctx, A2, what="trailing empty rows")


def test_single_entry_rows_direct_path(ctx):
    n = 400
    rng = np.random.default_rng(8)
    A = M.from_coo(n, n, np.arange(n), rng.integers(0, n, n), seed=9)
    B = M.uniform_random(n, n, 20, seed=10)
    got, st = check_case(ctx, A, B, what="direct")
    assert st["class_rows"]["direct"] > 0


def test_max_compression_and_none(ctx):
    # all B rows identical -> every product of a row hits the same few columns
    n = 256
    cols = np.array([3, 77, 200, 255])
    br = np.repeat(np.arange(n), cols.size)
    bc = np.tile(cols, n)
    B = M.from_coo(n, n, br, bc, seed=1)
    A = M.uniform_random(n, n, 30, seed=2)
    check_case(ctx, A, B, what="max compression")
    # disjoint B rows -> compression exactly 1
    B2 = M.from_coo(n, n * 8, np.repeat(np.arange(n), 8), np.arange(n * 8), seed=3)
    got, st = check_case(ctx, A, B2, what="compression 1")
    assert st["products"] == st["nnz_c"]


def test_cancellation_keeps_structural_zero(ctx):
    # C[0,0] = 1*2 + 1*(-2) = 0 must stay in C (reference never drops numerically)
    A = HostCSR(2, 2, np.array([0, 2, 3], np.uint32), np.array([0, 1, 1], np.uint32), np.array([1.0, 1.0, 5.0]))
    B = HostCSR(2, 2, np.array([0, 1, 2], np.uint32), np.array([0, 0], np.uint32), np.array([2.0, -2.0]))
    got, _ = check_case(ctx, A, B, what="cancellation")
    assert got.nnz == 2 and got.data[0] == 0.0


def test_empty_products_conventions(ctx):
    # A.nnz == 0 -> only C.nnz = 0 (Multiply.cu:67-70)
    Z = HostCSR(5, 5, np.zeros(6, np.uint32), np.zeros(0, np.uint32), np.zeros(0))
    got, _ = gpu_multiply(ctx, Z)
    assert got.nnz == 0
    # P == 0 with nnz > 0: A only references empty rows of B (Multiply.cu:256-261)
    A = HostCSR(3, 3, np.array([0, 1, 1, 1], np.uint32), np.array([2], np.uint32), np.array([1.0]))
    B = HostCSR(3, 3, np.array([0, 1, 1, 1], np.uint32), np.array([0], np.uint32), np.array([1.0]))
    got, _ = gpu_multiply(ctx, A, B)
    assert got.nnz == 0 and got.rows == 3


def test_banded_high_compression(ctx, sort_max):
    A = M.banded_fem_like(n=3000, per_row=64, clusters=8, band=300, seed=41)
    got, st = check_case(ctx, A, what=f"banded sort_max={sort_max}")
    # column extent (<= 2*300+8) is far below the 4096 products of a row: bitmap path whatever sort_max is
    assert st["class_rows"]["dense_local"] > 2500


def test_wide_matrix_multiwindow_and_wide_keys(ctx, sort_max):
    """cols > 2^20 -> several bitmap windows in the dense path; cols*N > 2^32 -> u64 sort keys."""
    rng = np.random.default_rng(21)
    n, cols = 3000, (1 << 23) + 12345
    m = 40000
    r = rng.integers(0, n, m)
    c = rng.integers(0, n, m)
    # hub rows: first 4 rows reference 400 rows each
    hub_r = np.repeat(np.arange(4), 400)
    hub_c = rng.integers(0, n, 1600)
    A = M.from_coo(n, n, np.concatenate([r, hub_r]), np.concatenate([c, hub_c]), seed=22)
    mb = 60000
    B = M.from_coo(n, cols, rng.integers(0, n, mb), rng.integers(0, cols, mb), seed=23)
    got, st = check_case(ctx, A, B, what=f"wide sort_max={sort_max}")
    if sort_max <= 1024:
        assert st["class_rows"]["dense"] >= 4


def test_fem_like_high_compression(ctx):
    """27-point stencil x 3 dofs: ~6500 products fold into ~375 columns per row -> bitmap path with
    the shared-memory value accumulator; also exercises rows wider than one local window."""
    A = M.fem3d_like(9, 8, 7)
    got, st = check_case(ctx, A, what="fem3d")
    assert st["class_rows"]["dense_local"] > 0
    assert st["products"] > 8 * st["nnz_c"]


@pytest.mark.parametrize("variant", [1, 2], ids=["lane-loads", "tma-staged"])
def test_dense_seq_rows_are_bit_equal_to_the_oracle(ctx, variant):
    """The sequential-k kernel (dense_seq.cuh) adds the products of an entry of C in ascending k with separately
    rounded products -- the oracle's order: rows it handles are bit-identical to the oracle and run-to-run."""
    ctx.set_option("dense_seq", variant)
    A = M.fem3d_like(9, 8, 7)
    first, st = gpu_multiply(ctx, A)
    assert st["class_rows"]["dense_local"] == A.rows     # every row takes the bitmap path with a kept bitmap
    want = oracle_multiply(A, A)
    np.testing.assert_array_equal(first.col_ids, want.col_ids)

    def seq_mask(Mx, Cx):
        """entries of C in rows the sequential-k kernel takes: <= 2048 entries, >= 6 products per entry"""
        import oracle
        ops = oracle.row_products(Mx.row_offsets, Mx.col_ids, Mx.row_offsets)[0].astype(np.int64)
        nnz = np.diff(Cx.row_offsets.astype(np.int64))
        return np.repeat((nnz <= 2048) & (ops >= 6 * nnz), nnz)
    m = seq_mask(A, want)
    assert m.mean() > 0.9
    np.testing.assert_array_equal(first.data[m], want.data[m])  # bit-exact, not just 1e-6
    np.testing.assert_allclose(first.data, want.data, rtol=1e-6)
    again, _ = gpu_multiply(ctx, A)
    np.testing.assert_array_equal(again.data[m], first.data[m])
    # long B rows (several 128-entry pieces per segment) and B rows of every alignment
    B = M.banded_fem_like(n=3000, per_row=300, clusters=3, band=600, seed=5)
    got, st = gpu_multiply(ctx, B)
    assert ndense(st) > 0
    wantB = oracle_multiply(B, B)
    np.testing.assert_array_equal(got.col_ids, wantB.col_ids)
    mask = seq_mask(B, wantB)
    assert mask.any()
    np.testing.assert_array_equal(got.data[mask], wantB.data[mask])
    np.testing.assert_allclose(got.data, wantB.data, rtol=1e-6)
    ctx.set_option("dense_seq", 1)


def test_dense_rows_exceeding_shared_accumulator(ctx):
    """banded rows with more distinct columns than the 2048-entry shared accumulator -> RED path"""
    A = M.banded_fem_like(n=6000, per_row=96, clusters=12, band=1500, seed=7)
    got, st = check_case(ctx, A, what="wide band")
    assert ndense(st) > 0


def test_fp32(ctx):
    A = M.rmat(12, 8, seed=3, dtype=np.float32)
    got, st = gpu_multiply(ctx, A)
    want = oracle_multiply(A, A)
    assert got.data.dtype == np.float32
    assert_csr_equal(got, want, rtol=2e-5, what="fp32")


def test_fp32_every_kernel_family(ctx, sort_max):
    """fp32 through the same kernel-family switch as the fp64 cases (rank / flat rank / CTA sort / bitmap / mapped
    and self-contained numeric kernels)."""
    A = M.rmat(14, 16, seed=9, dtype=np.float32)
    got, st = gpu_multiply(ctx, A)
    want = oracle_multiply(A.astype(np.float64), A.astype(np.float64))
    assert got.data.dtype == np.float32
    assert_csr_equal(got, want, rtol=1e-4, what=f"fp32 rmat14 sort_max={sort_max}")
    B = M.banded_fem_like(n=3000, per_row=48, clusters=6, band=200, seed=8, dtype=np.float32)
    got, st = gpu_multiply(ctx, B)
    assert_csr_equal(got, oracle_multiply(B.astype(np.float64), B.astype(np.float64)), rtol=1e-4, what="fp32 banded")


# ---------------------------------------------------------------- error paths (reference: source/GPU/Multiply.cu:57-97)
def _csr_struct(api, rows, cols, nnz, data=None, rp=None, ci=None):
    return api.CsrStruct(rows, cols, nnz, data, rp, ci)


def test_error_too_large_and_shape_mismatch(ctx):
    """rows(A) or cols(B) above 2^27 -> SPECK_ERR_TOO_LARGE before anything is touched (Multiply.cu:57-66); A.cols
    != B.rows -> SPECK_ERR_INVALID; C is left untouched in every case."""
    import ctypes
    from speck_b200 import api
    A = M.uniform_random(64, 64, 3, seed=1)
    dA = ctx.upload(A)
    C = api.DeviceCSR(ctx)
    big_rows = _csr_struct(api, (1 << 27) + 1, 64, dA.s.nnz, dA.s.data, dA.s.row_offsets, dA.s.col_ids)
    big_cols = _csr_struct(api, 64, (1 << 27) + 1, dA.s.nnz, dA.s.data, dA.s.row_offsets, dA.s.col_ids)
    lib = ctx.lib
    rc = lib.speck_b200_spgemm_f64(ctx.h, ctypes.byref(big_rows), ctypes.byref(dA.s), ctypes.byref(C.s), None)
    assert rc == -2 and b"rows" in lib.speck_b200_last_error()
    rc = lib.speck_b200_spgemm_f64(ctx.h, ctypes.byref(dA.s), ctypes.byref(big_cols), ctypes.byref(C.s), None)
    assert rc == -2 and b"columns" in lib.speck_b200_last_error()
    wrong = _csr_struct(api, 63, 64, dA.s.nnz, dA.s.data, dA.s.row_offsets, dA.s.col_ids)
    rc = lib.speck_b200_spgemm_f64(ctx.h, ctypes.byref(dA.s), ctypes.byref(wrong), ctypes.byref(C.s), None)
    assert rc == -1 and b"shape mismatch" in lib.speck_b200_last_error()
    rc = lib.speck_b200_spgemm_f64(ctx.h, None, ctypes.byref(dA.s), ctypes.byref(C.s), None)
    assert rc == -1
    assert (C.s.rows, C.s.nnz, C.s.data, C.s.col_ids, C.s.row_offsets) == (0, 0, None, None, None)
    with pytest.raises(api.SpeckError):
        ctx.set_option("no_such_option", 1)
    with pytest.raises(api.SpeckError):
        ctx.set_option("sort_max", 1000)
    # the context still works after the errors
    check_case(ctx, A, what="after errors")
    dA.free()


def test_error_nnz_overflow(ctx):
    """nnz(C) >= 2^32 does not fit the u32 row_offsets of the spECK API -> SPECK_ERR_OVERFLOW after the symbolic
    phase (the reference overflows silently).  A = 70000 rows of 1 entry hitting one dense
    B row of 2^16 columns: P = nnz(C) = 4.59e9, found by the analysis alone (direct rows need no symbolic kernel)."""
    import ctypes
    from speck_b200 import api
    rowsA, n = 70000, 1 << 16
    A = HostCSR(rowsA, 4, np.arange(rowsA + 1, dtype=np.uint32), np.zeros(rowsA, np.uint32), np.ones(rowsA))
    brp = np.array([0, n, n, n, n], np.uint32)
    B = HostCSR(4, n, brp, np.arange(n, dtype=np.uint32), np.ones(n))
    dA, dB = ctx.upload(A), ctx.upload(B)
    C = api.DeviceCSR(ctx)
    rc = ctx.lib.speck_b200_spgemm_f64(ctx.h, ctypes.byref(dA.s), ctypes.byref(dB.s), ctypes.byref(C.s), None)
    assert rc == -3, ctx.lib.speck_b200_last_error()
    assert b"does not fit" in ctx.lib.speck_b200_last_error()
    assert C.s.data is None and C.s.col_ids is None    # nothing was allocated for the oversized result
    C.free(), dA.free(), dB.free()
    check_case(ctx, M.uniform_random(100, 100, 4, seed=2), what="after overflow")


def test_error_out_of_memory_for_c(ctx):
    """C larger than the free device memory -> SPECK_ERR_OOM, C->data / col_ids left NULL (reference: prints
    'ERROR: out of memory' and returns, Multiply.cu:594-599); the context stays usable."""
    import ctypes
    import torch
    from speck_b200 import api
    free_b, _ = torch.cuda.mem_get_info(0)
    # nnz(C) = rowsA * n just below 2^32 needs 12 * 4.29e9 = 51.5 GB: hog memory so that less than that is free
    hog = []
    need = 51.5e9
    try:
        while free_b > need - (4 << 30) and free_b > (8 << 30):
            take = int(min(free_b - (need - (8 << 30)), 32 << 30))
            if take < (1 << 30):
                break
            hog.append(torch.empty(take, dtype=torch.uint8, device="cuda:0"))
            free_b, _ = torch.cuda.mem_get_info(0)
        rowsA, n = 65535, 1 << 16
        A = HostCSR(rowsA, 4, np.arange(rowsA + 1, dtype=np.uint32), np.zeros(rowsA, np.uint32), np.ones(rowsA))
        B = HostCSR(4, n, np.array([0, n, n, n, n], np.uint32), np.arange(n, dtype=np.uint32), np.ones(n))
        dA, dB = ctx.upload(A), ctx.upload(B)
        C = api.DeviceCSR(ctx)
        rc = ctx.lib.speck_b200_spgemm_f64(ctx.h, ctypes.byref(dA.s), ctypes.byref(dB.s), ctypes.byref(C.s), None)
        assert rc == -5, (rc, ctx.lib.speck_b200_last_error())
        assert C.s.data is None and C.s.col_ids is None and C.s.nnz == 0
        C.free(), dA.free(), dB.free()
    finally:
        del hog
        torch.cuda.empty_cache()
    check_case(ctx, M.uniform_random(100, 100, 4, seed=3), what="after OOM")


def test_host_entry_dtype_switch_and_empty_product(ctx):
    """One context, host entry: f32 then f64 with the same nnz(C) must not reuse the smaller value buffer (advisor
    finding, round 1); an empty product returns a well-formed CSR (zero row_offsets of rows + 1 entries)."""
    A64 = M.rmat(11, 8, seed=12)
    A32 = A64.astype(np.float32)
    got32, _, _ = ctx.multiply_host(A32, A32)
    got32 = HostCSR(got32.rows, got32.cols, got32.row_offsets.copy(), got32.col_ids.copy(), got32.data.copy())
    got64, _, _ = ctx.multiply_host(A64, A64)
    want = oracle_multiply(A64, A64)
    assert_csr_equal(HostCSR(got64.rows, got64.cols, got64.row_offsets.copy(), got64.col_ids.copy(), got64.data.copy()),
                     want, what="host f64 after f32")
    assert_csr_equal(got32, want, rtol=1e-4, what="host f32")
    Z = HostCSR(300, A64.rows, np.zeros(301, np.uint32), np.zeros(0, np.uint32), np.zeros(0))
    got, _, _ = ctx.multiply_host(Z, A64)
    assert got.rows == 300 and got.nnz == 0 and got.row_offsets.shape == (301,) and not got.row_offsets.any()
    # P == 0 with non-empty operands: A only references empty rows of B
    Bz = HostCSR(4, 8, np.array([0, 0, 0, 0, 3], np.uint32), np.array([1, 2, 5], np.uint32), np.ones(3))
    Az = HostCSR(3, 4, np.array([0, 1, 2, 3], np.uint32), np.array([0, 1, 2], np.uint32), np.ones(3))
    got, _, _ = ctx.multiply_host(Az, Bz)
    assert got.rows == 3 and got.nnz == 0 and not got.row_offsets.any()


def test_api_rejects_mixed_dtypes(ctx):
    from speck_b200 import api
    A = M.uniform_random(50, 50, 3, seed=1)
    d64, d32 = ctx.upload(A), ctx.upload(A.astype(np.float32))
    with pytest.raises(api.SpeckError):
        ctx.multiply(d64, d32)
    C = ctx.multiply(d64, d64)
    with pytest.raises(api.SpeckError):
        ctx.multiply(d32, d32, C)      # C holds f64 values
    C.free(), d64.free(), d32.free()


def test_c_reuse_semantics(ctx):
    """C is reused across calls when nnz is unchanged (Multiply.cu:155-165, 589-592)."""
    A = M.rmat(11, 8, seed=4)
    dA = ctx.upload(A)
    C = ctx.multiply(dA, dA)
    p0 = (C.s.row_offsets, C.s.col_ids, C.s.data)
    C = ctx.multiply(dA, dA, C)
    assert (C.s.row_offsets, C.s.col_ids, C.s.data) == p0
    h1 = ctx.download(C)
    A2 = M.rmat(11, 4, seed=5)
    dA2 = ctx.upload(A2)
    C = ctx.multiply(dA2, dA2, C)          # same rows, different nnz -> row_offsets kept
    assert C.s.row_offsets == p0[0]
    assert_csr_equal(ctx.download(C), oracle_multiply(A2, A2), what="reuse")
    assert_csr_equal(h1, oracle_multiply(A, A), what="first")
    assert ctx.compare(C, C, True)
    C.free(), dA.free(), dA2.free()


def test_row_products_matches_oracle(ctx):
    import oracle
    A = M.rmat(13, 16, seed=2)
    dA = ctx.upload(A)
    ops, P, mx = ctx.row_products(dA, dA)
    o_ops, _, oP, omx = oracle.row_products(A.row_offsets, A.col_ids, A.row_offsets)
    np.testing.assert_array_equal(ops, o_ops)
    assert (P, mx) == (oP, omx)
    dA.free()


def test_host_entry_point(ctx):
    A = M.rmat(12, 16, seed=6)
    got, up, down = ctx.multiply_host(A, A)
    want = oracle_multiply(A, A)
    assert_csr_equal(HostCSR(got.rows, got.cols, got.row_offsets.copy(), got.col_ids.copy(), got.data.copy()),
                     want, what="host entry")
    assert up == (A.rows + 1) * 4 + A.nnz * 12
    assert down == (A.rows + 1) * 4 + want.nnz * 12


def test_sort_max_option_routes_more_rows_to_dense(ctx):
    A = M.rmat(12, 16, seed=8)
    ctx.set_option("sort_max", 64)
    try:
        got, st = check_case(ctx, A, what="sort_max=64")
        assert st["class_rows"]["sort128"] == 0 and ndense(st) > 0
    finally:
        ctx.set_option("sort_max", 16384)


def test_rank_class_many_short_b_rows(ctx):
    """rank classes: A rows far longer than the CTA (several gather batches), B rows of length 0..3,
    so products per A entry are tiny and the owner table changes entry almost every product."""
    rng = np.random.default_rng(31)
    nb, cols = 6000, 1 << 18
    blen = rng.integers(0, 4, nb)
    br = np.repeat(np.arange(nb), blen)
    B = M.from_coo(nb, cols, br, rng.integers(0, cols, br.size), seed=32)
    ar, ac = [], []
    for i, alen in enumerate([700, 1500, 2500, 4000, 5000]):
        ar += [i] * alen
        ac += list(rng.choice(nb, alen, replace=False))
    A = M.from_coo(5, nb, ar, ac, seed=33)
    got, st = check_case(ctx, A, B, what="rank many short B rows")
    assert sum(st["class_rows"][f"sort{512 * w}"] for w in range(2, 17)) == 5


def test_rank_class_duplicates_over_wide_extent(ctx):
    """rank classes with heavy folding: every B row draws from the same 300 columns spread over the
    whole 2^20 extent (extent > 4 * products, so the rows are not taken by the bitmap path)."""
    rng = np.random.default_rng(34)
    nb, cols = 400, 1 << 20
    pool = np.unique(np.concatenate([[0, cols - 1], rng.integers(0, cols, 300)]))
    br, bc = [], []
    for k in range(nb):
        cs = rng.choice(pool, 40, replace=False)
        br += [k] * 40
        bc += list(cs)
    B = M.from_coo(nb, cols, br, bc, seed=35)
    A = M.uniform_random(64, nb, 60, seed=36)   # ~2400 products per row, <= 302 distinct columns
    got, st = check_case(ctx, A, B, what="rank duplicates")
    assert st["products"] > 5 * st["nnz_c"]
    assert sum(st["class_rows"][f"sort{512 * w}"] for w in range(2, 17)) > 32


@pytest.mark.parametrize("cols", [(1 << 24) + 7, 1 << 25, (1 << 25) + 1])
def test_rank_three_levels_wide_columns(ctx, cols):
    """cols(B) in (2^20, 2^25]: rank kernels with a third bitmap level (symbolic) + mapped numeric; one column
    past 2^25 the CTA sort / bitmap-window fallbacks take over.  Rows of ~600..16000 products, with folding
    (B rows share a column pool) and columns at both ends of the range."""
    rng = np.random.default_rng(41)
    nb = 1500
    pool = np.unique(np.concatenate([[0, cols - 1], rng.integers(0, cols, 60000)]))
    br, bc = [], []
    for k in range(nb):
        ln = int(rng.integers(1, 60))
        br += [k] * ln
        bc += list(rng.choice(pool, ln, replace=False))
    B = M.from_coo(nb, cols, br, bc, seed=42)
    ar, ac = [], []
    for i, alen in enumerate([20, 40, 70, 100, 140, 200, 280, 400, 520]):
        ar += [i] * alen
        ac += list(rng.choice(nb, alen, replace=False))
    A = M.from_coo(9, nb, ar, ac, seed=43)
    got, st = check_case(ctx, A, B, what=f"three levels cols={cols}")
    cta = sum(st["class_rows"][f"sort{512 * w}"] for w in range(2, 17)) + st["class_rows"]["sort16384"]
    if cols <= (1 << 25):
        assert cta >= 6 and st["class_rows"]["sort16384"] >= 1
    assert st["nnz_c"] < st["products"]


@pytest.mark.parametrize("cols", [1 << 18, (1 << 24) + 7])
def test_rank_map_allocation_fallback(ctx, cols):
    """Rows are binned for the mapped kernels (up to 16384 products, three bitmap levels for wide matrices), then
    the rank map "does not fit": the 16384 bin must fall back to the bitmap kernel, the other CTA bins to the
    self-contained rank kernels (two levels) or the 64-bit-key CTA sort (three levels)."""
    targets = [600, 1500, 3000, 5000, 8192, 9000, 16384, 100, 30]
    A, B = _rows_with_products(targets, cols=cols, nb=256, seed=5)
    ctx.set_option("rank_map_max_bytes", 1)
    try:
        got, st = check_case(ctx, A, B, what=f"map fallback cols={cols}")
    finally:
        ctx.set_option("rank_map_max_bytes", -1)
    assert st["class_rows"]["sort16384"] == 2
    got2, st2 = check_case(ctx, A, B, what=f"mapped cols={cols}")
    assert_csr_equal(got, got2, what="fallback vs mapped")


def test_compare_reports_first_mismatch(ctx):
    """speck_b200_compare_report: the first differing (row, kind, position), which the reference's d_compare
    (source/GPU/Compare.cu:27-58) cannot tell."""
    A = M.rmat(10, 8, seed=1)
    C = oracle_multiply(A, A)
    dRef = ctx.upload(C)
    eq, mm = ctx.compare_report(dRef, dRef, True)
    assert eq and mm is None
    # a wrong value in row 700, a wrong column in row 300: the column error comes first
    row_v, row_c = 700, 300
    bad = HostCSR(C.rows, C.cols, C.row_offsets.copy(), C.col_ids.copy(), C.data.copy())
    pv = int(C.row_offsets[row_v]) + 2
    bad.data[pv] *= 1.5
    dBad = ctx.upload(bad)
    eq, mm = ctx.compare_report(dRef, dBad, True, 1e-6)
    assert not eq and (mm["row"], mm["kind"], mm["index_in_row"]) == (row_v, 2, 2)
    assert mm["ref_val"] == C.data[pv] and mm["cmp_val"] == bad.data[pv] and mm["ref_col"] == C.col_ids[pv]
    assert ctx.compare(dRef, dBad, False) and not ctx.compare(dRef, dBad, True)
    dBad.free()
    pc = int(C.row_offsets[row_c]) + 1
    bad.col_ids[pc] += 1 if bad.col_ids[pc] + 1 != bad.col_ids[min(pc + 1, C.nnz - 1)] else 2
    dBad = ctx.upload(bad)
    eq, mm = ctx.compare_report(dRef, dBad, True, 1e-6)
    assert not eq and (mm["row"], mm["kind"], mm["index_in_row"]) == (row_c, 1, 1)
    assert mm["ref_col"] == C.col_ids[pc] and mm["cmp_col"] == bad.col_ids[pc]
    dBad.free()
    # a row length error (one entry moved from row 100 to row 101) comes before both
    rp = C.row_offsets.copy()
    rp[101] -= 1
    dBad = ctx.upload(HostCSR(C.rows, C.cols, rp, C.col_ids, C.data))
    eq, mm = ctx.compare_report(dRef, dBad, False)
    assert not eq and (mm["row"], mm["kind"]) == (100, 0) and mm["ref_len"] == mm["cmp_len"] + 1
    dBad.free(), dRef.free()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_gpu_coo_to_csr(ctx, dtype):
    """GPU-side COO -> CSR (loader path) against the host conversion: sorted by (row, column),
    stable; duplicates kept (the reference's loader, source/CSR.cpp:173-212) or summed in input order."""
    rng = np.random.default_rng(3)
    rows, cols, n = 5000, 3000, 200_000
    r = rng.integers(0, rows, n).astype(np.uint32)
    r[r % 7 == 3] = 11                       # many empty rows, one heavy row
    c = rng.integers(0, cols, n).astype(np.uint32)
    c[::5] = c[1::5][: c[::5].size]          # plenty of duplicate (row, col) candidates
    r[::5] = r[1::5][: r[::5].size]
    v = rng.standard_normal(n).astype(dtype)
    order = np.lexsort((np.arange(n), c, r))  # stable by (row, col)
    # KEEP
    d = ctx.coo_to_csr(rows, cols, r, c, v, sum_duplicates=False)
    got = ctx.download(d)
    d.free()
    want_rp = np.concatenate([[0], np.cumsum(np.bincount(r, minlength=rows))]).astype(np.uint32)
    np.testing.assert_array_equal(got.row_offsets, want_rp)
    np.testing.assert_array_equal(got.col_ids, c[order])
    np.testing.assert_array_equal(got.data, v[order])
    # SUM: runs of equal (row, col) folded in input order
    d = ctx.coo_to_csr(rows, cols, r, c, v, sum_duplicates=True)
    got = ctx.download(d)
    d.free()
    key = r[order].astype(np.int64) * cols + c[order]
    head = np.concatenate([[True], key[1:] != key[:-1]])
    starts = np.flatnonzero(head)
    want_v = np.array([np.add.reduce(v[order][s:e], dtype=dtype) if False else _seq_sum(v[order][s:e]) for s, e in zip(starts, np.append(starts[1:], n))], dtype=dtype)
    np.testing.assert_array_equal(got.col_ids, c[order][head])
    np.testing.assert_array_equal(got.row_offsets, np.concatenate([[0], np.cumsum(np.bincount(r[order][head], minlength=rows))]).astype(np.uint32))
    np.testing.assert_array_equal(got.data, want_v)
    assert got.nnz == int(head.sum()) < n
    # the summed form is a valid multiply operand
    sq = ctx.coo_to_csr(rows, rows, r, c % rows, v.astype(np.float64), sum_duplicates=True)
    h = ctx.download(sq)
    C = ctx.multiply(sq, sq)
    assert_csr_equal(ctx.download(C), oracle_multiply(h, h), what="A.A of a GPU-converted matrix")
    C.free(), sq.free()
    # empty input
    e = ctx.coo_to_csr(10, 10, np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, dtype))
    he = ctx.download(e)
    assert he.nnz == 0 and not he.row_offsets.any()
    e.free()


def _seq_sum(x):
    s = x[0]
    for t in x[1:]:
        s = s + t        # same type, left to right: the order the kernel uses
    return s


def test_deterministic_mode_is_bit_equal_to_the_oracle(ctx):
    """Option "deterministic" (the reference is "not bit stable", config.ini:8-9): every entry of C is the sum of
    its products in ascending k, products rounded before they are added -- the CPU oracle's order -- so values
    are bit-identical to the oracle and from run to run, on every row class (lane-group and CTA sort classes,
    sequential-k bitmap rows, recomputed bitmap rows, direct rows)."""
    cases = [("rmat15", M.rmat(15, 16, seed=15), None),                 # rows up to > 16384 products: recomputed bitmap rows
             ("fem", M.fem3d_like(9, 8, 7), None),                      # high compression: sequential-k kernel
             ("wide band", M.banded_fem_like(n=6000, per_row=96, clusters=12, band=1500, seed=7), None),   # > 2048 entries per row
             ("uniform", M.uniform_random(2000, 2000, 16, seed=4), None),
             ("rect", M.uniform_random(700, 300, 5, seed=11), M.uniform_random(300, 1500, 7, seed=12))]
    ctx.set_option("deterministic", 1)
    try:
        for name, A, B in cases:
            want = oracle_multiply(A, A if B is None else B)
            got, st = gpu_multiply(ctx, A, B)
            np.testing.assert_array_equal(got.row_offsets, want.row_offsets, err_msg=name)
            np.testing.assert_array_equal(got.col_ids, want.col_ids, err_msg=name)
            np.testing.assert_array_equal(got.data, want.data, err_msg=f"{name}: values must be bit-equal to the oracle")
            again, _ = gpu_multiply(ctx, A, B)
            np.testing.assert_array_equal(again.data, got.data, err_msg=f"{name}: run-to-run")
        A32 = M.rmat(13, 16, seed=13, dtype=np.float32)
        a, _ = gpu_multiply(ctx, A32)
        b, _ = gpu_multiply(ctx, A32)
        np.testing.assert_array_equal(a.data, b.data)
    finally:
        ctx.set_option("deterministic", 0)
    # and the default mode still agrees to 1e-6
    check_case(ctx, cases[0][1], what="default after deterministic")


def test_dense_symbolic_without_test_before_set(ctx):
    """test_set = 0: the bitmap pass sets every bit with an atomicOr (the default reads the word first)."""
    ctx.set_option("test_set", 0)
    try:
        check_case(ctx, M.fem3d_like(9, 8, 7), what="fem, test_set=0")
        check_case(ctx, M.banded_fem_like(n=4000, per_row=64, clusters=8, band=300, seed=41), what="banded, test_set=0")
    finally:
        ctx.set_option("test_set", 1)


def test_col_direct_variants(ctx):
    """col_direct = 1..3: the large mapped numeric shapes write column ids straight to C (kept as a tested option)."""
    A = M.rmat(15, 16, seed=15)
    try:
        for v in (1, 2, 3):
            ctx.set_option("col_direct", v)
            check_case(ctx, A, what=f"col_direct={v}")
    finally:
        ctx.set_option("col_direct", 0)


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_big_split_variants(ctx, variant):
    """big_split = 1..3: rows of 4097..16384 products take several CTAs per row, each staging one range of the row's
    positions (map_split.cuh): class edges, rows that need several batches of A entries, rows that fold into a few
    columns (the upper ranges stay empty), an R-MAT with hub rows."""
    rng = np.random.default_rng(77)
    try:
        ctx.set_option("big_split", variant)
        A, B = _rows_with_products([4095, 4096, 4097, 6000, 8191, 8192, 8193, 12345, 16383, 16384, 16385, 700],
                                   cols=1 << 18, nb=256)
        got, st = check_case(ctx, A, B, what=f"big_split={variant} class edges")
        assert st["class_rows"]["sort16384"] == 4 and st["class_rows"]["sort8192"] >= 2
        # several batches of A entries (more entries than threads), B rows of 0..5 entries
        nb, cols = 9000, 1 << 19
        blen = rng.integers(0, 6, nb)
        br = np.repeat(np.arange(nb), blen)
        B = M.from_coo(nb, cols, br, rng.integers(0, cols, br.size), seed=78)
        ar, ac = [], []
        for i, alen in enumerate([1800, 2500, 3300, 5000, 6400]):
            ar += [i] * alen
            ac += list(rng.choice(nb, alen, replace=False))
        A = M.from_coo(5, nb, ar, ac, seed=79)
        got, st = check_case(ctx, A, B, what=f"big_split={variant} many short B rows")
        assert st["class_rows"]["sort16384"] >= 2
        # heavy folding: ~6000 / ~12000 products per row onto <= 302 columns spread over 2^20
        nb, cols = 400, 1 << 20
        pool = np.unique(np.concatenate([[0, cols - 1], rng.integers(0, cols, 300)]))
        br, bc = [], []
        for k in range(nb):
            br += [k] * 40
            bc += list(rng.choice(pool, 40, replace=False))
        B = M.from_coo(nb, cols, br, bc, seed=80)
        ar, ac = [], []
        for i in range(24):
            alen = 150 if i % 2 else 300
            ar += [i] * alen
            ac += list(rng.choice(nb, alen, replace=False))
        A = M.from_coo(24, nb, ar, ac, seed=81)
        got, st = check_case(ctx, A, B, what=f"big_split={variant} folding rows")
        assert st["class_rows"]["sort16384"] == 12 and st["products"] > 5 * st["nnz_c"]
        check_case(ctx, M.rmat(15, 16, seed=15), what=f"big_split={variant} rmat15")
        A32 = M.rmat(14, 16, seed=16, dtype=np.float32)
        got, _ = gpu_multiply(ctx, A32)
        A64 = A32.astype(np.float64)
        assert_csr_equal(got, oracle_multiply(A64, A64), rtol=1e-4, what=f"big_split={variant} fp32")
    finally:
        ctx.set_option("big_split", 0)


@pytest.mark.parametrize("ctas", [1, 4, 16])
def test_sym_mix_plan(ctx, ctas):
    """sym_mix: the 128 / 256 / 512-product sort classes run as capped grids (every CTA loops over row groups) next to
    the rank kernels instead of after them."""
    try:
        ctx.set_option("sym_mix", ctas)
        for A in (M.rmat(15, 16, seed=15), M.rmat(13, 8, seed=5), M.uniform_random(30000, 30000, 12, seed=6)):
            got, st = check_case(ctx, A, what=f"sym_mix={ctas}")
        assert st["class_rows"]["sort256"] > 1000
    finally:
        ctx.set_option("sym_mix", 0)


@pytest.mark.parametrize("narrow", [0, 1, 2, 3])
def test_narrow_lane_groups(ctx, narrow):
    """narrow_groups: rows of <= 64 / <= 128 products share a warp (4 / 8 / 16 lanes per row, 2..8 products per lane) in
    the sort-symbolic and mapped-numeric lane-group kernels; rows whose A entries exceed the group width take several
    batches."""
    try:
        ctx.set_option("narrow_groups", narrow)
        targets = list(range(1, 140)) + [255, 256, 257]
        A, B = _rows_with_products(targets, cols=1 << 16, nb=64)
        check_case(ctx, A, B, what=f"narrow={narrow} every product count up to 139")
        # many A entries with tiny B rows: 20..120 entries per row, B rows of 0..2 entries
        rng = np.random.default_rng(91)
        nb, cols = 5000, 1 << 18
        blen = rng.integers(0, 3, nb)
        br = np.repeat(np.arange(nb), blen)
        B = M.from_coo(nb, cols, br, rng.integers(0, cols, br.size), seed=92)
        ar, ac = [], []
        for i in range(400):
            alen = int(rng.integers(20, 121))
            ar += [i] * alen
            ac += list(rng.choice(nb, alen, replace=False))
        A = M.from_coo(400, nb, ar, ac, seed=93)
        check_case(ctx, A, B, what=f"narrow={narrow} many short B rows")
        # folding inside small rows
        A = M.uniform_random(3000, 300, 6, seed=94)
        B = M.uniform_random(300, 40, 5, seed=95)
        got, st = check_case(ctx, A, B, what=f"narrow={narrow} folding")
        check_case(ctx, M.rmat(14, 4, seed=17), what=f"narrow={narrow} rmat14 ef4")
        A32 = M.rmat(13, 4, seed=18, dtype=np.float32)
        got, _ = gpu_multiply(ctx, A32)
        A64 = A32.astype(np.float64)
        assert_csr_equal(got, oracle_multiply(A64, A64), rtol=1e-4, what=f"narrow={narrow} fp32")
    finally:
        ctx.set_option("narrow_groups", 2)


@pytest.mark.parametrize("sort_max_value", [16384, 1024, 64])
def test_tiered_analysis(ctx, sort_max_value):
    """tiered_analysis=1: the analysis gathers only B's row_offsets and fetches column extents in a second pass for
    the rows that use them; same results on every row class."""
    ctx.set_option("tiered_analysis", 1)
    ctx.set_option("sort_max", sort_max_value)
    try:
        check_case(ctx, M.rmat(14, 16, seed=14), what="tiered rmat14")
        check_case(ctx, M.banded_fem_like(n=4000, per_row=64, clusters=8, band=300, seed=41), what="tiered banded")
        check_case(ctx, M.fem3d_like(9, 8, 7), what="tiered fem")
        check_case(ctx, M.webbase_like(n=50000, seed=3), what="tiered web-like")
        A = M.uniform_random(700, 300, 5, seed=11)
        B = M.uniform_random(300, 1 << 22, 7, seed=12)     # wide: three-level rank kernels / 64-bit keys
        check_case(ctx, A, B, what="tiered wide")
    finally:
        ctx.set_option("tiered_analysis", 0)
        ctx.set_option("sort_max", 16384)


def test_hub_rows_of_a_take_the_cta_analysis(ctx, sort_max):
    """A rows with >= 1024 entries are analysed by one CTA each (k_analyze_long); mixed with ordinary and empty rows,
    hub rows referencing empty and long B rows."""
    rng = np.random.default_rng(8)
    n = 6000
    r = list(rng.integers(0, n, 20000))
    c = list(rng.integers(0, n, 20000))
    for hub, cnt in ((5, 1024), (77, 1023), (4000, 3000), (5999, 5500)):
        cols = rng.choice(n, cnt, replace=False)
        r += [hub] * cnt
        c += list(cols)
    keep = [i for i in range(len(r)) if c[i] % 11 != 3]      # every 11th B row (= A row here) stays empty
    A = M.from_coo(n, n, np.array(r)[keep], np.array(c)[keep], seed=9)
    lens = np.diff(A.row_offsets.astype(np.int64))
    assert (lens >= 1024).sum() >= 2
    got, st = check_case(ctx, A, what=f"hub rows sort_max={sort_max}")
    assert st["max_row_products"] >= 5000
