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


@pytest.fixture(params=[(16384, 1, 1, 1, 8, 0), (16384, 1, 1, 1, 16, 1), (16384, 1, 1, 0, 8, 0), (8192, 1, 0, 1, 8, 0),
                        (8192, 0, 1, 1, 8, 0), (8192, 0, 0, 1, 8, 0), (1024, 1, 1, 1, 8, 1), (64, 1, 1, 1, 8, 0),
                        (64, 1, 0, 1, 8, 0)],
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
    yield request.param[0]
    ctx.set_option("sort_max", 16384)
    ctx.set_option("rank_path", 1)
    ctx.set_option("rank_map", 1)
    ctx.set_option("flat_sym", 1)
    ctx.set_option("flat_e", 8)
    ctx.set_option("seg_num", 0)


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
