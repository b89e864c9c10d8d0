"""Deterministic synthetic CSR workloads (host side, numpy only).

The generators follow SURVEY.md section 8d / Appendix D: every matrix is
duplicate-free with ascending columns per row, which is the input precondition
of the reference (include/common.cuh:395-400 reads the first/last column of a
B row as its min/max; HashSpGEMM.cuh:554-564 copies B rows verbatim).
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class HostCSR:
    """Host mirror of the reference's CSR<T> (include/CSR.h:66-72): u32 indices."""
    rows: int
    cols: int
    row_offsets: np.ndarray  # u32[rows+1]
    col_ids: np.ndarray      # u32[nnz]
    data: np.ndarray         # f64/f32[nnz]

    @property
    def nnz(self):
        return int(self.col_ids.shape[0])

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.data, self.col_ids.astype(np.int64),
                              self.row_offsets.astype(np.int64)), shape=(self.rows, self.cols))

    def row_slice(self, r0, r1):
        """Rows [r0, r1) with re-based row_offsets (multi-GPU shard of A)."""
        lo, hi = int(self.row_offsets[r0]), int(self.row_offsets[r1])
        rp = (self.row_offsets[r0:r1 + 1].astype(np.int64) - lo).astype(np.uint32)
        return HostCSR(r1 - r0, self.cols, rp, self.col_ids[lo:hi].copy(), self.data[lo:hi].copy())

    def astype(self, dt):
        return HostCSR(self.rows, self.cols, self.row_offsets, self.col_ids, self.data.astype(dt))


def from_coo(rows, cols, r, c, v=None, seed=0, dtype=np.float64):
    """Dedupe + sort (row, col); values U[0.5, 1.5) from `seed` unless given."""
    r = np.asarray(r, np.int64)
    c = np.asarray(c, np.int64)
    key = r * cols + c
    if v is None:
        key = np.unique(key)
        v = np.random.default_rng(seed).random(key.shape[0]) + 0.5
    else:
        key, idx = np.unique(key, return_index=True)
        v = np.asarray(v)[idx]
    rr = key // cols
    cc = (key % cols).astype(np.uint32)
    rp = np.zeros(rows + 1, np.int64)
    rp[1:] = np.bincount(rr, minlength=rows)
    rp = np.cumsum(rp)
    return HostCSR(rows, cols, rp.astype(np.uint32), cc, np.asarray(v, dtype))


def from_scipy(m, dtype=np.float64):
    m = m.tocsr()
    m.sort_indices()
    return HostCSR(m.shape[0], m.shape[1], m.indptr.astype(np.uint32),
                   m.indices.astype(np.uint32), m.data.astype(dtype))


def tiny8(dtype=np.float64):
    """Config #1: row i has entries at {i, (i+1)%8, (3i+2)%8}, value 1+i+0.125*j."""
    r, c, v = [], [], []
    for i in range(8):
        for j in sorted({i, (i + 1) % 8, (3 * i + 2) % 8}):
            r.append(i), c.append(j), v.append(1.0 + i + 0.125 * j)
    return from_coo(8, 8, r, c, v, dtype=dtype)


def _rmat_keys_native(rng, scale, m, ab, c_norm, a_norm):
    """Replay of the numpy loop below by libspeck_host.so (host/rmat_gen.cpp): same PCG64 stream, same keys,
    ~20x faster.  Returns None when the host library has not been built."""
    import ctypes
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libspeck_host.so")
    if not os.path.exists(path):
        return None
    try:
        lib = ctypes.CDLL(path)
        fn = lib.speck_host_rmat_keys
    except (OSError, AttributeError):
        return None
    u64 = ctypes.c_uint64
    fn.restype = u64
    fn.argtypes = [u64, u64, u64, u64, ctypes.c_int, u64, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                   ctypes.c_void_p]
    st = rng.bit_generator.state
    if st["bit_generator"] != "PCG64":
        return None
    s, inc = st["state"]["state"], st["state"]["inc"]
    mask = (1 << 64) - 1
    keys = np.empty(m, np.uint64)
    n = fn(s >> 64, s & mask, inc >> 64, inc & mask, scale, m, ab, c_norm, a_norm, keys.ctypes.data)
    return keys[:int(n)].astype(np.int64)


def rmat(scale, ef, a=0.45, b=0.15, c=0.15, d=0.25, seed=20, val_seed=None, dtype=np.float64, native=True):
    """R-MAT exactly as SURVEY Appendix D (numpy default_rng = PCG64).
    config #2: rmat(20, 16, seed=20) -> nnz 16 768 028, P 586 218 280."""
    rng = np.random.default_rng(seed)
    n = 1 << scale
    m = n * ef
    ab = a + b
    c_norm = c / (c + d)
    a_norm = a / (a + b)
    key = _rmat_keys_native(rng, scale, m, ab, c_norm, a_norm) if native else None
    if key is None:
        r = np.zeros(m, np.int64)
        cc = np.zeros(m, np.int64)
        for i in range(scale):
            ii = rng.random(m) > ab
            jj = rng.random(m) > (c_norm * ii + a_norm * (~ii))
            r |= ii.astype(np.int64) << i
            cc |= jj.astype(np.int64) << i
        key = np.unique(r * n + cc)
        del r, cc
    vals = np.random.default_rng(seed + 1 if val_seed is None else val_seed).random(key.shape[0]) + 0.5
    rows = key // n
    cols = (key % n).astype(np.uint32)
    rp = np.zeros(n + 1, np.int64)
    rp[1:] = np.bincount(rows, minlength=n)
    rp = np.cumsum(rp)
    return HostCSR(n, n, rp.astype(np.uint32), cols, vals.astype(dtype))


def uniform_random(rows, cols, nnz_per_row, seed=0, dtype=np.float64):
    rng = np.random.default_rng(seed)
    m = int(rows * nnz_per_row)
    return from_coo(rows, cols, rng.integers(0, rows, m), rng.integers(0, cols, m), seed=seed + 1,
                    dtype=dtype)


def banded_fem_like(n=62451, per_row=64, clusters=8, band=700, seed=41, dtype=np.float64):
    """cant_like: clusters of consecutive columns inside a +-band."""
    rng = np.random.default_rng(seed)
    w = per_row // clusters
    rows = np.repeat(np.arange(n, dtype=np.int64), per_row)
    starts = rng.integers(-band, band - w, size=(n, clusters)) + np.arange(n)[:, None]
    cols = (starts[:, :, None] + np.arange(w)[None, None, :]).reshape(-1)
    cols = np.clip(cols, 0, n - 1)
    return from_coo(n, n, rows, cols, seed=seed + 100, dtype=dtype)


def fem3d_like(nx=27, ny=27, nz=29, dof=3, seed=44, dtype=np.float64):
    """cant-like FEM matrix: nx*ny*nz grid nodes, `dof` unknowns per node, 27-point stencil coupling
    all dofs of neighbouring nodes (<= 81 entries per row).  A.A is the radius-2 stencil (<= 375
    entries per row out of <= 6561 products: compression ~15, like SuiteSparse cant: n 62 451,
    nnz 4.0 M, P 269.5 M, nnz(C) 17.4 M).  Default size: 63 423 rows."""
    nodes = nx * ny * nz
    idx = np.arange(nodes, dtype=np.int64)
    x, y, z = idx % nx, (idx // nx) % ny, idx // (nx * ny)
    rows, cols = [], []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                ok = (x + dx >= 0) & (x + dx < nx) & (y + dy >= 0) & (y + dy < ny) & (z + dz >= 0) & (z + dz < nz)
                src = idx[ok]
                dst = src + dx + dy * nx + dz * nx * ny
                for a in range(dof):
                    for b in range(dof):
                        rows.append(src * dof + a)
                        cols.append(dst * dof + b)
    n = nodes * dof
    return from_coo(n, n, np.concatenate(rows), np.concatenate(cols), seed=seed + 100, dtype=dtype)


def econ_like(n=206500, per_row=5.2, seed=42, dtype=np.float64):
    """mac_econ_like: diagonal + uniform random columns (compression ~1.1)."""
    rng = np.random.default_rng(seed)
    m = int(n * per_row)
    r = np.concatenate([np.arange(n, dtype=np.int64), rng.integers(0, n, m)])
    c = np.concatenate([np.arange(n, dtype=np.int64), rng.integers(0, n, m)])
    return from_coo(n, n, r, c, seed=seed + 100, dtype=dtype)


def circuit_like(n=170998, seed=43, dtype=np.float64):
    """scircuit_like: diagonal + near-diagonal + 0.1% heavy rows/cols (100-350 entries)."""
    rng = np.random.default_rng(seed)
    diag = np.arange(n, dtype=np.int64)
    m = int(n * 3.6)
    r1 = rng.integers(0, n, m)
    c1 = np.clip(r1 + rng.integers(-40, 41, m), 0, n - 1)
    heavy = rng.choice(n, max(1, n // 1000), replace=False)
    hr, hc = [], []
    for h in heavy:
        k = int(rng.integers(100, 351))
        t = rng.integers(0, n, k)
        hr.append(np.full(k, h)), hc.append(t)      # heavy row
        hr.append(t), hc.append(np.full(k, h))      # heavy column
    r = np.concatenate([diag, r1] + hr)
    c = np.concatenate([diag, c1] + hc)
    return from_coo(n, n, r, c, seed=seed + 100, dtype=dtype)


def webbase_like(n=1000005, seed=3, dtype=np.float64):
    """webbase1m_like: power-law out-degree (Zipf s~2, capped), 70% of columns within
    +-1000 of the row, nnz ~3.1 M."""
    rng = np.random.default_rng(seed)
    deg = np.minimum(rng.zipf(2.0, n), 4700).astype(np.int64)
    scale = 3.1e6 / deg.sum()
    if scale < 1.0:
        deg = np.maximum(1, (deg * scale).astype(np.int64))
    rows = np.repeat(np.arange(n, dtype=np.int64), deg)
    m = rows.shape[0]
    local = rng.random(m) < 0.7
    cols = np.where(local, rows + rng.integers(-1000, 1001, m), rng.integers(0, n, m))
    cols = np.clip(cols, 0, n - 1)
    return from_coo(n, n, rows, cols, seed=seed + 100, dtype=dtype)


def product_balanced_cuts(a: HostCSR, b_row_offsets, parts):
    """Contiguous row cuts of A balanced by products: prefix sum of
    rowOperations cut at g*P/G.  -> int64[parts+1] row boundaries."""
    blen = np.diff(b_row_offsets.astype(np.int64))
    per_nz = blen[a.col_ids]
    csum = np.concatenate([[0], np.cumsum(per_nz)])
    row_ops_prefix = csum[a.row_offsets.astype(np.int64)]  # products before row i
    total = int(row_ops_prefix[-1])
    targets = (np.arange(1, parts) * total) // parts
    cuts = np.searchsorted(row_ops_prefix, targets, side="left")
    return np.concatenate([[0], cuts, [a.rows]]).astype(np.int64)
