"""ctypes binding of the C ABI in include/speck_b200.h.

Mirrors the reference's host-side objects for the hot path:
    Context      <-> spECK::spECKConfig            (include/spECKConfig.h:8-53)
    DeviceCSR    <-> dCSR<T>                       (include/dCSR.h:9-22), convert() up/down
    Context.multiply(A, B, C) <-> spECK::MultiplyspECK<T,4,1024,...>(A, B, C, config, timings)
                                                   (include/Multiply.h:15-16)
    Context.compare(ref, cmp) <-> spECK::Compare   (include/Compare.h:5-6)

There is no CPU fallback: if the CUDA library is missing or no B200 is visible, every
entry point raises SpeckError.
"""
import ctypes
import os

import numpy as np

from .matrices import HostCSR

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPECK_B200_LIB") or os.path.join(_HERE, "_lib", "libspeck_b200.so")   # override: kernel experiments

NUM_CLASSES = 32
BIN_NAMES = (["direct"] + [f"sort{4 << c}" for c in range(8)] + [f"sort{512 * w}" for w in range(2, 17)]
             + ["sort16384", "dense_local", "dense"])


class SpeckError(RuntimeError):
    pass


class CsrStruct(ctypes.Structure):
    _fields_ = [("rows", ctypes.c_size_t), ("cols", ctypes.c_size_t), ("nnz", ctypes.c_size_t),
                ("data", ctypes.c_void_p), ("row_offsets", ctypes.c_void_p),
                ("col_ids", ctypes.c_void_p)]


class TimingsStruct(ctypes.Structure):
    _fields_ = [("measure_all", ctypes.c_int), ("measure_complete", ctypes.c_int)] + [
        (n, ctypes.c_float) for n in
        ("init", "count_products", "load_balance_counting", "global_maps_counting",
         "spgemm_counting", "alloc_c", "load_balance_numeric", "global_maps_numeric",
         "spgemm_numeric", "sorting", "cleanup", "complete")]


class StatsStruct(ctypes.Structure):
    _fields_ = [("products", ctypes.c_uint64), ("nnz_c", ctypes.c_uint64),
                ("max_row_products", ctypes.c_uint32),
                ("class_rows", ctypes.c_uint32 * NUM_CLASSES),
                ("kernel_launches", ctypes.c_uint32),
                ("ms_analysis", ctypes.c_float), ("ms_symbolic", ctypes.c_float),
                ("ms_scan", ctypes.c_float), ("ms_numeric", ctypes.c_float),
                ("ms_total", ctypes.c_float), ("workspace_bytes", ctypes.c_uint64)]


MAX_SHARDS = 16


class ShardInfoStruct(ctypes.Structure):
    _fields_ = [("shards", ctypes.c_int), ("concatenated", ctypes.c_int),
                ("cuts", ctypes.c_uint32 * (MAX_SHARDS + 1)), ("products", ctypes.c_uint64 * MAX_SHARDS),
                ("nnz_c", ctypes.c_uint64 * MAX_SHARDS), ("ms_device", ctypes.c_float * MAX_SHARDS),
                ("ms_setup", ctypes.c_float), ("ms_multiply", ctypes.c_float), ("ms_concat", ctypes.c_float)]


class MismatchStruct(ctypes.Structure):
    _fields_ = [("row", ctypes.c_uint64), ("kind", ctypes.c_uint32), ("index_in_row", ctypes.c_uint32),
                ("ref_len", ctypes.c_uint32), ("cmp_len", ctypes.c_uint32), ("ref_col", ctypes.c_uint32),
                ("cmp_col", ctypes.c_uint32), ("ref_val", ctypes.c_double), ("cmp_val", ctypes.c_double)]


EXPORTS = [
    "speck_b200_abi_version", "speck_b200_last_error", "speck_b200_create", "speck_b200_destroy",
    "speck_b200_sm_count", "speck_b200_spgemm_f64", "speck_b200_spgemm_f32",
    "speck_b200_spgemm_host_f64", "speck_b200_spgemm_host_f32", "speck_b200_get_stats",
    "speck_b200_row_products", "speck_b200_compare_f64", "speck_b200_compare_f32",
    "speck_b200_malloc", "speck_b200_free", "speck_b200_memcpy_h2d", "speck_b200_memcpy_d2h",
    "speck_b200_free_csr", "speck_b200_synchronize", "speck_b200_stream", "speck_b200_set_option",
    "speck_b200_compare_report_f64", "speck_b200_compare_report_f32", "speck_b200_partition_rows",
    "speck_b200_coo_to_csr_f64", "speck_b200_coo_to_csr_f32", "speck_b200_sharded_create_f64", "speck_b200_sharded_create_f32",
    "speck_b200_sharded_multiply", "speck_b200_sharded_concat", "speck_b200_sharded_slab",
    "speck_b200_sharded_destroy", "speck_b200_ipc_export", "speck_b200_ipc_open", "speck_b200_ipc_close",
    "speck_b200_push_slab_f64", "speck_b200_push_slab_f32",
]

_lib = None


def load_library():
    """Load libspeck_b200.so; raises SpeckError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SpeckError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(make -C speck_b200/csrc). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    P = ctypes.POINTER
    vp = ctypes.c_void_p
    lib.speck_b200_last_error.restype = ctypes.c_char_p
    lib.speck_b200_create.argtypes = [ctypes.c_int, P(vp)]
    lib.speck_b200_destroy.argtypes = [vp]
    lib.speck_b200_sm_count.argtypes = [vp]
    for n in ("speck_b200_spgemm_f64", "speck_b200_spgemm_f32"):
        getattr(lib, n).argtypes = [vp, P(CsrStruct), P(CsrStruct), P(CsrStruct), P(TimingsStruct)]
    for n in ("speck_b200_spgemm_host_f64", "speck_b200_spgemm_host_f32"):
        getattr(lib, n).argtypes = [vp, P(CsrStruct), P(CsrStruct), P(CsrStruct),
                                    P(ctypes.c_uint64), P(ctypes.c_uint64)]
    lib.speck_b200_get_stats.argtypes = [vp, P(StatsStruct)]
    lib.speck_b200_row_products.argtypes = [vp, P(CsrStruct), P(CsrStruct), vp,
                                            P(ctypes.c_uint64), P(ctypes.c_uint32)]
    for n in ("speck_b200_compare_f64", "speck_b200_compare_f32"):
        getattr(lib, n).argtypes = [vp, P(CsrStruct), P(CsrStruct), ctypes.c_int, ctypes.c_double]
    for n in ("speck_b200_compare_report_f64", "speck_b200_compare_report_f32"):
        getattr(lib, n).argtypes = [vp, P(CsrStruct), P(CsrStruct), ctypes.c_int, ctypes.c_double, P(MismatchStruct)]
    for n in ("speck_b200_coo_to_csr_f64", "speck_b200_coo_to_csr_f32"):
        getattr(lib, n).argtypes = [vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, vp, vp, vp, ctypes.c_int, P(CsrStruct)]
    lib.speck_b200_malloc.argtypes = [vp, P(vp), ctypes.c_size_t]
    lib.speck_b200_free.argtypes = [vp, vp]
    lib.speck_b200_memcpy_h2d.argtypes = [vp, vp, vp, ctypes.c_size_t]
    lib.speck_b200_memcpy_d2h.argtypes = [vp, vp, vp, ctypes.c_size_t]
    lib.speck_b200_free_csr.argtypes = [vp, P(CsrStruct)]
    lib.speck_b200_synchronize.argtypes = [vp]
    lib.speck_b200_stream.argtypes = [vp]
    lib.speck_b200_stream.restype = vp
    lib.speck_b200_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_longlong]
    lib.speck_b200_partition_rows.argtypes = [vp, P(CsrStruct), P(CsrStruct), ctypes.c_int, P(ctypes.c_uint32),
                                              P(ctypes.c_uint64)]
    for n in ("speck_b200_sharded_create_f64", "speck_b200_sharded_create_f32"):
        getattr(lib, n).argtypes = [P(vp), ctypes.c_int, P(CsrStruct), P(CsrStruct), P(vp)]
    lib.speck_b200_sharded_multiply.argtypes = [vp, P(ShardInfoStruct)]
    lib.speck_b200_sharded_concat.argtypes = [vp, P(CsrStruct), P(ShardInfoStruct)]
    lib.speck_b200_sharded_slab.argtypes = [vp, ctypes.c_int, P(CsrStruct), P(CsrStruct)]
    lib.speck_b200_sharded_destroy.argtypes = [vp]
    lib.speck_b200_ipc_export.argtypes = [vp, vp, ctypes.c_char_p]
    lib.speck_b200_ipc_open.argtypes = [vp, ctypes.c_char_p, P(vp)]
    lib.speck_b200_ipc_close.argtypes = [vp, vp]
    for n in ("speck_b200_push_slab_f64", "speck_b200_push_slab_f32"):
        getattr(lib, n).argtypes = [vp, P(CsrStruct), ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int, vp, vp, vp,
                                    P(ctypes.c_float)]
    _lib = lib
    return lib


def _check(rc):
    if rc < 0:
        raise SpeckError(f"speck_b200 error {rc}: {load_library().speck_b200_last_error().decode()}")
    return rc


class DeviceCSR:
    """Device CSR (the reference's dCSR<T>): three cudaMalloc'ed arrays + sizes."""

    def __init__(self, ctx, dtype=np.float64):
        self.ctx = ctx
        self.dtype = np.dtype(dtype)
        self.s = CsrStruct(0, 0, 0, None, None, None)

    rows = property(lambda self: self.s.rows)
    cols = property(lambda self: self.s.cols)
    nnz = property(lambda self: self.s.nnz)

    owned = True

    @classmethod
    def from_pointers(cls, ctx, rows, cols, nnz, row_offsets_ptr, col_ids_ptr, data_ptr, dtype=np.float64):
        """Borrowed view over device memory owned by someone else (e.g. torch tensors that
        received B through an NCCL broadcast)."""
        d = cls(ctx, dtype)
        d.s = CsrStruct(rows, cols, nnz, data_ptr, row_offsets_ptr, col_ids_ptr)
        d.owned = False
        return d

    def free(self):
        if self.owned and self.ctx is not None and self.ctx.h:
            _check(self.ctx.lib.speck_b200_free_csr(self.ctx.h, ctypes.byref(self.s)))

    def view(self, r0=None, r1=None):
        """Borrowed struct (for passing as A/B)."""
        return self.s


class Context:
    def __init__(self, device=0):
        self.lib = load_library()
        self.h = ctypes.c_void_p()
        _check(self.lib.speck_b200_create(int(device), ctypes.byref(self.h)))
        self.device = device

    def close(self):
        if self.h:
            self.lib.speck_b200_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def sm_count(self):
        return self.lib.speck_b200_sm_count(self.h)

    def set_option(self, key, value):
        _check(self.lib.speck_b200_set_option(self.h, key.encode(), int(value)))

    # ---- convert(dCSR&, const CSR&) / convert(CSR&, const dCSR&)  (source/dCSR.cpp:50-76)
    def _alloc(self, nbytes):
        p = ctypes.c_void_p()
        _check(self.lib.speck_b200_malloc(self.h, ctypes.byref(p), nbytes))
        return p

    # ---- one process per GPU: the concatenated C lives on one device, the others push their slabs into it
    IPC_HANDLE_BYTES = 64

    def ipc_export(self, dptr) -> bytes:
        """64-byte CUDA IPC handle of a device allocation made by this library (speck_b200_malloc)."""
        buf = ctypes.create_string_buffer(self.IPC_HANDLE_BYTES)
        _check(self.lib.speck_b200_ipc_export(self.h, dptr, buf))
        return buf.raw

    def ipc_open(self, handle: bytes):
        p = ctypes.c_void_p()
        _check(self.lib.speck_b200_ipc_open(self.h, ctypes.create_string_buffer(handle, self.IPC_HANDLE_BYTES), ctypes.byref(p)))
        return p

    def ipc_close(self, dptr):
        _check(self.lib.speck_b200_ipc_close(self.h, dptr))

    def push_slab(self, slab: DeviceCSR, nnz_base, row_base, last, dst_row_offsets, dst_col_ids, dst_data):
        """Slab of C -> its place in the concatenated C (local or IPC-opened peer arrays); returns the device ms."""
        fn = self.lib.speck_b200_push_slab_f32 if slab.dtype == np.float32 else self.lib.speck_b200_push_slab_f64
        ms = ctypes.c_float(0)
        _check(fn(self.h, ctypes.byref(slab.s), int(nnz_base), int(row_base), int(bool(last)), dst_row_offsets, dst_col_ids,
                  dst_data, ctypes.byref(ms)))
        return float(ms.value)

    def upload(self, m: HostCSR) -> DeviceCSR:
        d = DeviceCSR(self, m.data.dtype)
        rp = np.ascontiguousarray(m.row_offsets, np.uint32)
        ci = np.ascontiguousarray(m.col_ids, np.uint32)
        v = np.ascontiguousarray(m.data)
        d.s.rows, d.s.cols, d.s.nnz = m.rows, m.cols, m.nnz
        d.s.row_offsets = self._alloc(rp.nbytes)
        d.s.col_ids = self._alloc(ci.nbytes)
        d.s.data = self._alloc(v.nbytes)
        _check(self.lib.speck_b200_memcpy_h2d(self.h, d.s.row_offsets, rp.ctypes.data, rp.nbytes))
        _check(self.lib.speck_b200_memcpy_h2d(self.h, d.s.col_ids, ci.ctypes.data, ci.nbytes))
        _check(self.lib.speck_b200_memcpy_h2d(self.h, d.s.data, v.ctypes.data, v.nbytes))
        return d

    def download(self, d: DeviceCSR) -> HostCSR:
        rows, nnz = d.s.rows, d.s.nnz
        rp = np.zeros(rows + 1, np.uint32)
        ci = np.zeros(nnz, np.uint32)
        v = np.zeros(nnz, d.dtype)
        if d.s.row_offsets:
            _check(self.lib.speck_b200_memcpy_d2h(self.h, rp.ctypes.data, d.s.row_offsets, rp.nbytes))
        if nnz:
            _check(self.lib.speck_b200_memcpy_d2h(self.h, ci.ctypes.data, d.s.col_ids, ci.nbytes))
            _check(self.lib.speck_b200_memcpy_d2h(self.h, v.ctypes.data, d.s.data, v.nbytes))
        return HostCSR(rows, d.s.cols, rp, ci, v)

    # ---- MultiplyspECK
    def multiply(self, A: DeviceCSR, B: DeviceCSR, C: DeviceCSR = None, timings=None) -> DeviceCSR:
        if A.dtype != B.dtype:
            raise SpeckError(f"A is {A.dtype} but B is {B.dtype}: both operands must have the same value type")
        if C is None:
            C = DeviceCSR(self, A.dtype)
        elif C.dtype != A.dtype:
            if C.s.data or C.s.nnz:
                raise SpeckError(f"C holds {C.dtype} values but A and B are {A.dtype}: free C first (its value buffer "
                                 "would be reused with the wrong element size)")
            C.dtype = A.dtype
        fn = self.lib.speck_b200_spgemm_f32 if A.dtype == np.float32 else self.lib.speck_b200_spgemm_f64
        t = timings if timings is not None else TimingsStruct()
        _check(fn(self.h, ctypes.byref(A.s), ctypes.byref(B.s), ctypes.byref(C.s), ctypes.byref(t)))
        return C

    def multiply_host(self, A: HostCSR, B: HostCSR):
        """End-to-end on host buffers -> (HostCSR view into the context's pinned output buffers,
        h2d_bytes, d2h_bytes)."""
        dt = np.dtype(A.data.dtype)
        fn = self.lib.speck_b200_spgemm_host_f32 if dt == np.float32 else self.lib.speck_b200_spgemm_host_f64

        def st(m):
            return CsrStruct(m.rows, m.cols, m.nnz, m.data.ctypes.data, m.row_offsets.ctypes.data,
                             m.col_ids.ctypes.data)
        sa = st(A)
        sb_ = sa if B is A else st(B)
        sc = CsrStruct()
        up, down = ctypes.c_uint64(0), ctypes.c_uint64(0)
        _check(fn(self.h, ctypes.byref(sa), ctypes.byref(sb_), ctypes.byref(sc), ctypes.byref(up),
                  ctypes.byref(down)))
        rows, nnz = sc.rows, sc.nnz
        if sc.row_offsets:
            rp = np.ctypeslib.as_array(ctypes.cast(sc.row_offsets, ctypes.POINTER(ctypes.c_uint32)), (rows + 1,))
        else:
            rp = np.zeros(rows + 1, np.uint32)
        if nnz:
            ci = np.ctypeslib.as_array(ctypes.cast(sc.col_ids, ctypes.POINTER(ctypes.c_uint32)), (nnz,))
            ct = ctypes.c_float if dt == np.float32 else ctypes.c_double
            v = np.ctypeslib.as_array(ctypes.cast(sc.data, ctypes.POINTER(ct)), (nnz,))
        else:
            ci, v = np.zeros(0, np.uint32), np.zeros(0, dt)
        return HostCSR(rows, sc.cols, rp, ci, v), int(up.value), int(down.value)

    def stats(self):
        s = StatsStruct()
        _check(self.lib.speck_b200_get_stats(self.h, ctypes.byref(s)))
        return {
            "products": int(s.products), "nnz_c": int(s.nnz_c),
            "max_row_products": int(s.max_row_products),
            "class_rows": {BIN_NAMES[i]: int(s.class_rows[i]) for i in range(len(BIN_NAMES))},
            "kernel_launches": int(s.kernel_launches),
            "ms_analysis": s.ms_analysis, "ms_symbolic": s.ms_symbolic, "ms_scan": s.ms_scan,
            "ms_numeric": s.ms_numeric, "ms_total": s.ms_total,
            "workspace_bytes": int(s.workspace_bytes),
        }

    def row_products(self, A: DeviceCSR, B: DeviceCSR):
        """-> (rowOps u32[rows], P, max) ; readOperations (include/common.cuh:321-459)."""
        rows = A.rows
        d = self._alloc(max(rows, 1) * 4)
        P, mx = ctypes.c_uint64(0), ctypes.c_uint32(0)
        try:
            _check(self.lib.speck_b200_row_products(self.h, ctypes.byref(A.s), ctypes.byref(B.s), d,
                                                    ctypes.byref(P), ctypes.byref(mx)))
            out = np.zeros(rows, np.uint32)
            if rows:
                _check(self.lib.speck_b200_memcpy_d2h(self.h, out.ctypes.data, d, rows * 4))
        finally:
            self.lib.speck_b200_free(self.h, d)
        return out, int(P.value), int(mx.value)

    def partition_rows(self, A: DeviceCSR, B: DeviceCSR, parts):
        """-> (cuts uint32[parts+1], products uint64[parts]): contiguous row cuts of A balanced by products,
        computed on the device (analysis -> 64-bit scan -> search)."""
        cuts = (ctypes.c_uint32 * (parts + 1))()
        prods = (ctypes.c_uint64 * parts)()
        _check(self.lib.speck_b200_partition_rows(self.h, ctypes.byref(A.s), ctypes.byref(B.s), int(parts), cuts, prods))
        return np.array(cuts[:], np.uint32), np.array(prods[:], np.uint64)

    def compare(self, ref: DeviceCSR, cmp: DeviceCSR, compare_data=False, rel_tol=1e-6):
        fn = self.lib.speck_b200_compare_f32 if ref.dtype == np.float32 else self.lib.speck_b200_compare_f64
        return bool(_check(fn(self.h, ctypes.byref(ref.s), ctypes.byref(cmp.s), int(compare_data), float(rel_tol))))

    def coo_to_csr(self, rows, cols, row_ids, col_ids, values, sum_duplicates=False) -> DeviceCSR:
        """GPU-side COO -> CSR of host COO arrays (uploaded here; the conversion itself runs on the device)."""
        r = np.ascontiguousarray(row_ids, np.uint32)
        c = np.ascontiguousarray(col_ids, np.uint32)
        v = np.ascontiguousarray(values)
        if v.dtype not in (np.float32, np.float64):
            v = v.astype(np.float64)
        n = r.size
        bufs = [self._alloc(max(a.nbytes, 4)) for a in (r, c, v)]
        try:
            for b, a in zip(bufs, (r, c, v)):
                if a.nbytes:
                    _check(self.lib.speck_b200_memcpy_h2d(self.h, b, a.ctypes.data, a.nbytes))
            d = DeviceCSR(self, v.dtype)
            fn = self.lib.speck_b200_coo_to_csr_f32 if v.dtype == np.float32 else self.lib.speck_b200_coo_to_csr_f64
            rc = fn(self.h, int(rows), int(cols), int(n), bufs[0], bufs[1], bufs[2], 1 if sum_duplicates else 0, ctypes.byref(d.s))
            if rc < 0:
                raise SpeckError(f"speck_b200_coo_to_csr failed with status {rc}")
        finally:
            for b in bufs:
                self.lib.speck_b200_free(self.h, b)
        return d

    def compare_report(self, ref: DeviceCSR, cmp: DeviceCSR, compare_data=False, rel_tol=1e-6):
        """-> (equal, first difference or None); kinds: 0 row length, 1 column id, 2 value, 3 shape / nnz."""
        fn = self.lib.speck_b200_compare_report_f32 if ref.dtype == np.float32 else self.lib.speck_b200_compare_report_f64
        m = MismatchStruct()
        eq = bool(_check(fn(self.h, ctypes.byref(ref.s), ctypes.byref(cmp.s), int(compare_data), float(rel_tol), ctypes.byref(m))))
        return eq, (None if eq else {f: getattr(m, f) for f, _ in MismatchStruct._fields_})

    def synchronize(self):
        _check(self.lib.speck_b200_synchronize(self.h))


class ShardPlan:
    """Row-sharded multiply over several devices of one box driven from one process (include/speck_b200.h:
    speck_b200_sharded_*): HOST A and B in, product-balanced slabs of A on the contexts' devices, B replicated
    over NVLink peer copies, concurrent slab multiplies, optional concatenation on the first device."""

    def __init__(self, ctxs, A: HostCSR, B: HostCSR = None):
        self.ctxs = list(ctxs)
        self.lib = self.ctxs[0].lib
        self.dtype = np.dtype(A.data.dtype)
        B = A if B is None else B
        self._keep = (A, B)

        def st(m):
            return CsrStruct(m.rows, m.cols, m.nnz, m.data.ctypes.data, m.row_offsets.ctypes.data, m.col_ids.ctypes.data)
        arr = (ctypes.c_void_p * len(self.ctxs))(*[c.h for c in self.ctxs])
        self.h = ctypes.c_void_p()
        fn = self.lib.speck_b200_sharded_create_f32 if self.dtype == np.float32 else self.lib.speck_b200_sharded_create_f64
        sa, sb_ = st(A), st(B)
        _check(fn(arr, len(self.ctxs), ctypes.byref(sa), ctypes.byref(sb_), ctypes.byref(self.h)))
        self.info = ShardInfoStruct()

    def multiply(self):
        _check(self.lib.speck_b200_sharded_multiply(self.h, ctypes.byref(self.info)))
        return self.summary()

    def concat(self, C: DeviceCSR = None) -> DeviceCSR:
        if C is None:
            C = DeviceCSR(self.ctxs[0], self.dtype)
        _check(self.lib.speck_b200_sharded_concat(self.h, ctypes.byref(C.s), ctypes.byref(self.info)))
        return C

    def slab(self, g):
        """Borrowed (A slab, C slab) device views on device g."""
        a, c = DeviceCSR(self.ctxs[g], self.dtype), DeviceCSR(self.ctxs[g], self.dtype)
        a.owned = c.owned = False
        _check(self.lib.speck_b200_sharded_slab(self.h, int(g), ctypes.byref(a.s), ctypes.byref(c.s)))
        return a, c

    def summary(self):
        i, n = self.info, self.info.shards
        return {"shards": n, "cuts": list(i.cuts[:n + 1]), "products": list(i.products[:n]), "nnz_c": list(i.nnz_c[:n]),
                "ms_device": list(i.ms_device[:n]), "ms_setup": i.ms_setup, "ms_multiply": i.ms_multiply,
                "ms_concat": i.ms_concat, "concatenated": bool(i.concatenated)}

    def close(self):
        if self.h:
            self.lib.speck_b200_sharded_destroy(self.h)
            self.h = ctypes.c_void_p()


def numeric_bytes(rows_a, nnz_a, products, nnz_c, val_bytes=8):
    """Algorithmic bytes of the numeric phase:
    8(rA+1) + 12 nnzA + 8 nnzA + 12 P + 12 nnzC for fp64 / u32 indices."""
    iv = 4 + val_bytes
    return 8 * (rows_a + 1) + iv * nnz_a + 8 * nnz_a + iv * products + iv * nnz_c


def total_bytes(rows_a, nnz_a, products, nnz_c, val_bytes=8):
    ana = 4 * (rows_a + 1) + 12 * nnz_a + 12 * rows_a
    sym = 4 * (rows_a + 1) + 12 * nnz_a + 4 * products + 4 * rows_a
    scan = 8 * (rows_a + 1)
    return ana + sym + scan + numeric_bytes(rows_a, nnz_a, products, nnz_c, val_bytes)
