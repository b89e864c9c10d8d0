"""ctypes binding of oracle/_build/liboracle.so (see spgemm_oracle.c)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the C oracle with gcc (seconds)."""
    src = os.path.join(_HERE, "spgemm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "_build/liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


def _load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = ctypes.CDLL(_SO)
    u64 = ctypes.c_uint64
    lib.oracle_row_products.argtypes = [u64, _u32p, _u32p, _u32p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.POINTER(u64), ctypes.POINTER(ctypes.c_uint32)]
    lib.oracle_row_products.restype = None
    lib.oracle_symbolic.argtypes = [u64, u64, _u32p, _u32p, _u32p, _u32p, _u32p]
    lib.oracle_symbolic.restype = u64
    lib.oracle_numeric_f64.argtypes = [u64, u64, _u32p, _u32p, _f64p, _u32p, _u32p, _f64p,
                                       _u32p, _u32p, _f64p]
    lib.oracle_numeric_f64.restype = None
    lib.oracle_numeric_f32.argtypes = [u64, u64, _u32p, _u32p, _f32p, _u32p, _u32p, _f32p,
                                       _u32p, _u32p, _f32p]
    lib.oracle_numeric_f32.restype = None
    lib.oracle_compare_f64.argtypes = [u64, _u32p, _u32p, ctypes.c_void_p, _u32p, _u32p,
                                       ctypes.c_void_p, ctypes.c_double,
                                       ctypes.POINTER(u64), ctypes.POINTER(ctypes.c_double)]
    lib.oracle_compare_f64.restype = ctypes.c_int
    lib.oracle_num_threads.restype = ctypes.c_int
    lib.oracle_set_threads.argtypes = [ctypes.c_int]
    _lib = lib
    return lib


def num_threads():
    return _load().oracle_num_threads()


def set_threads(n):
    _load().oracle_set_threads(int(n))


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def row_products(a_rp, a_ci, b_rp):
    """-> (row_ops u32[rows], row_max u32[rows], P:int, max_row_products:int)
    restates readOperations (include/common.cuh:321-459)."""
    lib = _load()
    a_rp, a_ci, b_rp = _c(a_rp, np.uint32), _c(a_ci, np.uint32), _c(b_rp, np.uint32)
    rows = a_rp.shape[0] - 1
    ops = np.zeros(max(rows, 1), np.uint32)
    mx = np.zeros(max(rows, 1), np.uint32)
    total = ctypes.c_uint64(0)
    gmax = ctypes.c_uint32(0)
    lib.oracle_row_products(rows, a_rp, a_ci, b_rp, ops.ctypes.data, mx.ctypes.data,
                            ctypes.byref(total), ctypes.byref(gmax))
    return ops[:rows], mx[:rows], int(total.value), int(gmax.value)


def symbolic(a_rp, a_ci, b_rp, b_ci, cols_b):
    """-> (c_rp u32[rows+1], nnzC:int)"""
    lib = _load()
    a_rp, a_ci = _c(a_rp, np.uint32), _c(a_ci, np.uint32)
    b_rp, b_ci = _c(b_rp, np.uint32), _c(b_ci, np.uint32)
    rows = a_rp.shape[0] - 1
    c_rp = np.zeros(rows + 1, np.uint32)
    if a_ci.size == 0:
        a_ci = np.zeros(1, np.uint32)
    if b_ci.size == 0:
        b_ci = np.zeros(1, np.uint32)
    nnz = lib.oracle_symbolic(rows, int(cols_b), a_rp, a_ci, b_rp, b_ci, c_rp)
    return c_rp, int(nnz)


def spgemm(a_rp, a_ci, a_v, b_rp, b_ci, b_v, cols_b):
    """C = A.B, column-sorted rows, structural zeros kept. -> (c_rp, c_ci, c_v)"""
    lib = _load()
    dt = np.float32 if np.asarray(a_v).dtype == np.float32 else np.float64
    a_rp, a_ci, a_v = _c(a_rp, np.uint32), _c(a_ci, np.uint32), _c(a_v, dt)
    b_rp, b_ci, b_v = _c(b_rp, np.uint32), _c(b_ci, np.uint32), _c(b_v, dt)
    rows = a_rp.shape[0] - 1
    if a_ci.size == 0:
        a_ci, a_v = np.zeros(1, np.uint32), np.zeros(1, dt)
    if b_ci.size == 0:
        b_ci, b_v = np.zeros(1, np.uint32), np.zeros(1, dt)
    c_rp = np.zeros(rows + 1, np.uint32)
    nnz = lib.oracle_symbolic(rows, int(cols_b), a_rp, a_ci, b_rp, b_ci, c_rp)
    if nnz >= 2 ** 32:
        raise OverflowError("nnz(C) does not fit the u32 row_offsets of the spECK API")
    c_ci = np.zeros(max(nnz, 1), np.uint32)
    c_v = np.zeros(max(nnz, 1), dt)
    fn = lib.oracle_numeric_f32 if dt == np.float32 else lib.oracle_numeric_f64
    fn(rows, int(cols_b), a_rp, a_ci, a_v, b_rp, b_ci, b_v, c_rp, c_ci, c_v)
    return c_rp, c_ci[:nnz], c_v[:nnz]


def compare(rp_a, ci_a, v_a, rp_b, ci_b, v_b, rel_tol=1e-6):
    """restates d_compare (source/GPU/Compare.cu:11-62).
    -> (code, where, max_rel); code 0 = equal, 1 row_ptr, 2 col_idx, 3 values."""
    lib = _load()
    rp_a, rp_b = _c(rp_a, np.uint32), _c(rp_b, np.uint32)
    if rp_a.shape != rp_b.shape:
        return 1, 0, 0.0
    ci_a, ci_b = _c(ci_a, np.uint32), _c(ci_b, np.uint32)
    rows = rp_a.shape[0] - 1
    nnz_a, nnz_b = int(rp_a[rows]), int(rp_b[rows])
    if nnz_a != nnz_b or ci_a.size < nnz_a or ci_b.size < nnz_b:
        # let the C routine find the first differing row_ptr entry
        pass
    if ci_a.size == 0:
        ci_a = np.zeros(1, np.uint32)
    if ci_b.size == 0:
        ci_b = np.zeros(1, np.uint32)
    pa = pb = None
    if v_a is not None and v_b is not None:
        va, vb = _c(v_a, np.float64), _c(v_b, np.float64)
        if va.size == 0:
            va = np.zeros(1)
        if vb.size == 0:
            vb = np.zeros(1)
        pa, pb = va.ctypes.data, vb.ctypes.data
    where = ctypes.c_uint64(0)
    mr = ctypes.c_double(0.0)
    code = lib.oracle_compare_f64(rows, rp_a, ci_a, pa, rp_b, ci_b, pb, float(rel_tol),
                                  ctypes.byref(where), ctypes.byref(mr))
    return int(code), int(where.value), float(mr.value)
