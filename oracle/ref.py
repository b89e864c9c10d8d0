"""ctypes binding of oracle/_ref/libspeck_ref_{stock,tuned}.so: the UNMODIFIED reference spECK,
compiled in place from /root/reference for sm_100 by oracle/build_ref.sh.  Test / baseline
infrastructure only (needs a GPU)."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}


def available(variant="stock"):
    return os.path.exists(os.path.join(_HERE, "_ref", f"libspeck_ref_{variant}.so"))


def _load(variant):
    if variant in _libs:
        return _libs[variant]
    path = os.path.join(_HERE, "_ref", f"libspeck_ref_{variant}.so")
    lib = ctypes.CDLL(path)
    u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
    f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
    sz = ctypes.c_size_t
    lib.ref_speck_multiply_f64.argtypes = [
        sz, sz, sz, u32p, u32p, f64p, sz, sz, sz, u32p, u32p, f64p,
        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_uint64),
        ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
        ctypes.c_void_p, ctypes.c_void_p]
    lib.ref_speck_multiply_f64.restype = ctypes.c_int
    if hasattr(lib, "ref_speck_multiply_f32"):
        f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
        lib.ref_speck_multiply_f32.argtypes = [
            sz, sz, sz, u32p, u32p, f32p, sz, sz, sz, u32p, u32p, f32p,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_uint64),
            ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
            ctypes.c_void_p, ctypes.c_void_p]
        lib.ref_speck_multiply_f32.restype = ctypes.c_int
    lib.ref_speck_free.argtypes = [ctypes.c_void_p]
    _libs[variant] = lib
    return lib


STAGES = ["init", "countProducts", "loadBalanceCounting", "globalMapsCounting", "spGEMMCounting", "allocC",
          "loadBalanceNumeric", "globalMapsNumeric", "spGEMMNumeric", "sorting", "cleanup", "complete"]


def multiply(A, B=None, warmup=1, iters=1, variant="stock", stages=False, fetch=True):
    """Reference spECK C = A.B (fp64, or fp32 when A holds float32 values).
    -> dict(rp, ci, v, nnz, times_ms[iters], stage_ms{})"""
    lib = _load(variant)
    B = A if B is None else B
    dt = np.float32 if A.data.dtype == np.float32 else np.float64
    fn = lib.ref_speck_multiply_f32 if dt == np.float32 else lib.ref_speck_multiply_f64
    a = (np.ascontiguousarray(A.row_offsets, np.uint32), np.ascontiguousarray(A.col_ids, np.uint32),
         np.ascontiguousarray(A.data, dt))
    b = a if B is A else (np.ascontiguousarray(B.row_offsets, np.uint32), np.ascontiguousarray(B.col_ids, np.uint32),
                          np.ascontiguousarray(B.data, dt))
    nnz = ctypes.c_uint64(0)
    prp, pci, pv = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    times = np.zeros(max(iters, 1), np.float32)
    st = np.zeros(12, np.float32)
    rc = fn(A.rows, A.cols, A.nnz, a[0], a[1], a[2], B.rows, B.cols, B.nnz, b[0], b[1], b[2],
                                    int(warmup), int(iters), int(bool(stages)), ctypes.byref(nnz),
                                    ctypes.byref(prp), ctypes.byref(pci), ctypes.byref(pv),
                                    times.ctypes.data, st.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"reference spECK failed with code {rc}")
    n = int(nnz.value)
    out = {"nnz": n, "times_ms": times[:iters].copy(), "stage_ms": dict(zip(STAGES, st.tolist())) if stages else None}
    if fetch:
        out["rp"] = np.ctypeslib.as_array(ctypes.cast(prp, ctypes.POINTER(ctypes.c_uint32)), (A.rows + 1,)).copy()
        out["ci"] = np.ctypeslib.as_array(ctypes.cast(pci, ctypes.POINTER(ctypes.c_uint32)), (max(n, 1),))[:n].copy()
        ct = ctypes.c_float if dt == np.float32 else ctypes.c_double
        out["v"] = np.ctypeslib.as_array(ctypes.cast(pv, ctypes.POINTER(ct)), (max(n, 1),))[:n].copy()
    for p in (prp, pci, pv):
        lib.ref_speck_free(p)
    return out
