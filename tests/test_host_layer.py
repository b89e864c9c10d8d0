"""CPU tests (no GPU) of the boundary and the host layer: the C-ABI library loads and exports every
symbol include/speck_b200.h declares (no compute calls), the loaders of include/speck_hostio.h
round-trip BASELINE config #1 (tiny8.mtx -> CSR -> .hicsr -> oracle -> golden), the ini parser
follows the reference's six live keys, product-balanced row cuts (multi-GPU) are sane."""
import ctypes
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import oracle
from speck_b200 import api, matrices as M
from speck_b200.matrices import HostCSR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
HOSTLIB = os.path.join(ROOT, "speck_b200", "_lib", "libspeck_host.so")


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(speck_(?:b200|host)_\w+)\s*\(", src)))


def test_cabi_exports_every_declared_symbol():
    lib = api.load_library()
    names = _declared("speck_b200.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/speck_b200.h but not exported"
    assert set(names) == set(api.EXPORTS)
    assert lib.speck_b200_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(api.SpeckError):
        api.Context(0)


def _hostlib():
    if not os.path.exists(HOSTLIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "speck_b200", "host"), "../_lib/libspeck_host.so"])
    lib = ctypes.CDLL(HOSTLIB)
    lib.speck_host_last_error.restype = ctypes.c_char_p
    for n in _declared("speck_hostio.h"):
        assert hasattr(lib, n), n
    return lib


def _load(lib, fn, path):
    rows, cols, nnz = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
    rp, ci, v = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    rc = getattr(lib, fn)(path.encode(), ctypes.byref(rows), ctypes.byref(cols), ctypes.byref(nnz),
                          ctypes.byref(rp), ctypes.byref(ci), ctypes.byref(v))
    if rc != 0:
        raise RuntimeError(lib.speck_host_last_error().decode())
    r, n = rows.value, nnz.value
    out = HostCSR(r, cols.value,
                  np.ctypeslib.as_array(ctypes.cast(rp, ctypes.POINTER(ctypes.c_uint32)), (r + 1,)).copy(),
                  np.ctypeslib.as_array(ctypes.cast(ci, ctypes.POINTER(ctypes.c_uint32)), (max(n, 1),))[:n].copy(),
                  np.ctypeslib.as_array(ctypes.cast(v, ctypes.POINTER(ctypes.c_double)), (max(n, 1),))[:n].copy())
    for p in (rp, ci, v):
        lib.speck_host_free.argtypes = [ctypes.c_void_p]
        lib.speck_host_free(p)
    return out


def test_config1_plumbing_mtx_hicsr_oracle_golden(tmp_path):
    lib = _hostlib()
    A = _load(lib, "speck_host_load_mtx_f64", os.path.join(GOLD, "tiny8.mtx"))
    T = M.tiny8()
    np.testing.assert_array_equal(A.row_offsets, T.row_offsets)
    np.testing.assert_array_equal(A.col_ids, T.col_ids)
    np.testing.assert_array_equal(A.data, T.data)
    path = str(tmp_path / "tiny8.mtxd_.hicsr")
    u32p = np.ctypeslib.ndpointer(np.uint32)
    lib.speck_host_store_hicsr_f64.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t,
                                               u32p, u32p, np.ctypeslib.ndpointer(np.float64)]
    assert lib.speck_host_store_hicsr_f64(path.encode(), A.rows, A.cols, A.nnz, A.row_offsets, A.col_ids, A.data) == 0
    raw = open(path, "rb").read()
    # byte layout of the reference (source/CSR.cpp:27-137): 80-byte header, 16-byte state, arrays
    assert raw[:9] == b"Hi\x01Compsd" and len(raw) == 80 + 16 + A.nnz * 12 + (A.rows + 1) * 4
    assert struct.unpack_from("<8Q", raw, 16) == (8, 0, 4, 0, 4, 8, 8, A.nnz)
    assert struct.unpack_from("<d", raw, 80)[0] == 1.0
    A2 = _load(lib, "speck_host_load_hicsr_f64", path)
    np.testing.assert_array_equal(A2.col_ids, A.col_ids)
    np.testing.assert_array_equal(A2.data, A.data)
    rp, ci, v = oracle.spgemm(A2.row_offsets, A2.col_ids, A2.data, A2.row_offsets, A2.col_ids, A2.data, A2.cols)
    z = np.load(os.path.join(GOLD, "tiny8_expected.npz"))
    np.testing.assert_array_equal(rp, z["c_rp"])
    np.testing.assert_array_equal(ci, z["c_ci"])
    np.testing.assert_allclose(v, z["c_v"], rtol=1e-14)


def test_mtx_symmetric_pattern_duplicates_and_errors(tmp_path):
    lib = _hostlib()
    p = tmp_path / "s.mtx"
    p.write_text("%%MatrixMarket matrix coordinate pattern symmetric\n% c\n3 3 3\n1 1\n3 1\n2 3\n")
    A = _load(lib, "speck_host_load_mtx_f64", str(p))
    assert A.nnz == 5 and list(A.row_offsets) == [0, 2, 3, 5] and list(A.col_ids) == [0, 2, 2, 0, 1]
    assert np.all(A.data == 1.0)
    d = tmp_path / "d.mtx"   # duplicates are kept, adjacent after the sort (CSR.cpp:188-202)
    d.write_text("%%MatrixMarket matrix coordinate real general\n2 2 3\n1 2 1.5\n1 2 2.5\n2 1 3\n")
    D = _load(lib, "speck_host_load_mtx_f64", str(d))
    assert D.nnz == 3 and list(D.col_ids) == [1, 1, 0]
    for bad in ("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n",
                "%%MatrixMarket matrix coordinate real skew-symmetric\n2 2 1\n2 1 1\n",
                "%%MatrixMarket matrix coordinate real general\n2 2 1\n3 1 1\n"):
        b = tmp_path / "bad.mtx"
        b.write_text(bad)
        with pytest.raises(RuntimeError):
            _load(lib, "speck_host_load_mtx_f64", str(b))
    with pytest.raises(RuntimeError):
        _load(lib, "speck_host_load_hicsr_f64", str(d))   # not a .hicsr file


def test_config_ini_keys(tmp_path):
    lib = _hostlib()
    ini = tmp_path / "config.ini"
    ini.write_text("; comment\nTrackCompleteTimes=true\ntrackindividualtimes = false\nCompareResult=yes\n"
                   "IterationsWarmUp=7\nIterationsExecution = 3 ; trailing\n# InputFile=/x\nDevice=2\n")
    gi = lambda k, f: lib.speck_host_config_get_int(str(ini).encode(), k.encode(), f)
    gb = lambda k, f: lib.speck_host_config_get_bool(str(ini).encode(), k.encode(), f)
    assert gi("IterationsWarmUp", 5) == 7 and gi("iterationsexecution", 10) == 3 and gi("Device", 0) == 2
    assert gb("CompareResult", 0) == 1 and gb("TrackIndividualTimes", 1) == 0 and gb("TrackCompleteTimes", 0) == 1
    buf = ctypes.create_string_buffer(64)
    lib.speck_host_config_get_string(str(ini).encode(), b"InputFile", b"fallback.mtx", buf, 64)
    assert buf.value == b"fallback.mtx"
    assert lib.speck_host_config_get_int(None, b"IterationsWarmUp", 5) == 5   # defaults of Executor.cpp:15-16


def test_product_balanced_cuts():
    A = M.rmat(12, 8, seed=9)
    cuts = M.product_balanced_cuts(A, A.row_offsets, 4)
    assert cuts[0] == 0 and cuts[-1] == A.rows and np.all(np.diff(cuts) >= 0)
    ops, _, P, _ = oracle.row_products(A.row_offsets, A.col_ids, A.row_offsets)
    per = [int(ops[cuts[g]:cuts[g + 1]].sum()) for g in range(4)]
    assert sum(per) == P and max(per) < 1.5 * P / 4
    S = A.row_slice(int(cuts[1]), int(cuts[2]))
    assert S.row_offsets[0] == 0 and S.nnz == int(A.row_offsets[cuts[2]] - A.row_offsets[cuts[1]])


def test_runspeck_binary_and_static_lib_exist():
    lib = os.path.join(ROOT, "speck_b200", "_lib")
    if not os.path.exists(os.path.join(lib, "runspECK")):
        pytest.skip("host layer not built (python -c 'import __graft_entry__ as g; g.build()')")
    out = subprocess.run([os.path.join(lib, "runspECK")], capture_output=True, text=True)
    assert out.returncode != 0 and "no .mtx file path set" in out.stdout
    syms = subprocess.run(["nm", "-C", os.path.join(lib, "libspECKLib.a")], capture_output=True, text=True).stdout
    assert "spECK::MultiplyspECK<double, 4, 1024, 49152, 49152>" in syms
    assert "spECK::MultiplyspECK<float, 4, 1024, 49152, 49152>" in syms


def test_every_library_option_is_documented_in_the_header():
    """speck_b200_set_option keys (capi.cu) <-> the option list in include/speck_b200.h."""
    import re
    src = open(os.path.join(ROOT, "speck_b200", "csrc", "capi.cu")).read()
    hdr = open(os.path.join(ROOT, "include", "speck_b200.h")).read()
    keys = set(re.findall(r'strcmp\(key, "([a-z_]+)"\)', src))
    assert len(keys) >= 20
    missing = sorted(k for k in keys if f'"{k}"' not in hdr)
    assert not missing, f"options without documentation in include/speck_b200.h: {missing}"


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the CPU oracle port on the arm's workload) runs without a GPU and prints the
    contract's JSON line: same metric / unit as our arm, cpu_baseline describing the run, e2e equal to the value."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "rmat16",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "GFLOPS" and d["higher_is_better"] is True
    assert d["metric"].startswith("SpGEMM GFLOPS") and d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
