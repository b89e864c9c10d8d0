"""GPU: the reference-API host layer end to end -- runspECK <matrix.mtx> [config.ini] (reference
source/runspECK.cpp, Executor.cpp): same stdout contract, .hicsr cache written next to the .mtx,
CompareResult=true checks every iteration against the cuSPARSE-12 comparator
(speck_b200/host/cusparse_shim.cu) and must not print "Error: Matrix incorrect"."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from speck_b200 import matrices as M

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "speck_b200", "_lib", "runspECK")


def write_mtx(path, A):
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write(f"{A.rows} {A.cols} {A.nnz}\n")
        rows = np.repeat(np.arange(A.rows), np.diff(A.row_offsets.astype(np.int64)))
        for r, c, v in zip(rows, A.col_ids, A.data):
            f.write(f"{r + 1} {c + 1} {float(v)!r}\n")


def run(args, cwd):
    if not os.path.exists(EXE):
        pytest.skip("runspECK not built (python -c 'import __graft_entry__ as g; g.build()')")
    return subprocess.run([EXE] + args, capture_output=True, text=True, cwd=cwd, timeout=300)


def test_runspeck_tiny8_stdout_contract_and_cache(tmp_path):
    mtx = tmp_path / "tiny8.mtx"
    shutil.copy(os.path.join(ROOT, "tests", "golden", "tiny8.mtx"), mtx)
    ini = tmp_path / "config.ini"
    ini.write_text("TrackCompleteTimes=true\nTrackIndividualTimes=false\nCompareResult=true\n"
                   "IterationsWarmUp=2\nIterationsExecution=3\n")
    out = run([str(mtx), str(ini)], tmp_path)
    assert out.returncode == 0, out.stdout + out.stderr
    z = np.load(os.path.join(ROOT, "tests", "golden", "tiny8_expected.npz"))
    assert "Matrix: 8x8: 22 nonzeros" in out.stdout
    assert f" var-SpGEMM -> NNZ: {int(z['c_rp'][-1])}" in out.stdout
    assert re.search(r" var-SpGEMM SpGEMM: [0-9.e+-]+ ms", out.stdout)
    assert "Error: Matrix incorrect" not in out.stdout
    assert "could not load csr file" in out.stdout and "write csr file for future use" in out.stdout
    assert os.path.exists(str(mtx) + "d_.hicsr")
    out2 = run([str(mtx), str(ini)], tmp_path)      # second run loads the cache
    assert "successfully loaded: " in out2.stdout and "Error: Matrix incorrect" not in out2.stdout


def test_runspeck_compare_against_cusparse_on_mixed_classes(tmp_path):
    """R-MAT scale 11 and a banded matrix: lane-group sort, CTA sort and bitmap rows all occur;
    cuSPARSE must agree on every row length and column index in every iteration."""
    for name, A in (("rmat11", M.rmat(11, 16, seed=5)),
                    ("banded", M.banded_fem_like(n=1500, per_row=48, clusters=6, band=200, seed=3))):
        mtx = tmp_path / f"{name}.mtx"
        write_mtx(mtx, A)
        ini = tmp_path / "c.ini"
        ini.write_text("CompareResult=true\nIterationsWarmUp=1\nIterationsExecution=2\nTrackIndividualTimes=true\n")
        out = run([str(mtx), str(ini)], tmp_path)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "Error: Matrix incorrect" not in out.stdout, name
        assert "var-SpGEMM -> NNZ:" in out.stdout and "spGEMMNumeric" in out.stdout


def test_runspeck_rectangular_uses_transpose(tmp_path):
    """non-square A -> the driver multiplies A . A^T (DataLoader.cpp:64-69) via the csr2csc shim"""
    A = M.uniform_random(300, 120, 5, seed=8)
    mtx = tmp_path / "rect.mtx"
    write_mtx(mtx, A)
    ini = tmp_path / "c.ini"
    ini.write_text("CompareResult=true\nIterationsWarmUp=1\nIterationsExecution=1\n")
    out = run([str(mtx), str(ini)], tmp_path)
    assert out.returncode == 0, out.stdout + out.stderr
    S = A.to_scipy()
    C = (S @ S.T).tocsr()
    assert f"var-SpGEMM -> NNZ: {C.nnz}" in out.stdout
    assert "Error: Matrix incorrect" not in out.stdout


def test_runspeck_missing_file_message(tmp_path):
    out = run([str(tmp_path / "nope.mtx")], tmp_path)
    assert out.returncode != 0 and "could not load mtx file" in out.stdout


def test_runspeck_devices_key_runs_the_sharded_path(tmp_path):
    """`Devices=` (new ini key): A is cut into product-balanced slabs, one per listed device, the slabs of C are
    concatenated on the first device and checked against cuSPARSE.  On a one-GPU box the list repeats device 0
    (contexts may share a device); on a multi-GPU box it names distinct devices."""
    import torch
    ndev = torch.cuda.device_count()
    devices = ",".join(str(g if ndev >= 3 else 0) for g in range(3))
    A = M.rmat(11, 16, seed=5)
    mtx = tmp_path / "rmat11.mtx"
    write_mtx(mtx, A)
    ini = tmp_path / "config.ini"
    ini.write_text(f"CompareResult=true\nIterationsWarmUp=1\nIterationsExecution=2\nDevices={devices}\n")
    out = run([str(mtx), str(ini)], tmp_path)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "Error: Matrix incorrect" not in out.stdout and "ERROR" not in out.stdout, out.stdout
    import oracle
    rp, _, _ = oracle.spgemm(A.row_offsets, A.col_ids, A.data, A.row_offsets, A.col_ids, A.data, A.cols)
    assert f" var-SpGEMM -> NNZ: {int(rp[-1])}" in out.stdout
    assert re.search(r" var-SpGEMM SpGEMM: [0-9.e+-]+ ms", out.stdout)
    assert len(re.findall(r"device \d+: rows \[\d+, \d+\)", out.stdout)) == 3
    assert "3 devices" in out.stdout and "var-SpGEMM concat:" in out.stdout


def test_runspeck_gpu_convert_writes_the_same_cache(tmp_path):
    """GpuConvert=true: COO -> CSR of the parsed .mtx on the device; the .hicsr cache must be byte-identical to the
    one the host conversion writes (same order, duplicates kept)."""
    A = M.rmat(10, 8, seed=3)
    outs = {}
    for mode in ("false", "true"):
        d = tmp_path / mode
        d.mkdir()
        mtx = d / "m.mtx"
        write_mtx(mtx, A)
        ini = d / "config.ini"
        ini.write_text(f"CompareResult=true\nIterationsWarmUp=1\nIterationsExecution=1\nGpuConvert={mode}\n")
        out = run([str(mtx), str(ini)], d)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "Error: Matrix incorrect" not in out.stdout
        outs[mode] = open(str(mtx) + "d_.hicsr", "rb").read()
    assert outs["true"] == outs["false"]
