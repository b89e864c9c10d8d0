"""GPU: seeded randomised parity -- many small A.B products of random shape, density and structure under random
combinations of the library's kernel-selection options, each compared with the CPU oracle (indices bit-exact, values
1e-6; bit-exact values in deterministic mode)."""
import numpy as np
import pytest

from speck_b200 import matrices as M
from speck_b200.matrices import HostCSR
from helpers import assert_csr_equal, gpu_multiply, oracle_multiply

pytestmark = pytest.mark.gpu

OPTIONS = {"sort_max": [16384, 8192, 1024, 64], "rank_path": [0, 1], "rank_map": [0, 1], "flat_sym": [0, 1],
           "flat_e": [8, 16], "seg_num": [0, 1], "dense_seq": [0, 1, 2], "test_set": [0, 1], "spin_wait": [0, 1],
           "tiered_analysis": [0, 1], "col_direct": [0, 0, 1, 2, 3], "big_split": [0, 1, 2, 3], "sym_mix": [0, 0, 1, 3],
           "flat_min_class": [8, 8, 7, 6], "narrow_groups": [0, 1, 2, 3],
           "deterministic": [0, 0, 0, 1]}
DEFAULTS = {"sort_max": 16384, "rank_path": 1, "rank_map": 1, "flat_sym": 1, "flat_e": 8, "seg_num": 0, "dense_seq": 1,
            "test_set": 1, "spin_wait": 1, "tiered_analysis": 0, "col_direct": 0, "big_split": 0, "sym_mix": 0, "flat_min_class": 8, "narrow_groups": 2,
            "deterministic": 0}


def random_matrix(rng, rows, cols, kind):
    if kind == "uniform":
        d = rng.choice([0.5, 2, 8, 30])
        m = int(rows * d) + 1
        r, c = rng.integers(0, rows, m), rng.integers(0, cols, m)
    elif kind == "powerlaw":     # a few hub rows and hub columns
        m = rows * 6 + 1
        r = (rows * rng.random(m) ** 3).astype(np.int64)
        c = (cols * rng.random(m) ** 2.5).astype(np.int64)
    elif kind == "banded":
        per = int(rng.choice([8, 24, 60]))
        band = int(rng.choice([20, 200, 1500]))
        r = np.repeat(np.arange(rows), per)
        c = np.clip(r * cols // max(rows, 1) + rng.integers(-band, band + 1, r.size), 0, cols - 1)
    else:                        # "blocks": dense blocks -> heavy folding
        nb = max(1, rows // 40)
        r = np.repeat(np.arange(rows), 30)
        c = np.clip((r // 40) * (cols // nb) + rng.integers(0, min(60, cols), r.size), 0, cols - 1)
    keep = rng.random(r.size) > rng.choice([0.0, 0.3])          # thin out, creates empty rows
    if rng.random() < 0.3:
        keep &= (r % 5 != 2)                                      # whole empty rows
    return M.from_coo(rows, cols, r[keep], c[keep], seed=int(rng.integers(1 << 30)))


@pytest.mark.parametrize("seed", range(40))
def test_random_products_under_random_options(ctx, seed):
    rng = np.random.default_rng(1000 + seed)
    opts = {k: int(rng.choice(v)) for k, v in OPTIONS.items()}
    try:
        for k, v in opts.items():
            ctx.set_option(k, v)
        for _ in range(3):
            rows = int(rng.choice([1, 7, 60, 500, 3000]))
            inner = int(rng.choice([1, 9, 300, 2500]))
            cols = int(rng.choice([1, 33, 700, 5000, 1 << 21]))
            A = random_matrix(rng, rows, inner, str(rng.choice(["uniform", "powerlaw", "banded", "blocks"])))
            B = random_matrix(rng, inner, cols, str(rng.choice(["uniform", "powerlaw", "banded", "blocks"])))
            want = oracle_multiply(A, B)
            got, _ = gpu_multiply(ctx, A, B)
            what = f"seed {seed} opts {opts} A {A.rows}x{A.cols}/{A.nnz} B {B.rows}x{B.cols}/{B.nnz}"
            if A.nnz == 0 or B.nnz == 0:
                # the device entry follows the reference (source/GPU/Multiply.cu:67-70): only nnz is reset, the
                # shape of a fresh C stays untouched
                assert got.nnz == 0 and want.nnz == 0, what
                continue
            assert_csr_equal(got, want, rtol=1e-6, what=what)
            if opts["deterministic"]:
                np.testing.assert_array_equal(got.data, want.data, err_msg=what)
    finally:
        for k, v in DEFAULTS.items():
            ctx.set_option(k, v)
