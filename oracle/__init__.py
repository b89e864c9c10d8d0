"""CPU oracle for the spECK SpGEMM hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  See oracle/spgemm_oracle.c for the reference
file:line each function restates and for the parity-pinning status.
"""
from .oracle import (  # noqa: F401
    build, row_products, spgemm, symbolic, compare, num_threads, set_threads,
)
