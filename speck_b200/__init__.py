"""speck_b200 -- B200-native CSR x CSR SpGEMM behind the spECK API.

Host-side Python mirror of the C ABI in include/speck_b200.h (ctypes).  The
compute path is the CUDA library speck_b200/csrc -> speck_b200/_lib/libspeck_b200.so;
there is no CPU fallback: importing `speck_b200.api` without the built library, or
calling it without a GPU, raises.
"""
from .matrices import HostCSR  # noqa: F401
