"""quack_b200 -- B200-native per-read statistics accumulation for quack (IGBB/quack).

The product is the C-ABI shared library built from quack_b200/csrc (CUDA for sm_100a + C/C++ host
code) and declared in include/quack_b200.h, plus the `quack` host program built from
quack_b200/host.  This Python package is only the thin ctypes mirror the tests and bench.py use;
it never computes statistics itself and raises if the CUDA library is missing.
"""
from .build import build, lib_path  # noqa: F401

__all__ = ["build", "lib_path"]
