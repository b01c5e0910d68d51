"""polar-codes-with-bch-kernel_b200: B200-native (sm_100a) Kaneko/BCH Monte-Carlo hot path.

The product is libpkb200.so (CUDA kernels + C ABI, csrc/) and the C++ host mirror of the
reference interface (host/).  This Python package is plumbing only: ctypes bindings for
tests and bench.py and the multi-GPU sweep driver (sweep.py).
"""
from .capi import *  # noqa: F401,F403
from .capi import Code, Kaneko, PkError, lib  # noqa: F401
