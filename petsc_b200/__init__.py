"""petsc_b200 -- B200-native Krylov hot path behind PETSc's Mat/Vec/KSP/PC interface.

Product = libpetscb200.so (hand-written sm_100a kernels + C ABI, include/petscb200.h), bound into PETSc by
petsc_plugin/libpetscb200plugin.so.  This Python package is only the ctypes door to the C ABI for tests and bench.py; it has
no CPU fallback and raises if the CUDA library is missing.
"""
from . import _capi  # noqa: F401

__all__ = ["_capi"]
