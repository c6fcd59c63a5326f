"""cudecomp_b200 -- B200-native pencil-decomposition transpose engine behind the cuDecomp C ABI.

The product is cudecomp_b200/lib/libcudecomp.so (C++/CUDA, sources in csrc/, headers in ../include). This package is
only the ctypes view of its C ABI, used by the tests and the benchmark:

    from cudecomp_b200 import capi as cd
    cd.MPI_Init(); res, handle = cd.cudecompInit(cd.MPI_COMM_WORLD)

Importing `capi` fails loudly when the library has not been built; there is no CPU fallback.
"""
from .build import build_library, LIB_PATH  # noqa: F401

__all__ = ["build_library", "LIB_PATH"]
