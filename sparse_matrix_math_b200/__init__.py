"""sparse_matrix_math_b200 -- B200-native Krylov solve path behind the API of vasil-pashov/sparse_matrix_math.

The product is libsmm_b200.so (hand-written sm_100a CUDA behind the C ABI of include/smm_b200.h) and the
drop-in C++17 header include/sparse_matrix_math.h (namespace SMM).  This Python package is a thin ctypes
binding over the same C ABI, used by tests/ and bench.py; it mirrors the reference's names
(TripletMatrix, CSRMatrix, ConjugateGradient, BiCGSymmetric, ConjugateGradientSquared, BiCGStab,
SolverStatus, loadMatrix).  There is no CPU fallback: importing works without a GPU (so the build and
symbol checks run anywhere), every compute call fails loudly without one.
"""
from .binding import (  # noqa: F401
    ABI_SYMBOLS,
    BiCGStab,
    BiCGSymmetric,
    ConjugateGradient,
    ConjugateGradientSqared,
    ConjugateGradientSquared,
    CSRMatrix,
    DeviceVector,
    IC0Preconditioner,
    ILU0Preconditioner,
    JacobiPreconditioner,
    MatrixLoadStatus,
    SGSPreconditioner,
    SmmError,
    SolveInfo,
    SolverPreconditioner,
    SolverStatus,
    TripletMatrix,
    device_info,
    dot,
    kernel_launch_count,
    lib,
    lib_path,
    loadMatrix,
    REDUCE_FAST,
    REDUCE_REFERENCE_SERIAL,
    REDUCE_REFERENCE_TREE,
    DRIVER_AUTO,
    DRIVER_GRAPH_CHUNKED,
    DRIVER_GRAPH_WHILE,
    DRIVER_STREAM,
    DRIVER_PERSISTENT,
)
