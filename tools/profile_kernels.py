#!/usr/bin/env python
"""Launch every hot kernel other than the stencil SpMV a few times on its BASELINE-sized input, for one ncu capture:
the irregular-row SpMV on the power-law matrix (config 4), the SGS sweeps on convection-diffusion 256^3 (config 3),
the fused vector kernels of a BiCGStab iteration and the reference-order dot product.

    ncu --set full --clock-control none --import-source on -k regex:'spmv_kernel|sgs_sweep|sgs_tile|dot_tree|vec_kernel' \
        -c 14 -o gpurun_out/r01e_kernels python tools/profile_kernels.py
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sparse_matrix_math_b200 as smm  # noqa: E402
from sparse_matrix_math_b200 import binding as B  # noqa: E402

L = smm.lib()
which = sys.argv[1] if len(sys.argv) > 1 else "all"

if which in ("all", "powerlaw"):
    A = smm.CSRMatrix.generate(B.GEN_POWERLAW, 8388608)
    x = smm.DeviceVector(A.rows); y = smm.DeviceVector(A.rows)
    B._check(L.smm_gen_xstar_dev(A.rows, 0, 0xB200, x.ptr, None), "x")
    for _ in range(3):
        A.spmv_dev(B.OP_ASSIGN, None, x.ptr, y.ptr)
    L.smm_sync()
    print("powerlaw spmv: rows", A.rows, "nnz", A.nnz, flush=True)
    del A, x, y

if which in ("all", "sgs"):
    A = smm.CSRMatrix.generate(B.GEN_CONVDIFF3D, 256, 256, 256, 0.5)
    M = A.getPreconditioner(smm.SolverPreconditioner.SYMMETRIC_GAUS_SEIDEL)
    n = A.rows
    rhs = smm.DeviceVector(n); x = smm.DeviceVector(n)
    B._check(L.smm_gen_xstar_dev(n, 0, 1, rhs.ptr, None), "x")
    for _ in range(2):
        M.apply_dev(rhs.ptr, x.ptr)
    L.smm_sync()
    print("sgs apply: rows", n, "levels", M.levels(), flush=True)
    out = C.c_float()
    for mode in (B.REDUCE_REFERENCE_TREE, B.REDUCE_FAST):
        B._check(L.smm_dot_dev(n, rhs.ptr, x.ptr, mode, C.byref(out), None), "dot")
    # two preconditioned BiCGStab iterations: every fused vector kernel of that solver
    b = smm.DeviceVector(n)
    A.spmv_dev(B.OP_ASSIGN, None, rhs.ptr, b.ptr)
    x.zero()
    o, h = B._options(B.REDUCE_FAST, B.DRIVER_STREAM, 0, 0)
    info = B._Info()
    B._check(L.smm_solve_bicgstab_dev(A.handle, M.handle, b.ptr, x.ptr, 2, 0.0, C.byref(o), C.byref(info), None), "bicgstab")
    print("bicgstab+sgs 2 iterations:", info.iterations, flush=True)

if which in ("all", "dots"):
    # the reference-order dot on a long vector: lane-per-node kernel (dot_tree_rows_kernel<16>) at 134 M elements
    n = 134217728
    a = smm.DeviceVector(n); b2 = smm.DeviceVector(n)
    B._check(L.smm_gen_xstar_dev(n, 0, 1, a.ptr, None), "x")
    B._check(L.smm_gen_xstar_dev(n, 0, 2, b2.ptr, None), "x")
    out = C.c_float()
    for _ in range(2):
        B._check(L.smm_dot_dev(n, a.ptr, b2.ptr, B.REDUCE_REFERENCE_TREE, C.byref(out), None), "dot")
    print("tree dot 134M:", out.value, flush=True)
