#!/usr/bin/env python
"""ConjugateGradient on z-slabs of the 512^3 Poisson grid on ONE GPU: per-iteration time and per-kernel times against the slab
thickness -- the fixed and the per-row part of an iteration (what bounds strong scaling: a rank of an N-GPU run owns such a slab).
    python tools/cg_size_sweep.py [nz ...]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sparse_matrix_math_b200 as smm
from sparse_matrix_math_b200 import binding as B
L = smm.lib()
sizes = [int(a) for a in sys.argv[1:]] or [16, 32, 64, 128, 256, 512]
for nz in sizes:
    A = smm.CSRMatrix.generate(B.GEN_CONVDIFF3D, 512, 512, nz, 0.0)
    n = A.rows
    ones = smm.DeviceVector(n); ones.upload(__import__("numpy").ones(n, "float32"))
    b = smm.DeviceVector(n); x = smm.DeviceVector(n)
    A.spmv_dev(B.OP_ASSIGN, None, ones.ptr, b.ptr)
    best = None
    for rep in range(3):
        x.zero()
        o, _ = B._options(B.REDUCE_FAST, B.DRIVER_AUTO, 200, 0)
        info = B._Info()
        B._check(L.smm_solve_cg_dev(A.handle, b.ptr, x.ptr, x.ptr, 200, 0.0, C.byref(o), C.byref(info), None), "cg")
        t = info.seconds_solve / info.iterations
        best = t if best is None or t < best else best
    ms = [C.c_float(), C.c_float(), C.c_float()]
    B._check(L.smm_profile_cg_iteration(A.handle, 50, C.byref(ms[0]), C.byref(ms[1]), C.byref(ms[2]), None), "profile")
    k = [m.value * 1e3 for m in ms]
    print(f"512x512x{nz:<4d} rows {n:>10d}  iteration {best*1e6:8.1f} us  kernels back to back: spmv+dot {k[0]:7.1f}  r {k[1]:6.1f}  p,x {k[2]:6.1f}  sum {sum(k):8.1f} us", flush=True)
    del A, b, x, ones
