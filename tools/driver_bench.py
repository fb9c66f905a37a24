#!/usr/bin/env python
"""Compare the iteration drivers (graph chunked / graph while / stream) on a small, latency-bound problem.
    python tools/driver_bench.py [grid2d=1024]"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sparse_matrix_math_b200 as smm
from sparse_matrix_math_b200 import binding as B

n2 = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
A = smm.CSRMatrix.generate(B.GEN_POISSON2D, n2, n2)
n = A.rows
ones = smm.DeviceVector(n, np.ones(n, np.float32)); b = smm.DeviceVector(n); x = smm.DeviceVector(n)
A.spmv_dev(B.OP_ASSIGN, None, ones.ptr, b.ptr)
L = smm.lib()
for name, drv, ce in [("chunked/32", B.DRIVER_GRAPH_CHUNKED, 32), ("chunked/256", B.DRIVER_GRAPH_CHUNKED, 256), ("while", B.DRIVER_GRAPH_WHILE, 0), ("stream/256", B.DRIVER_STREAM, 256),
                      ("persistent", B.DRIVER_PERSISTENT, 0)]:
    for rep in range(2):
        x.zero()
        o, _ = B._options(B.REDUCE_FAST, drv, ce, 0)
        info = B._Info()
        B._check(L.smm_solve_cg_dev(A.handle, b.ptr, x.ptr, x.ptr, 1000, 0.0, C.byref(o), C.byref(info), None), "cg")
    print(f"{name:12s} iterations {info.iterations} {info.seconds_solve*1e6/info.iterations:7.2f} us/iteration  {info.iterations/info.seconds_solve:9.0f} it/s")
