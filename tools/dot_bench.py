#!/usr/bin/env python
"""Time the dot-product kernels of every reduction mode on resident vectors:  python tools/dot_bench.py [n ...]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sparse_matrix_math_b200 as smm
from sparse_matrix_math_b200 import binding as B
L = smm.lib()
for n in [int(a) for a in sys.argv[1:]] or [1 << 21, 1 << 24]:
    a = smm.DeviceVector(n); b = smm.DeviceVector(n)
    B._check(L.smm_gen_xstar_dev(n, 0, 1, a.ptr, None), "x")
    B._check(L.smm_gen_xstar_dev(n, 0, 2, b.ptr, None), "x")
    out = C.c_float()
    for name, mode, x, y in (("fast a.b", B.REDUCE_FAST, a, b), ("tree a.b", B.REDUCE_REFERENCE_TREE, a, b),
                             ("serial a.a (sum of squares, exact parallel)", B.REDUCE_REFERENCE_SERIAL, a, a),
                             ("serial a.b (one thread)", B.REDUCE_REFERENCE_SERIAL, a, b)):
        reps = 3 if "one thread" in name else 10
        B._check(L.smm_dot_dev(n, x.ptr, y.ptr, mode, C.byref(out), None), "dot")
        L.smm_sync()
        t = time.perf_counter()
        for _ in range(reps):
            B._check(L.smm_dot_dev(n, x.ptr, y.ptr, mode, C.byref(out), None), "dot")
        L.smm_sync()
        dt = (time.perf_counter() - t) / reps
        print(f"n {n:10d}  {name:46s} {dt*1e6:10.1f} us  {8*n/dt/1e9 if x is not y else 4*n/dt/1e9:8.1f} GB/s  value {out.value:.6g}")
