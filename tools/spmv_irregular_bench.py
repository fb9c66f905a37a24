#!/usr/bin/env python
"""Time rMult on the power-law matrix of BASELINE config 4 (irregular-row kernel):  python tools/spmv_irregular_bench.py [rows] [reps]
Knobs (environment): SMM_B200_SPMV_DEPTH = 2 | 4, SMM_B200_SPMV_HINTS = 0 | 1."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sparse_matrix_math_b200 as smm
from sparse_matrix_math_b200 import binding as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8388608
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
A = smm.CSRMatrix.generate(B.GEN_POWERLAW, n)
x = smm.DeviceVector(n); y = smm.DeviceVector(n)
B._check(smm.lib().smm_gen_xstar_dev(n, 0, 0xB200, x.ptr, None), "x")
for _ in range(3):
    A.spmv_dev(B.OP_ASSIGN, None, x.ptr, y.ptr)
smm.lib().smm_sync()
t = time.perf_counter()
for _ in range(reps):
    A.spmv_dev(B.OP_ASSIGN, None, x.ptr, y.ptr)
smm.lib().smm_sync()
dt = (time.perf_counter() - t) / reps
nbytes = 8 * A.nnz + 4 * (n + 1) + 4 * n + 4 * n
print(f"powerlaw rows {n} nnz {A.nnz}: rMult {dt*1e3:.3f} ms  {nbytes/dt/1e9:.0f} GB/s algorithmic  depth={os.environ.get('SMM_B200_SPMV_DEPTH','2')} hints={os.environ.get('SMM_B200_SPMV_HINTS','0')}")
