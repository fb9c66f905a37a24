#!/usr/bin/env python
"""SGS apply on the 2D 5-point Poisson (tile-level vs row-level schedule):  python tools/sgs2d_bench.py [n] [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sparse_matrix_math_b200 as smm
from sparse_matrix_math_b200 import binding as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
A = smm.CSRMatrix.generate(B.GEN_POISSON2D, n, n)
M = A.getPreconditioner(smm.SolverPreconditioner.SYMMETRIC_GAUS_SEIDEL)
rhs = smm.DeviceVector(A.rows); x = smm.DeviceVector(A.rows)
B._check(smm.lib().smm_gen_xstar_dev(A.rows, 0, 1, rhs.ptr, None), "x")
M.apply_dev(rhs.ptr, x.ptr)
smm.lib().smm_sync()
t = time.perf_counter()
for _ in range(reps):
    M.apply_dev(rhs.ptr, x.ptr)
smm.lib().smm_sync()
dt = (time.perf_counter() - t) / reps
print(f"2d {n}x{n} levels {M.levels()} tile levels {M.tile_levels()} apply {dt*1e3:.3f} ms  tiles={os.environ.get('SMM_B200_SGS_TILES','1')}")
