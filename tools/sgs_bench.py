#!/usr/bin/env python
"""Time SGSPreconditioner::apply on the GPU (wall clock around synchronised batches).
    python tools/sgs_bench.py [grid] [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sparse_matrix_math_b200 as smm
from sparse_matrix_math_b200 import binding as B

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ny = int(sys.argv[3]) if len(sys.argv) > 3 else grid
nz = int(sys.argv[4]) if len(sys.argv) > 4 else grid
A = smm.CSRMatrix.generate(B.GEN_CONVDIFF3D, grid, ny, nz, 0.5)
M = A.getPreconditioner(smm.SolverPreconditioner.SYMMETRIC_GAUS_SEIDEL)
n = A.rows
rhs = smm.DeviceVector(n); x = smm.DeviceVector(n)
B._check(smm.lib().smm_gen_xstar_dev(n, 0, 1, rhs.ptr, None), "x")
M.apply_dev(rhs.ptr, x.ptr)
smm.lib().smm_sync()
t = time.perf_counter()
for _ in range(reps):
    M.apply_dev(rhs.ptr, x.ptr)
smm.lib().smm_sync()
dt = (time.perf_counter() - t) / reps
bytes_apply = 8 * A.nnz + 44 * n
try:
    import ctypes
    st = (ctypes.c_ulonglong * 4)()
    smm.lib().smm_debug_line_stats(st)
    print(f"line schedule over {reps + 1} applies: lane-steps that missed an operand {st[0]}, their polls {st[1]}, gate polls {st[2]}")
except AttributeError:
    pass
print(f"grid {grid}x{ny}x{nz} levels {M.levels()} tile levels {M.tile_levels()} apply {dt*1e3:.3f} ms  ({dt*1e6/(2*M.levels()[0]):.2f} us per level)  {bytes_apply/dt/1e9:.0f} GB/s  ctas_per_sm={os.environ.get('SMM_B200_SGS_CTAS_PER_SM','4')}")
