#!/usr/bin/env python
"""Set-up time of getPreconditioner() (host analysis + layout + upload):  python tools/precond_setup_time.py [grid]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sparse_matrix_math_b200 as smm
from sparse_matrix_math_b200 import binding as B
g = int(sys.argv[1]) if len(sys.argv) > 1 else 256
A = smm.CSRMatrix.generate(B.GEN_CONVDIFF3D, g, g, g, 0.5)
smm.lib().smm_sync()
for kind, name in ((smm.SolverPreconditioner.SYMMETRIC_GAUS_SEIDEL, "SGS"), (smm.SolverPreconditioner.ILU0, "ILU(0)")):
    t = time.perf_counter(); M = A.getPreconditioner(kind); dt = time.perf_counter() - t
    print(f"{g}^3 {name} set-up {dt:.2f} s  tile levels {M.tile_levels()}  ({os.cpu_count()} host threads)")
    del M
