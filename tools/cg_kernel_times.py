#!/usr/bin/env python
"""Per-kernel device times of one fused CG iteration on a generated stencil (smm_profile_cg_iteration).
    python tools/cg_kernel_times.py 2d 1024   |   python tools/cg_kernel_times.py 3d 128"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sparse_matrix_math_b200 as smm
from sparse_matrix_math_b200 import binding as B
kind, n = sys.argv[1], int(sys.argv[2])
A = smm.CSRMatrix.generate(B.GEN_POISSON2D, n, n) if kind == "2d" else smm.CSRMatrix.generate(B.GEN_CONVDIFF3D, n, n, n, 0.0)
ms = [C.c_float(), C.c_float(), C.c_float()]
for _ in range(2):
    B._check(smm.lib().smm_profile_cg_iteration(A.handle, 200, C.byref(ms[0]), C.byref(ms[1]), C.byref(ms[2]), None), "profile")
b_spmv = 8 * A.nnz + 12 * A.rows; b_xr = 24 * A.rows; b_p = 12 * A.rows
print(f"{kind} {n}: rows {A.rows} nnz {A.nnz}  spmv+dot {ms[0].value*1e3:.1f} us ({b_spmv/ms[0].value/1e6:.0f} GB/s)  "
      f"xr {ms[1].value*1e3:.1f} us ({b_xr/ms[1].value/1e6:.0f} GB/s)  p {ms[2].value*1e3:.1f} us ({b_p/ms[2].value/1e6:.0f} GB/s)  "
      f"sum {sum(m.value for m in ms)*1e3:.1f} us")
