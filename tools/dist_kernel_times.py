#!/usr/bin/env python
"""Per-kernel times of the multi-GPU CG iteration on every rank's slab, peers not taking part (SMM_B200_DIST_DEBUG=7), next to the
single-GPU kernels on the same rows: what the distributed forms of the kernels cost by themselves.
    SMM_B200_DIST_DEBUG=7 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_kernel_times.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sparse_matrix_math_b200 as smm
from sparse_matrix_math_b200 import binding as B, dist as D

rank, world, local = D.init_process_group()
torch.cuda.set_device(local)
L = smm.lib()
B._check(L.smm_set_device(local), "dev")
grid = 512
rows = grid ** 3
rb, re = D.row_partition(rows, world, align=grid * grid)[rank]
A = D.generate_rows(B.GEN_CONVDIFF3D, grid, grid, grid, 0.0, rb, re)
M = D.DistMatrix(A, rows, rb, re, rank, world, D.all_gather_object)
L.smm_dist_profile_cg_iteration.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p]
ms = [C.c_float(), C.c_float(), C.c_float()]
B._check(L.smm_dist_profile_cg_iteration(M.handle, 50, C.byref(ms[0]), C.byref(ms[1]), C.byref(ms[2]), None), "dist profile")
d = [m.value * 1e3 for m in ms]
# the same rows as a stand-alone matrix (nz / world planes): the single-GPU kernels
S = smm.CSRMatrix.generate(B.GEN_CONVDIFF3D, grid, grid, grid // world, 0.0)
B._check(L.smm_profile_cg_iteration(S.handle, 50, C.byref(ms[0]), C.byref(ms[1]), C.byref(ms[2]), None), "profile")
s1 = [m.value * 1e3 for m in ms]
print(f"rank {rank}/{world} rows {re - rb}: dist kernels spmv+dot {d[0]:7.1f}  r {d[1]:6.1f}  p,x(+push) {d[2]:6.1f}  sum {sum(d):7.1f} us | single-GPU kernels {s1[0]:7.1f} {s1[1]:6.1f} {s1[2]:6.1f} sum {sum(s1):7.1f} us", flush=True)
torch.distributed.barrier()
