#!/usr/bin/env python
"""Run the five BASELINE.json configurations at full size on one B200 and report iterations, rate, residuals and the
per-iteration roofline fraction (profiles/r01_config_report.json).  Not the driver's benchmark (that is bench.py,
config 5); this is the per-config evidence DESIGN.md cites.

    python tools/config_report.py [--configs 1,2,3,4,5] [--modes fast,tree] [--out profiles/r01_config_report.json]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import sparse_matrix_math_b200 as smm  # noqa: E402
from sparse_matrix_math_b200 import binding as B  # noqa: E402

PEAK = 6546.2
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def solve_dev(fn, A, pre, b, x, maxit, eps, mode, driver=B.DRIVER_AUTO, check_every=0, hist=0):
    o, h = B._options(mode, driver, check_every, hist)
    info = B._Info()
    L = smm.lib()
    if fn == "cg":
        rc = L.smm_solve_cg_dev(A.handle, b.ptr, x.ptr, x.ptr, maxit, eps, C.byref(o), C.byref(info), None)
    elif fn == "bicgsym":
        rc = L.smm_solve_bicgsym_dev(A.handle, b.ptr, x.ptr, maxit, eps, C.byref(o), C.byref(info), None)
    elif fn == "cgs":
        rc = L.smm_solve_cgs_dev(A.handle, b.ptr, x.ptr, maxit, eps, C.byref(o), C.byref(info), None)
    else:
        rc = L.smm_solve_bicgstab_dev(A.handle, None if pre is None else pre.handle, b.ptr, x.ptr, maxit, eps, C.byref(o), C.byref(info), None)
    B._check(rc, fn)
    return B.SolveInfo(info, h)


def rhs(A, kind):
    n = A.rows
    xs = smm.DeviceVector(n)
    if kind == "ones":
        xs.upload(np.ones(n, np.float32))
    else:
        B._check(smm.lib().smm_gen_xstar_dev(n, 0, 0xB200, xs.ptr, None), "xstar")
    b = smm.DeviceVector(n)
    A.spmv_dev(B.OP_ASSIGN, None, xs.ptr, b.ptr)
    return xs, b


def true_residual(A, b, x):
    r = smm.DeviceVector(A.rows)
    A.spmv_dev(B.OP_SUB, b.ptr, x.ptr, r.ptr)
    rr = C.c_float()
    B._check(smm.lib().smm_dot_dev(A.rows, r.ptr, r.ptr, B.REDUCE_FAST, C.byref(rr), None), "dot")
    return float(np.sqrt(rr.value))


def run(name, fn, A, pre, rhs_kind, eps, modes, bytes_per_it, ref_iters=None, maxit=-1):
    out = {"config": name, "solver": fn + ("+" + type(pre).__name__.replace("Preconditioner", "").lower() if pre is not None else ""), "rows": A.rows, "nnz": A.nnz, "eps": eps,
           "reference_mt_iterations": ref_iters, "runs": []}
    xs, b = rhs(A, rhs_kind)
    for mode in modes:
        m = {"fast": B.REDUCE_FAST, "tree": B.REDUCE_REFERENCE_TREE}[mode]
        x = smm.DeviceVector(A.rows)
        x.zero()
        t = time.perf_counter()
        info = solve_dev(fn, A, pre, b, x, maxit, eps, m)
        wall = time.perf_counter() - t
        xh = x.download()
        xsh = xs.download()
        rate = info.iterations / info.seconds_solve if info.seconds_solve > 0 else None
        rec = {"mode": mode, "status": info.status.name, "iterations": info.iterations, "solver_residual": info.residual,
               "seconds_solve": info.seconds_solve, "wall_s": wall, "it_per_s": rate,
               "true_residual_l2": true_residual(A, b, x), "max_abs_error": float(np.max(np.abs(xh - xsh))) if np.all(np.isfinite(xh)) else None,
               "kernel_launches": info.kernel_launches}
        if rate and bytes_per_it:
            rec["iteration_gbs"] = bytes_per_it * rate / 1e9
            rec["frac_of_measured_peak"] = rec["iteration_gbs"] / PEAK
        out["runs"].append(rec)
        print(json.dumps({"config": name, **rec}), flush=True)
    return out


def reference_iterations():
    """Iteration counts of the reference's multithreaded build at full size (recorded by tests/golden/make_fullsize_golden.py)."""
    try:
        d = json.load(open(os.path.join(ROOT, "tests", "golden", "fullsize_reference.json")))
        return {k: v["iterations"] for k, v in d.items()}
    except Exception:
        return {}


def main():
    REF = reference_iterations()
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--modes", default="fast,tree")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_config_report.json"))
    args = ap.parse_args()
    todo = [int(c) for c in args.configs.split(",")]
    modes = args.modes.split(",")
    report = {"device": smm.device_info(), "peak_gbs": PEAK, "configs": []}
    if 1 in todo:
        A = smm.CSRMatrix.generate(B.GEN_POISSON2D, 1024, 1024)
        report["configs"].append(run("1: CG, 2D 5-point Poisson 1024^2, eps 1e-6, b=A*1", "cg", A, None, "ones", 1e-6, modes,
                                     8 * A.nnz + 48 * A.rows, ref_iters=REF.get("1")))
        del A
    if 2 in todo:
        A = smm.CSRMatrix.generate(B.GEN_CONVDIFF3D, 128, 128, 128, 0.5)
        report["configs"].append(run("2: BiCGStab, 3D convection-diffusion 128^3, eps 1e-6, b=A*x*", "bicgstab", A, None, "xstar", 1e-6, modes,
                                     16 * A.nnz + 84 * A.rows, ref_iters=REF.get("2")))
        M = A.getPreconditioner(smm.SolverPreconditioner.SYMMETRIC_GAUS_SEIDEL)
        report["configs"].append(run("2': BiCGStab+SGS, 128^3, eps 1e-6", "bicgstab", A, M, "xstar", 1e-6, modes,
                                     2 * (16 * A.nnz + 64 * A.rows) + 16 * A.nnz + 84 * A.rows, ref_iters=REF.get("2s")))
        del M, A
    if 3 in todo:
        A = smm.CSRMatrix.generate(B.GEN_CONVDIFF3D, 256, 256, 256, 0.5)
        t = time.perf_counter()
        M = A.getPreconditioner(smm.SolverPreconditioner.SYMMETRIC_GAUS_SEIDEL)
        print(json.dumps({"sgs_analysis_s": time.perf_counter() - t, "levels": M.levels()}), flush=True)
        report["configs"].append(run("3: BiCGStab+SGS, 3D convection-diffusion 256^3, eps 1e-6, b=A*x*", "bicgstab", A, M, "xstar", 1e-6, modes,
                                     2 * (16 * A.nnz + 64 * A.rows) + 16 * A.nnz + 84 * A.rows, ref_iters=REF.get("3")))
        del M
        # extension: the zero-fill incomplete LU the reference only sketches (dead code there), same sweeps on its own factor
        t = time.perf_counter()
        M = A.getPreconditioner(smm.SolverPreconditioner.ILU0)
        print(json.dumps({"ilu0_setup_s": time.perf_counter() - t, "code": M.init_code, "tile_levels": M.tile_levels()}), flush=True)
        report["configs"].append(run("3': BiCGStab+ILU0 (extension), 3D convection-diffusion 256^3, eps 1e-6, b=A*x*", "bicgstab", A, M, "xstar", 1e-6, ["fast"],
                                     2 * (16 * A.nnz + 64 * A.rows) + 16 * A.nnz + 84 * A.rows))
        del M, A
    if 4 in todo:
        A = smm.CSRMatrix.generate(B.GEN_POWERLAW, 8388608)
        for fn, bpi in (("cgs", 16 * A.nnz + 80 * A.rows), ("bicgsym", 8 * A.nnz + 48 * A.rows)):
            report["configs"].append(run(f"4: {fn} sweep, power-law 8.4M rows (fixed 8 iterations)", fn, A, None, "xstar", 0.0, ["fast"], bpi, maxit=8))
        del A
    if 5 in todo:
        A = smm.CSRMatrix.generate(B.GEN_CONVDIFF3D, 512, 512, 512, 0.0)
        report["configs"].append(run("5: CG, 3D 7-point Poisson 512^3, eps 1e-6, b=A*1", "cg", A, None, "ones", 1e-6, modes,
                                     8 * A.nnz + 48 * A.rows, ref_iters=REF.get("5")))
        del A
    json.dump(report, open(args.out, "w"), indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
