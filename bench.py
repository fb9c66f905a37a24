#!/usr/bin/env python
"""bench.py -- headline benchmark of the Krylov hot path (BASELINE.json).

Workload (N = 1 and N > 1): BASELINE.json configs[4], the configuration the metric's target is quoted on --
ConjugateGradient, float, 3D 7-point Poisson 512^3 (134,217,728 rows, 937,951,232 stored entries), b = A*1, x0 = 0,
row-sharded over the N GPUs of one box (strong scaling: the total problem is fixed).  A "step" is one solver call
that executes exactly --iters CG iterations (eps = 0 never passes the stopping test, maxIterations = --iters), so
metric = Krylov iterations per second = K * iters / time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--grid 512] [--iters 50]

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for how every field is obtained.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "krylov_iterations_per_sec"
UNIT = "it/s"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


TRAFFIC_FILES = ("r02e_spmv_rows_kernel_full.txt", "r02_spmv_rows_kernel_full.txt", "r01h_spmv_rows_kernel_full.txt")
TRAFFIC_SOURCE = "profiles/%s: committed ncu --set full capture of this kernel in this command (a constant, not re-measured in this run)"


def traffic_file():
    for name in TRAFFIC_FILES:
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            return p
    return None


def ncu_traffic_bytes(kernel_prefix, grid):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed `ncu --set full` capture of this
    same command (the newest of TRAFFIC_FILES under profiles/, 512^3); None for any other problem size."""
    p = traffic_file()
    if grid != 512 or p is None:
        return None
    rd = wr = None
    unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    for line in open(p):
        f = line.split()
        if len(f) >= 3 and f[0] == "dram__bytes_read.sum" and rd is None:
            rd = float(f[1]) * unit.get(f[2], 1.0)
        if len(f) >= 3 and f[0] == "dram__bytes_write.sum" and wr is None:
            wr = float(f[1]) * unit.get(f[2], 1.0)
    return None if rd is None or wr is None else rd + wr


def stencil_nnz(n):
    return 7 * n ** 3 - 6 * n ** 2


def bytes_cg_iteration(rows, nnz):
    """fused minimum of one CG iteration (SpMV + dot | x,r update | p update)."""
    return 8 * nnz + 48 * rows + 4


def bytes_cg_iteration_two_pass(rows, nnz):
    """What the iteration moves since round 2: SpMV + dot (8 nnz + 10 n: 16-bit row starts) | r update + r.r (12 n) | x and p update,
    p read once (20 n)."""
    return bytes_spmv_dot_moved(rows, nnz) + 32 * rows


def bytes_spmv_dot(rows, nnz):
    """The dominant kernel: Ap = A p with p.Ap in its epilogue: values+positions, start, p (gathered once), Ap."""
    return 8 * nnz + 4 * (rows + 1) + 4 * rows + 4 * rows


def bytes_spmv_dot_moved(rows, nnz):
    """What the rows kernel reads and writes since it takes the row starts from start16 (2 bytes per row, 257 entries per group of
    256 rows, relative to the group's staging window) instead of start[] (4 bytes per row)."""
    return 8 * nnz + 2 * (rows + rows // 256 + 1) + 4 * rows + 4 * rows


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                if t0 - 0.1 <= ts <= t1 + 0.3:
                    sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
                    for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (oracle/_ref, SMM_MULTITHREADING build) on the host cores
# ----------------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _cpu_libs():
    """The CPU arm always uses every host core this process may run on: torch.distributed.run exports OMP_NUM_THREADS=1 to
    its ranks, which is a default for GPU workers, not a statement about the reference's parallel path."""
    import oracle_lib as ol
    n = host_cores()
    olib = ol.oracle()
    olib.smm_oracle_set_threads.argtypes = [C.c_int]
    olib.smm_oracle_set_threads(n)
    if ol.ref_available():
        rlib = ol.ref(1)
        rlib.smm_ref_set_threads.argtypes = [C.c_int]
        rlib.smm_ref_set_threads(n)
        return ol, rlib, "reference"
    return ol, None, "port"


def cpu_problem(grid):
    """Build the grid^3 Poisson CSR on the host inside the reference library (no std::map); returns closures."""
    import oracle_lib as ol
    _, rlib, kind = _cpu_libs()
    olib = ol.oracle()
    olib.smm_oracle_stencil_nnz.restype = C.c_int64
    olib.smm_oracle_stencil_nnz.argtypes = [C.c_int] * 4
    olib.smm_oracle_gen_stencil.argtypes = [C.c_int] * 4 + [C.c_float] * 3 + [C.c_void_p] * 3
    rows = grid ** 3
    nnz = olib.smm_oracle_stencil_nnz(grid, grid, grid, 1)
    if rlib is not None:
        rlib.smm_ref_alloc_int.restype = C.c_void_p
        rlib.smm_ref_alloc_int.argtypes = [C.c_int64]
        rlib.smm_ref_alloc_float.restype = C.c_void_p
        rlib.smm_ref_alloc_float.argtypes = [C.c_int64]
        rlib.smm_ref_csr_adopt.restype = C.c_void_p
        rlib.smm_ref_csr_adopt.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        start = rlib.smm_ref_alloc_int(rows + 1)
        pos = rlib.smm_ref_alloc_int(nnz)
        val = rlib.smm_ref_alloc_float(nnz)
        olib.smm_oracle_gen_stencil(grid, grid, grid, 1, -1.0, 6.0, -1.0, start, pos, val)
        h = rlib.smm_ref_csr_adopt(rows, rows, start, pos, val)
        ones = np.ones(rows, np.float32)
        b = np.zeros(rows, np.float32)
        rlib.smm_ref_spmv(h, 0, None, ones, b)
        del ones

        def run(iters):
            x = np.zeros(rows, np.float32)
            t = time.perf_counter()
            st = rlib.smm_ref_cg(h, b, x, x, iters, 0.0)
            dt = time.perf_counter() - t
            assert st == 2, st       # MAX_ITERATIONS_REACHED: exactly `iters` iterations ran
            return dt

        threads = rlib.smm_ref_threads()
        return run, threads, kind, lambda: rlib.smm_ref_csr_destroy(h)
    # port: the C restatement (OpenMP)
    start = np.zeros(rows + 1, np.int32); pos = np.zeros(nnz, np.int32); val = np.zeros(nnz, np.float32)
    olib.smm_oracle_gen_stencil(grid, grid, grid, 1, -1.0, 6.0, -1.0, start.ctypes.data_as(C.c_void_p), pos.ctypes.data_as(C.c_void_p), val.ctypes.data_as(C.c_void_p))
    m = ol.CSR(rows, rows, start, pos, val, 0)
    b = ol.spmv(m, 0, None, np.ones(rows, np.float32))

    def run(iters):
        t = time.perf_counter()
        o = ol.solve("cg", m, b, np.zeros(rows, np.float32), iters, 0.0, 1)
        dt = time.perf_counter() - t
        assert o["status"] == 2
        return dt

    return run, olib.smm_oracle_threads(), kind, lambda: None


def pick_cpu_grid(grid):
    """The full problem needs ~12 GB of host memory per 512^3; fall back to a smaller cube if the box is small."""
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 32 << 30
    need = lambda g: 8 * stencil_nnz(g) + 4 * 10 * g ** 3
    g = grid
    while g > 64 and need(g) * 1.3 > avail:
        g //= 2
    return g


def cpu_baseline(grid, budget_s=15.0):
    g = pick_cpu_grid(grid)
    run, threads, kind, free = cpu_problem(g)
    t2 = run(2)                                   # includes r0 = b - A x0 and the first dot
    t1 = run(1)
    per_it = max(t2 - t1, 1e-9)
    k = int(max(3, min(40, budget_s / per_it)))
    tk = run(k)
    rate = (k - 1) / max(tk - t1, 1e-9)           # marginal iterations: start-up cost excluded, as on the GPU side
    free()
    scale = (g ** 3) / float(grid ** 3)
    sample = f"CG {g}^3 Poisson, {k} iterations (marginal rate over iterations 2..{k})"
    if g != grid:
        sample += f"; host memory too small for {grid}^3: rate scaled by rows ratio {scale:.4f}"
    return {"value": rate * scale, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample}


def workload_config(grid, iters):
    """The keys both arms share (the driver compares them): the workload, not how an arm executes it."""
    return {"workload": f"ConjugateGradient float, 3D 7-point Poisson {grid}^3, b=A*1, x0=0 (BASELINE configs[4])",
            "grid": grid, "rows": grid ** 3, "nnz": stencil_nnz(grid), "iterations_per_step": iters, "eps": 0.0}


def run_reference(args):
    """The reference's own ConjugateGradient (oracle/_ref, SMM_MULTITHREADING build) on all host cores.  Same workload as
    the GPU arm (a step = one solver call of --iters iterations); each timed step executes a BOUNDED SAMPLE of it -- k
    iterations, k sized so that steps + warmup finish in ~2.5 minutes -- and the value is the marginal iteration rate
    (start-up r0 = b - A x0, measured once, subtracted), which is what a --iters-iteration step sustains."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    g = pick_cpu_grid(args.grid)
    run, threads, kind, free = cpu_problem(g)
    t2, t1 = run(2), run(1)
    t1 = min(t1, run(1))                                      # start-up: r0, p, r.r and one iteration
    per_it = max(t2 - t1, 1e-6)
    k = int(max(3, min(args.iters, 150.0 / ((args.steps + args.warmup) * per_it))))
    for _ in range(args.warmup):
        run(k)
    t0 = time.perf_counter()
    total = 0.0
    for _ in range(args.steps):
        total += run(k)
    wall = time.perf_counter() - t0
    free()
    scale = (g ** 3) / float(args.grid ** 3)
    marginal = args.steps * (k - 1) / max(total - args.steps * t1, 1e-9) * scale
    inclusive = args.steps * k / total * scale
    sample = (f"CG {g}^3 Poisson on {threads} host threads: every step runs {k} of the workload's {args.iters} iterations; value = marginal rate "
              f"(start-up of {t1 * 1e3:.0f} ms per call subtracted; including it: {inclusive:.3f} it/s)")
    if g != args.grid:
        sample += f"; host memory too small for {args.grid}^3: rate scaled by rows ratio {scale:.4f}"
    line = {
        "impl": "reference", "metric": METRIC, "value": marginal, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.grid, args.iters),
        "cpu_baseline": {"value": marginal, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                         "sample_iterations_per_step": k, "cpu_grid": g,
                         "flags": "-O3 -fopenmp -DSMM_MULTITHREADING, no -march (the prebuilt .so must run on any host; FMA contraction would change the reference's bits)"},
        "e2e": {"value": marginal, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
def golden_fullsize():
    try:
        return json.load(open(os.path.join(ROOT, "tests", "golden", "fullsize_reference.json")))
    except Exception:
        return {}


def x_checksum(x):
    """Order-independent checksum of a float vector's bit patterns (what tests/golden/fullsize_reference.json records)."""
    bits = x.view(np.uint32).astype(np.uint64)
    return int(bits.sum() & np.uint64(0xFFFFFFFFFFFFFFFF)), int(np.bitwise_xor.reduce(bits))


def parity_record(ref, status, iterations, residual, sum_bits, xor_bits, max_abs_error):
    """Compare a REFERENCE_TREE-mode solve with the golden record of the reference's multithreaded arithmetic."""
    rb = int(np.float32(residual).view(np.uint32))
    rec = {"mode": "reference tree (H:305-328 summation order)", "status": int(status), "iterations": int(iterations), "residual_bits": rb,
           "x_checksum": [int(sum_bits), int(xor_bits)], "max_abs_error": float(max_abs_error)}
    if ref:
        rec["golden_iterations"] = ref["iterations"]
        rec["matches_golden"] = bool(int(status) == ref["status"] and int(iterations) == ref["iterations"] and rb == ref["residual_bits"]
                                     and int(sum_bits) == ref["x"]["sum_bits"] and int(xor_bits) == ref["x"]["xor_bits"]
                                     and float(max_abs_error) == ref["max_abs_error"])
    else:
        rec["matches_golden"] = None
    return rec


# fused-minimum algorithmic bytes per iteration
def bytes_per_iteration(solver, rows, nnz, sgs=False):
    base = {"cg": 8 * nnz + 48 * rows, "bicgsym": 8 * nnz + 48 * rows, "cgs": 16 * nnz + 80 * rows, "bicgstab": 16 * nnz + 84 * rows}[solver]
    return base + (2 * (8 * nnz + 32 * rows) if sgs else 0)          # + two applies of B_sgs = 8 nnz + 32 n


def bytes_spmv(rows, cols, nnz):
    return 8 * nnz + 4 * (rows + 1) + 4 * cols + 4 * rows


def run_config(smm, B, L, name, key, solver, A, M, rhs_kind, eps, maxit, modes, peak, golden, setup_s=None):
    """One BASELINE configuration at full size: iterations, rate, the section 8(d) roofline fraction per mode, and the
    bit-exact comparison with the golden record in the reference-order mode.  Device-timed (CUDA events inside the call)."""
    n, nnz = A.rows, A.nnz
    xs = smm.DeviceVector(n)
    if rhs_kind == "ones":
        xs.upload(np.ones(n, np.float32))
    else:
        B._check(L.smm_gen_xstar_dev(n, 0, 0xB200, xs.ptr, None), "xstar")
    b = smm.DeviceVector(n)
    A.spmv_dev(B.OP_ASSIGN, None, xs.ptr, b.ptr)
    xs_h = xs.download()
    bpi = bytes_per_iteration(solver, n, nnz, sgs=M is not None)
    out = {"config": name, "rows": n, "nnz": nnz, "eps": eps, "algorithmic_bytes_per_iteration": bpi, "runs": []}
    if setup_s is not None:
        out["preconditioner_setup_s"] = round(setup_s, 4)
    ref = golden.get(key) if key else None
    if ref:
        out["reference_mt_iterations"] = ref["iterations"]
    # SpMV alone (rMult): effective GB/s over section 8(d)'s B_spmv
    y = smm.DeviceVector(n)
    reps = 20
    t_spmv = spmv_ms(smm, B, L, A, xs, y, reps)
    out["spmv"] = {"ms": t_spmv, "effective_gbs": bytes_spmv(n, A.cols, nnz) / (t_spmv * 1e-3) / 1e9,
                   "frac": bytes_spmv(n, A.cols, nnz) / (t_spmv * 1e-3) / 1e9 / peak}
    del y
    if solver in ("cgs", "bicgsym") and key is None:
        # Config 4's gathers have no locality (hash-placed columns): every stored entry costs its own 32-byte sector of x, and an SM's
        # L1 looks up one gathered sector per clock.  That, not HBM, is the ceiling of this SpMV: entries / (SMs x clock); a kernel
        # that only streams index + value and gathers (tools/gather_ceiling.cu) measures 0.731 ms on this matrix.
        sms, clk = smm.device_info()["sm_count"], 1.965e9
        bound_ms = nnz / (sms * clk) * 1e3
        out["spmv"]["second_ceiling"] = {"what": "L1 tag stage: one gathered 32-byte sector per clock and SM (x gathers without locality: one sector per stored entry)",
                                         "sectors": int(nnz), "bound_ms": bound_ms, "frac": bound_ms / t_spmv,
                                         "measured_ceiling_ms": 0.731, "measured_ceiling_source": "profiles/r02_gather_ceiling.txt (stream + gather only, same matrix)",
                                         "frac_of_measured_ceiling": 0.731 / t_spmv}
    for mode in modes:
        m = {"fast": B.REDUCE_FAST, "tree": B.REDUCE_REFERENCE_TREE}[mode]
        x = smm.DeviceVector(n)
        best = None
        for rep in range(2):                                  # second run: graphs instantiated, clocks up
            x.zero()
            o, _ = B._options(m, B.DRIVER_AUTO, 0, 0)
            info = B._Info()
            if solver == "cg":
                rc = L.smm_solve_cg_dev(A.handle, b.ptr, x.ptr, x.ptr, maxit, eps, C.byref(o), C.byref(info), None)
            elif solver == "bicgsym":
                rc = L.smm_solve_bicgsym_dev(A.handle, b.ptr, x.ptr, maxit, eps, C.byref(o), C.byref(info), None)
            elif solver == "cgs":
                rc = L.smm_solve_cgs_dev(A.handle, b.ptr, x.ptr, maxit, eps, C.byref(o), C.byref(info), None)
            else:
                rc = L.smm_solve_bicgstab_dev(A.handle, None if M is None else M.handle, b.ptr, x.ptr, maxit, eps, C.byref(o), C.byref(info), None)
            B._check(rc, solver)
            if best is None or info.seconds_solve < best:
                best = info.seconds_solve
            if n > (1 << 26):
                break                                         # the 512^3 solves take seconds: once
        xh = x.download()
        finite = bool(np.all(np.isfinite(xh)))
        rate = info.iterations / best if best > 0 else None
        rec = {"mode": mode, "status": int(info.status), "iterations": int(info.iterations), "solver_residual": float(info.residual),
               "seconds_solve": best, "it_per_s": rate, "max_abs_error": float(np.max(np.abs(xh - xs_h))) if finite else None,
               "driver": int(info.driver_mode), "kernel_launches": int(info.kernel_launches)}
        if rate:
            rec["iteration_gbs"] = bpi * rate / 1e9
            rec["frac_of_peak"] = rec["iteration_gbs"] / peak
        if setup_s is not None:
            rec["solve_plus_setup_s"] = best + setup_s              # getPreconditioner() is host analysis + layout: part of a first solve
        if ref and mode == "fast":
            rec["iterations_vs_reference"] = info.iterations / ref["iterations"]
        if mode == "tree" and ref:
            sb, xb = x_checksum(xh)
            rec["parity"] = parity_record(ref, info.status, info.iterations, info.residual, sb, xb, float(np.max(np.abs(xh - xs_h))) if finite else float("nan"))
        out["runs"].append(rec)
        del x
    return out


def spmv_ms(smm, B, L, A, xs, y, reps):
    """Average device time of rMult (smm_spmv_dev) on the library's own stream, bracketed by host-timed synchronisation
    of `reps` back-to-back launches (launch overhead is hidden behind the queue for anything but the smallest matrix)."""
    for _ in range(3):
        A.spmv_dev(B.OP_ASSIGN, None, xs.ptr, y.ptr)
    B._check(L.smm_sync(), "sync")
    t = time.perf_counter()
    for _ in range(reps):
        A.spmv_dev(B.OP_ASSIGN, None, xs.ptr, y.ptr)
    B._check(L.smm_sync(), "sync")
    return (time.perf_counter() - t) * 1e3 / reps


def configs_report(smm, B, L, peak, which=(1, 2, 3, 4), modes=("fast", "tree")):
    """BASELINE.json configs[0..3] at full size (config 5 is the headline above)."""
    golden = golden_fullsize()
    out = []
    SGS = smm.SolverPreconditioner.SYMMETRIC_GAUS_SEIDEL
    if 1 in which:
        A = smm.CSRMatrix.generate(B.GEN_POISSON2D, 1024, 1024)
        out.append(run_config(smm, B, L, "1: ConjugateGradient, 2D 5-point Poisson 1024^2, eps 1e-6, b=A*1", "1", "cg", A, None, "ones", 1e-6, -1, modes, peak, golden))
        del A
    if 2 in which:
        A = smm.CSRMatrix.generate(B.GEN_CONVDIFF3D, 128, 128, 128, 0.5)
        out.append(run_config(smm, B, L, "2: BiCGStab, 3D convection-diffusion 128^3, no preconditioner, eps 1e-6, b=A*x*", "2", "bicgstab", A, None, "xstar", 1e-6, -1, modes, peak, golden))
        del A
    if 3 in which:
        A = smm.CSRMatrix.generate(B.GEN_CONVDIFF3D, 256, 256, 256, 0.5)
        t = time.perf_counter()
        M = A.getPreconditioner(SGS)
        setup = time.perf_counter() - t
        out.append(run_config(smm, B, L, "3: BiCGStab + getPreconditioner() (SGS), 3D convection-diffusion 256^3, eps 1e-6, b=A*x*", "3", "bicgstab", A, M, "xstar", 1e-6, -1, modes, peak, golden, setup_s=setup))
        del M, A
    if 4 in which:
        A = smm.CSRMatrix.generate(B.GEN_POWERLAW, 8388608)
        for fn in ("cgs", "bicgsym"):
            out.append(run_config(smm, B, L, f"4: {fn} SpMV-bound sweep, power-law rows (8.4 M rows, 12..34755 entries per row), 100 iterations, eps 0, b=A*x*",
                                  None, fn, A, None, "xstar", 0.0, 100, ("fast",), peak, golden))
        del A
    return out


def run_b200(args):
    import torch

    import sparse_matrix_math_b200 as smm
    from sparse_matrix_math_b200 import binding as B

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch N>1 with torch.distributed.run")
    torch.cuda.set_device(local)
    L = smm.lib()
    B._check(L.smm_set_device(local), "smm_set_device")
    if world > 1:
        from sparse_matrix_math_b200 import dist
        return dist.bench(args, METRIC, UNIT)

    grid, iters = args.grid, args.iters
    rows, nnz = grid ** 3, stencil_nnz(grid)
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    t_setup = time.perf_counter()
    A = smm.CSRMatrix.generate(B.GEN_CONVDIFF3D, grid, grid, grid, 0.0)
    assert (A.rows, A.nnz) == (rows, nnz)
    ones = torch.ones(rows, dtype=torch.float32, device="cuda")
    b = torch.empty(rows, dtype=torch.float32, device="cuda")
    x = torch.zeros(rows, dtype=torch.float32, device="cuda")
    A.spmv_dev(B.OP_ASSIGN, None, ones.data_ptr(), b.data_ptr(), stream=sp)     # b = A * 1
    del ones
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup

    opts = B._Options()
    opts.reduction_mode = B.REDUCE_REFERENCE_TREE if args.reduction == "tree" else B.REDUCE_FAST
    opts.driver_mode = {"auto": B.DRIVER_AUTO, "chunked": B.DRIVER_GRAPH_CHUNKED, "while": B.DRIVER_GRAPH_WHILE, "stream": B.DRIVER_STREAM}[args.driver]
    opts.check_every = max(iters, 1)
    info = B._Info()

    def step_dev():
        x.zero_()
        B._check(L.smm_solve_cg_dev(A.handle, b.data_ptr(), x.data_ptr(), x.data_ptr(), iters, 0.0, C.byref(opts), C.byref(info), sp), "smm_solve_cg_dev")
        assert info.iterations == iters and info.status == 2, (info.iterations, info.status)

    for _ in range(args.warmup):
        step_dev()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    launches0 = smm.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tm0 = sampler.mark()
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
    e1.record(stream)
    torch.cuda.synchronize()
    tm1 = sampler.mark()
    dev_ms = e0.elapsed_time(e1)
    launches = smm.kernel_launch_count() - launches0
    solve_ms = dev_ms / args.steps
    value = args.steps * iters / (dev_ms * 1e-3)
    final_rr = float(info.residual)

    # --- per-kernel times of the fused CG iteration (same kernels, same arguments), CUDA events on this stream
    ms = [C.c_float(), C.c_float(), C.c_float()]
    B._check(L.smm_profile_cg_iteration(A.handle, max(10, min(50, iters)), C.byref(ms[0]), C.byref(ms[1]), C.byref(ms[2]), sp), "smm_profile_cg_iteration")
    ms_spmv, ms_xr, ms_p = (m.value for m in ms)
    clocks = sampler.stop(tm0, tm1)

    # --- end to end through the host-pointer C ABI call (what the drop-in header's ConjugateGradient makes):
    # pinned host b, x0, x; H2D of b and x0 and D2H of x inside the timed region, every step
    hb = torch.empty(rows, dtype=torch.float32, pin_memory=True)
    hx = torch.zeros(rows, dtype=torch.float32, pin_memory=True)
    hb.copy_(b)
    torch.cuda.synchronize()
    hb_p, hx_p = C.c_void_p(hb.data_ptr()), C.c_void_p(hx.data_ptr())
    e2e_steps = max(1, min(args.steps, 3))

    def step_host():
        hx.zero_()
        B._check(L.smm_solve_cg(A.handle, hb_p, hx_p, hx_p, iters, 0.0, C.byref(opts), C.byref(info)), "smm_solve_cg")
        assert info.iterations == iters

    step_host()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_value = e2e_steps * iters / e2e_s
    x_host_check = float(hx[rows // 2])

    peak, peak_src = measured_peak_gbs()
    spmv_bytes = bytes_spmv_dot(rows, nnz)
    achieved = spmv_bytes / (ms_spmv * 1e-3) / 1e9
    iter_bytes = bytes_cg_iteration(rows, nnz)
    iter_gbs = iter_bytes * value / 1e9
    kernel_sum = ms_spmv + ms_xr + ms_p

    # --- parity where the driver runs it: the same problem solved to convergence (eps 1e-6) in the reference's summation
    # order, compared bit for bit with the golden record of the reference's multithreaded arithmetic
    parity = None
    if not args.no_parity:
        x.zero_()
        opts_t = B._Options()
        opts_t.reduction_mode = B.REDUCE_REFERENCE_TREE
        info_t = B._Info()
        B._check(L.smm_solve_cg_dev(A.handle, b.data_ptr(), x.data_ptr(), x.data_ptr(), -1, 1e-6, C.byref(opts_t), C.byref(info_t), sp), "smm_solve_cg_dev (parity)")
        torch.cuda.synchronize()
        xh = x.cpu().numpy()
        sb, xb = x_checksum(xh)
        parity = parity_record(golden_fullsize().get("5") if grid == 512 else None, info_t.status, info_t.iterations, info_t.residual, sb, xb,
                               float(np.max(np.abs(xh - 1.0))))
        parity["it_per_s"] = info_t.iterations / info_t.seconds_solve if info_t.seconds_solve > 0 else None
        if parity["it_per_s"]:
            parity["frac_of_peak"] = bytes_cg_iteration(rows, nnz) * parity["it_per_s"] / 1e9 / measured_peak_gbs()[0]
        del xh

    del hb, hx, x, b
    A = None
    torch.cuda.empty_cache()
    configs = None
    if not args.no_configs:
        configs = configs_report(smm, B, L, measured_peak_gbs()[0])

    cpu = None
    if not args.no_cpu:
        cpu = cpu_baseline(grid, args.cpu_budget)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": solve_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {**workload_config(grid, iters),
                   "parallelism": "1 GPU", "driver": args.driver, "reductions": "fast (fused, deterministic two-stage)" if args.reduction == "fast" else "reference tree (bit-identical to the reference's multithreaded build)",
                   "l2": f"working set {(8 * nnz + 24 * rows) / 1e9:.1f} GB >> 126 MB L2 (no flush needed)",
                   "setup_s": round(setup_s, 3)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * rows, "d2h_bytes_per_step": 4 * rows,
                "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps, "api": "smm_solve_cg (host pointers, pinned)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "spmv_rows_kernel<1> (Ap = A p, p.Ap fused; TMA-staged, 1 lane per row)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic_bytes("spmv_rows", grid),
                     "traffic_source": (TRAFFIC_SOURCE % os.path.basename(traffic_file())) if grid == 512 and traffic_file() else None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": spmv_bytes, "ms_per_launch": ms_spmv,
                     "bytes_moved_per_launch": bytes_spmv_dot_moved(rows, nnz),
                     "frac_bytes_moved": bytes_spmv_dot_moved(rows, nnz) / (ms_spmv * 1e-3) / 1e9 / peak,
                     "note": "achieved / frac use SURVEY 8(d)'s 8 nnz + 4 (rows + 1) + 8 rows; the kernel reads its row starts as 16-bit offsets "
                             "(2 bytes per row), so it moves 2 bytes per row less than that; peak is the measured COPY bandwidth (reads + writes), "
                             "which a read-dominated stream can exceed",
                     "share_of_iteration": ms_spmv / kernel_sum if kernel_sum > 0 else None},
        "iteration": {"algorithmic_bytes": iter_bytes, "achieved_gbs": iter_gbs, "frac_of_peak": iter_gbs / peak,
                      "bytes_moved_two_pass": bytes_cg_iteration_two_pass(rows, nnz),
                      "frac_of_peak_bytes_moved": bytes_cg_iteration_two_pass(rows, nnz) * value / 1e9 / peak,
                      "note": "algorithmic_bytes is SURVEY 8(d)'s 8 nnz + 48 n; the vector passes read p once (r update | x and p update) and the SpMV reads 16-bit row starts: 8 nnz + 42 n are moved",
                      "ms_spmv_dot": ms_spmv, "ms_r_update": ms_xr, "ms_px_update": ms_p,
                      "ms_per_iteration": solve_ms / iters, "final_rr": final_rr, "x_mid": x_host_check},
        "spmv_effective_gbs": achieved,
        "parity": parity,
        "configs": configs,
        "fast_mode_note": "fast-mode iteration counts are at or below the reference's (fewer is allowed: SURVEY 7 hard part 1; the reference's own two "
                          "builds differ by 2x); the reference-order mode reproduces its counts and bits exactly (parity / configs[].runs[].parity)",
        "cpu_baseline": cpu,
    }
    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--iters", type=int, default=200, help="CG iterations per step")
    ap.add_argument("--driver", default="auto", choices=["auto", "chunked", "while", "stream"])
    ap.add_argument("--reduction", default="fast", choices=["fast", "tree"],
                    help="tree: the reference's summation order (bit-identical to its multithreaded build), also across GPUs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the reference-order solve to convergence")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs 1-4 (N = 1 only)")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
