"""Multi-GPU layer: one process per GPU (torchrun), contiguous row blocks, P2P halo exchange and fused P2P scalar
all-reduce inside the CUDA kernels (csrc/dist.cu, csrc/dist_device.cuh).  torch.distributed is only the plumbing that
carries the CUDA IPC handles and row ranges between the processes at set-up, and the barrier / max-over-ranks of the
benchmark timing; nothing on the iteration path goes through it.
"""
import ctypes as C
import json
import os
import time

import numpy as np

from . import binding as B

MAX_RANKS = 8


# ---- host-side partition logic (pure Python: exercised by the gloo tests on CPU) -----------------------------------
def row_partition(global_rows, nranks, align=1):
    """Contiguous blocks of rows, as equal as possible, boundaries rounded to a multiple of `align`."""
    cuts = [0]
    for r in range(1, nranks):
        c = (global_rows * r) // nranks
        c = (c // align) * align
        cuts.append(max(c, cuts[-1]))
    cuts.append(global_rows)
    return [(cuts[r], cuts[r + 1]) for r in range(nranks)]


def tbb_partition(global_rows, nranks):
    """Row blocks that are the depth-log2(nranks) nodes of the reference's reduction tree over [0, global_rows)
    (tbb::parallel_deterministic_reduce halves a range at lo + (hi - lo) / 2, H:308-320).  With this partition the
    REFERENCE_TREE mode distributes bit-exactly: every rank sums its own subtree, the ranks are joined pairwise."""
    assert nranks & (nranks - 1) == 0, "power-of-two number of ranks"
    parts = [(0, global_rows)]
    while len(parts) < nranks:
        parts = [half for lo, hi in parts for half in ((lo, lo + (hi - lo) // 2), (lo + (hi - lo) // 2, hi))]
    return parts


def nnz_partition(start, nranks):
    """Contiguous blocks of rows with (nearly) equal numbers of stored entries; start = global CSR row pointer."""
    start = np.asarray(start, np.int64)
    rows, nnz = len(start) - 1, int(start[-1])
    cuts = [0]
    for r in range(1, nranks):
        c = int(np.searchsorted(start, (nnz * r) // nranks, side="left"))
        cuts.append(min(max(c, cuts[-1]), rows))
    cuts.append(rows)
    return [(cuts[r], cuts[r + 1]) for r in range(nranks)]


def window_of(row_begin, row_end, positions):
    """[lo, hi) of the global vector the rows read; lo rounded so that row_begin - lo is a multiple of 4."""
    lo, hi = row_begin, row_end
    if len(positions):
        lo = min(lo, int(np.min(positions)))
        hi = max(hi, int(np.max(positions)) + 1)
    lo = row_begin - ((row_begin - lo + 3) // 4) * 4
    return lo, hi


def halo_plan(rank, ranges):
    """ranges[r] = (row_begin, row_end, lo, hi).  Returns (sends, sources): sends = [(peer, src_off, dst_off, len)] of this
    rank's owned entries that fall into a peer's window; sources = ranks whose owned rows fall into this rank's window."""
    rb, re, lo, hi = ranges[rank]
    sends, sources = [], []
    for r, (prb, pre, plo, phi) in enumerate(ranges):
        if r == rank:
            continue
        a, b = max(rb, plo), min(re, phi)
        if b > a:
            sends.append((r, a - lo, a - plo, b - a))
        a2, b2 = max(prb, lo), min(pre, hi)
        if b2 > a2:
            sources.append(r)
    return sends, sources


def slice_rows(start, positions, values, row_begin, row_end):
    """Rows [row_begin,row_end) of a host CSR, global column indices kept."""
    start = np.asarray(start, np.int64)
    k0, k1 = int(start[row_begin]), int(start[row_end])
    return (start[row_begin:row_end + 1] - k0).astype(np.int32), np.ascontiguousarray(positions[k0:k1], np.int32), \
        np.ascontiguousarray(values[k0:k1], np.float32)


# ---- device side --------------------------------------------------------------------------------------------------
def _bind(L):
    if getattr(L, "_dist_bound", False):
        return
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    L.smm_dist_create.argtypes = [i32, i32, i64, i64, i64, vp, C.POINTER(vp)]
    L.smm_dist_info.argtypes = [vp, C.POINTER(i64), vp]
    L.smm_dist_connect.argtypes = [vp, C.POINTER(i64), vp]
    L.smm_dist_spmv_dev.argtypes = [vp, vp, vp, vp]
    L.smm_dist_solve_cg.argtypes = [vp, vp, vp, vp, i32, C.c_float, C.POINTER(B._Options), C.POINTER(B._Info), vp]
    for name in ("smm_dist_solve_bicgsym", "smm_dist_solve_cgs", "smm_dist_solve_bicgstab"):
        getattr(L, name).argtypes = [vp, vp, vp, i32, C.c_float, C.POINTER(B._Options), C.POINTER(B._Info), vp]
    L.smm_dist_error.argtypes = [vp, C.POINTER(i32)]
    L.smm_dist_destroy.argtypes = [vp]
    L.smm_gen_csr_rows.argtypes = [i32, i32, i32, i32, C.c_float, i64, i64, C.POINTER(vp)]
    L._dist_bound = True


class DistMatrix:
    """This rank's row block of a global matrix, connected to its peers."""

    def __init__(self, local, global_rows, row_begin, row_end, rank, nranks, gather):
        """local: CSRMatrix holding rows [row_begin,row_end) with GLOBAL column indices (it is re-indexed in place).
        gather(obj) -> list of every rank's obj (e.g. torch.distributed.all_gather_object)."""
        L = B.lib()
        _bind(L)
        self.L, self.local, self.rank, self.nranks = L, local, rank, nranks
        self.global_rows, self.row_begin, self.row_end = global_rows, row_begin, row_end
        h = C.c_void_p()
        B._check(L.smm_dist_create(rank, nranks, global_rows, row_begin, row_end, local.handle, C.byref(h)), "smm_dist_create")
        self.handle = h.value
        ranges = (C.c_int64 * 4)()
        ipc = (C.c_char * 64)()
        B._check(L.smm_dist_info(self.handle, ranges, ipc), "smm_dist_info")
        self.ranges = tuple(int(v) for v in ranges)
        if nranks > 1:
            everyone = gather((self.ranges, bytes(ipc.raw)))
            all_ranges = (C.c_int64 * (4 * nranks))(*[v for r, _ in everyone for v in r])
            all_handles = (C.c_char * (64 * nranks)).from_buffer_copy(b"".join(hd for _, hd in everyone))
            B._check(L.smm_dist_connect(self.handle, all_ranges, all_handles), "smm_dist_connect")
            self.all_ranges = [tuple(r) for r, _ in everyone]
        else:
            self.all_ranges = [self.ranges]

    @property
    def n_local(self):
        return self.row_end - self.row_begin

    def spmv_dev(self, x_ptr, y_ptr, stream=None):
        B._check(self.L.smm_dist_spmv_dev(self.handle, x_ptr, y_ptr, stream), "smm_dist_spmv_dev")

    def solve_cg_dev(self, b_ptr, x0_ptr, x_ptr, max_iterations, eps, stream=None, driver_mode=B.DRIVER_AUTO, check_every=0,
                     reduction_mode=B.REDUCE_FAST):
        """reduction_mode REDUCE_REFERENCE_TREE needs the row blocks of tbb_partition (bit-identical to the reference's
        multithreaded build on any power-of-two number of GPUs)."""
        o = B._Options()
        o.reduction_mode, o.driver_mode, o.check_every = reduction_mode, driver_mode, check_every
        info = B._Info()
        B._check(self.L.smm_dist_solve_cg(self.handle, b_ptr, x0_ptr, x_ptr, int(max_iterations), float(eps), C.byref(o), C.byref(info), stream),
                 "smm_dist_solve_cg")
        return B.SolveInfo(info)

    def solve_dev(self, solver, b_ptr, x_ptr, max_iterations, eps, stream=None, driver_mode=B.DRIVER_AUTO, check_every=0,
                  reduction_mode=B.REDUCE_FAST):
        """solver in {"bicgsym", "cgs", "bicgstab"}; x is initial guess and result (device slices of this rank).
        REDUCE_REFERENCE_TREE: bicgsym and cgs, with the row blocks of tbb_partition."""
        o = B._Options()
        o.reduction_mode, o.driver_mode, o.check_every = reduction_mode, driver_mode, check_every
        info = B._Info()
        fn = getattr(self.L, "smm_dist_solve_" + solver)
        B._check(fn(self.handle, b_ptr, x_ptr, int(max_iterations), float(eps), C.byref(o), C.byref(info), stream), "smm_dist_solve_" + solver)
        return B.SolveInfo(info)

    def error(self):
        e = C.c_int()
        B._check(self.L.smm_dist_error(self.handle, C.byref(e)), "smm_dist_error")
        return e.value

    def close(self):
        if getattr(self, "handle", None):
            self.L.smm_dist_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def generate_rows(kind, nx, ny, nz, c, row_begin, row_end):
    L = B.lib()
    _bind(L)
    m = B.CSRMatrix()
    h = C.c_void_p()
    B._check(L.smm_gen_csr_rows(kind, nx, max(ny, 1), max(nz, 1), float(c), row_begin, row_end, C.byref(h)), "smm_gen_csr_rows")
    m.handle = h.value
    m._read_shape()
    return m


def init_process_group():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    B._check(B.lib().smm_set_device(local), "smm_set_device")
    return rank, world, local


def all_gather_object(obj):
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


# ---- bench.py, N > 1 ------------------------------------------------------------------------------------------------
def bench(args, metric, unit):
    """CG on the 512^3 Poisson problem split into z-slabs over the N GPUs of one box (strong scaling)."""
    import torch
    import torch.distributed as dist

    from bench import ClockSampler, bytes_cg_iteration, golden_fullsize, measured_peak_gbs, parity_record, stencil_nnz, workload_config, x_checksum

    rank, world, local = init_process_group()
    grid, iters = args.grid, args.iters
    rows = grid ** 3
    plane = grid * grid
    rb, re = row_partition(rows, world, align=plane)[rank]
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    t_setup = time.perf_counter()
    A = generate_rows(B.GEN_CONVDIFF3D, grid, grid, grid, 0.0, rb, re)
    D = DistMatrix(A, rows, rb, re, rank, world, all_gather_object)
    n = re - rb
    ones = torch.ones(n, dtype=torch.float32, device="cuda")
    b = torch.empty(n, dtype=torch.float32, device="cuda")
    x = torch.zeros(n, dtype=torch.float32, device="cuda")
    dist.barrier()
    D.spmv_dev(ones.data_ptr(), b.data_ptr(), sp)            # b = A * 1 (halo of the ones vector exchanged)
    del ones
    torch.cuda.synchronize()
    dist.barrier()
    setup_s = time.perf_counter() - t_setup
    drv = {"auto": B.DRIVER_AUTO, "chunked": B.DRIVER_GRAPH_CHUNKED, "while": B.DRIVER_GRAPH_WHILE, "stream": B.DRIVER_STREAM}[args.driver]
    red = B.REDUCE_REFERENCE_TREE if getattr(args, "reduction", "fast") == "tree" else B.REDUCE_FAST

    def step():
        x.zero_()
        info = D.solve_cg_dev(b.data_ptr(), x.data_ptr(), x.data_ptr(), iters, 0.0, sp, driver_mode=drv, check_every=max(iters, 1), reduction_mode=red)
        assert info.iterations == iters and int(info.status) == 2, (info.iterations, info.status)
        return info

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    dist.barrier()
    torch.cuda.synchronize()
    launches0 = B.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tm0 = time.time()
    e0.record(stream)
    info = None
    for _ in range(args.steps):
        info = step()
    e1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    tm1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)                 # max over ranks, device-timed
    dev_ms = float(ms.item())
    launches = B.kernel_launch_count() - launches0
    value = args.steps * iters / (dev_ms * 1e-3)

    # end to end: this rank's slices of b, x0 come from pinned host memory and x goes back, every step
    hb = torch.empty(n, dtype=torch.float32, pin_memory=True)
    hx0 = torch.zeros(n, dtype=torch.float32, pin_memory=True)
    hx = torch.empty(n, dtype=torch.float32, pin_memory=True)
    hb.copy_(b)
    torch.cuda.synchronize()
    e2e_steps = max(1, min(args.steps, 3))

    def step_host():
        b.copy_(hb, non_blocking=True)
        x.copy_(hx0, non_blocking=True)
        D.solve_cg_dev(b.data_ptr(), x.data_ptr(), x.data_ptr(), iters, 0.0, sp, driver_mode=drv, check_every=max(iters, 1), reduction_mode=red)
        hx.copy_(x, non_blocking=True)
        torch.cuda.synchronize()

    step_host()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    dist.barrier()
    e2e = torch.tensor([time.perf_counter() - t0], device="cuda")
    dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e.item())
    # parity where the driver runs it: the same system solved to convergence (eps 1e-6) in the reference's summation order --
    # every rank sums its own node of the reference's reduction tree, the ranks are joined pairwise -- compared bit for bit
    # with the golden record of the reference's multithreaded arithmetic (iterations, residual bits, checksum of x)
    parity = None
    if not getattr(args, "no_parity", False) and world & (world - 1) == 0 and tbb_partition(rows, world)[rank] == (rb, re):
        x.zero_()
        info_t = D.solve_cg_dev(b.data_ptr(), x.data_ptr(), x.data_ptr(), -1, 1e-6, sp, driver_mode=drv, reduction_mode=B.REDUCE_REFERENCE_TREE)
        torch.cuda.synchronize()
        xh = x.cpu().numpy()
        sb, xb = x_checksum(xh)
        parts = all_gather_object((sb, xb, float(np.max(np.abs(xh - 1.0))), int(info_t.status), info_t.iterations, info_t.residual, info_t.seconds_solve))
        if rank == 0:
            sum_bits, xor_bits = 0, 0
            for p_ in parts:
                sum_bits = (sum_bits + p_[0]) & 0xFFFFFFFFFFFFFFFF
                xor_bits ^= p_[1]
            same = all(p_[3:6] == parts[0][3:6] for p_ in parts)           # every rank took the same branches
            parity = parity_record(golden_fullsize().get("5") if grid == 512 else None, parts[0][3], parts[0][4], parts[0][5], sum_bits, xor_bits,
                                   max(p_[2] for p_ in parts))
            parity["ranks_agree"] = same
            secs = max(p_[6] for p_ in parts)
            parity["it_per_s"] = parts[0][4] / secs if secs > 0 else None
        del xh
    err = D.error()
    clocks = sampler.stop(tm0, tm1) if rank == 0 else None

    if rank == 0:
        nnz = stencil_nnz(grid)
        peak, peak_src = measured_peak_gbs()
        iter_bytes = bytes_cg_iteration(rows, nnz)
        gbs = iter_bytes * value / 1e9
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {**workload_config(grid, iters),
                       "parallelism": f"{world} GPUs, z-slab row blocks, P2P halo exchange + fused P2P scalar all-reduce (no NCCL in the loop)",
                       "rows_per_gpu": n, "driver": args.driver, "reductions": getattr(args, "reduction", "fast"),
                       "l2": f"per-GPU working set {(8 * nnz + 24 * rows) / world / 1e9:.2f} GB >> 126 MB L2 (no flush needed)",
                       "setup_s": round(setup_s, 3)},
            "e2e": {"value": e2e_steps * iters / e2e_s, "unit": unit, "h2d_bytes_per_step": 8 * rows, "d2h_bytes_per_step": 4 * rows,
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "api": "DistMatrix.solve_cg_dev (smm_dist_solve_cg) with per-rank pinned host slices copied in and out"},
            "gpu_launches": int(launches) * world,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "whole CG iteration, aggregate over GPUs", "achieved": gbs, "peak": peak * world,
                         "unit": "GB/s", "frac": gbs / (peak * world), "traffic": None, "peak_source": peak_src + f" x {world} GPUs",
                         "algorithmic_bytes_per_launch": iter_bytes},
            "iteration": {"ms_per_iteration": dev_ms / args.steps / iters, "final_rr": float(info.residual), "comm_error": err},
            "parity": parity,
            "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    D.close()
    dist.destroy_process_group()
