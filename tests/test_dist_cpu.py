"""Host-side logic of the multi-GPU path on CPU: partitioning, halo windows, send plans, and a 2-process gloo run that
moves real halo data with the plan and reproduces the global SpMV / CG results of the oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import matgen
import oracle_lib as ol
from sparse_matrix_math_b200 import dist as smd


def test_row_partition_covers_and_aligns():
    for rows, p, align in [(100, 3, 1), (512 ** 3, 8, 512 * 512), (7, 8, 1), (64, 2, 16), (10, 1, 4)]:
        parts = smd.row_partition(rows, p, align)
        assert parts[0][0] == 0 and parts[-1][1] == rows and len(parts) == p
        for (a, b), (c, d) in zip(parts, parts[1:]):
            assert b == c and a <= b
        for a, b in parts[:-1]:
            assert b % align == 0
    parts = smd.row_partition(512 ** 3, 8, 512 * 512)
    assert all(b - a == 512 ** 3 // 8 for a, b in parts)


def test_nnz_partition_balances_entries():
    m = matgen.powerlaw(5000)
    parts = smd.nnz_partition(m.start, 4)
    assert parts[0][0] == 0 and parts[-1][1] == m.rows
    sizes = [int(m.start[b] - m.start[a]) for a, b in parts]
    assert max(sizes) - min(sizes) <= 2 * int(np.diff(m.start).max())


def test_window_and_halo_plan_for_a_stencil():
    g = matgen.poisson3d(6, 6, 12)                      # 432 rows, plane = 36
    parts = smd.row_partition(g.rows, 3, 36)
    ranges = []
    for rb, re in parts:
        s, p, v = smd.slice_rows(g.start, g.positions, g.values, rb, re)
        lo, hi = smd.window_of(rb, re, p)
        assert (rb - lo) % 4 == 0 and lo <= rb and hi >= re
        ranges.append((rb, re, lo, hi))
    # middle rank: one plane from each neighbour
    sends, sources = smd.halo_plan(1, ranges)
    assert sorted(sources) == [0, 2]
    assert sorted((peer, ln) for peer, _, _, ln in sends) == [(0, 36), (2, 36)]
    sends0, sources0 = smd.halo_plan(0, ranges)
    assert sources0 == [1] and [(p, ln) for p, _, _, ln in sends0] == [(1, 36)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = matgen.convdiff3d(7, 0.5, 5, 8)            # non-symmetric stencil, 280 rows
        parts = smd.nnz_partition(g.start, world)
        rb, re = parts[rank]
        start, pos, val = smd.slice_rows(g.start, g.positions, g.values, rb, re)
        lo, hi = smd.window_of(rb, re, pos)
        gathered = [None] * world
        dist.all_gather_object(gathered, (rb, re, lo, hi))
        sends, sources = smd.halo_plan(rank, gathered)
        local = ol.CSR(re - rb, hi - lo, start, pos - lo, val)
        x = matgen.xstar(g.rows)

        def exchange(ext):
            reqs = []
            bufs = {}
            for s in sources:
                prb, pre, _, _ = gathered[s]
                a, b = max(prb, lo), min(pre, hi)
                bufs[s] = (torch.empty(b - a), a - lo)
                reqs.append(dist.irecv(bufs[s][0], src=s))
            for peer, src_off, _, ln in sends:
                reqs.append(dist.isend(torch.from_numpy(ext[src_off:src_off + ln].copy()), dst=peer))
            for r in reqs:
                r.wait()
            for s, (t, off) in bufs.items():
                ext[off:off + len(t)] = t.numpy()

        ext = np.full(hi - lo, np.nan, np.float32)
        ext[rb - lo:re - lo] = x[rb:re]
        exchange(ext)
        y_local = ol.spmv(local, 0, None, np.nan_to_num(ext, nan=1e30))   # any halo entry the plan missed would poison y
        y_ref = ol.spmv(g, 0, None, x)[rb:re]
        ok = y_local.tobytes() == y_ref.tobytes()
        out[rank] = ok
    finally:
        dist.destroy_process_group()


def test_two_rank_halo_exchange_reproduces_global_spmv():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_tbb_partition_reproduces_the_reference_tree():
    """dist.tbb_partition: the blocks are nodes of tbb::parallel_deterministic_reduce's split tree (H:308-320), so the
    per-block reference dots joined pairwise are the reference dot of the whole vector, bit for bit."""
    import numpy as np

    import oracle_lib as ol
    from sparse_matrix_math_b200 import dist as smd
    rng = np.random.default_rng(5)
    for n in (65536, 89999, 100003, 1 << 18):
        a = rng.standard_normal(n).astype(np.float32)
        b = rng.standard_normal(n).astype(np.float32)
        want = np.float32(ol.dot(a, b, True))
        for P in (1, 2, 4, 8):
            parts = smd.tbb_partition(n, P)
            assert parts[0][0] == 0 and parts[-1][1] == n and all(parts[i][1] == parts[i + 1][0] for i in range(P - 1))
            if min(hi - lo for lo, hi in parts) <= 8192:
                continue                                     # the blocks would be smaller than a leaf of the tree
            v = [np.float32(ol.dot(a[lo:hi], b[lo:hi], True)) for lo, hi in parts]
            while len(v) > 1:
                v = [np.float32(v[2 * k] + v[2 * k + 1]) for k in range(len(v) // 2)]
            assert v[0].tobytes() == want.tobytes(), (n, P)
