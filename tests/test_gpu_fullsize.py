"""GPU parity at the FULL sizes of BASELINE.json's configurations (pytest -m gpu).

* REFERENCE_TREE mode against tests/golden/fullsize_reference.json -- the CPU oracle's results for the same inputs in
  the reference's multithreaded arithmetic (recorded by tests/golden/make_fullsize_golden.py): same status, same
  iteration count, same residual bits, same checksum of x.  Config 1's and config 5's counts (2265, 1570) are also the
  numbers measured with the real reference header during the survey.
* FAST mode: status, the solver's own stopping quantity, accuracy of x, iteration count against the documented bar.
* config 4 (power-law, 8.4 M rows, ~2e8 entries): (a) against the CPU oracle at FULL size -- the device-generated CSR is
  downloaded and handed to oracle/smm_oracle.c: rMult / rMultSub bit-identical in exact mode and within 1e-5 in both
  SURVEY 8(d) norms in the fast mode (the long rows are tree-summed there); x, residual and iteration count after 8
  iterations of ConjugateGradientSquared and BiCGSymmetric bit-identical to the oracle's multithreaded arithmetic in
  REFERENCE_TREE mode; (b) size-independent properties (generator statistics, linearity, exact vs fast row
  accumulation, fast vs reference-order solver steps).
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
REF = json.load(open(os.path.join(HERE, "golden", "fullsize_reference.json")))


@pytest.fixture(scope="module")
def smm():
    import sparse_matrix_math_b200 as s
    s.device_info()
    return s


def checksum(x):
    bits = x.view(np.uint32).astype(np.uint64)
    return int(bits.sum() & np.uint64(0xFFFFFFFFFFFFFFFF)), int(np.bitwise_xor.reduce(bits))


def make(smm, key):
    from sparse_matrix_math_b200 import binding as B
    if key == "1":
        return smm.CSRMatrix.generate(B.GEN_POISSON2D, 1024, 1024)
    if key in ("2", "2s"):
        return smm.CSRMatrix.generate(B.GEN_CONVDIFF3D, 128, 128, 128, 0.5)
    if key == "3":
        return smm.CSRMatrix.generate(B.GEN_CONVDIFF3D, 256, 256, 256, 0.5)
    if key == "5":
        return smm.CSRMatrix.generate(B.GEN_CONVDIFF3D, 512, 512, 512, 0.0)
    raise KeyError(key)


def solve(smm, A, key, mode, M=None):
    from sparse_matrix_math_b200 import binding as B
    n = A.rows
    xs = smm.DeviceVector(n)
    if key in ("1", "5"):
        xs.upload(np.ones(n, np.float32))
    else:
        B._check(smm.lib().smm_gen_xstar_dev(n, 0, 0xB200, xs.ptr, None), "xstar")
    b = smm.DeviceVector(n)
    A.spmv_dev(B.OP_ASSIGN, None, xs.ptr, b.ptr)
    x = smm.DeviceVector(n)
    x.zero()
    o, _ = B._options(mode, B.DRIVER_AUTO, 0, 0)
    info = B._Info()
    L = smm.lib()
    if key in ("1", "5"):
        rc = L.smm_solve_cg_dev(A.handle, b.ptr, x.ptr, x.ptr, -1, 1e-6, C.byref(o), C.byref(info), None)
    else:
        rc = L.smm_solve_bicgstab_dev(A.handle, None if M is None else M.handle, b.ptr, x.ptr, -1, 1e-6, C.byref(o), C.byref(info), None)
    B._check(rc, "solve")
    return B.SolveInfo(info), x.download(), xs.download()


# FAST-mode iteration bars, each justified in DESIGN.md section 2: (lower, upper) as fractions of the reference count
FAST_BARS = {"1": (0.70, 1.05), "2": (0.90, 1.10), "2s": (0.90, 1.10), "3": (0.85, 1.10), "5": (0.90, 1.05)}


@pytest.mark.parametrize("key", [k for k in ["1", "2", "2s", "3", "5"] if k in REF])
def test_fullsize_reference_order_is_bit_exact_and_fast_mode_converges(smm, key):
    ref = REF[key]
    A = make(smm, key)
    assert (A.rows, A.nnz) == (ref["rows"], ref["nnz"])
    M = A.getPreconditioner(smm.SolverPreconditioner.SYMMETRIC_GAUS_SEIDEL) if key in ("2s", "3") else None
    info, x, xs = solve(smm, A, key, smm.REDUCE_REFERENCE_TREE, M)
    assert int(info.status) == ref["status"]
    assert info.iterations == ref["iterations"]
    assert int(np.float32(info.residual).view(np.uint32)) == ref["residual_bits"]
    assert checksum(x) == (ref["x"]["sum_bits"], ref["x"]["xor_bits"])
    assert float(np.max(np.abs(x - xs))) == ref["max_abs_error"]
    # throughput mode
    info, x, xs = solve(smm, A, key, smm.REDUCE_FAST, M)
    assert int(info.status) == 0
    assert info.residual <= (1e-6 if key in ("2", "2s", "3") else np.float32(1e-6) * np.float32(1e-6))
    lo, hi = FAST_BARS[key]
    assert lo * ref["iterations"] <= info.iterations <= hi * ref["iterations"], (info.iterations, ref["iterations"])
    assert float(np.max(np.abs(x - xs))) <= max(2.0 * ref["max_abs_error"], 5e-5)


@pytest.mark.parametrize("grid", [(219, 219, 219), (343, 343, 343), (401, 397, 263)])
def test_tree_mode_cg_r_update_fused_into_the_dot(smm, grid):
    """ConjugateGradient in the reference-tree mode on long vectors: the r update (H:2366-2368) rides on the tree dot that follows
    it (dot_tree_rows_kernel<R, true>).  Same bits as the separate update kernel + dot (SMM_B200_DOT_UPDATE=0), for 2 / 8 / 16
    nodes per warp and node boundaries that are not multiples of four elements; the 512^3 golden run covers aligned nodes."""
    from sparse_matrix_math_b200 import binding as B
    A = smm.CSRMatrix.generate(B.GEN_CONVDIFF3D, *grid, 0.0)
    n = A.rows
    xs = smm.DeviceVector(n)
    B._check(smm.lib().smm_gen_xstar_dev(n, 0, 0xB200, xs.ptr, None), "xstar")
    b = smm.DeviceVector(n)
    A.spmv_dev(B.OP_ASSIGN, None, xs.ptr, b.ptr)
    got = []
    for fused in (True, False):
        if not fused:
            os.environ["SMM_B200_DOT_UPDATE"] = "0"
        try:
            x = smm.DeviceVector(n)
            x.zero()
            o, _ = B._options(smm.REDUCE_REFERENCE_TREE, B.DRIVER_AUTO, 0, 0)
            info = B._Info()
            B._check(smm.lib().smm_solve_cg_dev(A.handle, b.ptr, x.ptr, x.ptr, 12, 0.0, C.byref(o), C.byref(info), None), "solve")
            got.append((info.iterations, int(np.float32(info.residual).view(np.uint32)), info.kernel_launches, checksum(x.download())))
        finally:
            os.environ.pop("SMM_B200_DOT_UPDATE", None)
    assert got[0][0] == got[1][0] == 12
    assert got[0][1] == got[1][1] and got[0][3] == got[1][3]
    assert got[0][2] < got[1][2]                                      # one kernel fewer per iteration


def test_config4_powerlaw_properties(smm):
    from sparse_matrix_math_b200 import binding as B
    n = 8388608
    A = smm.CSRMatrix.generate(B.GEN_POWERLAW, n)
    assert A.rows == n and 1.9e8 < A.nnz < 2.1e8                       # ~2e8 entries, mean row length ~24
    start = np.empty(n + 1, np.int32)
    B._check(smm.lib().smm_csr_download(A.handle, start.ctypes.data_as(C.c_void_p), None, None), "download")
    lens = np.diff(start)
    assert lens.min() == 12 and lens.max() > 4096 and 23.0 < lens.mean() < 24.5
    # the generator is the same function as tests/matgen.py (bit-identical arrays are checked at 20k rows elsewhere);
    # here: row lengths agree with matgen for every row of the full-size matrix
    import matgen
    assert np.array_equal(lens, matgen.powerlaw_row_lengths(n))
    rng = np.random.default_rng(4)
    xa = rng.uniform(-1, 1, n).astype(np.float32); xb = rng.uniform(-1, 1, n).astype(np.float32)
    da, db_, dz = smm.DeviceVector(n, xa), smm.DeviceVector(n, xb), smm.DeviceVector(n, (np.float32(0.5) * xa + np.float32(2) * xb).astype(np.float32))
    ya, yb, yz, ye = smm.DeviceVector(n), smm.DeviceVector(n), smm.DeviceVector(n), smm.DeviceVector(n)
    A.spmv_dev(B.OP_ASSIGN, None, da.ptr, ya.ptr)
    A.spmv_dev(B.OP_ASSIGN, None, db_.ptr, yb.ptr)
    A.spmv_dev(B.OP_ASSIGN, None, dz.ptr, yz.ptr)
    A.spmv_dev(B.OP_ASSIGN, None, da.ptr, ye.ptr, exact=True)          # left-to-right accumulation in every row
    ya_h, yb_h, yz_h, ye_h = ya.download(), yb.download(), yz.download(), ye.download()
    # |A| row sums are < 3 (diag 2 + sum |off| < 1), |x| <= 2.5
    assert np.max(np.abs(yz_h - (0.5 * ya_h.astype(np.float64) + 2.0 * yb_h.astype(np.float64)))) <= 1e-5 * 3 * 2.5
    assert np.max(np.abs(ya_h - ye_h)) <= 1e-5 * 3                       # tree-summed long rows vs the reference order
    # rows streamed through shared memory are accumulated left to right in both modes (bit-identical); only the rows of
    # the few chunks that hold a long row take the warp-per-row tree
    assert np.mean(ya_h == ye_h) > 0.9
    # CGS and BiCGSymmetric sweeps: fast vs reference-order reductions agree after a few iterations
    xs = smm.DeviceVector(n)
    B._check(smm.lib().smm_gen_xstar_dev(n, 0, 0xB200, xs.ptr, None), "xstar")
    b = smm.DeviceVector(n)
    A.spmv_dev(B.OP_ASSIGN, None, xs.ptr, b.ptr)
    L = smm.lib()
    for fn in (L.smm_solve_cgs_dev, L.smm_solve_bicgsym_dev):
        res = []
        for mode in (B.REDUCE_FAST, B.REDUCE_REFERENCE_TREE):
            x = smm.DeviceVector(n); x.zero()
            o, _ = B._options(mode, B.DRIVER_AUTO, 0, 0)
            info = B._Info()
            B._check(fn(A.handle, b.ptr, x.ptr, 3, 0.0, C.byref(o), C.byref(info), None), "solve")
            assert info.iterations == 3
            res.append(x.download())
        assert np.max(np.abs(res[0] - res[1])) < 1e-5
        assert np.max(np.abs(res[0] - xs.download())) < 1e-2


def test_config4_powerlaw_fullsize_against_the_oracle(smm):
    """SpMV parity and x after k <= 10 iterations against the oracle at 8.4 M rows / ~197 M entries.
    Row sums of the reference are strictly sequential (H:1484-1489); the longest row here has 34755 entries."""
    import matgen
    import oracle_lib as ol
    from sparse_matrix_math_b200 import binding as B
    n = 8388608
    A = smm.CSRMatrix.generate(B.GEN_POWERLAW, n)
    start, positions, values = A.download()
    g = ol.CSR(n, n, start, positions, values, 0)
    assert g.nnz == A.nnz and int(np.diff(start).max()) > 30000
    xs_h = matgen.xstar(n)
    xs = smm.DeviceVector(n)
    B._check(smm.lib().smm_gen_xstar_dev(n, 0, 0xB200, xs.ptr, None), "xstar")
    assert xs.download().tobytes() == xs_h.tobytes()                     # device generator == tests/matgen.py

    # ---- rMult (H:1501-1505) and rMultSub (H:1512-1515) ----
    y_ref = ol.spmv(g, 0, None, xs_h)
    gabs = ol.CSR(n, n, start, positions, np.abs(values), 0)
    scale = ol.spmv(gabs, 0, None, np.abs(xs_h))                         # sum_k |a_ik| |x_k|, the denominator of 8(d)'s row norm
    del gabs
    assert scale.min() > 0
    y, ye = smm.DeviceVector(n), smm.DeviceVector(n)
    A.spmv_dev(B.OP_ASSIGN, None, xs.ptr, ye.ptr, exact=True)
    assert ye.download().tobytes() == y_ref.tobytes()                    # exact mode: the reference's bits in every row
    A.spmv_dev(B.OP_ASSIGN, None, xs.ptr, y.ptr)
    yf = y.download()
    assert float(np.max(np.abs(yf - y_ref) / scale)) <= 1e-5
    assert float(np.linalg.norm((yf - y_ref).astype(np.float64)) / np.linalg.norm(y_ref.astype(np.float64))) <= 1e-5
    lhs_h = np.random.default_rng(44).uniform(-1, 1, n).astype(np.float32)
    lhs = smm.DeviceVector(n, lhs_h)
    s_ref = ol.spmv(g, 2, lhs_h, xs_h)
    A.spmv_dev(B.OP_SUB, lhs.ptr, xs.ptr, ye.ptr, exact=True)
    assert ye.download().tobytes() == s_ref.tobytes()
    A.spmv_dev(B.OP_SUB, lhs.ptr, xs.ptr, y.ptr)
    sf = y.download()
    assert float(np.max(np.abs(sf - s_ref) / (scale + np.abs(lhs_h)))) <= 1e-5
    assert float(np.linalg.norm((sf - s_ref).astype(np.float64)) / np.linalg.norm(s_ref.astype(np.float64))) <= 1e-5
    del lhs, ye, y

    # ---- 8 iterations of CGS (H:2109-2178) and BiCGSymmetric (H:2021-2102): b = A x*, x0 = 0, eps = 0 ----
    b = smm.DeviceVector(n, y_ref)
    L = smm.lib()
    for name, fn in (("cgs", L.smm_solve_cgs_dev), ("bicgsym", L.smm_solve_bicgsym_dev)):
        o = ol.solve(name, g, y_ref, np.zeros(n, np.float32), 8, 0.0, 1)
        assert o["iterations"] == 8
        for mode in (B.REDUCE_REFERENCE_TREE, B.REDUCE_FAST):
            x = smm.DeviceVector(n); x.zero()
            opts, _ = B._options(mode, B.DRIVER_AUTO, 0, 0)
            info = B._Info()
            B._check(fn(A.handle, b.ptr, x.ptr, 8, 0.0, C.byref(opts), C.byref(info), None), name)
            xh = x.download()
            assert info.iterations == 8 and info.status == o["status"]
            if mode == B.REDUCE_REFERENCE_TREE:                          # the multithreaded reference's bits
                assert xh.tobytes() == o["x"].tobytes(), name
                assert np.float32(info.residual).tobytes() == np.float32(o["residual"]).tobytes(), name
            else:                                                        # throughput mode: same iterate up to rounding
                assert float(np.max(np.abs(xh - o["x"]))) <= 1e-5, name
            assert float(np.max(np.abs(xh - xs_h))) < 1e-2
