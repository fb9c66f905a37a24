"""Pin oracle/smm_oracle.c (the CPU restatement every GPU parity test is judged by).

Three anchors:
  1. the known answers of the reference's own tests (test/cpp/csr.cpp, triplet.cpp, cg.cpp, bicgstab.cpp ...),
  2. tests/golden/golden_v1.npz: outputs of the REAL reference header (serial and SMM_MULTITHREADING builds)
     recorded by tests/golden/make_golden.py -- bit-exact comparison, iteration counts included,
  3. (only where oracle/_ref was built, i.e. in the build container) live bit-exact comparison against the
     reference on fresh random inputs.
"""
import os

import numpy as np
import pytest

import matgen
import oracle_lib as ol

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ASSETS = ["mesh1e1", "mesh1em1", "mesh1em6", "sherman1"]
GENERATED = ["poisson2d_96x100", "convdiff3d_22", "powerlaw_9000"]


def gold_csr(golden, key):
    rows, cols, fas = golden[f"{key}/shape"]
    return ol.CSR(rows, cols, golden[f"{key}/start"], golden[f"{key}/positions"], golden[f"{key}/values"], fas)


# ---------------------------------------------------------------------------------------------
# 1. known answers from the reference's tests
# ---------------------------------------------------------------------------------------------
def csr_5x4():
    # test/cpp/csr.cpp:263-275
    ents = [(0, 0, 4.5), (0, 2, 3.2), (1, 0, 3.1), (1, 1, 2.9), (1, 3, 0.9), (2, 1, 1.7), (2, 2, 3.0), (3, 0, 3.5), (3, 1, 0.4), (3, 3, 1.0)]
    r, c, v = zip(*ents)
    return ol.triplets_to_csr(5, 4, r, c, v)


SPMV_CASES = [
    # (op, mult, lhs, expected)  test/cpp/csr.cpp:278-369 (rMultAdd) and :390-521 (rMultSub)
    (1, [1, 2, 3, 4], [0, 0, 0, 0, 0], [14.1, 12.5, 12.4, 8.3, 0]),
    (1, [0, 0, 0, 0], [5, 6, 7, 8, 9], [5, 6, 7, 8, 9]),
    (1, [1, 0, 3, 4], [5, 6, 7, 8, 10], [19.1, 12.7, 16.0, 15.5, 10]),
    (2, [1, 2, 3, 4], [0, 0, 0, 0, 0], [-14.1, -12.5, -12.4, -8.3, 0]),
    (2, [0, 0, 0, 0], [5, 6, 7, 8, 9], [5, 6, 7, 8, 9]),
    (2, [1, 0, 3, 4], [5, 6, 7, 8, 10], [-9.1, -0.7, -2.0, 0.5, 10]),
    (0, [1, 2, 3, 4], None, [14.1, 12.5, 12.4, 8.3, 0]),
]


@pytest.mark.parametrize("op,mult,lhs,expected", SPMV_CASES)
@pytest.mark.parametrize("inplace", [False, True])
def test_spmv_known_answers(op, mult, lhs, expected, inplace):
    m = csr_5x4()
    assert m.nnz == 10 and list(m.start) == [0, 2, 5, 7, 10, 10]
    mult = np.array(mult, np.float32)
    lhs_a = None if lhs is None else np.array(lhs, np.float32)
    if inplace and lhs_a is not None:
        out = ol.spmv(m, op, lhs_a, mult, out=lhs_a)          # out aliases lhs (csr.cpp:296, :398)
    else:
        keep = None if lhs_a is None else lhs_a.copy()
        out = ol.spmv(m, op, lhs_a, mult)
        if keep is not None:
            assert np.array_equal(keep, lhs_a)                 # "lhs not modified" (csr.cpp:410-414)
    assert np.allclose(out, np.array(expected, np.float32), rtol=1e-6, atol=0)


def test_spmv_empty_matrix():
    # csr.cpp:278-289: A == 0 -> out = lhs
    m = ol.triplets_to_csr(5, 4, [], [], [])
    assert m.nnz == 0 and m.first_active_start == 5
    out = ol.spmv(m, 1, np.array([5, 6, 7, 8, 9], np.float32), np.array([1, 2, 3, 4], np.float32))
    assert list(out) == [5, 6, 7, 8, 9]


def test_triplet_duplicates_sum_in_call_order():
    # test/cpp/triplet.cpp:24-96: repeated addEntry sums and does not grow nnz; explicit zeros are kept
    m = ol.triplets_to_csr(3, 3, [2, 0, 2, 2, 1], [1, 0, 1, 1, 2], [1e8, 1.0, 1.0, -1e8, 0.0])
    assert m.nnz == 3
    assert list(m.start) == [0, 1, 2, 3] and list(m.positions) == [0, 2, 1]
    assert m.values[2] == np.float32(np.float32(np.float32(1e8) + np.float32(1.0)) - np.float32(1e8))   # == 0, not 1
    assert m.values[1] == 0.0


def test_ic0_known_answer(golden):
    # test/cpp/cg.cpp:28-60
    trow = [0, 0, 1, 1, 2, 3, 3, 3, 4, 4, 4]
    tcol = [3, 0, 1, 4, 2, 0, 3, 4, 1, 3, 4]
    tval = [4, 10, 9, 5, 12, 4, 15, 7, 5, 7, 8]
    m = ol.triplets_to_csr(5, 5, trow, tcol, tval)
    rc, ic0 = ol.ic0_factorize(m)
    assert rc == 0
    x = ol.ic0_apply(m, ic0, np.ones(5, np.float32))
    assert np.allclose(x, [0.0995763, 0.0646186, 0.0833333, 0.0010593, 0.0836864], rtol=1e-4)
    assert x.tobytes() == golden["ic0_5x5/apply_ones"].tobytes()
    assert ic0[: m.nnz].tobytes() == golden["ic0_5x5/factor"].tobytes()


def test_load_symmetric_known_answer():
    # test/cpp/csr.cpp:788-826 (explicit zero kept: nnz 8)
    st, m = ol.load_mtx(os.path.join(GOLD, "load_symmetric_test.mtx"))
    assert st == 0 and (m.rows, m.cols, m.nnz) == (5, 5, 8)
    dense = np.zeros((5, 5), np.float32)
    for r in range(5):
        for k in range(m.start[r], m.start[r + 1]):
            dense[r, m.positions[k]] = m.values[k]
    ref = np.zeros((5, 5), np.float32)
    ref[0, 0], ref[1, 1], ref[1, 4], ref[4, 1], ref[2, 2], ref[4, 4] = 3, 12, 34, 34, np.float32(-0.3), -4
    assert np.array_equal(dense, ref)


@pytest.mark.parametrize("text,status", [
    ("%%NotMatrixMarket matrix coordinate real symmetric\n1 1 1\n1 1 1\n", 4),
    ("%%MatrixMarket tensor coordinate real symmetric\n1 1 1\n1 1 1\n", 5),
    ("%%MatrixMarket matrix array real symmetric\n1 1 1\n1 1 1\n", 6),
    ("%%MatrixMarket matrix coordinate complex symmetric\n1 1 1\n1 1 1\n", 7),
    ("%%MatrixMarket matrix coordinate real general\n1 1 1\n1 1 1\n", 8),
    ("%%MatrixMarket matrix coordinate real symmetric\n% c\nx y z\n", 3),
    ("%%MatrixMarket MATRIX Coordinate INTEGER Symmetric\n%c\n\n2 2 2\n1 1 2\n2 1 -1", 0),
])
def test_loader_status_codes(tmp_path, text, status):
    # H:2531-2609 error paths; tokens other than the banner are case-insensitive
    p = tmp_path / "m.mtx"
    p.write_text(text)
    st, m = ol.load_mtx(str(p))
    assert st == status
    if status == 0:
        assert m.nnz == 3 and list(m.values) == [2, -1, -1]
    assert ol.load_mtx(str(tmp_path / "missing.mtx"))[0] == 1


# ---------------------------------------------------------------------------------------------
# 2. golden outputs of the real reference
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("key", ["load_symmetric_test"] + ASSETS)
def test_loader_matches_reference_csr(golden, key):
    st, m = ol.load_mtx(os.path.join(GOLD, key + ".mtx"))
    g = gold_csr(golden, key)
    assert st == 0
    assert np.array_equal(m.start, g.start) and np.array_equal(m.positions, g.positions)
    assert m.values.tobytes() == g.values.tobytes()
    assert m.first_active_start == g.first_active_start


@pytest.mark.parametrize("key", GENERATED)
def test_generators_are_stable(golden, key):
    g = gold_csr(golden, key)
    m = {"poisson2d_96x100": lambda: matgen.poisson2d(96, 100), "convdiff3d_22": lambda: matgen.convdiff3d(22),
         "powerlaw_9000": lambda: matgen.powerlaw(9000)}[key]()
    assert np.array_equal(m.start, g.start) and np.array_equal(m.positions, g.positions)
    assert m.values.tobytes() == g.values.tobytes()


@pytest.mark.parametrize("key", ASSETS + GENERATED)
@pytest.mark.parametrize("mt", [0, 1])
def test_kernels_match_reference(golden, key, mt):
    m = gold_csr(golden, key)
    tag = "mt" if mt else "st"
    b = golden[f"{key}/b"]
    assert ol.spmv(m, 2, b, b).tobytes() == golden[f"{key}/{tag}/spmv_sub"].tobytes()
    assert np.float32(ol.dot(b, b, mt)) == golden[f"{key}/{tag}/dot_bb"]
    rc, y = ol.sgs_apply(m, b)
    assert rc == int(golden[f"{key}/{tag}/sgs_rc"]) and y.tobytes() == golden[f"{key}/{tag}/sgs_b"].tobytes()


def _solver_cases(golden_path=os.path.join(GOLD, "golden_v1.npz")):
    names = [k[: -len("/status_iterations_eps")] for k in np.load(golden_path).files if k.endswith("/status_iterations_eps")]
    return sorted(names)


@pytest.mark.parametrize("name", _solver_cases())
def test_solvers_match_reference(golden, name):
    key, tag, solver = name.split("/")
    pre = solver.endswith("_sgs")
    solver = solver.replace("_sgs", "")
    m = gold_csr(golden, key)
    st, it, eps = golden[name + "/status_iterations_eps"]
    ic0 = ol.ic0_factorize(m)[1] if solver == "cg_ic0" else None
    o = ol.solve(solver, m, golden[f"{key}/b"], np.zeros(m.rows, np.float32), -1, np.float32(eps), tag == "mt", precond=int(pre), ic0=ic0)
    assert o["status"] == int(st)
    assert o["iterations"] == int(it)
    assert o["x"].tobytes() == golden[name + "/x"].tobytes()


def test_asset_iteration_table(golden):
    # SURVEY.md section 8(c) table (float, eps 1e-4, b = row sums, x0 = 0)
    table = {"mesh1e1": (13, 13, 7, 8, 3, 5), "mesh1em1": (24, 24, 15, 17, 5, 8), "mesh1em6": (13, 13, 8, 8, 3, 5)}
    for key, row in table.items():
        for solver, it in zip(["cg", "bicgsym", "cgs", "bicgstab", "bicgstab_sgs", "cg_ic0"], row):
            for tag in ("st", "mt"):
                assert int(golden[f"{key}/{tag}/{solver}/status_iterations_eps"][1]) == it
                assert np.abs(golden[f"{key}/{tag}/{solver}/x"] - 1).max() < 1e-4      # cg.cpp:23-25 criterion


# ---------------------------------------------------------------------------------------------
# 3. live comparison with the reference (build container only)
# ---------------------------------------------------------------------------------------------
needs_ref = pytest.mark.skipif(not ol.ref_available(), reason="oracle/_ref not built (no /root/reference here)")


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_live_triplets_and_kernels(seed):
    rng = np.random.default_rng(seed)
    rows, cols, n = 200, 150, 5000
    trow = rng.integers(0, rows, n).astype(np.int32)
    tcol = rng.integers(0, cols, n).astype(np.int32)
    trow[trow == 7] = 8                                       # an empty row
    tval = rng.uniform(-1, 1, n).astype(np.float32)
    m = ol.triplets_to_csr(rows, cols, trow, tcol, tval)
    start = np.zeros(rows + 1, np.int32); pos = np.zeros(n, np.int32); val = np.zeros(n, np.float32)
    import ctypes as C
    fas = C.c_int()
    nnz = ol.ref(0).smm_ref_triplets_to_csr(rows, cols, n, trow, tcol, tval, start, pos, val, C.byref(fas))
    assert nnz == m.nnz and np.array_equal(start, m.start) and np.array_equal(pos[:nnz], m.positions)
    assert val[:nnz].tobytes() == m.values.tobytes() and fas.value == m.first_active_start
    x = rng.uniform(-1, 1, cols).astype(np.float32)
    lhs = rng.uniform(-1, 1, rows).astype(np.float32)
    for mt in (0, 1):
        R = ol.RefCSR(m, mt)
        for op in (0, 1, 2):
            assert R.spmv(op, lhs, x).tobytes() == ol.spmv(m, op, lhs, x).tobytes()


@needs_ref
@pytest.mark.parametrize("n", [1, 100, 8192, 8193, 16385, 100003])
def test_live_dot_tree(n):
    rng = np.random.default_rng(n)
    a = rng.uniform(-1, 1, n).astype(np.float32)
    b = rng.uniform(-1, 1, n).astype(np.float32)
    for mt in (0, 1):
        assert np.float32(ol.dot(a, b, mt)) == np.float32(ol.ref(mt).smm_ref_dot(n, a, b))


@needs_ref
@pytest.mark.parametrize("solver", ["cg", "bicgsym", "cgs", "bicgstab"])
@pytest.mark.parametrize("maxit", [0, 1, 3, -7])
def test_live_iteration_cap_quirks(solver, maxit):
    # do-while solvers run the body once even for maxIterations <= 0 and then report MAX_ITERATIONS_REACHED (H:2098)
    m = matgen.poisson2d(20, 20)
    b = ol.spmv(m, 0, None, matgen.xstar(m.rows))
    for mt in (0, 1):
        o = ol.solve(solver, m, b, np.zeros(m.rows, np.float32), maxit, 1e-6, mt)
        st, x = ol.RefCSR(m, mt).solve(solver, b, np.zeros(m.rows, np.float32), maxit, 1e-6)
        assert st == o["status"] and x.tobytes() == o["x"].tobytes()
