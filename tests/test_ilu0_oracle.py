"""ILU(0) is an EXTENSION -- the reference's ILU0Preconditioner is dead code (factorize() cannot succeed,
apply() is undefined), so parity is UNPINNED by the reference.  The oracle restates the algorithm H:1723-1790 describes;
these CPU tests pin that restatement by the defining properties of a zero-fill LU instead:
  * (L U)_ij == a_ij on A's pattern (to rounding), * exact LU (== a direct solve) when the pattern admits no fill,
  * the error codes of validate(), * as a preconditioner it cuts BiCGStab's iteration count.
BiCGStab instantiated with the IC0Preconditioner, which the reference's template allows, IS pinned: against oracle/_ref."""
import numpy as np
import pytest
import scipy.sparse as sp

import matgen
import oracle_lib as ol


def factors(m, lu):
    a = sp.csr_matrix((lu[: m.nnz].astype(np.float64), m.positions, m.start), shape=(m.rows, m.cols))
    L = sp.tril(a, -1) + sp.identity(m.rows)
    U = sp.triu(a, 0)
    return L.tocsr(), U.tocsr()


@pytest.mark.parametrize("gen", ["convdiff3d_9", "poisson2d_17x13", "powerlaw_600"])
def test_ilu0_reproduces_a_on_its_pattern(gen):
    m = {"convdiff3d_9": lambda: matgen.convdiff3d(9, 0.5), "poisson2d_17x13": lambda: matgen.poisson2d(17, 13),
         "powerlaw_600": lambda: matgen.powerlaw(600)}[gen]()
    rc, lu = ol.ilu0_factorize(m)
    assert rc == 0
    L, U = factors(m, lu)
    prod = (L @ U).tocsr()
    a = sp.csr_matrix((m.values.astype(np.float64), m.positions, m.start), shape=(m.rows, m.cols))
    mask = a.copy(); mask.data[:] = 1.0
    on_pattern = prod.multiply(mask) - a
    assert abs(on_pattern).max() <= 2e-6 * abs(a).max()
    # apply == U^-1 L^-1 rhs
    rhs = matgen.xstar(m.rows)
    x = ol.ilu0_apply(m, lu, rhs)
    y = sp.linalg.spsolve_triangular(L, rhs.astype(np.float64), lower=True, unit_diagonal=True)
    want = sp.linalg.spsolve_triangular(U, y, lower=False)
    assert np.max(np.abs(x - want)) <= 1e-5 * np.max(np.abs(want))


def test_ilu0_is_exact_lu_without_fill():
    n = 200                                             # tridiagonal: LU has no fill, so ILU(0) solves A x = b
    rows = np.repeat(np.arange(n), 3)[1:-1]
    cols = (np.repeat(np.arange(n), 3) + np.tile([-1, 0, 1], n))[1:-1]
    vals = np.tile([-1.0, 2.5, -0.5], n)[1:-1]
    m = ol.triplets_to_csr(n, n, rows, cols, vals)
    rc, lu = ol.ilu0_factorize(m)
    assert rc == 0
    xs = matgen.xstar(n)
    b = ol.spmv(m, 0, None, xs)
    assert np.max(np.abs(ol.ilu0_apply(m, lu, b) - xs)) < 1e-5


def test_ilu0_error_codes():
    assert ol.ilu0_factorize(ol.triplets_to_csr(3, 3, [1, 2], [1, 2], [1, 1]))[0] == 1          # leading empty row (H:1734-1737)
    assert ol.ilu0_factorize(ol.triplets_to_csr(3, 3, [0, 1, 1, 2], [0, 0, 2, 2], [1, 1, 1, 1]))[0] == 1   # missing diagonal
    assert ol.ilu0_factorize(ol.triplets_to_csr(2, 2, [0, 0, 1, 1], [0, 1, 0, 1], [1, 1, 1, 1]))[0] == 2   # pivot 1 - 1*1 = 0
    assert ol.ilu0_factorize(ol.triplets_to_csr(0, 0, [], [], []))[0] == 0


def test_ilu0_preconditions_bicgstab():
    m = matgen.convdiff3d(14, 0.5)
    b = ol.spmv(m, 0, None, matgen.xstar(m.rows))
    plain = ol.solve("bicgstab", m, b, np.zeros(m.rows, np.float32), -1, 1e-5, 1)
    lu = ol.ilu0_factorize(m)[1]
    pre = ol.solve("bicgstab", m, b, np.zeros(m.rows, np.float32), -1, 1e-5, 1, precond=2, factor=lu)
    assert pre["status"] == 0 and pre["residual"] <= 1e-5 and pre["iterations"] < plain["iterations"]
    assert np.max(np.abs(pre["x"] - matgen.xstar(m.rows))) < 1e-3


@pytest.mark.parametrize("mt", [0, 1])
@pytest.mark.parametrize("key", ["mesh1e1", "mesh1em1", "poisson2d_96x100"])
def test_bicgstab_over_ic0_matches_the_reference(golden, key, mt):
    """BiCGStab<IC0Preconditioner, float> is a legal instantiation of the reference's template (H:2191-2199)."""
    if not ol.ref_available():
        pytest.skip("oracle/_ref not built")
    rows, cols, fas = golden[f"{key}/shape"]
    g = ol.CSR(rows, cols, golden[f"{key}/start"], golden[f"{key}/positions"], golden[f"{key}/values"], fas)
    b = golden[f"{key}/b"]
    f = ol.ic0_factorize(g)[1]
    o = ol.solve("bicgstab", g, b, np.zeros(g.rows, np.float32), -1, 1e-4, mt, precond=3, factor=f)
    r = ol.RefCSR(g, mt)
    st, x = r.solve("bicgstab", b, np.zeros(g.rows, np.float32), -1, 1e-4, precond=3)[:2]
    assert o["status"] == st == 0 and o["x"].tobytes() == x.tobytes()
