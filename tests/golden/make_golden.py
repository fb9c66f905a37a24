"""Regenerate tests/golden/*.npz and the .mtx loader fixtures from the REAL reference.

Run in the build container only (needs /root/reference and `make -C oracle ref`):
    python tests/golden/make_golden.py

What is recorded comes from oracle/_ref/libsmm_ref_{st,mt}.so, i.e. the unmodified reference header
(plus the 2-line scope fix that lets GCC compile ConjugateGradientSquared), serial build and
-DSMM_MULTITHREADING build.  Iteration counts are not reported by the reference's API; they are found
with the public API alone: the smallest maxIterations whose result is bit-identical to the
maxIterations=-1 result (and, for ConjugateGradient, the smallest that returns SUCCESS).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import matgen  # noqa: E402
import oracle_lib as ol  # noqa: E402

ASSETS = "/root/reference/test/assets/"
ASSET_FILES = {
    "load_symmetric_test": "load_symmetric_test.mtx",
    "mesh1e1": "mesh1e1_structural_48_48_177.mtx",
    "mesh1em1": "mesh1em1_structural_48_48_177.mtx",
    "mesh1em6": "mesh1em6_structural_48_48_177.mtx",
    "sherman1": "sherman1_1000_1000_2375.mtx",
}
SOLVERS = [("cg", 0), ("bicgsym", 0), ("cgs", 0), ("bicgstab", 0), ("bicgstab", 1), ("cg_ic0", 0)]


def rowsum(m):
    # test/include/test_common.h:13-22 sumColumsPerRow: sequential float accumulation per row
    b = np.zeros(m.rows, np.float32)
    for r in range(m.rows):
        acc = np.float32(0)
        for k in range(m.start[r], m.start[r + 1]):
            acc = np.float32(acc + m.values[k])
        b[r] = acc
    return b


def find_iterations(R, solver, pre, b, x0, eps, hint):
    """Smallest maxIterations reproducing the unconstrained run (public API only)."""
    st_full, x_full = R.solve(solver, b, x0, -1, eps, precond=pre)

    def same(k):
        st, x = R.solve(solver, b, x0, k, eps, precond=pre)
        return st == st_full and x.tobytes() == x_full.tobytes()

    lo, hi = max(hint - 2, 0), hint + 2
    assert same(hi), (solver, hint)
    k = hi
    while k > 0 and same(k - 1):
        k -= 1
    assert k >= lo
    return st_full, x_full, k


def write_mtx_lower(path, m, comment):
    """Re-emit a symmetric matrix as MatrixMarket (lower triangle, 9 significant digits: float round trip)."""
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real symmetric\n")
        f.write("% " + comment + "\n")
        ents = [(r, int(m.positions[k]), m.values[k]) for r in range(m.rows) for k in range(m.start[r], m.start[r + 1]) if m.positions[k] <= r]
        f.write(f"{m.rows} {m.cols} {len(ents)}\n")
        for r, c, v in ents:
            f.write(f"{r + 1} {c + 1} {np.float32(v):.9g}\n")


def main():
    assert ol.ref_available(), "run `make -C oracle ref` first"
    out = {}
    meta = []
    mats = {}
    for key, fn in ASSET_FILES.items():
        st, m = ol.ref_load_matrix(ASSETS + fn)
        assert st == 0
        mats[key] = m
        write_mtx_lower(os.path.join(HERE, key + ".mtx"), m, f"re-emitted from the CSR the reference loads for its asset {fn}")
    mats["poisson2d_96x100"] = matgen.poisson2d(96, 100)
    mats["convdiff3d_22"] = matgen.convdiff3d(22)
    mats["powerlaw_9000"] = matgen.powerlaw(9000)

    for key, m in mats.items():
        out[f"{key}/start"] = m.start
        out[f"{key}/positions"] = m.positions
        out[f"{key}/values"] = m.values
        out[f"{key}/shape"] = np.array([m.rows, m.cols, m.first_active_start], np.int32)
        if key == "load_symmetric_test":
            continue
        generated = key not in ASSET_FILES
        if generated:
            xs = matgen.xstar(m.rows)
            b = ol.RefCSR(m, 0).spmv(0, None, xs)
            eps = 1e-5
        else:
            b = rowsum(m)
            eps = 1e-4
        out[f"{key}/b"] = b
        x0 = np.zeros(m.rows, np.float32)
        for mt in (0, 1):
            R = ol.RefCSR(m, mt)
            tag = "mt" if mt else "st"
            # SpMV / dot / SGS straight from the reference
            out[f"{key}/{tag}/spmv_sub"] = R.spmv(2, b, b)
            out[f"{key}/{tag}/dot_bb"] = np.float32(R.lib.smm_ref_dot(m.rows, b, b))
            rc, y = R.sgs_apply(b)
            out[f"{key}/{tag}/sgs_rc"] = np.int32(rc)
            out[f"{key}/{tag}/sgs_b"] = y
            for solver, pre in SOLVERS:
                if solver == "cg_ic0" and (generated or key == "sherman1"):
                    continue  # O(rows^2) factorisation / not SPD
                if solver in ("cg", "bicgsym") and key in ("convdiff3d_22", "powerlaw_9000"):
                    continue  # non-symmetric
                o = ol.solve(solver, m, b, x0, -1, eps, mt, precond=pre, ic0=ol.ic0_factorize(m)[1] if solver == "cg_ic0" else None)
                st, x, it = find_iterations(R, solver, pre, b, x0, eps, o["iterations"])
                name = f"{key}/{tag}/{solver}{'_sgs' if pre else ''}"
                out[name + "/x"] = x
                out[name + "/status_iterations_eps"] = np.array([st, it, eps], np.float64)
                meta.append((name, st, it))
                print(f"{name:45s} status {st} iterations {it:5d}  (oracle port: {o['iterations']}, x bit-equal {o['x'].tobytes() == x.tobytes()})")
    # IC0 known answer of test/cpp/cg.cpp:28-60, recomputed by the reference
    trow = [0, 0, 1, 1, 2, 3, 3, 3, 4, 4, 4]
    tcol = [3, 0, 1, 4, 2, 0, 3, 4, 1, 3, 4]
    tval = [4, 10, 9, 5, 12, 4, 15, 7, 5, 7, 8]
    m = ol.triplets_to_csr(5, 5, trow, tcol, tval)
    rc, ic0, x = ol.RefCSR(m, 0).ic0(np.ones(5, np.float32))
    out["ic0_5x5/apply_ones"] = x
    out["ic0_5x5/factor"] = ic0[: m.nnz]
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote", os.path.join(HERE, "golden_v1.npz"), os.path.getsize(os.path.join(HERE, "golden_v1.npz")), "bytes")


if __name__ == "__main__":
    main()
