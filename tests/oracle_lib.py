"""ctypes bindings for the CHECKERS: oracle/libsmm_oracle.so (the C restatement) and, when present,
oracle/_ref/libsmm_ref_{st,mt}.so (the real reference header compiled by oracle/Makefile).

TEST INFRASTRUCTURE ONLY -- imported by tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke(); never by sparse_matrix_math_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


class Info(C.Structure):
    _fields_ = [("status", C.c_int), ("iterations", C.c_int), ("residual", C.c_float), ("precond_error", C.c_int)]


def build_oracle():
    """(Re)build the C restatement; cheap, so tests call it unconditionally."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True)


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "libsmm_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        lib = C.CDLL(path)
        lib.smm_oracle_triplets_to_csr.restype = C.c_int
        lib.smm_oracle_triplets_to_csr.argtypes = [C.c_int, C.c_int, C.c_int64, _i32p, _i32p, _f32p, _i32p, _i32p, _f32p, C.POINTER(C.c_int)]
        lib.smm_oracle_spmv.restype = None
        lib.smm_oracle_spmv.argtypes = [C.c_int, _i32p, _i32p, _f32p, C.c_int, C.c_void_p, _f32p, _f32p]
        lib.smm_oracle_dot.restype = C.c_float
        lib.smm_oracle_dot.argtypes = [C.c_int, _f32p, _f32p, C.c_int]
        lib.smm_oracle_sgs_apply.restype = C.c_int
        lib.smm_oracle_sgs_apply.argtypes = [C.c_int, _i32p, _i32p, _f32p, C.c_int, _f32p, _f32p]
        lib.smm_oracle_ic0_factorize.restype = C.c_int
        lib.smm_oracle_ic0_factorize.argtypes = [C.c_int, _i32p, _i32p, _f32p, _f32p]
        lib.smm_oracle_ic0_apply.restype = C.c_int
        lib.smm_oracle_ic0_apply.argtypes = [C.c_int, _i32p, _i32p, _f32p, _f32p, _f32p]
        common = [C.c_int, _i32p, _i32p, _f32p]
        tail = [C.c_int, C.c_float, C.c_int, C.POINTER(Info), C.c_void_p, C.c_int]
        lib.smm_oracle_cg.restype = None
        lib.smm_oracle_cg.argtypes = common + [_f32p, _f32p, _f32p] + tail
        lib.smm_oracle_bicgsym.restype = None
        lib.smm_oracle_bicgsym.argtypes = common + [_f32p, _f32p] + tail
        lib.smm_oracle_cgs.restype = None
        lib.smm_oracle_cgs.argtypes = common + [_f32p, _f32p] + tail
        lib.smm_oracle_bicgstab.restype = None
        lib.smm_oracle_bicgstab.argtypes = common + [C.c_int, C.c_int, _f32p, _f32p] + tail
        lib.smm_oracle_bicgstab_pc.restype = None
        lib.smm_oracle_bicgstab_pc.argtypes = common + [C.c_int, C.c_int, C.c_void_p, _f32p, _f32p] + tail
        lib.smm_oracle_ilu0_factorize.restype = C.c_int
        lib.smm_oracle_ilu0_factorize.argtypes = [C.c_int, _i32p, _i32p, _f32p, C.c_int, _f32p]
        lib.smm_oracle_ilu0_apply.restype = C.c_int
        lib.smm_oracle_ilu0_apply.argtypes = [C.c_int, _i32p, _i32p, _f32p, _f32p, _f32p]
        lib.smm_oracle_jacobi_apply.restype = C.c_int
        lib.smm_oracle_jacobi_apply.argtypes = [C.c_int, _i32p, _i32p, _f32p, _f32p, _f32p]
        lib.smm_oracle_cg_ic0.restype = None
        lib.smm_oracle_cg_ic0.argtypes = common + [_f32p, _f32p, _f32p, _f32p] + tail
        lib.smm_oracle_load_mtx.restype = C.c_int
        lib.smm_oracle_load_mtx.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64),
                                            C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.POINTER(C.c_float))]
        lib.smm_oracle_free.restype = None
        lib.smm_oracle_free.argtypes = [C.c_void_p]
        lib.smm_oracle_threads.restype = C.c_int
        _oracle = lib
    return _oracle


# ------------------------------------------------------------------------------------------------
# numpy-level wrappers around the oracle
# ------------------------------------------------------------------------------------------------
class CSR:
    """Plain container: start[rows+1] int32, positions[nnz] int32, values[nnz] float32."""

    def __init__(self, rows, cols, start, positions, values, first_active_start=None):
        self.rows, self.cols = int(rows), int(cols)
        self.start = np.ascontiguousarray(start, np.int32)
        self.positions = np.ascontiguousarray(positions, np.int32)
        self.values = np.ascontiguousarray(values, np.float32)
        if first_active_start is None:
            nz = np.nonzero(self.start[1:] != 0)[0]
            first_active_start = int(nz[0]) if len(nz) else self.rows
        self.first_active_start = int(first_active_start)

    @property
    def nnz(self):
        return int(self.start[self.rows])

    def _pad(self):
        # ctypes ndpointer rejects 0-length views in some numpy builds: keep at least one element
        pos = self.positions if self.positions.size else np.zeros(1, np.int32)
        val = self.values if self.values.size else np.zeros(1, np.float32)
        return pos, val


def triplets_to_csr(rows, cols, trow, tcol, tval):
    trow = np.ascontiguousarray(trow, np.int32)
    tcol = np.ascontiguousarray(tcol, np.int32)
    tval = np.ascontiguousarray(tval, np.float32)
    n = len(trow)
    start = np.zeros(rows + 1, np.int32)
    pos = np.zeros(max(n, 1), np.int32)
    val = np.zeros(max(n, 1), np.float32)
    fas = C.c_int(0)
    pad = lambda a, dt: a if a.size else np.zeros(1, dt)
    nnz = oracle().smm_oracle_triplets_to_csr(rows, cols, n, pad(trow, np.int32), pad(tcol, np.int32), pad(tval, np.float32),
                                              start, pos, val, C.byref(fas))
    return CSR(rows, cols, start, pos[:nnz].copy(), val[:nnz].copy(), fas.value)


def spmv(m, op, lhs, mult, out=None):
    mult = np.ascontiguousarray(mult, np.float32)
    if out is None:
        out = np.zeros(m.rows, np.float32)
    lhs_p = None if lhs is None else lhs.ctypes.data_as(C.c_void_p)
    pos, val = m._pad()
    oracle().smm_oracle_spmv(m.rows, m.start, pos, val, op, lhs_p, mult if mult.size else np.zeros(1, np.float32), out if out.size else np.zeros(1, np.float32))
    return out


def dot(a, b, mt):
    return float(oracle().smm_oracle_dot(len(a), np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32), 1 if mt else 0))


def sgs_apply(m, rhs):
    x = np.zeros(m.rows, np.float32)
    pos, val = m._pad()
    rc = oracle().smm_oracle_sgs_apply(m.rows, m.start, pos, val, m.first_active_start, np.ascontiguousarray(rhs, np.float32), x)
    return rc, x


def ic0_factorize(m):
    ic0 = np.zeros(max(m.nnz, 1), np.float32)
    rc = oracle().smm_oracle_ic0_factorize(m.rows, m.start, m.positions, m.values, ic0)
    return rc, ic0


def ic0_apply(m, ic0, rhs):
    x = np.zeros(m.rows, np.float32)
    oracle().smm_oracle_ic0_apply(m.rows, m.start, m.positions, ic0, np.ascontiguousarray(rhs, np.float32), x)
    return x


def ilu0_factorize(m):
    """EXTENSION (parity unpinned by the reference): returns (rc, factor)."""
    lu = np.zeros(max(m.nnz, 1), np.float32)
    pos, val = m._pad()
    rc = oracle().smm_oracle_ilu0_factorize(m.rows, m.start, pos, val, m.first_active_start, lu)
    return rc, lu


def ilu0_apply(m, lu, rhs):
    x = np.zeros(m.rows, np.float32)
    oracle().smm_oracle_ilu0_apply(m.rows, m.start, m.positions, lu, np.ascontiguousarray(rhs, np.float32), x)
    return x


def jacobi_apply(m, rhs):
    """EXTENSION: x = D^-1 rhs.  Returns (rc, x)."""
    x = np.zeros(m.rows, np.float32)
    pos, val = m._pad()
    rc = oracle().smm_oracle_jacobi_apply(m.rows, m.start, pos, val, np.ascontiguousarray(rhs, np.float32), x)
    return rc, x


def _hist(cap):
    if not cap:
        return None, None
    h = np.full(cap, np.nan, np.float32)
    return h, h.ctypes.data_as(C.c_void_p)


def solve(solver, m, b, x0, max_iterations, eps, mt, precond=0, ic0=None, history_cap=0, factor=None):
    """Run an oracle solver.  Returns dict(status, iterations, residual, precond_error, x, history).
    bicgstab: precond 0 none, 1 SGS, 2 ILU(0) (factor = ilu0_factorize), 3 IC(0) (factor = ic0_factorize), 4 Jacobi."""
    lib = oracle()
    info = Info()
    b = np.ascontiguousarray(b, np.float32).copy()
    x0 = np.ascontiguousarray(x0, np.float32).copy()
    x = x0.copy()
    h, hp = _hist(history_cap)
    pos, val = m._pad()
    args = (m.rows, m.start, pos, val)
    tail = (int(max_iterations), float(eps), 1 if mt else 0, C.byref(info), hp, history_cap)
    if solver == "cg":
        lib.smm_oracle_cg(*args, b, x0, x, *tail)
    elif solver == "cg_ic0":
        lib.smm_oracle_cg_ic0(*args, ic0, b, x0, x, *tail)
    elif solver == "bicgsym":
        lib.smm_oracle_bicgsym(*args, b, x, *tail)
    elif solver == "cgs":
        lib.smm_oracle_cgs(*args, b, x, *tail)
    elif solver == "bicgstab":
        fp = None if factor is None else np.ascontiguousarray(factor, np.float32).ctypes.data_as(C.c_void_p)
        lib.smm_oracle_bicgstab_pc(*args, m.first_active_start, precond, fp, b, x, *tail)
    else:
        raise ValueError(solver)
    return dict(status=info.status, iterations=info.iterations, residual=float(info.residual),
                precond_error=info.precond_error, x=x, history=h)


def load_mtx(path):
    rows, cols, n = C.c_int(), C.c_int(), C.c_int64()
    tr, tc, tv = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(C.c_float)()
    lib = oracle()
    st = lib.smm_oracle_load_mtx(path.encode(), C.byref(rows), C.byref(cols), C.byref(n), C.byref(tr), C.byref(tc), C.byref(tv))
    try:
        if st != 0:
            return st, None
        k = n.value
        trow = np.ctypeslib.as_array(tr, (k,)).copy() if k else np.zeros(0, np.int32)
        tcol = np.ctypeslib.as_array(tc, (k,)).copy() if k else np.zeros(0, np.int32)
        tval = np.ctypeslib.as_array(tv, (k,)).copy() if k else np.zeros(0, np.float32)
    finally:
        lib.smm_oracle_free(tr); lib.smm_oracle_free(tc); lib.smm_oracle_free(tv)
    return 0, triplets_to_csr(rows.value, cols.value, trow, tcol, tval)


# ------------------------------------------------------------------------------------------------
# the real reference (optional: only where oracle/_ref was built)
# ------------------------------------------------------------------------------------------------
_ref = {}


def ref_available():
    return all(os.path.exists(os.path.join(ORACLE_DIR, "_ref", f"libsmm_ref_{k}.so")) for k in ("st", "mt"))


def ref(mt):
    key = "mt" if mt else "st"
    if key not in _ref:
        lib = C.CDLL(os.path.join(ORACLE_DIR, "_ref", f"libsmm_ref_{key}.so"))
        lib.smm_ref_triplets_to_csr.restype = C.c_int
        lib.smm_ref_triplets_to_csr.argtypes = [C.c_int, C.c_int, C.c_int64, _i32p, _i32p, _f32p, _i32p, _i32p, _f32p, C.POINTER(C.c_int)]
        lib.smm_ref_csr_create.restype = C.c_void_p
        lib.smm_ref_csr_create.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f32p]
        lib.smm_ref_csr_destroy.argtypes = [C.c_void_p]
        lib.smm_ref_spmv.argtypes = [C.c_void_p, C.c_int, C.c_void_p, _f32p, _f32p]
        lib.smm_ref_dot.restype = C.c_float
        lib.smm_ref_dot.argtypes = [C.c_int, _f32p, _f32p]
        lib.smm_ref_sgs_apply.restype = C.c_int
        lib.smm_ref_sgs_apply.argtypes = [C.c_void_p, _f32p, _f32p]
        lib.smm_ref_ic0.restype = C.c_int
        lib.smm_ref_ic0.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.smm_ref_cg.restype = C.c_int
        lib.smm_ref_cg.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, C.c_int, C.c_float]
        lib.smm_ref_cg_ic0.restype = C.c_int
        lib.smm_ref_cg_ic0.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, C.c_int, C.c_float]
        for name in ("smm_ref_bicgsym", "smm_ref_cgs"):
            getattr(lib, name).restype = C.c_int
            getattr(lib, name).argtypes = [C.c_void_p, _f32p, _f32p, C.c_int, C.c_float]
        lib.smm_ref_bicgstab.restype = C.c_int
        lib.smm_ref_bicgstab.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, C.c_int, C.c_float]
        lib.smm_ref_load_matrix.restype = C.c_int
        lib.smm_ref_load_matrix.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.smm_ref_csr_export.argtypes = [C.c_void_p, _i32p, _i32p, _f32p]
        lib.smm_ref_threads.restype = C.c_int
        _ref[key] = lib
    return _ref[key]


class RefCSR:
    """A CSRMatrix<float> living inside the reference library."""

    def __init__(self, m, mt):
        self.lib = ref(mt)
        self.m = m
        pos, val = m._pad()
        self.h = self.lib.smm_ref_csr_create(m.rows, m.cols, m.start, pos, val)

    def close(self):
        if self.h:
            self.lib.smm_ref_csr_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def spmv(self, op, lhs, mult, out=None):
        if out is None:
            out = np.zeros(self.m.rows, np.float32)
        lhs_p = None if lhs is None else lhs.ctypes.data_as(C.c_void_p)
        self.lib.smm_ref_spmv(self.h, op, lhs_p, np.ascontiguousarray(mult, np.float32), out)
        return out

    def sgs_apply(self, rhs):
        x = np.zeros(self.m.rows, np.float32)
        rc = self.lib.smm_ref_sgs_apply(self.h, np.ascontiguousarray(rhs, np.float32), x)
        return rc, x

    def ic0(self, rhs=None):
        ic0 = np.zeros(max(self.m.nnz, 1), np.float32)
        x = np.zeros(self.m.rows, np.float32)
        r = None if rhs is None else np.ascontiguousarray(rhs, np.float32)
        rc = self.lib.smm_ref_ic0(self.h, ic0.ctypes.data_as(C.c_void_p),
                                  None if r is None else r.ctypes.data_as(C.c_void_p),
                                  None if r is None else x.ctypes.data_as(C.c_void_p))
        return rc, ic0, x

    def solve(self, solver, b, x0, max_iterations, eps, precond=0):
        b = np.ascontiguousarray(b, np.float32).copy()
        x0 = np.ascontiguousarray(x0, np.float32).copy()
        x = x0.copy()
        if solver == "cg":
            st = self.lib.smm_ref_cg(self.h, b, x0, x, max_iterations, eps)
        elif solver == "cg_ic0":
            st = self.lib.smm_ref_cg_ic0(self.h, b, x0, x, max_iterations, eps)
        elif solver == "bicgsym":
            st = self.lib.smm_ref_bicgsym(self.h, b, x, max_iterations, eps)
        elif solver == "cgs":
            st = self.lib.smm_ref_cgs(self.h, b, x, max_iterations, eps)
        elif solver == "bicgstab":
            st = self.lib.smm_ref_bicgstab(self.h, precond, b, x, max_iterations, eps)
        else:
            raise ValueError(solver)
        return st, x


def ref_load_matrix(path, mt=False):
    lib = ref(mt)
    h = C.c_void_p()
    rows, cols, nnz, fas = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    st = lib.smm_ref_load_matrix(path.encode(), C.byref(h), C.byref(rows), C.byref(cols), C.byref(nnz), C.byref(fas))
    if st != 0:
        return st, None
    start = np.zeros(rows.value + 1, np.int32)
    pos = np.zeros(max(nnz.value, 1), np.int32)
    val = np.zeros(max(nnz.value, 1), np.float32)
    lib.smm_ref_csr_export(h, start, pos, val)
    lib.smm_ref_csr_destroy(h)
    return 0, CSR(rows.value, cols.value, start, pos[:nnz.value].copy(), val[:nnz.value].copy(), fas.value)
