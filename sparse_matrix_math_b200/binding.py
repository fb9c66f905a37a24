"""ctypes binding of libsmm_b200.so (include/smm_b200.h) with the reference's vocabulary on top."""
import ctypes as C
import enum
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

REDUCE_FAST, REDUCE_REFERENCE_TREE, REDUCE_REFERENCE_SERIAL = 0, 1, 2
DRIVER_AUTO, DRIVER_GRAPH_CHUNKED, DRIVER_GRAPH_WHILE, DRIVER_STREAM, DRIVER_PERSISTENT = 0, 1, 2, 3, 4
OP_ASSIGN, OP_ADD, OP_SUB = 0, 1, 2
GEN_POISSON2D, GEN_CONVDIFF3D, GEN_POWERLAW = 0, 1, 2


class SmmError(RuntimeError):
    pass


class SolverStatus(enum.IntEnum):          # H:2010-2014
    SUCCESS = 0
    DIVERGED = 1
    MAX_ITERATIONS_REACHED = 2


class SolverPreconditioner(enum.IntEnum):  # H:1002-1006 (+ README spelling)
    NONE = 0
    SYMMETRIC_GAUS_SEIDEL = 1
    SYMMETRIC_GAUSS_SEIDEL = 1
    ILU0 = 2
    JACOBI = 3                      # extension, not in the reference


class MatrixLoadStatus(enum.IntEnum):      # H:2507-2522
    SUCCESS = 0
    FAILED_TO_OPEN_FILE = 1
    FAILED_TO_OPEN_FILE_UNKNOWN_FORMAT = 2
    FAILED_TO_PARSE_FILE = 3
    PARSE_ERROR_MMX_FILE_MISSING_BANNER = 4
    PARSE_ERROR_MMX_FILE_UNSUPPORTED_TYPE = 5
    PARSE_ERROR_MMX_FILE_UNSUPPORTED_FORMAT = 6
    PARSE_ERROR_MMX_FILE_UNSUPPORTED_EL_TYPE = 7
    PARSE_ERROR_MMX_FILE_UNSUPPORTED_STRUCTURE = 8


class _Options(C.Structure):
    _fields_ = [("reduction_mode", C.c_int), ("driver_mode", C.c_int), ("check_every", C.c_int), ("history_cap", C.c_int),
                ("history", C.POINTER(C.c_float)), ("reserved", C.c_int * 4)]


class _Info(C.Structure):
    _fields_ = [("status", C.c_int), ("iterations", C.c_int), ("residual", C.c_float), ("precond_error", C.c_int),
                ("seconds_solve", C.c_double), ("seconds_total", C.c_double), ("reduction_mode", C.c_int),
                ("driver_mode", C.c_int), ("kernel_launches", C.c_longlong), ("reserved", C.c_int * 4)]


class SolveInfo:
    def __init__(self, info, history=None):
        self.status = SolverStatus(info.status)
        self.iterations = info.iterations
        self.residual = float(info.residual)
        self.precond_error = info.precond_error
        self.seconds_solve = info.seconds_solve
        self.seconds_total = info.seconds_total
        self.reduction_mode = info.reduction_mode
        self.driver_mode = info.driver_mode
        self.kernel_launches = info.kernel_launches
        self.history = history

    def __repr__(self):
        return (f"SolveInfo(status={self.status.name}, iterations={self.iterations}, residual={self.residual:.6g}, "
                f"seconds_solve={self.seconds_solve:.6f}, launches={self.kernel_launches})")


def lib_path():
    return os.path.join(HERE, "libsmm_b200.so")


def header_symbols():
    """Every function declared in include/smm_b200.h (the ABI contract the tests check the .so against)."""
    text = open(os.path.join(ROOT, "include", "smm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(smm_[a-z0-9_]+)\s*\(", text)))


ABI_SYMBOLS = header_symbols()

_lib = None
_vp, _i32, _f32 = C.c_void_p, C.c_int, C.c_float


def lib():
    """Load libsmm_b200.so.  Fails loudly when it has not been built -- there is no other code path."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise SmmError(f"{path} is missing: build it with `python -m sparse_matrix_math_b200.build` "
                           "(the package has no CPU fallback)")
        L = C.CDLL(path)
        L.smm_last_error.restype = C.c_char_p
        L.smm_kernel_launch_count.restype = C.c_longlong
        L.smm_csr_create.argtypes = [_i32, _i32, _vp, _vp, _vp, C.POINTER(_vp)]
        L.smm_csr_create_dev.argtypes = [_i32, _i32, _vp, _vp, _vp, _i32, C.POINTER(_vp)]
        L.smm_csr_update_values.argtypes = [_vp, _vp]
        L.smm_csr_destroy.argtypes = [_vp]
        L.smm_csr_shape.argtypes = [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(C.c_int64), C.POINTER(_i32)]
        L.smm_csr_download.argtypes = [_vp, _vp, _vp, _vp]
        L.smm_csr_device_arrays.argtypes = [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]
        L.smm_spmv.argtypes = [_vp, _i32, _vp, _vp, _vp]
        L.smm_spmv_dev.argtypes = [_vp, _i32, _vp, _vp, _vp, _i32, _vp]
        L.smm_dot.argtypes = [C.c_int64, _vp, _vp, _i32, C.POINTER(_f32)]
        L.smm_dot_dev.argtypes = [C.c_int64, _vp, _vp, _i32, C.POINTER(_f32), _vp]
        L.smm_precond_sgs_create.argtypes = [_vp, C.POINTER(_vp)]
        L.smm_precond_apply.argtypes = [_vp, _vp, _vp, C.POINTER(_i32)]
        L.smm_precond_apply_dev.argtypes = [_vp, _vp, _vp, C.POINTER(_i32), _vp]
        L.smm_precond_levels.argtypes = [_vp, C.POINTER(_i32), C.POINTER(_i32)]
        L.smm_precond_tile_levels.argtypes = [_vp, C.POINTER(_i32), C.POINTER(_i32)]
        L.smm_precond_schedule.argtypes = [_vp]
        L.smm_precond_layout_fingerprint.argtypes = [_vp, C.POINTER(C.c_uint64)]
        L.smm_precond_destroy.argtypes = [_vp]
        L.smm_precond_ic0_create.argtypes = [_vp, C.POINTER(_i32), C.POINTER(_vp)]
        L.smm_precond_ic0_factor.argtypes = [_vp, _vp]
        L.smm_precond_ilu0_create.argtypes = [_vp, C.POINTER(_i32), C.POINTER(_vp)]
        L.smm_precond_jacobi_create.argtypes = [_vp, C.POINTER(_vp)]
        L.smm_precond_kind.argtypes = [_vp]
        op, ip = C.POINTER(_Options), C.POINTER(_Info)
        L.smm_solve_cg.argtypes = [_vp, _vp, _vp, _vp, _i32, _f32, op, ip]
        L.smm_solve_cg_ic0.argtypes = [_vp, _vp, _vp, _vp, _vp, _i32, _f32, op, ip]
        L.smm_solve_cg_ic0_dev.argtypes = [_vp, _vp, _vp, _vp, _vp, _i32, _f32, op, ip, _vp]
        L.smm_solve_bicgsym.argtypes = [_vp, _vp, _vp, _i32, _f32, op, ip]
        L.smm_solve_cgs.argtypes = [_vp, _vp, _vp, _i32, _f32, op, ip]
        L.smm_solve_bicgstab.argtypes = [_vp, _vp, _vp, _vp, _i32, _f32, op, ip]
        L.smm_solve_cg_dev.argtypes = [_vp, _vp, _vp, _vp, _i32, _f32, op, ip, _vp]
        L.smm_solve_bicgsym_dev.argtypes = [_vp, _vp, _vp, _i32, _f32, op, ip, _vp]
        L.smm_solve_cgs_dev.argtypes = [_vp, _vp, _vp, _i32, _f32, op, ip, _vp]
        L.smm_solve_bicgstab_dev.argtypes = [_vp, _vp, _vp, _vp, _i32, _f32, op, ip, _vp]
        L.smm_gen_csr.argtypes = [_i32, _i32, _i32, _i32, _f32, C.c_uint64, C.POINTER(_vp)]
        L.smm_gen_xstar_dev.argtypes = [C.c_int64, C.c_int64, C.c_uint64, _vp, _vp]
        L.smm_profile_cg_iteration.argtypes = [_vp, _i32, C.POINTER(_f32), C.POINTER(_f32), C.POINTER(_f32), _vp]
        L.smm_malloc_dev.argtypes = [C.c_size_t, C.POINTER(_vp)]
        L.smm_free_dev.argtypes = [_vp]
        L.smm_memcpy_h2d.argtypes = [_vp, _vp, C.c_size_t]
        L.smm_memcpy_d2h.argtypes = [_vp, _vp, C.c_size_t]
        L.smm_memset_dev.argtypes = [_vp, _i32, C.c_size_t]
        L.smm_device_info.argtypes = [C.POINTER(_i32), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.smm_device_count.argtypes = [C.POINTER(_i32)]
        L.smm_set_device.argtypes = [_i32]
        _lib = L
    return _lib


def _check(rc, what):
    if rc != 0:
        raise SmmError(f"{what} failed with code {rc}: {lib().smm_last_error().decode(errors='replace')}")


def kernel_launch_count():
    return int(lib().smm_kernel_launch_count())


def device_info():
    sm, l2, tot, free = _i32(), C.c_size_t(), C.c_size_t(), C.c_size_t()
    _check(lib().smm_device_info(C.byref(sm), C.byref(l2), C.byref(tot), C.byref(free)), "smm_device_info")
    return dict(sm_count=sm.value, l2_bytes=l2.value, total_mem=tot.value, free_mem=free.value)


def _f32arr(a):
    return np.ascontiguousarray(a, np.float32)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


class DeviceVector:
    """A float vector in HBM owned through the C ABI (smm_malloc_dev)."""

    def __init__(self, n, host=None):
        self.n = int(n)
        p = _vp()
        _check(lib().smm_malloc_dev(max(self.n, 1) * 4, C.byref(p)), "smm_malloc_dev")
        self.ptr = p.value
        if host is not None:
            self.upload(host)

    def upload(self, host):
        host = _f32arr(host)
        assert host.size == self.n
        if self.n:
            _check(lib().smm_memcpy_h2d(self.ptr, _ptr(host), self.n * 4), "smm_memcpy_h2d")

    def zero(self):
        if self.n:
            _check(lib().smm_memset_dev(self.ptr, 0, self.n * 4), "smm_memset_dev")

    def download(self):
        out = np.empty(self.n, np.float32)
        if self.n:
            _check(lib().smm_memcpy_d2h(_ptr(out), self.ptr, self.n * 4), "smm_memcpy_d2h")
        return out

    def free(self):
        if getattr(self, "ptr", None):
            lib().smm_free_dev(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class TripletMatrix:
    """SMM::TripletMatrix<float> (H:487-684): std::map keyed (row<<32)|col; duplicates sum in call order."""

    def __init__(self, rows=0, cols=0, num_triplets=0):
        self._rows, self._cols = int(rows), int(cols)
        self._data = {}

    def init(self, rows, cols, num_triplets=0):
        assert not self._data and self._rows == 0 and self._cols == 0
        self._rows, self._cols = int(rows), int(cols)

    def deinit(self):
        self._rows = self._cols = 0
        self._data.clear()

    def addEntry(self, row, col, value):
        key = (int(row) << 32) | int(col)
        v = np.float32(value)
        if key in self._data:
            self._data[key] = np.float32(self._data[key] + v)   # H:616
        else:
            self._data[key] = v                                 # H:614

    def updateEntry(self, row, col, value):
        key = (int(row) << 32) | int(col)
        if key in self._data:
            self._data[key] = np.float32(value)
            return True
        return False

    def getValue(self, row, col):
        return float(self._data.get((int(row) << 32) | int(col), np.float32(0)))

    def getNonZeroCount(self):
        return len(self._data)

    def getDenseRowCount(self):
        return self._rows

    def getDenseColCount(self):
        return self._cols

    def __imul__(self, scalar):
        s = np.float32(scalar)
        for k in self._data:
            self._data[k] = np.float32(self._data[k] * s)
        return self

    def __iter__(self):
        for key in sorted(self._data):
            yield key >> 32, key & 0xFFFFFFFF, float(self._data[key])

    def to_csr_arrays(self):
        """CSRMatrix::fillArrays, H:1606-1641."""
        keys = np.array(sorted(self._data), np.uint64)
        rows = (keys >> np.uint64(32)).astype(np.int64)
        start = np.zeros(self._rows + 1, np.int64)
        np.add.at(start, rows + 1, 1)
        np.cumsum(start, out=start)
        positions = (keys & np.uint64(0xFFFFFFFF)).astype(np.int32)
        values = np.array([self._data[int(k)] for k in keys], np.float32)
        return start.astype(np.int32), positions, values


class SGSPreconditioner:
    """CSRMatrix<float>::SGSPreconditioner (H:1172-1186), from CSRMatrix.getPreconditioner()."""

    def __init__(self, matrix):
        self.matrix = matrix
        h = _vp()
        _check(lib().smm_precond_sgs_create(matrix.handle, C.byref(h)), "smm_precond_sgs_create")
        self.handle = h.value

    def apply(self, rhs, x=None):
        """int apply(const T* rhs, T* x): returns (code, x)."""
        rhs = _f32arr(rhs)
        if x is None:
            x = np.zeros(self.matrix.rows, np.float32)
        rc = _i32()
        _check(lib().smm_precond_apply(self.handle, _ptr(rhs), _ptr(x), C.byref(rc)), "smm_precond_apply")
        return rc.value, x

    def apply_dev(self, rhs_ptr, x_ptr, stream=None):
        rc = _i32()
        _check(lib().smm_precond_apply_dev(self.handle, rhs_ptr, x_ptr, C.byref(rc), stream), "smm_precond_apply_dev")
        return rc.value

    def levels(self):
        f, b = _i32(), _i32()
        _check(lib().smm_precond_levels(self.handle, C.byref(f), C.byref(b)), "smm_precond_levels")
        return f.value, b.value

    def tile_levels(self):
        """(forward, backward) levels of the tile graph, (0, 0) when the sweeps run row by row."""
        f, b = _i32(), _i32()
        _check(lib().smm_precond_tile_levels(self.handle, C.byref(f), C.byref(b)), "smm_precond_tile_levels")
        return f.value, b.value

    def schedule(self):
        """0: rows in level order, 1: tiles, 2: lines (which schedule the triangular sweeps of this handle run)."""
        return int(lib().smm_precond_schedule(self.handle))

    def layout_fingerprint(self):
        """FNV-1a fingerprints of the layout arrays the sweep kernels read (smm_precond_layout_fingerprint): a tuple of 13."""
        out = (C.c_uint64 * 13)()
        _check(lib().smm_precond_layout_fingerprint(self.handle, out), "smm_precond_layout_fingerprint")
        return tuple(int(v) for v in out)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                lib().smm_precond_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class IC0Preconditioner(SGSPreconditioner):
    """CSRMatrix<float>::IC0Preconditioner (H:1214-1235): IC0Preconditioner(m); init(); apply(rhs, x)."""

    def __init__(self, matrix):
        self.matrix = matrix
        self.handle = None
        self.init_code = None

    def init(self):
        h, rc = _vp(), _i32()
        _check(lib().smm_precond_ic0_create(self.matrix.handle, C.byref(rc), C.byref(h)), "smm_precond_ic0_create")
        self.handle, self.init_code = h.value, rc.value
        return rc.value

    def factor(self):
        out = np.zeros(max(self.matrix.nnz, 1), np.float32)
        _check(lib().smm_precond_ic0_factor(self.handle, _ptr(out)), "smm_precond_ic0_factor")
        return out[: self.matrix.nnz]


class JacobiPreconditioner(SGSPreconditioner):
    """EXTENSION (not in the reference): diagonal preconditioner, apply is x = rhs / diag(A) element-wise."""

    def __init__(self, matrix):
        self.matrix = matrix
        h = _vp()
        _check(lib().smm_precond_jacobi_create(matrix.handle, C.byref(h)), "smm_precond_jacobi_create")
        self.handle = h.value


class ILU0Preconditioner(IC0Preconditioner):
    """EXTENSION: CSRMatrix<float>::ILU0Preconditioner (dead code in the reference, H:1188-1212, 1715-1790) made to
    work: ILU0Preconditioner(m); validate() factorises (0 ok, 1 structure, 2 pivot); apply(rhs, x); BiCGStab takes it."""

    def validate(self):
        h, rc = _vp(), _i32()
        _check(lib().smm_precond_ilu0_create(self.matrix.handle, C.byref(rc), C.byref(h)), "smm_precond_ilu0_create")
        self.handle, self.init_code = h.value, rc.value
        return rc.value

    init = validate


class CSRMatrix:
    """SMM::CSRMatrix<float> (H:1011-1302) with its arrays resident in HBM."""

    def __init__(self, triplet=None):
        self.handle = None
        self.rows = self.cols = 0
        self.nnz = 0
        self.first_active_start = 0
        if triplet is not None:
            self.init(triplet)

    # -- construction ---------------------------------------------------------------------------
    def init(self, triplet):
        start, positions, values = triplet.to_csr_arrays()
        return self.init_arrays(triplet.getDenseRowCount(), triplet.getDenseColCount(), start, positions, values)

    def init_arrays(self, rows, cols, start, positions, values):
        """Additive: direct CSR ingest (the reference can only go through TripletMatrix)."""
        self._release()
        start = np.ascontiguousarray(start, np.int32)
        positions = np.ascontiguousarray(positions, np.int32)
        values = _f32arr(values)
        h = _vp()
        _check(lib().smm_csr_create(int(rows), int(cols), _ptr(start), _ptr(positions), _ptr(values), C.byref(h)), "smm_csr_create")
        self.handle = h.value
        self._read_shape()
        return 0

    @classmethod
    def from_arrays(cls, rows, cols, start, positions, values):
        m = cls()
        m.init_arrays(rows, cols, start, positions, values)
        return m

    @classmethod
    def generate(cls, kind, nx, ny=0, nz=0, c=0.0, seed=0x5EED):
        m = cls()
        h = _vp()
        if kind != GEN_POWERLAW:
            ny, nz = max(int(ny), 1), max(int(nz), 1)
        _check(lib().smm_gen_csr(kind, int(nx), int(ny), int(nz), float(c), int(seed), C.byref(h)), "smm_gen_csr")
        m.handle = h.value
        m._read_shape()
        return m

    def _read_shape(self):
        r, c, n, f = _i32(), _i32(), C.c_int64(), _i32()
        _check(lib().smm_csr_shape(self.handle, C.byref(r), C.byref(c), C.byref(n), C.byref(f)), "smm_csr_shape")
        self.rows, self.cols, self.nnz, self.first_active_start = r.value, c.value, n.value, f.value

    def _release(self):
        if getattr(self, "handle", None):
            lib().smm_csr_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    # -- reference accessors -----------------------------------------------------------------------
    def getNonZeroCount(self):
        return self.nnz

    def getDenseRowCount(self):
        return self.rows

    def getDenseColCount(self):
        return self.cols

    def download(self):
        start = np.empty(self.rows + 1, np.int32)
        positions = np.empty(max(self.nnz, 1), np.int32)
        values = np.empty(max(self.nnz, 1), np.float32)
        _check(lib().smm_csr_download(self.handle, _ptr(start), _ptr(positions), _ptr(values)), "smm_csr_download")
        return start, positions[: self.nnz], values[: self.nnz]

    def update_values(self, values):
        values = _f32arr(values)
        assert values.size == self.nnz
        _check(lib().smm_csr_update_values(self.handle, _ptr(values)), "smm_csr_update_values")

    # -- SpMV: rMult / rMultAdd / rMultSub, H:1458-1515 -----------------------------------------------
    def _spmv(self, op, lhs, mult, out):
        mult = _f32arr(mult)
        if out is None:
            out = np.zeros(self.rows, np.float32)
        _check(lib().smm_spmv(self.handle, op, _ptr(lhs), _ptr(mult), _ptr(out)), "smm_spmv")
        return out

    def rMult(self, mult, out=None):
        return self._spmv(OP_ASSIGN, None, mult, out)

    def rMultAdd(self, lhs, mult, out=None):
        return self._spmv(OP_ADD, lhs, mult, out)

    def rMultSub(self, lhs, mult, out=None):
        return self._spmv(OP_SUB, lhs, mult, out)

    def spmv_dev(self, op, lhs_ptr, mult_ptr, out_ptr, exact=False, stream=None):
        _check(lib().smm_spmv_dev(self.handle, op, lhs_ptr, mult_ptr, out_ptr, 1 if exact else 0, stream), "smm_spmv_dev")

    # -- preconditioner factory, H:1643-1651 ----------------------------------------------------------
    def getPreconditioner(self, kind):
        kind = SolverPreconditioner(kind)
        if kind == SolverPreconditioner.NONE:
            return None                         # IDPreconditioner
        if kind == SolverPreconditioner.SYMMETRIC_GAUS_SEIDEL:
            return SGSPreconditioner(self)
        if kind == SolverPreconditioner.JACOBI:
            return JacobiPreconditioner(self)
        # the reference's factory returns void for ILU0 (H:1645-1651); here it hands out the working extension
        M = ILU0Preconditioner(self)
        M.validate()
        return M


def _options(reduction_mode, driver_mode, check_every, history_cap):
    o = _Options()
    o.reduction_mode, o.driver_mode, o.check_every = reduction_mode, driver_mode, check_every
    hist = None
    if history_cap:
        hist = np.full(history_cap, np.nan, np.float32)
        o.history_cap = history_cap
        o.history = hist.ctypes.data_as(C.POINTER(C.c_float))
    return o, hist


def _solve(fn, name, args, reduction_mode, driver_mode, check_every, history_cap):
    o, hist = _options(reduction_mode, driver_mode, check_every, history_cap)
    info = _Info()
    _check(fn(*args, C.byref(o), C.byref(info)), name)
    return SolveInfo(info, hist)


def ConjugateGradient(a, b, x0, x, maxIterations, eps, M=None, reduction_mode=REDUCE_FAST, driver_mode=DRIVER_AUTO,
                      check_every=0, history_cap=0):
    """SMM::ConjugateGradient (H:2316-2398); with M = IC0Preconditioner the PCG overload (H:2414-2505).  x may be x0.
    Returns SolveInfo (status as the reference returns it)."""
    b = _f32arr(b)
    assert x.dtype == np.float32 and x0.dtype == np.float32
    if M is not None:
        return _solve(lib().smm_solve_cg_ic0, "smm_solve_cg_ic0", (a.handle, M.handle, _ptr(b), _ptr(x0), _ptr(x), int(maxIterations), float(eps)),
                      reduction_mode, driver_mode, check_every, history_cap)
    return _solve(lib().smm_solve_cg, "smm_solve_cg", (a.handle, _ptr(b), _ptr(x0), _ptr(x), int(maxIterations), float(eps)),
                  reduction_mode, driver_mode, check_every, history_cap)


def BiCGSymmetric(a, b, x, maxIterations, eps, reduction_mode=REDUCE_FAST, driver_mode=DRIVER_AUTO, check_every=0, history_cap=0):
    """SMM::BiCGSymmetric (H:2021-2102); x is initial guess and result."""
    b = _f32arr(b)
    assert x.dtype == np.float32
    return _solve(lib().smm_solve_bicgsym, "smm_solve_bicgsym", (a.handle, _ptr(b), _ptr(x), int(maxIterations), float(eps)),
                  reduction_mode, driver_mode, check_every, history_cap)


def ConjugateGradientSquared(a, b, x, maxIterations, eps, reduction_mode=REDUCE_FAST, driver_mode=DRIVER_AUTO, check_every=0,
                             history_cap=0):
    """SMM::ConjugateGradientSquared (H:2109-2178)."""
    b = _f32arr(b)
    assert x.dtype == np.float32
    return _solve(lib().smm_solve_cgs, "smm_solve_cgs", (a.handle, _ptr(b), _ptr(x), int(maxIterations), float(eps)),
                  reduction_mode, driver_mode, check_every, history_cap)


ConjugateGradientSqared = ConjugateGradientSquared   # README.md:57 spelling


def BiCGStab(a, b, x, maxIterations, eps, preconditioner=None, reduction_mode=REDUCE_FAST, driver_mode=DRIVER_AUTO,
             check_every=0, history_cap=0):
    """SMM::BiCGStab (H:2191-2303); preconditioner None = IDPreconditioner / the 5-argument overload."""
    b = _f32arr(b)
    assert x.dtype == np.float32
    ph = None if preconditioner is None else preconditioner.handle
    return _solve(lib().smm_solve_bicgstab, "smm_solve_bicgstab", (a.handle, ph, _ptr(b), _ptr(x), int(maxIterations), float(eps)),
                  reduction_mode, driver_mode, check_every, history_cap)


def dot(a, b, reduction_mode=REDUCE_FAST):
    """Vector::operator* (H:305-328)."""
    a, b = _f32arr(a), _f32arr(b)
    out = _f32()
    _check(lib().smm_dot(a.size, _ptr(a), _ptr(b), reduction_mode, C.byref(out)), "smm_dot")
    return float(out.value)


def loadMatrix(path, out, extended=False):
    """SMM::loadMatrix (H:2648-2669) for .mtx: host-side parse (H:2531-2609) -> TripletMatrix -> CSRMatrix.
    extended=True is SMM::ext::loadMatrix (extension): general / skew-symmetric structure and pattern files too."""
    from .mmio import load_matrix_market
    dot_pos = path.rfind(".")
    ext = path[dot_pos + 1:] if dot_pos >= 0 else ""
    if ext != "mtx":
        return MatrixLoadStatus.FAILED_TO_OPEN_FILE_UNKNOWN_FORMAT
    t = out if isinstance(out, TripletMatrix) else TripletMatrix()
    st = load_matrix_market(path, t, extended)
    if st != MatrixLoadStatus.SUCCESS:
        return st
    if isinstance(out, CSRMatrix):
        out.init(t)
    return MatrixLoadStatus.SUCCESS
