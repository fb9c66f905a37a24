// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C-ABI shim around the REAL reference (vasil-pashov/sparse_matrix_math, include/sparse_matrix_math.h),
// compiled where it lies under /root/reference by oracle/Makefile into oracle/_ref/libsmm_ref_{st,mt}.so.
// Nothing of the reference is copied into this repository: the Makefile writes a patched copy of the
// header (2-line scope fix in ConjugateGradientSquared, H:2131/H:2171, without which GCC rejects the
// header) to a scratch directory OUTSIDE the repository and compiles this file against it; only the
// resulting shared libraries land in the git-ignored oracle/_ref/ directory.
//
// Used for (1) pinning oracle/smm_oracle.c (tests/test_oracle_pinned.py, tests/golden/make_golden.py),
// (2) the CPU baseline / `bench.py --impl reference` arm.  Never linked into the product library.
//
// `#define private public` (after the standard headers are in) gives this harness direct access to the
// CSR arrays so that 1e8-1e9-entry matrices can be ingested without the std::map based TripletMatrix.
#include <algorithm>
#include <cassert>
#include <cctype>
#include <cinttypes>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>
#if defined(SMM_MULTITHREADING)
#include <tbb/blocked_range.h>
#include <tbb/parallel_for.h>
#include <tbb/parallel_reduce.h>
#endif

#define private public
#include "sparse_matrix_math.h"  // patched copy of the reference header made by oracle/Makefile (kept outside the repo)
#undef private

#ifdef _OPENMP
#include <omp.h>
#endif

using Csr = SMM::CSRMatrix<float>;

extern "C" {

int smm_ref_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// torchrun exports OMP_NUM_THREADS=1 to every rank; the timing harness asks for all host cores explicitly
void smm_ref_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int smm_ref_multithreaded(void) {
#if defined(SMM_MULTITHREADING)
    return 1;
#else
    return 0;
#endif
}

// TripletMatrix::addEntry + CSRMatrix(const TripletMatrix&)
int smm_ref_triplets_to_csr(int rows, int cols, int64_t n, const int* trow, const int* tcol, const float* tval,
                            int* start, int* positions, float* values, int* first_active_start) {
    SMM::TripletMatrix<float> t(rows, cols);
    for (int64_t i = 0; i < n; ++i) t.addEntry(trow[i], tcol[i], tval[i]);
    Csr m;
    m.init(t);
    const int nnz = m.getNonZeroCount();
    std::memcpy(start, m.start.get(), sizeof(int) * (rows + 1));
    std::memcpy(positions, m.positions.get(), sizeof(int) * nnz);
    std::memcpy(values, m.values.get(), sizeof(float) * nnz);
    *first_active_start = m.firstActiveStart;
    return nnz;
}

// Direct CSR ingest (private-array access).  The arrays are copied.
void* smm_ref_csr_create(int rows, int cols, const int* start, const int* positions, const float* values) {
    Csr* m = new Csr();
    const int64_t nnz = start[rows];
    m->denseRowCount = rows;
    m->denseColCount = cols;
    m->start.reset(new int[rows + 1]);
    m->positions.reset(new int[nnz > 0 ? nnz : 1]);
    m->values.reset(new float[nnz > 0 ? nnz : 1]);
    std::memcpy(m->start.get(), start, sizeof(int) * (size_t(rows) + 1));
    std::memcpy(m->positions.get(), positions, sizeof(int) * size_t(nnz));
    std::memcpy(m->values.get(), values, sizeof(float) * size_t(nnz));
    int fas = rows;
    for (int i = 0; i < rows; ++i) {
        if (start[i + 1] != 0) { fas = i; break; }
    }
    m->firstActiveStart = fas;
    return m;
}

// Same, but takes ownership of arrays allocated with new[] by smm_ref_alloc_* (no copy; used for 512^3).
int* smm_ref_alloc_int(int64_t n) { return new int[n > 0 ? n : 1]; }
float* smm_ref_alloc_float(int64_t n) { return new float[n > 0 ? n : 1]; }
void* smm_ref_csr_adopt(int rows, int cols, int* start, int* positions, float* values) {
    Csr* m = new Csr();
    m->denseRowCount = rows;
    m->denseColCount = cols;
    m->start.reset(start);
    m->positions.reset(positions);
    m->values.reset(values);
    int fas = rows;
    for (int i = 0; i < rows; ++i) {
        if (start[i + 1] != 0) { fas = i; break; }
    }
    m->firstActiveStart = fas;
    return m;
}

void smm_ref_csr_destroy(void* h) { delete static_cast<Csr*>(h); }

// rMult / rMultAdd / rMultSub
void smm_ref_spmv(void* h, int op, const float* lhs, const float* mult, float* out) {
    Csr* m = static_cast<Csr*>(h);
    if (op == 0) m->rMult(mult, out);
    else if (op == 1) m->rMultAdd(lhs, mult, out);
    else m->rMultSub(lhs, mult, out);
}

// Vector::operator*
float smm_ref_dot(int n, const float* a, const float* b) {
    SMM::Vector<float> va, vb;
    va.data = const_cast<float*>(a); va.size = n;
    vb.data = const_cast<float*>(b); vb.size = n;
    const float r = va * vb;
    va.data = nullptr; va.size = 0;
    vb.data = nullptr; vb.size = 0;
    return r;
}

int smm_ref_sgs_apply(void* h, const float* rhs, float* x) {
    Csr* m = static_cast<Csr*>(h);
    const auto& M = m->getPreconditioner<SMM::SolverPreconditioner::SYMMETRIC_GAUS_SEIDEL>();
    return M.apply(rhs, x);
}

// IC0: factorize, copy the factor out (nnz floats), apply
int smm_ref_ic0(void* h, float* ic0_out, const float* rhs, float* x) {
    Csr* m = static_cast<Csr*>(h);
    Csr::IC0Preconditioner M(*m);
    const int rc = M.init();
    if (rc) return rc;
    if (ic0_out) std::memcpy(ic0_out, M.ic0Val.get(), sizeof(float) * m->getNonZeroCount());
    if (rhs && x) return M.apply(rhs, x);
    return 0;
}

int smm_ref_cg(void* h, const float* b, const float* x0, float* x, int maxIterations, float eps) {
    return int(SMM::ConjugateGradient<float>(*static_cast<Csr*>(h), b, x0, x, maxIterations, eps));
}

int smm_ref_cg_ic0(void* h, const float* b, const float* x0, float* x, int maxIterations, float eps) {
    Csr* m = static_cast<Csr*>(h);
    Csr::IC0Preconditioner M(*m);
    if (M.init()) return -1;
    return int(SMM::ConjugateGradient<float>(*m, b, x0, x, maxIterations, eps, M));
}

int smm_ref_bicgsym(void* h, float* b, float* x, int maxIterations, float eps) {
    return int(SMM::BiCGSymmetric<float>(*static_cast<Csr*>(h), b, x, maxIterations, eps));
}

int smm_ref_cgs(void* h, float* b, float* x, int maxIterations, float eps) {
    return int(SMM::ConjugateGradientSquared<float>(*static_cast<Csr*>(h), b, x, maxIterations, eps));
}

int smm_ref_bicgstab(void* h, int precond, float* b, float* x, int maxIterations, float eps) {
    Csr* m = static_cast<Csr*>(h);
    if (precond == 3) {                                        // the template over any object with apply(): IC0Preconditioner
        Csr::IC0Preconditioner M(*m);
        if (M.init()) return -1;
        return int(SMM::BiCGStab<Csr::IC0Preconditioner, float>(*m, b, x, maxIterations, eps, M));
    }
    if (precond) {
        using SGS = Csr::SGSPreconditioner;
        const SGS& M = m->getPreconditioner<SMM::SolverPreconditioner::SYMMETRIC_GAUS_SEIDEL>();
        return int(SMM::BiCGStab<SGS, float>(*m, b, x, maxIterations, eps, M));
    }
    return int(SMM::BiCGStab<float>(*m, b, x, maxIterations, eps));
}

// loadMatrix(path, CSRMatrix&): returns MatrixLoadStatus; on success *h receives a CSR handle.
int smm_ref_load_matrix(const char* path, void** h, int* rows, int* cols, int* nnz, int* first_active_start) {
    Csr* m = new Csr();
    const SMM::MatrixLoadStatus st = SMM::loadMatrix(path, *m);
    if (st != SMM::MatrixLoadStatus::SUCCESS) {
        delete m;
        *h = nullptr;
        return int(st);
    }
    *h = m;
    *rows = m->getDenseRowCount();
    *cols = m->getDenseColCount();
    *nnz = m->getNonZeroCount();
    *first_active_start = m->firstActiveStart;
    return 0;
}

void smm_ref_csr_export(void* h, int* start, int* positions, float* values) {
    Csr* m = static_cast<Csr*>(h);
    const int rows = m->getDenseRowCount();
    const int nnz = m->getNonZeroCount();
    std::memcpy(start, m->start.get(), sizeof(int) * (rows + 1));
    std::memcpy(positions, m->positions.get(), sizeof(int) * nnz);
    std::memcpy(values, m->values.get(), sizeof(float) * nnz);
}

}  // extern "C"
