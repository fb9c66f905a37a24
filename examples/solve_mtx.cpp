// solve_mtx.cpp -- the drop-in header in use: load a Matrix Market file, solve A x = A 1 with one of the library's
// Krylov solvers on the GPU, report status / iterations / error.  The code below is what a user of the reference
// (vasil-pashov/sparse_matrix_math) already has; only the include path and the link line change:
//
//   g++ -std=c++17 -O2 -Iinclude examples/solve_mtx.cpp -Lsparse_matrix_math_b200 -lsmm_b200
//       -Wl,-rpath,$PWD/sparse_matrix_math_b200 -L/usr/local/cuda/lib64 -Wl,-rpath,/usr/local/cuda/lib64 -o solve_mtx
//   ./solve_mtx tests/golden/sherman1.mtx bicgstab-sgs 1e-4
//
// solver: cg | bicgsym | cgs | bicgstab | bicgstab-sgs | bicgstab-ilu0 (extension) | cg-ic0
// A file the reference's loader rejects (general / skew-symmetric / pattern) is read with SMM::ext::loadMatrix.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "sparse_matrix_math.h"

int main(int argc, char** argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s matrix.mtx [solver] [eps] [reference-order: 0|1]\n", argv[0]);
        return 2;
    }
    const std::string solver = argc > 2 ? argv[2] : "cg";
    const float eps = argc > 3 ? static_cast<float>(std::atof(argv[3])) : 1e-4f;
    if (argc > 4 && std::atoi(argv[4]) != 0)                       // same bits and iteration count as the reference's
        SMM::b200::options().reduction_mode = SMM_REDUCE_REFERENCE_TREE;   // SMM_MULTITHREADING build

    SMM::CSRMatrix<float> a;
    SMM::MatrixLoadStatus st = SMM::loadMatrix(argv[1], a);
    if (st == SMM::MatrixLoadStatus::PARSE_ERROR_MMX_FILE_UNSUPPORTED_STRUCTURE ||
        st == SMM::MatrixLoadStatus::PARSE_ERROR_MMX_FILE_UNSUPPORTED_EL_TYPE)
        st = SMM::ext::loadMatrix(argv[1], a);                     // extension: general / skew-symmetric / pattern files
    if (st != SMM::MatrixLoadStatus::SUCCESS) {
        std::fprintf(stderr, "cannot load %s (status %d)\n", argv[1], static_cast<int>(st));
        return 1;
    }
    const int n = a.getDenseRowCount();
    SMM::Vector<float> ones(n, 1.0f), b(n, 0.0f), x(n, 0.0f);
    a.rMult(ones, b);                                              // b = A * 1, so the answer is known

    SMM::SolverStatus status = SMM::SolverStatus::DIVERGED;
    if (solver == "cg") {
        status = SMM::ConjugateGradient<float>(a, b, x, x, -1, eps);
    } else if (solver == "bicgsym") {
        status = SMM::BiCGSymmetric<float>(a, b, x, -1, eps);
    } else if (solver == "cgs") {
        status = SMM::ConjugateGradientSquared<float>(a, b, x, -1, eps);
    } else if (solver == "bicgstab") {
        status = SMM::BiCGStab<float>(a, b, x, -1, eps);
    } else if (solver == "bicgstab-sgs") {
        using SGS = SMM::CSRMatrix<float>::SGSPreconditioner;
        const SGS& m = a.getPreconditioner<SMM::SolverPreconditioner::SYMMETRIC_GAUS_SEIDEL>();
        status = SMM::BiCGStab<SGS, float>(a, b, x, -1, eps, m);
    } else if (solver == "bicgstab-ilu0") {
        using ILU0 = SMM::CSRMatrix<float>::ILU0Preconditioner;
        const ILU0& m = a.getPreconditioner<SMM::SolverPreconditioner::ILU0>();
        status = SMM::BiCGStab<ILU0, float>(a, b, x, -1, eps, m);
    } else if (solver == "cg-ic0") {
        SMM::CSRMatrix<float>::IC0Preconditioner m(a);
        if (m.init() != 0) { std::fprintf(stderr, "IC(0) factorisation failed\n"); return 1; }
        status = SMM::ConjugateGradient<float>(a, b, x, x, -1, eps, m);
    } else {
        std::fprintf(stderr, "unknown solver %s\n", solver.c_str());
        return 2;
    }
    float err = 0.0f;
    for (const float xi : x) err = std::fmax(err, std::fabs(xi - 1.0f));
    const SMM::SolveInfo& info = SMM::b200::lastSolveInfo();       // additive: the reference does not report these
    std::printf("%s: %d rows, %d entries, solver %s, status %d, %d iterations, residual %.3e, max |x - 1| = %.3e, %.3f ms on the device\n",
                argv[1], n, a.getNonZeroCount(), solver.c_str(), static_cast<int>(status), info.iterations, static_cast<double>(info.residual),
                static_cast<double>(err), 1e3 * info.secondsSolve);
    return status == SMM::SolverStatus::SUCCESS ? 0 : 1;
}
