/*
 * oracle/smm_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the Krylov-solve hot path of
 * vasil-pashov/sparse_matrix_math, written from the behaviour of
 * include/sparse_matrix_math.h ("H:n" below = line n of that header).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or
 * the reported CPU baseline.  Nothing under sparse_matrix_math_b200/ or
 * include/ links, imports or calls it.
 *
 * Parity is PINNED: tests/test_oracle_pinned.py checks every function here
 * against (a) the golden vectors of the reference's own tests
 * (test/cpp/csr.cpp, cg.cpp, bicgstab.cpp ...) and (b) outputs of the
 * reference itself, compiled unmodified-but-for-a-2-line-scope-fix into
 * oracle/_ref/ by oracle/Makefile (fixtures committed in tests/golden/).
 *
 * Arithmetic model: T = float, `a*x+b` is two roundings (H:27-37, the
 * reference's default build has no FMA contraction); this file is compiled
 * with -ffp-contract=off so gcc cannot fuse them either.
 */
#ifndef SMM_ORACLE_H
#define SMM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* SolverStatus, H:2010-2014 */
enum { SMM_ORACLE_SUCCESS = 0, SMM_ORACLE_DIVERGED = 1, SMM_ORACLE_MAX_ITERATIONS_REACHED = 2 };

/* Reduction flavours of Vector::operator* (H:305-328). */
enum {
    SMM_ORACLE_DOT_SERIAL = 0,   /* #else branch: left-to-right */
    SMM_ORACLE_DOT_TBB8192 = 1   /* SMM_MULTITHREADING: parallel_deterministic_reduce, grain 8192 */
};

typedef struct {
    int status;          /* SolverStatus */
    int iterations;      /* loop trips executed (SpMV A*p calls) */
    float residual;      /* last value the solver compared with eps (squared for CG/BiCGSym/CGS, L2 for BiCGStab) */
    int precond_error;   /* OR of non-zero apply() return codes (BiCGStab only) */
} smm_oracle_info;

/* TripletMatrix::addEntry (H:606-618) + CSRMatrix::fillArrays (H:1606-1641).
 * Triplets are given in CALL order; duplicates are summed in that order; explicit zeros kept.
 * start has rows+1 entries; positions/values must hold at least n_triplets entries.
 * Returns nnz (>=0). *first_active_start receives firstActiveStart (H:1622-1628). */
int smm_oracle_triplets_to_csr(int rows, int cols, int64_t n_triplets,
                               const int *trow, const int *tcol, const float *tval,
                               int *start, int *positions, float *values, int *first_active_start);

/* CSRMatrix::rMultOp (H:1458-1499). op: 0 = rMult (out = A*mult), 1 = rMultAdd, 2 = rMultSub.
 * out may alias lhs.  lhs is ignored for op 0. */
void smm_oracle_spmv(int rows, const int *start, const int *positions, const float *values,
                     int op, const float *lhs, const float *mult, float *out);

/* Vector::operator* (H:305-328). */
float smm_oracle_dot(int n, const float *a, const float *b, int dot_mode);

/* SGSPreconditioner::apply (H:1658-1713). Returns 0, or 1 on the reference's error exits. */
int smm_oracle_sgs_apply(int rows, const int *start, const int *positions, const float *values,
                         int first_active_start, const float *rhs, float *x);

/* IC0Preconditioner::factorize / apply (H:1839-1928, H:1802-1837). ic0 has nnz entries. */
int smm_oracle_ic0_factorize(int rows, const int *start, const int *positions, const float *values, float *ic0);
int smm_oracle_ic0_apply(int rows, const int *start, const int *positions, const float *ic0,
                         const float *rhs, float *x);

/* Solvers.  `mt` selects the SMM_MULTITHREADING build's arithmetic (dot tree; CG's separate r*r) when
 * non-zero, the serial build's otherwise.  history (may be NULL) receives the residual quantity after
 * every iteration, up to history_cap entries. */
void smm_oracle_cg(int rows, const int *start, const int *positions, const float *values,
                   const float *b, const float *x0, float *x, int max_iterations, float eps, int mt,
                   smm_oracle_info *info, float *history, int history_cap);          /* H:2316-2398 */
void smm_oracle_bicgsym(int rows, const int *start, const int *positions, const float *values,
                        const float *b, float *x, int max_iterations, float eps, int mt,
                        smm_oracle_info *info, float *history, int history_cap);     /* H:2021-2102 */
void smm_oracle_cgs(int rows, const int *start, const int *positions, const float *values,
                    const float *b, float *x, int max_iterations, float eps, int mt,
                    smm_oracle_info *info, float *history, int history_cap);         /* H:2109-2178 */
/* precond: 0 = IDPreconditioner, 1 = SGSPreconditioner */
void smm_oracle_bicgstab(int rows, const int *start, const int *positions, const float *values,
                         int first_active_start, int precond,
                         const float *b, float *x, int max_iterations, float eps, int mt,
                         smm_oracle_info *info, float *history, int history_cap);    /* H:2191-2283 */
/* precond: 0 / 1 as above, 2 = ILU(0) (extension), 3 = IC(0), 4 = Jacobi (extension); factor = the values of 2 / 3 (else NULL) */
void smm_oracle_bicgstab_pc(int rows, const int *start, const int *positions, const float *values,
                            int first_active_start, int precond, const float *factor,
                            const float *b, float *x, int max_iterations, float eps, int mt,
                            smm_oracle_info *info, float *history, int history_cap);
/* EXTENSION, parity unpinned by the reference (its ILU0Preconditioner is dead code, H:1188-1212, 1715-1790): the
 * zero-fill LU those lines describe, and the two triangular solves in the shape of IC0Preconditioner::apply. */
int smm_oracle_ilu0_factorize(int rows, const int *start, const int *positions, const float *values,
                              int first_active_start, float *ilu0);
int smm_oracle_ilu0_apply(int rows, const int *start, const int *positions, const float *ilu0,
                          const float *rhs, float *x);
int smm_oracle_jacobi_apply(int rows, const int *start, const int *positions, const float *values,
                            const float *rhs, float *x);
/* PCG with IC0 (H:2414-2505). */
void smm_oracle_cg_ic0(int rows, const int *start, const int *positions, const float *values,
                       const float *ic0, const float *b, const float *x0, float *x,
                       int max_iterations, float eps, int mt,
                       smm_oracle_info *info, float *history, int history_cap);

/* loadMatrixMarketMatrix (H:2531-2609) -> triplets in call order (mirrored entries included).
 * Returns MatrixLoadStatus (H:2507-2522). Caller frees *trow,*tcol,*tval with smm_oracle_free. */
int smm_oracle_load_mtx(const char *path, int *rows, int *cols, int64_t *n_triplets,
                        int **trow, int **tcol, float **tval);
void smm_oracle_free(void *p);

/* Benchmark inputs on the host (same bits as tests/matgen.py and the device generators): 7-point
 * convection-diffusion / Poisson stencil on an nx*ny*nz grid (use_z = 0: the 5-point 2D Poisson with diag 4).
 * start has rows+1 entries; positions/values hold smm_oracle_stencil_nnz() entries.  OpenMP-parallel. */
int64_t smm_oracle_stencil_nnz(int nx, int ny, int nz, int use_z);
void smm_oracle_gen_stencil(int nx, int ny, int nz, int use_z, float lo, float diag, float hi,
                            int *start, int *positions, float *values);
/* x*_i = (splitmix64(seed, i) >> 40) / 2^24 */
void smm_oracle_gen_xstar(int64_t n, uint64_t seed, float *x);

/* Number of OpenMP threads the row/vector loops use (1 if built without OpenMP). */
int smm_oracle_threads(void);
/* Override OMP_NUM_THREADS (torchrun exports 1 to every rank; the timing legs want all host cores). */
void smm_oracle_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
