/*
 * smm_b200.h -- C ABI of the B200-native Krylov hot path (libsmm_b200.so).
 *
 * The reference (vasil-pashov/sparse_matrix_math) is a header-only C++17 template library with no FFI
 * layer; its hot path is reached through the inline members / free functions of
 * include/sparse_matrix_math.h ("H:n" = line n of that header).  This ABI is what the drop-in header of
 * this repository (include/sparse_matrix_math.h, namespace SMM) binds for T = float; each entry point
 * names the reference interface it replaces.  Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions
 *   - every function returns 0 on success, a negative SMM_E_* code on failure (never throws, never falls
 *     back to the CPU: without a CUDA device the calls fail with SMM_E_CUDA); smm_last_error() gives text.
 *   - pointers are HOST pointers unless the function name ends in _dev; *_dev entry points take device
 *     pointers and a cudaStream_t passed as void* (NULL = the library's own stream).
 *   - a handle is bound to the CUDA device that was current when it was created.
 *   - handles may be used from one thread at a time.
 */
#ifndef SMM_B200_H
#define SMM_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMM_B200_ABI_VERSION 1

/* error codes */
enum {
    SMM_OK = 0,
    SMM_E_INVALID = -1,   /* bad argument */
    SMM_E_CUDA = -2,      /* CUDA runtime error (no device, launch failure, out of memory ...) */
    SMM_E_ALIAS = -3,     /* forbidden aliasing (rMult: mult == out, H:1503; SGS apply: rhs == x, H:1667) */
    SMM_E_STATE = -4,     /* handle in the wrong state */
    SMM_E_TIMEOUT = -5    /* device-side wait exceeded its bound (sync-free triangular solve, halo wait) */
};

/* SMM::SolverStatus, H:2010-2014 */
enum { SMM_SOLVER_SUCCESS = 0, SMM_SOLVER_DIVERGED = 1, SMM_SOLVER_MAX_ITERATIONS_REACHED = 2 };

/* rMultOp functor, H:1501-1515 */
enum { SMM_OP_ASSIGN = 0 /* rMult */, SMM_OP_ADD = 1 /* rMultAdd */, SMM_OP_SUB = 2 /* rMultSub */ };

/* How dot products / norms are summed (Vector::operator*, H:305-328).
 *   FAST            two-stage deterministic GPU tree, fused into the producing kernel (throughput mode)
 *   REFERENCE_TREE  bit-for-bit the SMM_MULTITHREADING build: tbb::parallel_deterministic_reduce, grain 8192
 *                   (and sequential row sums in SpMV), so iteration counts equal the reference's exactly
 *   REFERENCE_SERIAL bit-for-bit the serial build: left-to-right (slow; parity runs on small inputs only) */
enum { SMM_REDUCE_FAST = 0, SMM_REDUCE_REFERENCE_TREE = 1, SMM_REDUCE_REFERENCE_SERIAL = 2 };

/* How the iteration loop is driven.
 *   GRAPH_CHUNKED  one CUDA graph per iteration, enqueued `check_every` at a time; kernels no-op once the
 *                  device-side convergence flag is set; host polls the flag once per chunk
 *   GRAPH_WHILE    one launch: a conditional WHILE graph node loops on the device until the convergence test
 *   STREAM         plain stream launches, flag polled every check_every iterations (debug)
 *   PERSISTENT     ConjugateGradient only (fast reductions, one GPU, stencil-like matrices): the whole loop in ONE
 *                  cooperative kernel with grid barriers instead of kernel boundaries -- for problems that live in L2, where
 *                  an iteration is mostly launch latency.  Same bits as the graph drivers (it executes their kernels' CTAs as
 *                  virtual CTAs); any other solve asked for in this mode runs GRAPH_CHUNKED and says so in smm_solve_info. */
enum { SMM_DRIVER_AUTO = 0, SMM_DRIVER_GRAPH_CHUNKED = 1, SMM_DRIVER_GRAPH_WHILE = 2, SMM_DRIVER_STREAM = 3, SMM_DRIVER_PERSISTENT = 4 };

typedef struct smm_csr smm_csr_t;          /* device-resident CSRMatrix<float> (H:1243-1259) + analysis */
typedef struct smm_precond smm_precond_t;  /* CSRMatrix::SGSPreconditioner (H:1172-1186) + level analysis */

typedef struct {
    int reduction_mode;   /* SMM_REDUCE_* (default FAST) */
    int driver_mode;      /* SMM_DRIVER_* (default AUTO) */
    int check_every;      /* iterations between host polls, 0 = library default */
    int history_cap;      /* entries available in `history` (0 = none) */
    float* history;       /* HOST buffer: the residual quantity after every iteration (as the solver compares it) */
    int reserved[4];
} smm_solve_options;

typedef struct {
    int status;             /* SMM_SOLVER_* exactly as the reference function would return */
    int iterations;         /* loop trips executed (= SpMV A*p calls) */
    float residual;         /* last value compared with eps: ||r||^2 (CG, BiCGSymmetric, CGS), ||r||_2 (BiCGStab) */
    int precond_error;      /* OR of non-zero preconditioner apply codes (the reference only asserts on them) */
    double seconds_solve;   /* device time of the solve (CUDA events), vectors already resident */
    double seconds_total;   /* wall time of the call including host<->device copies of b, x0, x */
    int reduction_mode;     /* what was actually used */
    int driver_mode;
    long long kernel_launches; /* kernels launched inside the timed region */
    int reserved[4];
} smm_solve_info;

/* ---- library ------------------------------------------------------------------------------------ */
int smm_abi_version(void);
const char* smm_last_error(void);
int smm_device_count(int* count);
int smm_set_device(int device);
int smm_device_info(int* sm_count, size_t* l2_bytes, size_t* total_mem, size_t* free_mem);
/* total kernels launched by this library in this process (bench.py's gpu_launches) */
long long smm_kernel_launch_count(void);
int smm_sync(void);

/* ---- CSRMatrix<float>: storage, H:1243-1259; construction CSRMatrix(const TripletMatrix&) H:1314-1349 ----
 * The triplet -> CSR conversion stays on the host (include/sparse_matrix_math.h of this repo, bit-exact);
 * this uploads the three arrays and analyses row lengths for the SpMV kernels. */
int smm_csr_create(int rows, int cols, const int32_t* start, const int32_t* positions, const float* values,
                   smm_csr_t** out);
/* same, from DEVICE arrays; copy != 0 copies them, copy == 0 adopts them (freed with cudaFree on destroy).  Adopted arrays
 * change hands only when the call SUCCEEDS: on failure they are untouched and still the caller's. */
int smm_csr_create_dev(int rows, int cols, int32_t* start_dev, int32_t* positions_dev, float* values_dev,
                       int copy, smm_csr_t** out);
/* values changed on the host (operator*=, inplaceAdd, updateEntry, setValue ... H:1525-1604, H:846-849) */
int smm_csr_update_values(smm_csr_t* m, const float* values);
int smm_csr_destroy(smm_csr_t* m);
int smm_csr_shape(const smm_csr_t* m, int* rows, int* cols, int64_t* nnz, int* first_active_start);
/* device -> host copy of the arrays (for parity checks of device-generated matrices) */
int smm_csr_download(const smm_csr_t* m, int32_t* start, int32_t* positions, float* values);
/* raw device pointers of the resident arrays */
int smm_csr_device_arrays(const smm_csr_t* m, const int32_t** start_dev, const int32_t** positions_dev,
                          const float** values_dev);

/* ---- CSRMatrix::rMult / rMultAdd / rMultSub, H:1458-1515 ----
 * out[row] = op(lhs[row], sum_k values[k] * mult[positions[k]]); out may alias lhs; lhs ignored for ASSIGN.
 * exact != 0 forces left-to-right accumulation in every row (bit-identical to the reference). */
int smm_spmv(const smm_csr_t* m, int op, const float* lhs, const float* mult, float* out);
int smm_spmv_dev(const smm_csr_t* m, int op, const float* lhs_dev, const float* mult_dev, float* out_dev,
                 int exact, void* stream);

/* ---- Vector::operator* and secondNorm[Squared], H:287-328 ---- */
int smm_dot(int64_t n, const float* a, const float* b, int reduction_mode, float* out);
int smm_dot_dev(int64_t n, const float* a_dev, const float* b_dev, int reduction_mode, float* out_host, void* stream);

/* ---- CSRMatrix::getPreconditioner<SYMMETRIC_GAUS_SEIDEL>() H:1643-1651; SGSPreconditioner::apply H:1658-1713 ----
 * create builds the schedule of the two triangular sweeps once -- on the device from the resident CSR arrays for grid
 * stencils (tile schedule, ~0.03-0.16 s at 16.8 M rows), on the host for everything else (row levels); the reference has no
 * set-up.  apply returns the reference's int code through *rc (0 ok, 1 = empty row / missing or tiny diagonal / leading
 * empty rows). */
int smm_precond_sgs_create(const smm_csr_t* m, smm_precond_t** out);
int smm_precond_apply(const smm_precond_t* p, const float* rhs, float* x, int* rc);
int smm_precond_apply_dev(const smm_precond_t* p, const float* rhs_dev, float* x_dev, int* rc, void* stream);
int smm_precond_levels(const smm_precond_t* p, int* forward_levels, int* backward_levels);
/* Levels of the TILE graph when the sweeps run tile by tile (matrices whose column offsets are those of a natural-order
 * 2D / 3D grid stencil: rows are grouped into tiles of <= 64 that one warp solves in shared memory), 0 / 0 when the
 * row-level schedule is in use.  Diagnostic; additive. */
int smm_precond_tile_levels(const smm_precond_t* p, int* forward_levels, int* backward_levels);
/* Which schedule the sweeps of this handle run: 0 = row by row in level order (any matrix), 1 = tiles, 2 = lines (grid
 * stencils: a lane per grid line, a warp per patch of 32 lines, sgs_lines.cu; smm_precond_tile_levels then reports the
 * number of patch offsets).  The result of apply is bit-identical whatever the schedule.  Diagnostic; additive. */
int smm_precond_schedule(const smm_precond_t* p);
/* Fingerprints (FNV-1a, 64 bit) of the layout arrays the sweep kernels of this handle read, for checking that two ways of
 * building a handle (the set-up kernels of sgs_tiles_setup.cu and the host code of sgs_tiles.cu; SMM_B200_SGS_SETUP=host)
 * produce the same arrays bit for bit: out[0..11] = diagonal positions, forward / backward row order, backward-to-forward
 * positions, then per sweep operand positions, operand value indices, steps, push lists; out[12] = sizes and level counts.
 * Downloads the arrays: a test / diagnostic call.  Diagnostic; additive. */
int smm_precond_layout_fingerprint(const smm_precond_t* p, uint64_t out[13]);
/* CSRMatrix::IC0Preconditioner (H:1214-1235): construction + init() (factorize, H:1839-1928; *rc = its return code) and
 * apply (H:1802-1837, through smm_precond_apply[_dev]).  The factorisation is set-up code: row by row instead of the
 * reference's O(rows^2) scan, with bit-identical values -- on the device in the order of the forward tile schedule when the
 * matrix has one, else on the host; the two triangular solves of every apply run on the GPU with the same level-scheduled
 * sweeps as SGS. */
int smm_precond_ic0_create(const smm_csr_t* m, int* rc, smm_precond_t** out);
int smm_precond_ic0_factor(const smm_precond_t* p, float* factor_host);
/* EXTENSION -- CSRMatrix::ILU0Preconditioner (H:1188-1212, 1715-1790).  The reference declares it but its factorize()
 * cannot succeed and apply() is never defined (dead code), so there is nothing to be bit-identical to: this is the
 * zero-fill LU that code describes (row-wise IKJ in A's pattern, unit lower factor, multipliers formed with the
 * reciprocal pivot), factorised at create time (on the device for grid stencils, else on the host); apply = L y = rhs, U x = y on the GPU with the level-scheduled
 * sweeps.  *rc: 0 ok, 1 structurally unusable, 2 pivot not > 1e-6 in magnitude.  smm_precond_ic0_factor returns the
 * factor of either kind (strict L and U in A's pattern).  smm_solve_bicgstab accepts it as `precond`. */
int smm_precond_ilu0_create(const smm_csr_t* m, int* rc, smm_precond_t** out);
/* EXTENSION -- diagonal (Jacobi) preconditioner, not in the reference: apply is the element-wise x_i = rhs_i / a_ii on the
 * matrix's current values (*rc = 1 when |a_ii| < 1e-5 for some row, the guard of the SGS sweeps).  smm_solve_bicgstab
 * accepts it as `precond`. */
int smm_precond_jacobi_create(const smm_csr_t* m, smm_precond_t** out);
/* 0 Symmetric Gauss-Seidel, 1 IC(0), 2 ILU(0), 3 Jacobi */
int smm_precond_kind(const smm_precond_t* p);
int smm_precond_destroy(smm_precond_t* p);

/* ---- solvers ----
 * ConjugateGradient H:2316-2398; BiCGSymmetric H:2021-2102; ConjugateGradientSquared H:2109-2178;
 * BiCGStab H:2191-2303 (precond == NULL: the 5-argument overload / IDPreconditioner).
 * Argument meaning, clamping of maxIterations, status codes and stopping tests follow the reference
 * function exactly; opts may be NULL (defaults); info may be NULL. */
int smm_solve_cg(const smm_csr_t* a, const float* b, const float* x0, float* x, int maxIterations, float eps,
                 const smm_solve_options* opts, smm_solve_info* info);
/* ConjugateGradient(a, b, x0, x, maxIterations, eps, IC0Preconditioner) H:2414-2505 */
int smm_solve_cg_ic0(const smm_csr_t* a, const smm_precond_t* ic0, const float* b, const float* x0, float* x, int maxIterations,
                     float eps, const smm_solve_options* opts, smm_solve_info* info);
int smm_solve_cg_ic0_dev(const smm_csr_t* a, const smm_precond_t* ic0, const float* b_dev, const float* x0_dev, float* x_dev,
                         int maxIterations, float eps, const smm_solve_options* opts, smm_solve_info* info, void* stream);
int smm_solve_bicgsym(const smm_csr_t* a, const float* b, float* x, int maxIterations, float eps,
                      const smm_solve_options* opts, smm_solve_info* info);
int smm_solve_cgs(const smm_csr_t* a, const float* b, float* x, int maxIterations, float eps,
                  const smm_solve_options* opts, smm_solve_info* info);
int smm_solve_bicgstab(const smm_csr_t* a, const smm_precond_t* precond, const float* b, float* x,
                       int maxIterations, float eps, const smm_solve_options* opts, smm_solve_info* info);
/* device-resident variants: b/x0/x already in HBM (x0 may equal x) */
int smm_solve_cg_dev(const smm_csr_t* a, const float* b_dev, const float* x0_dev, float* x_dev, int maxIterations,
                     float eps, const smm_solve_options* opts, smm_solve_info* info, void* stream);
int smm_solve_bicgsym_dev(const smm_csr_t* a, const float* b_dev, float* x_dev, int maxIterations, float eps,
                          const smm_solve_options* opts, smm_solve_info* info, void* stream);
int smm_solve_cgs_dev(const smm_csr_t* a, const float* b_dev, float* x_dev, int maxIterations, float eps,
                      const smm_solve_options* opts, smm_solve_info* info, void* stream);
int smm_solve_bicgstab_dev(const smm_csr_t* a, const smm_precond_t* precond, const float* b_dev, float* x_dev,
                           int maxIterations, float eps, const smm_solve_options* opts, smm_solve_info* info,
                           void* stream);

/* ---- benchmark inputs generated in HBM (additive; the reference can only ingest through std::map) ----
 * bit-identical to tests/matgen.py: stencil7 = convdiff3d(nx,ny,nz,c) (c = 0: Poisson; nz = 1 and the y/z
 * couplings of a 5-point 2D Poisson are selected with kind) */
enum { SMM_GEN_POISSON2D = 0, SMM_GEN_CONVDIFF3D = 1, SMM_GEN_POWERLAW = 2 };
int smm_gen_csr(int kind, int nx, int ny, int nz, float c, uint64_t seed, smm_csr_t** out);
/* x*_i = (splitmix64(seed, i + offset) >> 40) / 2^24 written to a device vector */
int smm_gen_xstar_dev(int64_t n, int64_t offset, uint64_t seed, float* x_dev, void* stream);

/* ---- multi-GPU: one process per GPU, contiguous row blocks (additive; the reference is single-process) ----
 * Each rank builds the CSR of its rows [row_begin,row_end) with GLOBAL column indices (smm_csr_create /
 * smm_gen_csr_rows), then
 *   smm_dist_create   finds the window of the global vector those rows read and re-indexes the columns into it,
 *                     allocates the extended vector + reduction mailboxes + halo flags in one IPC-exported block;
 *   smm_dist_info     returns {row_begin,row_end,lo,hi} and the 64-byte CUDA IPC handle of that block, which the host
 *                     program all-gathers over its own transport (torch.distributed, MPI ...);
 *   smm_dist_connect  maps every peer's block and derives the halo send plan.
 * smm_dist_solve_cg then runs ConjugateGradient (H:2316-2398) on the partitioned system: b, x0, x are this rank's
 * DEVICE slices; halo exchange by P2P stores + flags, scalar reductions by a P2P all-reduce fused into the kernels'
 * epilogues (summed in rank order: identical bits, identical branches on every rank).  No NCCL inside the loop.
 * opts->reduction_mode == SMM_REDUCE_REFERENCE_TREE is accepted by smm_dist_solve_cg when the number of ranks is a power
 * of two and every rank's row block is a node of the reference's reduction tree over [0, global_rows) (the range halved at
 * lo + (hi - lo) / 2, H:308-320): the ranks' subtree sums are then joined pairwise and the solve is bit-identical to the
 * reference's SMM_MULTITHREADING build; any other partition is refused with SMM_E_INVALID.  The same holds for the other
 * three solvers below (BiCGStab's serial ||r||^2, H:2262-2267, is chained through the ranks). */
/* smm_dist_create re-indexes local's column indices IN PLACE (global -> window) and sets its column count to the window
 * length; `local` must outlive the smm_dist_t and must not be used for single-GPU calls afterwards.  On failure nothing the
 * function allocated is kept. */
typedef struct smm_dist smm_dist_t;
int smm_dist_create(int rank, int nranks, int64_t global_rows, int64_t row_begin, int64_t row_end, smm_csr_t* local,
                    smm_dist_t** out);
int smm_dist_info(const smm_dist_t* d, int64_t* ranges4, void* ipc_handle64);
int smm_dist_connect(smm_dist_t* d, const int64_t* all_ranges /* [nranks][4] */, const void* all_handles /* [nranks][64] */);
/* Single-process variant of smm_dist_connect: all[r] is rank r's handle, each created in THIS process on its own device
 * (smm_set_device before smm_csr_create / smm_dist_create).  The devices are made peer-accessible and address each other's
 * blocks directly: no IPC handles, no second process. */
int smm_dist_connect_local(smm_dist_t** all, int nranks);
/* y_local = A_local * x with the halo of x exchanged.  Calls need no barrier between them: every exchange is acknowledged
 * by its consumers, and a rank pushes only after the previous halo it sent has been read. */
int smm_dist_spmv_dev(smm_dist_t* d, const float* x_local_dev, float* y_local_dev, void* stream);
int smm_dist_solve_cg(smm_dist_t* d, const float* b_dev, const float* x0_dev, float* x_dev, int maxIterations, float eps,
                      const smm_solve_options* opts, smm_solve_info* info, void* stream);
/* the other unpreconditioned solvers on the partitioned system (BiCGSymmetric H:2021, ConjugateGradientSquared H:2109,
 * BiCGStab H:2294): every SpMV operand is staged into the extended vector and its halo exchanged first */
int smm_dist_solve_bicgsym(smm_dist_t* d, const float* b_dev, float* x_dev, int maxIterations, float eps,
                           const smm_solve_options* opts, smm_solve_info* info, void* stream);
int smm_dist_solve_cgs(smm_dist_t* d, const float* b_dev, float* x_dev, int maxIterations, float eps,
                       const smm_solve_options* opts, smm_solve_info* info, void* stream);
int smm_dist_solve_bicgstab(smm_dist_t* d, const float* b_dev, float* x_dev, int maxIterations, float eps,
                            const smm_solve_options* opts, smm_solve_info* info, void* stream);
int smm_dist_error(const smm_dist_t* d, int* error);
int smm_dist_destroy(smm_dist_t* d);
/* rows [row_begin,row_end) of a generated stencil matrix, global column indices (kinds POISSON2D / CONVDIFF3D) */
int smm_gen_csr_rows(int kind, int nx, int ny, int nz, float c, int64_t row_begin, int64_t row_end, smm_csr_t** out);

/* ---- multi-GPU in ONE process (additive): what the C++ drop-in header binds when SMM::b200::devices() > 1 ----
 * smm_group_create cuts a host CSR (square) into `ndevices` contiguous row blocks -- partition 0: equal numbers of stored
 * entries; partition 1: the nodes of the reference's reduction tree (H:308-320; power-of-two device counts), which the
 * reference-order reduction mode needs -- uploads block r to devices[r] (NULL: devices 0 .. ndevices-1), enables peer access
 * and connects the blocks (smm_dist_connect_local).  Every later call runs one host thread per device; all pointers are
 * HOST pointers of the global system, like the reference's functions.  No torch, no launcher, no IPC. */
typedef struct smm_group smm_group_t;
int smm_group_create(int rows, int cols, const int32_t* start, const int32_t* positions, const float* values, int ndevices,
                     const int* devices, int partition, smm_group_t** out);
int smm_group_info(const smm_group_t* g, int* ndevices, int64_t* row_cuts /* [ndevices + 1] */);
int smm_group_update_values(smm_group_t* g, const float* values);
/* CSRMatrix::rMult / rMultAdd / rMultSub, H:1501-1515 */
int smm_group_spmv(smm_group_t* g, int op, const float* lhs, const float* mult, float* out);
/* solver: 0 ConjugateGradient H:2316 (b, x0, x), 1 BiCGSymmetric H:2021, 2 ConjugateGradientSquared H:2109, 3 BiCGStab without
 * preconditioner H:2294 (b, x; x0 ignored).  Status, clamping and stopping tests as in smm_solve_*. */
int smm_group_solve(smm_group_t* g, int solver, const float* b, const float* x0, float* x, int maxIterations, float eps,
                    const smm_solve_options* opts, smm_solve_info* info);
int smm_group_destroy(smm_group_t* g);

/* ---- measurement hook (bench.py): average device time, in ms, of each of the three kernels of one fused CG
 * iteration (SpMV + p.Ap | r update + r.r | x and p update, p read once), `reps` launches each, CUDA events on `stream`;
 * ms_xr receives the r update's time, ms_p the x and p update's ---- */
int smm_profile_cg_iteration(const smm_csr_t* a, int reps, float* ms_spmv, float* ms_xr, float* ms_p, void* stream);
/* The same for this rank's kernels of the multi-GPU iteration (halo stores fused into the x,p update, halo wait), timed without
 * the peers taking part: needs SMM_B200_DIST_DEBUG=7 in the environment when the handle is created (tools/dist_kernel_times.py). */
int smm_dist_profile_cg_iteration(smm_dist_t* d, int reps, float* ms_spmv, float* ms_r, float* ms_px, void* stream);

/* ---- device memory helpers for hosts without their own allocator ---- */
int smm_malloc_dev(size_t bytes, void** ptr_dev);
int smm_free_dev(void* ptr_dev);
int smm_memcpy_h2d(void* dst_dev, const void* src, size_t bytes);
int smm_memcpy_d2h(void* dst, const void* src_dev, size_t bytes);
int smm_memset_dev(void* dst_dev, int byte, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* SMM_B200_H */
