// smm_internal.cuh -- shared declarations of libsmm_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <utility>

#include "smm_b200.h"

// ---------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------
void smm_set_error(const char* fmt, ...);
int smm_cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define SMM_CUDA(call)                                                            \
    do {                                                                          \
        cudaError_t _e = (call);                                                  \
        if (_e != cudaSuccess) return smm_cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define SMM_TRY(call)            \
    do {                         \
        int _rc = (call);        \
        if (_rc != SMM_OK) return _rc; \
    } while (0)

// kernels launched by this library: process-wide (smm_kernel_launch_count) and by the calling thread (a solve's own count;
// single-process multi-GPU solves run one host thread per device).  Launches recorded into a CUDA graph are not counted
// while they are captured but once per graph launch.
extern std::atomic<long long> g_smm_launches;
extern thread_local long long t_smm_launches;
extern thread_local bool t_smm_capturing;
#define SMM_COUNT_LAUNCH(n)                                                   \
    do {                                                                      \
        if (!t_smm_capturing) { g_smm_launches.fetch_add((n), std::memory_order_relaxed); t_smm_launches += (n); } \
    } while (0)

// per-device caches of launch attributes (a process may drive several GPUs, from several threads)
constexpr int SMM_MAX_DEVICES = 64;
extern std::mutex g_smm_attr_mu;

cudaStream_t smm_default_stream();

// ---------------------------------------------------------------------------------------------------
// tunables of the SpMV kernels
// ---------------------------------------------------------------------------------------------------
constexpr int SPMV_THREADS = 256;          // threads per CTA
constexpr int SPMV_CHUNK = 4096;           // nnz per CTA: CTA q owns the rows whose first entry lies in [q*CHUNK,(q+1)*CHUNK)
constexpr int SPMV_CAP = 6144;             // products staged in shared memory (24 KB): CHUNK + longest row streamed
constexpr int SPMV_MIN_STREAM_ROWS = 48;   // fewer rows than this in a CTA's range -> warp/CTA-per-row path
constexpr int SPMV_LONG_IN_STREAM = 192;   // rows longer than this inside a streamed range are summed by a warp
constexpr int SPMV_PERM_MAX_ROWS = 2048;   // chunks of at most this many rows have their rows ordered by length (row_perm)
constexpr int SPMV_CTA_ROW = 4096;         // rows longer than this are reduced by the whole CTA
constexpr int SPMV_ROWS_CAP = 2560;        // rows kernel: upper bound of the entries per TMA stage

// ---------------------------------------------------------------------------------------------------
// handles
// ---------------------------------------------------------------------------------------------------
struct smm_csr {
    int device = 0;
    int rows = 0, cols = 0;
    int64_t nnz = 0;
    int64_t nnz_alloc = 0;          // entries the positions/values allocations really hold (>= nnz; padded to 4 when owned)
    int first_active_start = 0;
    int32_t* start = nullptr;       // [rows+1]
    int32_t* positions = nullptr;   // [nnz]
    float* values = nullptr;        // [nnz]
    bool owns_arrays = true;
    unsigned long long values_version = 1;   // bumped by smm_csr_update_values (packed copies refresh lazily)
    // SpMV analysis: CTA q handles rows [block_row[q], block_row[q+1])
    int num_blocks = 0;
    int32_t* block_row = nullptr;   // [num_blocks+1]
    int32_t* row_perm = nullptr;    // [rows] or null: the rows of every chunk, longest first (irregular-row kernels; spmv.cu)
    uint16_t* start16 = nullptr;    // rows kernel: [groups][R + 1] row starts relative to the group's staging window (2 B per row instead of 4)
    int max_row_len = 0;
    int rows_kernel_lanes = 0;      // 0: product-staging kernel; V > 0: TMA rows kernel with V lanes per row
    int sm_count = 148;
    int rows_kernel_cap = 0;        // entries per TMA stage of the rows kernel (sized to the matrix)
    // reduction scratch shared by every kernel launched for this matrix (one solve at a time per handle)
    struct smm_workspace* ws = nullptr;
};

// Device-side scalar block of a solve.  Written only by "last block" epilogues (epilogue.cuh).
struct SolveState {
    int done;            // 1 once the stopping test fired (every later kernel is a no-op)
    int status;          // SMM_SOLVER_*
    int iterations;      // loop trips completed
    int max_iterations;
    float eps, eps2;
    float residual;      // last value compared with eps
    int precond_error;
    // Krylov scalars
    float rr;            // CG/BiCGSym: r.r ; CGS/BiCGStab: r.r0
    float denom;         // p.Ap / ap.r0
    float alpha, beta, omega;
    float as_s, as_as;   // BiCGStab
    float res2;          // CGS: r.r ; BiCGStab: ||r||^2
    float scratch[2];    // plain dot products (FIN_STORE)
    int history_cap;
    int pad0;
    float* history;      // device residual history (may be null)
    unsigned long long cond_handle;   // cudaGraphConditionalHandle of the WHILE driver (0 = none)
    struct DistComm* comm;            // multi-GPU: reductions are summed over ranks before the scalar step (null: one GPU)
    int x_owed;          // ConjugateGradient / BiCGSymmetric: the stopping test fired in the r update, the x update of that iteration is still to come (VEC_CG_PX / VEC_BICGSYM_PX)
    int pad[5];
};

// scratch of the reference-order dot products (dots.cu): per handle, so that two handles solving on different streams of
// one device never share the node buffer or the last-block ticket
struct smm_dot_scratch {
    float* nodes = nullptr;
    size_t nodes_cap = 0;
    unsigned int* ticket = nullptr;
    float* out = nullptr;        // 2 floats
    float* sq = nullptr;         // 2 floats: totals of sum_squares_serial_kernel on their way to dot_serial_kernel
};

struct smm_workspace {
    int device = 0;
    smm_dot_scratch dot;
    int sm_count = 148;
    // partial sums of fused reductions: [RED_SLOTS][2][partials_cap] floats
    float* partials = nullptr;
    size_t partials_cap = 0;
    unsigned int* tickets = nullptr;   // [RED_SLOTS] last-block counters (self-resetting)
    SolveState* state = nullptr;       // device
    SolveState* state_host = nullptr;  // pinned
    float* history = nullptr;          // device residual history
    int history_cap = 0;
    // work vectors of the solvers, grown on demand
    float* vec[10] = {nullptr};
    size_t vec_len = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    unsigned int* grid_barrier = nullptr;   // [2] arrival count and generation of the persistent iteration's grid barrier
};

constexpr int RED_SLOTS = 4;

int smm_workspace_get(const smm_csr* m, smm_workspace** out);
int smm_workspace_vectors(smm_workspace* ws, int count, size_t len);
void smm_workspace_free(smm_workspace* ws);

// ---------------------------------------------------------------------------------------------------
// kernels' host launchers (spmv.cu, vecops.cu)
// ---------------------------------------------------------------------------------------------------
struct SpmvArgs {
    const smm_csr* m = nullptr;
    int op = SMM_OP_ASSIGN;
    const float* lhs = nullptr;
    const float* mult = nullptr;
    float* out = nullptr;
    int exact = 0;           // sequential accumulation in every row (bit-identical to the reference)
    int reduce = 0;          // ReduceShape (epilogue.cuh): which dot products ride on the SpMV
    int finish = 0;          // FinishKind: what the last CTA does with the totals
    int slot = 0;            // reduction slot (partials + ticket)
    const float* aux = nullptr;   // second operand of the fused dot
    SolveState* state = nullptr;  // null only when reduce == RED_NONE and no early-out is wanted
    float* copy1 = nullptr;  // optional extra copies of out[row] (p = r, r0 = r, u = r ...)
    float* copy2 = nullptr;
    float* copy3 = nullptr;
    const void* halo_wait = nullptr;   // HaloWaitDev* (device): fused halo wait of the multi-GPU SpMV (rows kernel only)
};
int smm_launch_spmv(const SpmvArgs& a, cudaStream_t s);
int smm_spmv_rows_lanes(const smm_csr* m, int exact);

// fused element-wise kernels (vecops.cu); operand order per kind is documented at each functor
enum VecKind {
    VEC_CG_XR = 0,     // in: x p r Ap        out: x r     t0 = r.r
    VEC_CG_P,          // in: p r             out: p
    VEC_BICGSYM_XR,    // in: x p r ap        out: x r     t0 = r.r
    VEC_BICGSYM_P,     // in: p r             out: p
    VEC_CGS_QX,        // in: ap u x          out: q auq x
    VEC_CGS_UP,        // in: q r p           out: u p
    VEC_STAB_S,        // in: ap r            out: s
    VEC_STAB_XR,       // in: x p s as r0     out: x r     t0 = r.r, t1 = r.r0
    VEC_STAB_P,        // in: p ap r          out: p
    VEC_DOT2,          // in: a b                          t0 = a.b, t1 = a.a
    VEC_COPY3,         // in: a               out: o0 o1 o2
    VEC_CG_R,          // in: r Ap            out: r       t0 = r.r
    VEC_CG_PX,         // in: p r x           out: p x     (x += alpha p with the OLD p, then p = beta p + r)
    VEC_BICGSYM_R,     // in: r ap            out: r       t0 = r.r
    VEC_BICGSYM_PX,    // in: p r x           out: p x
};
struct VecArgs {
    const void* halo_push = nullptr;   // HaloPushDev* (device): out[0]'s boundary entries also go to the peers (VEC_CG_P / VEC_CG_PX / VEC_COPY3)
    long long n = 0;
    const float* in[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float* out[3] = {nullptr, nullptr, nullptr};
    SolveState* state = nullptr;
    int finish = 0;          // FinishKind
    int slot = 0;
    smm_workspace* ws = nullptr;
};
int smm_launch_vec(int kind, const VecArgs& a, cudaStream_t s);
int smm_vec_max_grid(const smm_workspace* ws);
int smm_vec_grid(const smm_workspace* ws, long long n, bool aligned);
// persistent CG iteration (spmv.cu): the whole loop of ConjugateGradient (H:2352-2396) in ONE cooperative kernel
bool smm_cg_persistent_fits(const smm_csr* m, const float* x, const float* r, const float* p, const float* ap);
int smm_launch_cg_persistent(const smm_csr* m, SolveState* state, float* x, float* r, float* p, float* ap, cudaStream_t s);

// dot products in the reference's summation orders (dots.cu)
int smm_tree_depth(long long n);
// ws: the handle's workspace (its own scratch); null: the stand-alone smm_dot's per-device scratch
int smm_dot_ref_prepare(long long n, smm_workspace* ws);
int smm_launch_dot_ref(int mode, long long n, int ndots, const float* a0, const float* b0, const float* a1, const float* b1,
                       SolveState* state, int finish, float* out_dev, cudaStream_t s, smm_workspace* ws, float* update_r = nullptr);
// update_r (== a0, with b0 = Ap): r = r - state->alpha * Ap rides on the dot, which is then r.r of the new r (ConjugateGradient in
// the reference-tree mode on long vectors; smm_dot_ref_update_applies says whether this form exists for the operands)
bool smm_dot_ref_update_applies(int mode, long long n, const float* r, const float* ap);
void smm_dot_scratch_free(smm_dot_scratch* sc);

// SGS preconditioner (sgs.cu)
int smm_sgs_apply_async(const smm_precond* p, const float* rhs_dev, float* x_dev, SolveState* state, cudaStream_t s);
int smm_sgs_apply_async_rc(const smm_precond* p, const float* rhs_dev, float* x_dev, SolveState* state, int* rc_dev, cudaStream_t s);
int smm_sgs_kernels_per_apply(const smm_precond* p);
extern "C" int smm_precond_kind(const smm_precond* p);   // 0 SGS, 1 IC(0)
int smm_csr_analyse(smm_csr* m, cudaStream_t s);
int smm_first_active_start(const smm_csr* m, int* out_host, cudaStream_t s);

// ---------------------------------------------------------------------------------------------------
// Programmatic dependent launch.  The kernels of an iteration form a chain (each reads what the previous one wrote, down
// to the scalars of the recurrence); launched with cudaLaunchAttributeProgrammaticStreamSerialization a kernel may start
// while its predecessor drains -- its CTAs become resident and run their prologue (shared-memory ring, barriers, index
// arithmetic) -- and blocks in smm_pdl_wait() until the predecessor has completed and its writes are visible.  A kernel
// calls smm_pdl_trigger() once its main loop is done, which is what lets the successor start early.  Same kernels, same
// results; only the launch / ramp-up latency between dependent kernels is taken off the critical path, which is what an
// iteration on an L2-resident problem mostly consists of.  (SMM_B200_PDL=0 launches the plain way.)
// ---------------------------------------------------------------------------------------------------
bool smm_pdl_enabled();
#ifdef __CUDACC__
template <class... KArgs, class... Args>
cudaError_t smm_launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at;
    cfg.numAttrs = smm_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
// every kernel that may be launched through smm_launch_chain calls this before it reads anything a predecessor wrote
__device__ __forceinline__ void smm_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void smm_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// ---------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
// `_smm_fma` of the reference (H:27-37) is a*x+b with TWO roundings; keep nvcc from contracting it.
__device__ __forceinline__ float smm_fma2(float a, float x, float b) { return __fadd_rn(__fmul_rn(a, x), b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming loads: read once, do not keep in L1
__device__ __forceinline__ int4 ldg_stream_i4(const int4* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ int ldg_stream_i(const int* p) {
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream_f(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// Deterministic block sum of up to 3 values; result valid in thread 0.  `sh` holds 3*32 floats.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) sh[i * 32 + warp] = v[i];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float t = lane < nw ? sh[i * 32 + lane] : 0.0f;
            v[i] = warp_sum(t);
        }
    }
}

// Last-block pattern: every CTA publishes NV partials, the CTA that takes the last ticket sums all of them in a
// fixed order (thread t adds partials t, t+T, ... then a block tree) and returns true in ALL its threads with the
// totals in v (valid in thread 0).  The ticket resets itself so the slot can be reused by the next launch.
template <int NV>
__device__ __forceinline__ bool grid_sum_last_block(float (&v)[NV], float* partials, size_t partials_stride,
                                                    unsigned int* ticket, float* sh, int* sh_flag) {
    block_sum<NV>(v, sh);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) partials[i * partials_stride + blockIdx.x] = v[i];
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        *sh_flag = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!*sh_flag) return false;
    __threadfence();
    float acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.0f;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] += __ldcg(&partials[i * partials_stride + b]);
    }
    __syncthreads();   // sh is reused
    block_sum<NV>(acc, sh);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = acc[i];
    if (threadIdx.x == 0) *ticket = 0u;
    return true;
}
#endif  // __CUDACC__
