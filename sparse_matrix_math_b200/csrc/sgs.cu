// sgs.cu -- Symmetric Gauss-Seidel preconditioner apply, the one preconditioner CSRMatrix::getPreconditioner()
// hands to BiCGStab in the reference (H:1643-1651, class H:1172-1186, apply H:1658-1713):
//     (D + L) y = rhs            forward substitution, rows ascending,  cols ascending inside a row
//     x_i = y_i - (sum_{j>i} a_ij x_j) / a_ii   backward, rows descending, cols DESCENDING inside a row
// on A's own CSR arrays (no factor storage).  The reference runs both sweeps serially; here they are
// level-scheduled and sync-free:
//   * create: one host pass over the structure computes the dependency level of every row in the lower and in the
//     upper triangle (level = 1 + max level of the rows it reads) and the two row orders sorted by level, each level
//     padded to a warp so that no lane ever waits on a lane of its own warp;
//     for each sweep the off-diagonal entries are also re-packed in that order as 32-row slices (sliced ELL: entry k
//     of lane l of slice s sits at slice_ptr[s] + 32 k + l), so a warp reads its rows' entries with coalesced loads
//     that depend on nothing but the thread index -- the only dependent accesses left are rhs[row] and x[col];
//   * apply: ONE launch per sweep, one thread per row in level order.  A thread accumulates its row in the
//     reference's order and, for every x[col] it needs, spins until the value has been published.  The output
//     vector doubles as the ready flag: it is pre-filled with a NaN payload no arithmetic instruction can produce.
//     CTAs take their logical index from an atomic ticket, so every producer a thread waits for belongs to a CTA
//     that has already started -- no deadlock whatever order the hardware dispatches CTAs in.  Levels overlap
//     freely (no grid barrier): the sweep is bounded by the dependency chain (levels x L2 round trip), not by
//     launch or barrier latency.  A poll bound turns a would-be hang into SMM_E_TIMEOUT.
// Per-row arithmetic is identical to the reference (two roundings per multiply-add, one division per sweep), so the
// result is bit-identical to SGSPreconditioner::apply for any schedule.
// Algorithmic bytes per apply: each stored entry once over the two sweeps (8 nnz) + start/diag index/order
// (about 24 n) + rhs, y, x traffic (about 20 n).
#include <stdlib.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "smm_internal.cuh"

struct smm_precond {
    int kind = 0;                    // 0: Symmetric Gauss-Seidel on A's values; 1: IC(0) on its own factor values
    float* factor = nullptr;         // IC(0): [nnz] factor in A's pattern (L below and on the diagonal, L^T above), ref H:1233-1234
    const smm_csr* m = nullptr;
    int rows = 0;
    bool valid = true;               // structure admits the sweeps (else apply returns the reference's code 1)
    int levels_fwd = 0, levels_bwd = 0;
    long long threads_fwd = 0, threads_bwd = 0;   // padded launch sizes
    int32_t* order_fwd = nullptr;    // [threads_fwd] row index or -1 (padding)
    int32_t* order_bwd = nullptr;    // [threads_bwd]
    int32_t* diag_pos = nullptr;     // [rows] index of a_ii in positions/values
    // sliced-ELL copies of the strict lower / upper triangles in sweep order (index 0: forward, 1: backward)
    long long* slice_ptr[2] = {nullptr, nullptr};   // [threads/32 + 1]
    int32_t* ecol[2] = {nullptr, nullptr};          // column or -1 (padding)
    int32_t* eidx[2] = {nullptr, nullptr};          // index into the CSR values (to refresh eval after value updates)
    float* eval[2] = {nullptr, nullptr};
    float* dval[2] = {nullptr, nullptr};            // [threads] a_ii of the thread's row
    long long esize[2] = {0, 0};
    unsigned long long values_version = ~0ull;      // version of m->values the packed copies were gathered from
    float* y = nullptr;              // [rows] forward result
    unsigned int* tickets = nullptr; // [2] logical CTA counters, [2] = abort flag, [3] = error bits
    float* io[2] = {nullptr, nullptr};   // staging for the host-pointer apply
};

namespace {

constexpr unsigned int SENTINEL = 0x7FC0DEADu;   // quiet NaN with a payload; GPU arithmetic only produces 0x7FFFFFFF
constexpr int SGS_THREADS = 128;
constexpr unsigned int POLL_LIMIT = 1u << 22;

__global__ void sgs_fill_kernel(float* __restrict__ y, float* __restrict__ x, long long n, unsigned int* tickets, const SolveState* st) {
    if (st != nullptr && st->done) return;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        y[i] = __uint_as_float(SENTINEL);
        x[i] = __uint_as_float(SENTINEL);
    }
    if (i == 0) { tickets[0] = 0u; tickets[1] = 0u; tickets[2] = 0u; tickets[3] = 0u; }
}

__device__ __forceinline__ float wait_value(const float* p, unsigned int* abort_flag) {
    unsigned int bits;
    unsigned int polls = 0;
    for (;;) {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(bits) : "l"(p));
        if (bits != SENTINEL) break;
        __nanosleep(polls < 16u ? 32u : 128u);                // back off: thousands of lanes may be waiting on L2
        if ((++polls & 255u) == 0u) {
            unsigned int a;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(a) : "l"(abort_flag));
            if (a != 0u || polls >= POLL_LIMIT) { atomicExch(abort_flag, 1u); break; }
        }
    }
    return __uint_as_float(bits);
}

__device__ __forceinline__ void publish(float* p, float v) {
    // a computed value can never equal the sentinel payload, so the store itself is the ready flag
    asm volatile("st.relaxed.gpu.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

struct SweepArgs {
    const int32_t* order;        // [nthreads] row or -1
    const long long* slice_ptr;  // [nthreads/32 + 1]
    const int32_t* ecol;
    const float* eval;
    const float* dval;           // [nthreads]
    long long nthreads;
};

__global__ void sgs_gather_values_kernel(const float* __restrict__ values, const int32_t* __restrict__ eidx, float* __restrict__ eval, long long n,
                                         const int32_t* __restrict__ order, const int32_t* __restrict__ diag_pos, float* __restrict__ dval, long long nthreads) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const int k = eidx[i]; eval[i] = k >= 0 ? values[k] : 0.0f; }
    if (i < nthreads) { const int r = order[i]; dval[i] = r >= 0 ? values[diag_pos[r]] : 1.0f; }
}

// IC0 = false: SGS sweeps (H:1658-1713).  IC0 = true: L y = rhs then L^T x = y (IC0Preconditioner::apply, H:1802-1837):
// `sum -= ic0[j] * x[col]`, one division by the diagonal of the factor per row and sweep.
template <bool FORWARD, bool IC0>
__global__ void __launch_bounds__(SGS_THREADS) sgs_sweep_kernel(const SweepArgs A, const float* __restrict__ rhs, float* y, float* x,
                                                               unsigned int* tickets, const SolveState* st) {
    if (st != nullptr && st->done) return;
    __shared__ unsigned int sh_bid;
    unsigned int* abort_flag = tickets + 2;
    // persistent CTAs: logical blocks are handed out in order by an atomic ticket, so (a) every row a thread waits for
    // belongs to a block that has already been claimed by a running CTA, and (b) the number of lanes that can be
    // spinning at any time is bounded by the grid, which is sized to stay a few levels deep at most
    const long long nblocks = (A.nthreads + SGS_THREADS - 1) / SGS_THREADS;
    const int lane = threadIdx.x & 31;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) sh_bid = atomicAdd(&tickets[FORWARD ? 0 : 1], 1u);
        __syncthreads();
        const long long bid = sh_bid;
        if (bid >= nblocks) break;
        const long long t = bid * SGS_THREADS + threadIdx.x;      // nthreads is a multiple of 32: whole warps in or out
        if (t >= A.nthreads) continue;
        const long long slice = t >> 5;
        const long long e0 = A.slice_ptr[slice], e1 = A.slice_ptr[slice + 1];
        const int width = (int)((e1 - e0) >> 5);
        const int row = A.order[t];
        const float d = A.dval[t];
        float acc;
        const float* src = FORWARD ? y : x;
        if (FORWARD) {
            if (!IC0 && row >= 0 && fabsf(d) < 1e-5) atomicOr(tickets + 3, 1u);  // H:1691-1693 (reported, not fatal here)
            acc = row >= 0 ? rhs[row] : 0.0f;                                     // H:1683 / H:1807
        } else if (IC0) {
            acc = row >= 0 ? wait_value(y + row, abort_flag) : 0.0f;              // T sum = x[row], H:1823
        } else {
            acc = 0.0f;                                                           // H:1702
        }
        for (int k = 0; k < width; k += 4) {
            int c[4];
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {                                         // coalesced, independent of any other row
                const bool in = k + j < width;
                c[j] = in ? A.ecol[e0 + (long long)(k + j) * 32 + lane] : -1;
                v[j] = in ? A.eval[e0 + (long long)(k + j) * 32 + lane] : 0.0f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (c[j] >= 0) {
                    const float xv = wait_value(src + c[j], abort_flag);
                    // forward: _smm_fma(-value, x[col], lhs) cols ascending (H:1685); backward: _smm_fma(value, x[col], lhs) cols descending (H:1704)
                    // IC0: sum -= ic0[j] * x[col] (H:1813, H:1829) -- the same bits as the forward SGS form
                    acc = (FORWARD || IC0) ? __fsub_rn(acc, __fmul_rn(v[j], xv)) : __fadd_rn(__fmul_rn(v[j], xv), acc);
                }
            }
        }
        if (row < 0) continue;
        if (FORWARD || IC0) {
            publish((FORWARD ? y : x) + row, __fdiv_rn(acc, d));                  // H:1694 / H:1818, H:1834
        } else {
            const float yr = wait_value(y + row, abort_flag);                     // own forward result (already published)
            publish(x + row, __fsub_rn(yr, __fdiv_rn(acc, d)));                   // H:1710
        }
    }
}

__global__ void sgs_status_kernel(const unsigned int* tickets, SolveState* st, int* rc_out) {
    // fold the apply's status into the solve (the reference only asserts on it, H:2235-2238) / report it to the caller
    const unsigned int aborted = tickets[2], bad_diag = tickets[3];
    int rc = 0;
    if (bad_diag) rc |= 1;
    if (aborted) rc |= 4;
    if (st != nullptr && !st->done && rc) st->precond_error |= rc;
    if (rc_out) *rc_out = rc;
}

// level analysis on the host: one pass per triangle
// pack the strict triangle of every row, in sweep order, as 32-row slices in thread order
void build_sell(bool forward, const std::vector<int32_t>& order, const std::vector<int32_t>& start, const std::vector<int32_t>& pos,
                const std::vector<int32_t>& diag, std::vector<long long>* slice_ptr, std::vector<int32_t>* ecol, std::vector<int32_t>* eidx) {
    const size_t nslices = order.size() / 32;
    slice_ptr->assign(nslices + 1, 0);
    for (size_t s = 0; s < nslices; ++s) {
        int w = 0;
        for (int l = 0; l < 32; ++l) {
            const int r = order[s * 32 + l];
            if (r < 0) continue;
            const int cnt = forward ? diag[r] - start[r] : start[r + 1] - 1 - diag[r];
            w = std::max(w, cnt);
        }
        (*slice_ptr)[s + 1] = (*slice_ptr)[s] + (long long)w * 32;
    }
    ecol->assign((size_t)(*slice_ptr)[nslices], -1);
    eidx->assign((size_t)(*slice_ptr)[nslices], -1);
    for (size_t s = 0; s < nslices; ++s) {
        const long long base = (*slice_ptr)[s];
        for (int l = 0; l < 32; ++l) {
            const int r = order[s * 32 + l];
            if (r < 0) continue;
            const int cnt = forward ? diag[r] - start[r] : start[r + 1] - 1 - diag[r];
            for (int k = 0; k < cnt; ++k) {
                const int src = forward ? start[r] + k : start[r + 1] - 1 - k;   // ascending / descending columns
                (*ecol)[(size_t)(base + (long long)k * 32 + l)] = pos[src];
                (*eidx)[(size_t)(base + (long long)k * 32 + l)] = src;
            }
        }
    }
}

void analyse(int rows, const std::vector<int32_t>& start, const std::vector<int32_t>& pos, int first_active_start, bool* valid,
             std::vector<int32_t>* diag, std::vector<int32_t>* order_f, std::vector<int32_t>* order_b, int* lf, int* lb) {
    *valid = first_active_start == 0 || rows == 0;                            // H:1668-1670
    diag->assign((size_t)rows, 0);
    std::vector<int32_t> lev((size_t)rows, 0);
    int maxl = -1;
    for (int r = 0; r < rows && *valid; ++r) {
        int k = start[r];
        const int e = start[r + 1];
        if (e == k) { *valid = false; break; }                               // H:1678-1680
        int l = 0;
        while (k < e && pos[k] < r) { l = std::max(l, lev[pos[k]] + 1); ++k; }
        if (k >= e || pos[k] != r) { *valid = false; break; }                 // H:1691 (col != row)
        (*diag)[r] = k;
        lev[r] = l;
        maxl = std::max(maxl, l);
    }
    if (!*valid) { *lf = *lb = 0; order_f->clear(); order_b->clear(); return; }
    auto build_order = [&](const std::vector<int32_t>& level, int nlev, bool descending, std::vector<int32_t>* out) {
        std::vector<long long> count((size_t)nlev + 1, 0);
        for (int r = 0; r < rows; ++r) count[(size_t)level[r] + 1]++;
        std::vector<long long> off((size_t)nlev + 1, 0);
        for (int l = 0; l < nlev; ++l) off[(size_t)l + 1] = off[l] + ((count[(size_t)l + 1] + 31) / 32) * 32;   // pad each level to a warp
        out->assign((size_t)off[nlev], -1);
        std::vector<long long> cur(off.begin(), off.end() - 1);
        if (!descending) { for (int r = 0; r < rows; ++r) (*out)[(size_t)cur[level[r]]++] = r; }
        else { for (int r = rows - 1; r >= 0; --r) (*out)[(size_t)cur[level[r]]++] = r; }
    };
    *lf = maxl + 1;
    build_order(lev, *lf, false, order_f);
    maxl = -1;
    for (int r = rows - 1; r >= 0; --r) {
        int l = 0;
        for (int k = start[r + 1] - 1; k > (*diag)[r]; --k) l = std::max(l, lev[pos[k]] + 1);   // lev[] of rows > r already hold backward levels
        lev[r] = l;
        maxl = std::max(maxl, l);
    }
    *lb = maxl + 1;
    build_order(lev, *lb, true, order_b);
}

// Zero-fill incomplete Cholesky in A's pattern, IC0Preconditioner::factorize (H:1839-1928), on the host (set-up code).
// The reference walks column by column and scans every later row for each column (O(rows^2)); this walks row by row.
// Every entry is computed from the same already-final entries with the same operations in the same order
// (sum += l_ik * l_jk over row j's columns k < i in ascending order, l_ji = (a_ji - sum) * (1 / l_ii),
// l_jj = sqrt(a_jj - sum of l_jk^2)), so the factor is bit-identical.  Returns 0, or 1 on a missing diagonal (H:1873-1876).
int ic0_factorize_host(int rows, const std::vector<int32_t>& start, const std::vector<int32_t>& pos, const std::vector<int32_t>& diag,
                       const std::vector<float>& a, std::vector<float>* out) {
    std::vector<float>& l = *out;
    l.assign(a.size(), 0.0f);
    std::vector<float> dinv((size_t)rows, 0.0f);
    for (int j = 0; j < rows; ++j) {
        const int rs = start[j], dj = diag[j];
        for (int e = rs; e < dj; ++e) {
            const int i = pos[e];                              // entry (j, i), i < j
            float sum = 0.0f;
            int ki = start[i];
            const int di = diag[i];
            for (int ek = rs; ek < e; ++ek) {                  // row j's columns k < i, ascending (H:1900-1907)
                const int k = pos[ek];
                while (ki < di && pos[ki] < k) ++ki;
                if (ki < di && pos[ki] == k) sum += l[ki] * l[ek];
            }
            l[e] = (a[e] - sum) * dinv[i];                     // H:1914
        }
        float dsum = 0.0f;
        for (int e = rs; e < dj; ++e) dsum += l[e] * l[e];    // H:1868-1872
        const float d = std::sqrt(a[dj] - dsum);               // H:1879
        l[dj] = d;
        dinv[j] = 1.0f / d;                                    // H:1883
    }
    // the transpose goes into the upper triangle of the same pattern (H:1916-1917)
    for (int j = 0; j < rows; ++j) {
        for (int e = start[j]; e < diag[j]; ++e) {
            const int i = pos[e];
            const int32_t* b = pos.data() + diag[i] + 1;
            const int32_t* en = pos.data() + start[i + 1];
            const int32_t* it = std::lower_bound(b, en, j);
            if (it != en && *it == j) l[(size_t)(it - pos.data())] = l[e];
        }
    }
    return 0;
}

}  // namespace

int smm_sgs_kernels_per_apply(const smm_precond* p) { return p && p->valid ? 4 : 1; }


// rhs_dev -> x_dev on stream s.  With `state` (inside a solve) the kernels no-op once state->done is set.
int smm_sgs_apply_async(const smm_precond* p, const float* rhs_dev, float* x_dev, SolveState* state, cudaStream_t s) {
    return smm_sgs_apply_async_rc(p, rhs_dev, x_dev, state, nullptr, s);
}

int smm_sgs_apply_async_rc(const smm_precond* p, const float* rhs_dev, float* x_dev, SolveState* state, int* rc_dev, cudaStream_t s) {
    if (!p) return SMM_E_INVALID;
    if (rhs_dev == x_dev) { smm_set_error("SGS apply: rhs must not alias x (H:1667)"); return SMM_E_ALIAS; }
    const smm_csr* m = p->m;
    if (p->rows == 0) return SMM_OK;
    if (!p->valid) {
        // SGSPreconditioner::apply returns 1 here (H:1668, 1678, 1691) and BiCGStab only asserts on it (H:2235-2238),
        // continuing on an unspecified vector; this build refuses instead of iterating on garbage
        smm_set_error("SGS preconditioner unusable for this matrix (leading empty rows, an empty row or a missing diagonal): apply() returns 1");
        return SMM_E_STATE;
    }
    const long long n = p->rows;
    sgs_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p->y, x_dev, n, p->tickets, state);
    static int ctas_per_sm = 0;
    if (!ctas_per_sm) { const char* e = getenv("SMM_B200_SGS_CTAS_PER_SM"); ctas_per_sm = e ? atoi(e) : 4; if (ctas_per_sm < 1) ctas_per_sm = 1; }
    // SGS reads A's current values; an IC(0) factor is frozen at init() like the reference's ic0Val
    const unsigned long long want_version = p->kind == 1 ? 0ull : m->values_version;
    if (p->values_version != want_version) {                   // matrix values changed since the packed copies were gathered
        smm_precond* pm = const_cast<smm_precond*>(p);
        for (int w = 0; w < 2; ++w) {
            const long long nt = w == 0 ? p->threads_fwd : p->threads_bwd;
            const long long n = std::max(p->esize[w], nt);
            sgs_gather_values_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p->kind == 1 ? p->factor : m->values, p->eidx[w], p->eval[w], p->esize[w],
                                                                                 w == 0 ? p->order_fwd : p->order_bwd, p->diag_pos, p->dval[w], nt);
        }
        SMM_COUNT_LAUNCH(2);
        pm->values_version = want_version;
    }
    const long long cap = (long long)m->sm_count * ctas_per_sm;
    const long long bf = (p->threads_fwd + SGS_THREADS - 1) / SGS_THREADS, bb = (p->threads_bwd + SGS_THREADS - 1) / SGS_THREADS;
    SweepArgs F{p->order_fwd, p->slice_ptr[0], p->ecol[0], p->eval[0], p->dval[0], p->threads_fwd};
    SweepArgs Bk{p->order_bwd, p->slice_ptr[1], p->ecol[1], p->eval[1], p->dval[1], p->threads_bwd};
    if (p->kind == 1) {
        sgs_sweep_kernel<true, true><<<(unsigned)(bf < cap ? bf : cap), SGS_THREADS, 0, s>>>(F, rhs_dev, p->y, x_dev, p->tickets, state);
        sgs_sweep_kernel<false, true><<<(unsigned)(bb < cap ? bb : cap), SGS_THREADS, 0, s>>>(Bk, rhs_dev, p->y, x_dev, p->tickets, state);
    } else {
        sgs_sweep_kernel<true, false><<<(unsigned)(bf < cap ? bf : cap), SGS_THREADS, 0, s>>>(F, rhs_dev, p->y, x_dev, p->tickets, state);
        sgs_sweep_kernel<false, false><<<(unsigned)(bb < cap ? bb : cap), SGS_THREADS, 0, s>>>(Bk, rhs_dev, p->y, x_dev, p->tickets, state);
    }
    sgs_status_kernel<<<1, 1, 0, s>>>(p->tickets, state, rc_dev);
    SMM_COUNT_LAUNCH(4);
    SMM_CUDA(cudaGetLastError());
    return SMM_OK;
}

extern "C" {

static int precond_create(const smm_csr_t* m, int kind, int* rc_out, smm_precond_t** out) {
    if (!m || !out) return SMM_E_INVALID;
    if (m->rows != m->cols) { smm_set_error("preconditioner: matrix must be square"); return SMM_E_INVALID; }
    SMM_CUDA(cudaSetDevice(m->device));
    smm_precond* p = new smm_precond();
    p->m = m;
    p->kind = kind;
    p->rows = m->rows;
    std::vector<int32_t> start((size_t)m->rows + 1), pos((size_t)m->nnz);
    SMM_CUDA(cudaDeviceSynchronize());
    SMM_CUDA(cudaMemcpy(start.data(), m->start, sizeof(int32_t) * start.size(), cudaMemcpyDeviceToHost));
    if (m->nnz) SMM_CUDA(cudaMemcpy(pos.data(), m->positions, sizeof(int32_t) * pos.size(), cudaMemcpyDeviceToHost));
    std::vector<int32_t> diag, of, ob;
    analyse(m->rows, start, pos, m->first_active_start, &p->valid, &diag, &of, &ob, &p->levels_fwd, &p->levels_bwd);
    if (rc_out) *rc_out = p->valid ? 0 : 1;
    if (kind == 1 && p->valid && m->nnz > 0) {
        std::vector<float> a((size_t)m->nnz), l;
        SMM_CUDA(cudaMemcpy(a.data(), m->values, sizeof(float) * a.size(), cudaMemcpyDeviceToHost));
        ic0_factorize_host(m->rows, start, pos, diag, a, &l);
        SMM_CUDA(cudaMalloc(&p->factor, sizeof(float) * l.size()));
        SMM_CUDA(cudaMemcpy(p->factor, l.data(), sizeof(float) * l.size(), cudaMemcpyHostToDevice));
    }
    SMM_CUDA(cudaMalloc(&p->tickets, 4 * sizeof(unsigned int)));
    SMM_CUDA(cudaMemset(p->tickets, 0, 4 * sizeof(unsigned int)));
    if (p->valid && m->rows > 0) {
        p->threads_fwd = (long long)of.size();
        p->threads_bwd = (long long)ob.size();
        SMM_CUDA(cudaMalloc(&p->order_fwd, sizeof(int32_t) * of.size()));
        SMM_CUDA(cudaMalloc(&p->order_bwd, sizeof(int32_t) * ob.size()));
        SMM_CUDA(cudaMalloc(&p->diag_pos, sizeof(int32_t) * diag.size()));
        SMM_CUDA(cudaMalloc(&p->y, sizeof(float) * (size_t)m->rows));
        SMM_CUDA(cudaMemcpy(p->order_fwd, of.data(), sizeof(int32_t) * of.size(), cudaMemcpyHostToDevice));
        SMM_CUDA(cudaMemcpy(p->order_bwd, ob.data(), sizeof(int32_t) * ob.size(), cudaMemcpyHostToDevice));
        SMM_CUDA(cudaMemcpy(p->diag_pos, diag.data(), sizeof(int32_t) * diag.size(), cudaMemcpyHostToDevice));
        for (int w = 0; w < 2; ++w) {
            std::vector<long long> sp;
            std::vector<int32_t> ec, ei;
            build_sell(w == 0, w == 0 ? of : ob, start, pos, diag, &sp, &ec, &ei);
            const size_t n = ec.size() ? ec.size() : 1;
            const size_t nt = (w == 0 ? of : ob).size();
            p->esize[w] = (long long)ec.size();
            SMM_CUDA(cudaMalloc(&p->slice_ptr[w], sizeof(long long) * sp.size()));
            SMM_CUDA(cudaMalloc(&p->ecol[w], sizeof(int32_t) * n));
            SMM_CUDA(cudaMalloc(&p->eidx[w], sizeof(int32_t) * n));
            SMM_CUDA(cudaMalloc(&p->eval[w], sizeof(float) * n));
            SMM_CUDA(cudaMalloc(&p->dval[w], sizeof(float) * (nt ? nt : 1)));
            SMM_CUDA(cudaMemcpy(p->slice_ptr[w], sp.data(), sizeof(long long) * sp.size(), cudaMemcpyHostToDevice));
            if (!ec.empty()) {
                SMM_CUDA(cudaMemcpy(p->ecol[w], ec.data(), sizeof(int32_t) * ec.size(), cudaMemcpyHostToDevice));
                SMM_CUDA(cudaMemcpy(p->eidx[w], ei.data(), sizeof(int32_t) * ei.size(), cudaMemcpyHostToDevice));
            }
        }
    }
    *out = p;
    return SMM_OK;
}

int smm_precond_sgs_create(const smm_csr_t* m, smm_precond_t** out) { return precond_create(m, 0, nullptr, out); }

// IC0Preconditioner(m) + init() (H:1216-1235, 1798-1800): *rc receives init()'s code (0 ok, 1 = the matrix has no
// usable diagonal / is not SPD structurally).  The factor is frozen at this point, like the reference's ic0Val.
int smm_precond_ic0_create(const smm_csr_t* m, int* rc, smm_precond_t** out) { return precond_create(m, 1, rc, out); }

// copy of the factor values (nnz floats, A's pattern) for parity checks
int smm_precond_ic0_factor(const smm_precond_t* p, float* factor_host) {
    if (!p || p->kind != 1 || !factor_host) return SMM_E_INVALID;
    if (!p->factor) return SMM_E_STATE;
    SMM_CUDA(cudaMemcpy(factor_host, p->factor, sizeof(float) * (size_t)p->m->nnz, cudaMemcpyDeviceToHost));
    return SMM_OK;
}

int smm_precond_kind(const smm_precond_t* p) { return p ? p->kind : -1; }

int smm_precond_apply_dev(const smm_precond_t* p, const float* rhs_dev, float* x_dev, int* rc, void* stream) {
    if (!p || (p->rows && (!rhs_dev || !x_dev))) return SMM_E_INVALID;
    if (rc) *rc = 0;
    if (!p->valid) { if (rc) *rc = 1; return SMM_OK; }       // the reference's error exits (H:1668, 1678, 1691): x is not produced
    if (p->rows == 0) return SMM_OK;
    SMM_CUDA(cudaSetDevice(p->m->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : smm_default_stream();
    int* rc_dev = reinterpret_cast<int*>(p->tickets) + 3;      // reuse: the status kernel writes the code over the error bits
    SMM_TRY(smm_sgs_apply_async_rc(p, rhs_dev, x_dev, nullptr, rc_dev, s));
    int code = 0;
    SMM_CUDA(cudaMemcpyAsync(&code, rc_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
    SMM_CUDA(cudaStreamSynchronize(s));
    if (code & 4) { smm_set_error("SGS apply: device-side wait exceeded its bound"); return SMM_E_TIMEOUT; }
    if (rc) *rc = code & 1;
    return SMM_OK;
}

int smm_precond_apply(const smm_precond_t* pc, const float* rhs, float* x, int* rc) {
    smm_precond* p = const_cast<smm_precond*>(pc);
    if (!p || (p->rows && (!rhs || !x))) return SMM_E_INVALID;
    if (rhs == x && p->rows) { smm_set_error("SGS apply: rhs must not alias x (H:1667)"); return SMM_E_ALIAS; }
    if (rc) *rc = 0;
    if (!p->valid) { if (rc) *rc = 1; return SMM_OK; }
    if (p->rows == 0) return SMM_OK;
    SMM_CUDA(cudaSetDevice(p->m->device));
    const size_t bytes = sizeof(float) * (size_t)p->rows;
    for (int i = 0; i < 2; ++i) if (!p->io[i]) SMM_CUDA(cudaMalloc(&p->io[i], bytes));
    SMM_CUDA(cudaMemcpy(p->io[0], rhs, bytes, cudaMemcpyHostToDevice));
    SMM_TRY(smm_precond_apply_dev(p, p->io[0], p->io[1], rc, nullptr));
    SMM_CUDA(cudaMemcpy(x, p->io[1], bytes, cudaMemcpyDeviceToHost));
    return SMM_OK;
}

int smm_precond_levels(const smm_precond_t* p, int* forward_levels, int* backward_levels) {
    if (!p) return SMM_E_INVALID;
    if (forward_levels) *forward_levels = p->levels_fwd;
    if (backward_levels) *backward_levels = p->levels_bwd;
    return SMM_OK;
}

int smm_precond_destroy(smm_precond_t* p) {
    if (!p) return SMM_OK;
    cudaFree(p->order_fwd); cudaFree(p->order_bwd); cudaFree(p->diag_pos); cudaFree(p->y); cudaFree(p->tickets); cudaFree(p->factor);
    for (int w = 0; w < 2; ++w) { cudaFree(p->slice_ptr[w]); cudaFree(p->ecol[w]); cudaFree(p->eidx[w]); cudaFree(p->eval[w]); cudaFree(p->dval[w]); }
    cudaFree(p->io[0]); cudaFree(p->io[1]);
    delete p;
    return SMM_OK;
}

}  // extern "C"
