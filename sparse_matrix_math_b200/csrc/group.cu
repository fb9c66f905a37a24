// group.cu -- single-process multi-GPU: the partitioned solvers of dist.cu behind ONE host call.
//
// The per-rank machinery (row blocks, extended vector, P2P halo pushes, all-reduce fused into the kernels' epilogues) is
// that of dist.cu; what differs is how the ranks come to exist.  There: one process per GPU, CUDA IPC handles carried by
// the host program's own transport.  Here: ONE process owns every GPU.  smm_group_create cuts a host CSR into contiguous
// row blocks, uploads block r to devices[r], makes the devices peer-accessible (smm_dist_connect_local) and every solve
// runs one host thread per device -- each thread drives its device exactly like a rank of the multi-process solve (its own
// stream, its own CUDA graphs), the kernels find each other through the peer-mapped mailboxes and flags.  This is what
// makes the N-GPU solve reachable from the C++ drop-in header: SMM::ConjugateGradient<float>(a, b, x0, x, ...) with
// SMM::b200::devices() = N (include/smm_b200.hpp) takes host pointers in and out like the reference's function
// (H:2316-2324) and needs neither torch nor a launcher.
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "dist.h"

int smm_solve_dist_cg_impl(smm_dist* d, const float* b_dev, const float* x0_dev, float* x_dev, int maxIterations, float eps,
                           const smm_solve_options* opts, smm_solve_info* info, cudaStream_t s);   // solvers.cu
int smm_solve_dist_impl(smm_dist* d, int solver, const float* b_dev, float* x_dev, int maxIterations, float eps,
                        const smm_solve_options* opts, smm_solve_info* info, cudaStream_t s);   // solvers.cu
int smm_solve_prepare(const smm_csr* a, const smm_solve_options* opts);                         // solvers.cu

struct smm_group {
    int n = 0;                                  // devices = ranks
    int rows = 0, cols = 0;
    std::vector<int> devices;
    std::vector<int64_t> cut;                   // [n + 1] row_begin of every rank, then rows
    std::vector<int64_t> nnz_cut;               // [n + 1] first stored entry of every rank
    std::vector<smm_csr_t*> local;
    std::vector<smm_dist_t*> dist;
    std::vector<cudaStream_t> stream;
    std::vector<float*> b, x, y;                // per-device slices (b / lhs, x / mult, out)
    int partition = 0;
};

namespace {

// f(r) on one host thread per device; returns the first non-zero code (and keeps that thread's error text)
template <class F>
int on_every_device(const smm_group* g, F f) {
    std::vector<int> rc((size_t)g->n, SMM_OK);
    std::vector<std::string> err((size_t)g->n);
    std::vector<std::thread> th;
    for (int r = 0; r < g->n; ++r) {
        th.emplace_back([&, r] {
            if (cudaSetDevice(g->devices[r]) != cudaSuccess) { rc[r] = SMM_E_CUDA; err[r] = "cudaSetDevice failed"; return; }
            rc[r] = f(r);
            if (rc[r] != SMM_OK) err[r] = smm_last_error();      // the error text is thread-local
        });
    }
    for (std::thread& t : th) t.join();
    for (int r = 0; r < g->n; ++r) if (rc[r] != SMM_OK) { smm_set_error("device %d (rank %d): %s", g->devices[r], r, err[r].c_str()); return rc[r]; }
    return SMM_OK;
}

// all host threads of a solve meet here after their allocations and before their first kernel: once the kernels of one
// device spin on another device's contribution, nothing that could serialise against a running kernel (cudaMalloc with peer
// mappings, cudaFree) may be left to do
struct HostBarrier {
    std::mutex mu;
    std::condition_variable cv;
    int waiting = 0, generation = 0;
    const int n;
    explicit HostBarrier(int n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        const int gen = generation;
        if (++waiting == n) { waiting = 0; ++generation; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != generation; });
    }
};

void group_free(smm_group* g) {
    if (!g) return;
    for (int r = 0; r < g->n; ++r) {
        cudaSetDevice(g->devices[r]);
        if (r < (int)g->dist.size() && g->dist[r]) smm_dist_destroy(g->dist[r]);
        if (r < (int)g->local.size() && g->local[r]) smm_csr_destroy(g->local[r]);
        if (r < (int)g->b.size()) { cudaFree(g->b[r]); cudaFree(g->x[r]); cudaFree(g->y[r]); }
        if (r < (int)g->stream.size() && g->stream[r]) cudaStreamDestroy(g->stream[r]);
    }
    delete g;
}

}  // namespace

extern "C" {

int smm_group_create(int rows, int cols, const int32_t* start, const int32_t* positions, const float* values, int ndevices, const int* devices,
                     int partition, smm_group_t** out) {
    if (!out || rows < 0 || cols < 0 || !start || ndevices < 1 || ndevices > SMM_MAX_RANKS || rows != cols) {
        smm_set_error("smm_group_create: bad arguments (square matrix, 1..%d devices)", SMM_MAX_RANKS);
        return SMM_E_INVALID;
    }
    int have = 0;
    SMM_CUDA(cudaGetDeviceCount(&have));
    int prev = 0;
    cudaGetDevice(&prev);
    smm_group* g = new smm_group();
    g->n = ndevices; g->rows = rows; g->cols = cols; g->partition = partition;
    for (int r = 0; r < ndevices; ++r) {
        const int dev = devices ? devices[r] : r;
        if (dev < 0 || dev >= have) { delete g; smm_set_error("smm_group_create: device %d of %d requested", dev, have); return SMM_E_INVALID; }
        g->devices.push_back(dev);
    }
    // row blocks: the nodes of the reference's reduction tree over [0, rows) (tbb::parallel_deterministic_reduce halves a
    // range at lo + (hi - lo) / 2, H:308-320; needed by the reference-order reduction mode) or equal numbers of stored entries
    const int64_t nnz = start[rows];
    g->cut.assign((size_t)ndevices + 1, 0);
    g->cut[ndevices] = rows;
    if (partition == 1) {
        if (ndevices & (ndevices - 1)) { delete g; smm_set_error("smm_group_create: the reference-tree partition needs a power-of-two number of devices"); return SMM_E_INVALID; }
        std::vector<int64_t> c = {0, rows};
        while ((int)c.size() - 1 < ndevices) {
            std::vector<int64_t> nx;
            for (size_t i = 0; i + 1 < c.size(); ++i) { nx.push_back(c[i]); nx.push_back(c[i] + (c[i + 1] - c[i]) / 2); }
            nx.push_back(rows);
            c.swap(nx);
        }
        g->cut = c;
    } else {
        for (int r = 1; r < ndevices; ++r) {
            const int64_t target = nnz * r / ndevices;
            int64_t lo = g->cut[r - 1], hi = rows;                   // first row whose first entry is >= target
            while (lo < hi) { const int64_t mid = lo + (hi - lo) / 2; if (start[mid] < target) lo = mid + 1; else hi = mid; }
            g->cut[r] = (lo / 4) * 4 > g->cut[r - 1] ? (lo / 4) * 4 : lo;   // 16-byte aligned slices where possible
        }
    }
    g->nnz_cut.resize((size_t)ndevices + 1);
    for (int r = 0; r <= ndevices; ++r) g->nnz_cut[r] = start[g->cut[r]];
    g->local.assign(ndevices, nullptr); g->dist.assign(ndevices, nullptr); g->stream.assign(ndevices, nullptr);
    g->b.assign(ndevices, nullptr); g->x.assign(ndevices, nullptr); g->y.assign(ndevices, nullptr);
    int rc = on_every_device(g, [&](int r) -> int {
        const int64_t rb = g->cut[r], re = g->cut[r + 1], k0 = g->nnz_cut[r];
        const int n = (int)(re - rb);
        std::vector<int32_t> st((size_t)n + 1);
        for (int i = 0; i <= n; ++i) st[i] = (int32_t)(start[rb + i] - k0);
        SMM_TRY(smm_csr_create(n, cols, st.data(), positions + k0, values + k0, &g->local[r]));   // global column indices
        SMM_TRY(smm_dist_create(r, g->n, rows, rb, re, g->local[r], &g->dist[r]));
        SMM_CUDA(cudaStreamCreateWithFlags(&g->stream[r], cudaStreamNonBlocking));
        const size_t bytes = sizeof(float) * (size_t)(n > 0 ? n : 1);
        SMM_CUDA(cudaMalloc(&g->b[r], bytes));
        SMM_CUDA(cudaMalloc(&g->x[r], bytes));
        SMM_CUDA(cudaMalloc(&g->y[r], bytes));
        return SMM_OK;
    });
    if (rc == SMM_OK && ndevices > 1) rc = smm_dist_connect_local(g->dist.data(), ndevices);
    cudaSetDevice(prev);
    if (rc != SMM_OK) { group_free(g); return rc; }
    *out = g;
    return SMM_OK;
}

int smm_group_destroy(smm_group_t* g) {
    int prev = 0;
    cudaGetDevice(&prev);
    group_free(g);
    cudaSetDevice(prev);
    return SMM_OK;
}

int smm_group_info(const smm_group_t* g, int* ndevices, int64_t* row_cuts) {
    if (!g) return SMM_E_INVALID;
    if (ndevices) *ndevices = g->n;
    if (row_cuts) for (int r = 0; r <= g->n; ++r) row_cuts[r] = g->cut[r];
    return SMM_OK;
}

// values changed on the host (CSRMatrix::operator*=, updateEntry ... H:1525-1604): every device refreshes its rows
int smm_group_update_values(smm_group_t* g, const float* values) {
    if (!g || !values) return SMM_E_INVALID;
    return on_every_device(g, [&](int r) -> int { return smm_csr_update_values(g->local[r], values + g->nnz_cut[r]); });
}

// CSRMatrix::rMult / rMultAdd / rMultSub (H:1501-1515) on the partitioned matrix; host vectors of the global system
int smm_group_spmv(smm_group_t* g, int op, const float* lhs, const float* mult, float* out) {
    if (!g || !mult || !out || op < SMM_OP_ASSIGN || op > SMM_OP_SUB || (op != SMM_OP_ASSIGN && !lhs)) { smm_set_error("smm_group_spmv: bad arguments"); return SMM_E_INVALID; }
    if (mult == out) { smm_set_error("rMult: mult and out must not alias (H:1503)"); return SMM_E_ALIAS; }
    return on_every_device(g, [&](int r) -> int {
        const int64_t rb = g->cut[r];
        const size_t n = (size_t)(g->cut[r + 1] - rb), bytes = n * sizeof(float);
        cudaStream_t s = g->stream[r];
        if (n) SMM_CUDA(cudaMemcpyAsync(g->x[r], mult + rb, bytes, cudaMemcpyHostToDevice, s));
        SMM_TRY(smm_dist_spmv_dev(g->dist[r], g->x[r], g->y[r], s));
        if (op == SMM_OP_ASSIGN) {
            if (n) SMM_CUDA(cudaMemcpyAsync(out + rb, g->y[r], bytes, cudaMemcpyDeviceToHost, s));
            SMM_CUDA(cudaStreamSynchronize(s));
        } else {                                            // out = lhs +- A mult: one rounding per row like H:1509 / H:1514; out may alias lhs
            std::vector<float> y(n);
            if (n) SMM_CUDA(cudaMemcpyAsync(y.data(), g->y[r], bytes, cudaMemcpyDeviceToHost, s));
            SMM_CUDA(cudaStreamSynchronize(s));
            for (size_t i = 0; i < n; ++i) out[rb + i] = op == SMM_OP_ADD ? lhs[rb + i] + y[i] : lhs[rb + i] - y[i];
        }
        return SMM_OK;
    });
}

// solver: 0 ConjugateGradient (H:2316), 1 BiCGSymmetric (H:2021), 2 ConjugateGradientSquared (H:2109), 3 BiCGStab without
// preconditioner (H:2294).  b, x0, x: HOST vectors of the global system (x0 == x allowed; ignored unless solver == 0).
int smm_group_solve(smm_group_t* g, int solver, const float* b, const float* x0, float* x, int maxIterations, float eps,
                    const smm_solve_options* opts, smm_solve_info* info) {
    if (!g || solver < 0 || solver > 3 || (g->rows && (!b || !x || (solver == 0 && !x0)))) { smm_set_error("smm_group_solve: bad arguments"); return SMM_E_INVALID; }
    if (opts && opts->reduction_mode != SMM_REDUCE_FAST && g->n > 1 && g->partition != 1) {
        smm_set_error("smm_group_solve: the reference-order reduction modes need the reference-tree partition (smm_group_create(..., partition = 1))");
        return SMM_E_INVALID;
    }
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<smm_solve_info> infos((size_t)g->n);
    const float* guess = solver == 0 ? x0 : x;
    smm_solve_options o;
    if (opts) o = *opts; else memset(&o, 0, sizeof o);
    HostBarrier ready(g->n);
    const int rc = on_every_device(g, [&](int r) -> int {
        const int64_t rb = g->cut[r];
        const size_t n = (size_t)(g->cut[r + 1] - rb), bytes = n * sizeof(float);
        cudaStream_t s = g->stream[r];
        const int prc = smm_solve_prepare(g->local[r], &o);                   // every allocation of the solve, up front
        ready.wait();
        if (prc != SMM_OK) return prc;
        smm_solve_options mine = o;
        if (r != 0) { mine.history = nullptr; mine.history_cap = 0; }        // the scalar state is identical on every rank: rank 0 reports it
        if (n) {
            SMM_CUDA(cudaMemcpyAsync(g->b[r], b + rb, bytes, cudaMemcpyHostToDevice, s));
            SMM_CUDA(cudaMemcpyAsync(g->x[r], guess + rb, bytes, cudaMemcpyHostToDevice, s));
        }
        if (solver == 0) SMM_TRY(smm_solve_dist_cg_impl(g->dist[r], g->b[r], g->x[r], g->x[r], maxIterations, eps, &mine, &infos[r], s));
        else SMM_TRY(smm_solve_dist_impl(g->dist[r], solver, g->b[r], g->x[r], maxIterations, eps, &mine, &infos[r], s));
        // ConjugateGradient returns before touching x when the initial residual already passes (H:2342-2344)
        const bool x_written = !(solver == 0 && infos[r].iterations == 0) || x == x0;
        if (n && x_written) SMM_CUDA(cudaMemcpyAsync(x + rb, g->x[r], bytes, cudaMemcpyDeviceToHost, s));
        SMM_CUDA(cudaStreamSynchronize(s));
        int err = 0;
        SMM_TRY(smm_dist_error(g->dist[r], &err));
        if (err) { smm_set_error("a bounded device-side wait between the GPUs expired"); return SMM_E_TIMEOUT; }
        return SMM_OK;
    });
    if (rc != SMM_OK) return rc;
    if (info) {
        *info = infos[0];
        for (int r = 1; r < g->n; ++r) {
            if (infos[r].seconds_solve > info->seconds_solve) info->seconds_solve = infos[r].seconds_solve;   // max over ranks, device-timed
            info->kernel_launches += infos[r].kernel_launches;
            if (infos[r].iterations != infos[0].iterations || infos[r].status != infos[0].status) { smm_set_error("the ranks disagree on the scalar state"); return SMM_E_STATE; }
        }
        info->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    return SMM_OK;
}

}  // extern "C"
