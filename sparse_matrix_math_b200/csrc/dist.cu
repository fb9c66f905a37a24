// dist.cu -- multi-GPU layer: one process per GPU, rows partitioned in contiguous blocks.
//
// Rank r owns rows [row_begin, row_end) of the global matrix and stores them as a local CSR whose columns index an
// EXTENDED vector: the contiguous window [lo, hi) of the global vector that its rows touch (owned entries in the
// middle, halo entries on both sides; lo is rounded so the owned part stays 16-byte aligned).  The extended vector,
// the reduction mailboxes and the halo flags live in one cudaMalloc block exported with CUDA IPC; peers map it and
// write into it directly over NVLink (dist_device.cuh).  There is no NCCL call and no host synchronisation inside an
// iteration: the scalar all-reduces are fused into the epilogues of the kernels that produce the partial sums, the
// halo exchange is one push kernel (P2P stores + release flags) and one wait kernel (acquire spin).
#include <string.h>

#include <vector>

#include "dist.h"
#include "epilogue.cuh"

namespace {

struct SegDev { float* dst; const float* src; long long len; long long first_block; };

__global__ void minmax_col_kernel(const int32_t* __restrict__ positions, long long nnz, int* __restrict__ mn, int* __restrict__ mx) {
    int lo = 0x7fffffff, hi = -1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (long long)gridDim.x * blockDim.x) {
        const int c = positions[i];
        lo = min(lo, c); hi = max(hi, c);
    }
    for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(mn, lo); atomicMax(mx, hi); }
}

__global__ void shift_cols_kernel(int32_t* __restrict__ positions, long long nnz, int shift) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (long long)gridDim.x * blockDim.x) positions[i] -= shift;
}

constexpr int PUSH_THREADS = 256;
constexpr int PUSH_ELEMS_PER_BLOCK = PUSH_THREADS * 4;

// copies every send segment into the peers' extended vectors, then raises this rank's flag on every destination
__global__ void __launch_bounds__(PUSH_THREADS) halo_push_kernel(const SegDev* __restrict__ segs, int nsegs, DistComm* comm, unsigned int* ticket,
                                                                 const int* __restrict__ dests, int ndests, const SolveState* st) {
    if (st != nullptr && st->done) return;
    __shared__ int last;
    for (int s = 0; s < nsegs; ++s) {
        const SegDev sg = segs[s];
        const long long b = (long long)blockIdx.x - sg.first_block;
        if (b < 0) continue;
        const long long base = b * PUSH_ELEMS_PER_BLOCK;
        if (base >= sg.len) continue;
        for (int k = threadIdx.x; k < PUSH_ELEMS_PER_BLOCK; k += PUSH_THREADS) {
            const long long i = base + k;
            if (i < sg.len) sg.dst[i] = sg.src[i];                 // peer-mapped store over NVLink
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int seq = comm->push_seq + 1u;
        for (int k = 0; k < ndests; ++k) st_release_sys_u32(comm->flags[dests[k]] + comm->rank, seq);
        comm->push_seq = seq;
        *ticket = 0u;
    }
}

// a launch with nothing to send still has to advance the exchange counter
__global__ void halo_push_empty_kernel(DistComm* comm, const SolveState* st) {
    if (st != nullptr && st->done) return;
    comm->push_seq = comm->push_seq + 1u;
}

// rows that read entries outside the owned part [own_lo, own_hi) of the extended vector
__global__ void halo_rows_kernel(const int32_t* __restrict__ start, const int32_t* __restrict__ positions, int rows, int own_lo, int own_hi,
                                 int* __restrict__ row_lo, int* __restrict__ row_hi) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    bool below = false, above = false;
    for (int k = start[r]; k < start[r + 1]; ++k) { const int c = positions[k]; below |= c < own_lo; above |= c >= own_hi; }
    if (below) atomicMax(row_lo, r + 1);
    if (above) atomicMin(row_hi, r);
}

// stand-alone SpMV (no reduction between two exchanges): wait until every destination has consumed the previous halo
__global__ void halo_credit_wait_kernel(DistComm* comm, const int* __restrict__ dests, int ndests) {
    const int k = threadIdx.x;
    if (k >= ndests) return;
    const unsigned int want = comm->ack_seq;
    const unsigned int* a = comm->acks[comm->rank] + dests[k];
    unsigned int polls = 0;
    while ((int)(ld_acquire_sys_u32(a) - want) < 0) {
        if (++polls >= SMM_DIST_POLL_LIMIT) { comm->error = 1; break; }
    }
}
// ... and tell every source that this rank's SpMV has read what they pushed
__global__ void halo_ack_kernel(DistComm* comm, const int* __restrict__ sources, int nsources) {
    const unsigned int seq = comm->ack_seq + 1u;
    __syncthreads();
    if ((int)threadIdx.x < nsources) st_release_sys_u32(comm->acks[sources[threadIdx.x]] + comm->rank, seq);
    if (threadIdx.x == 0) comm->ack_seq = seq;
}

__global__ void halo_wait_kernel(DistComm* comm, const int* __restrict__ sources, int nsources, SolveState* st) {
    if (st != nullptr && st->done) return;
    const int k = threadIdx.x;
    if (k >= nsources) return;
    const unsigned int want = comm->push_seq;                       // my own push of this exchange is already counted
    const unsigned int* f = comm->flags[comm->rank] + sources[k];
    unsigned int polls = 0;
    while ((int)(ld_acquire_sys_u32(f) - want) < 0) {
        if (++polls >= SMM_DIST_POLL_LIMIT) { comm->error = 1; if (st) { st->done = 1; st->precond_error |= 8; } break; }
    }
}

}  // namespace

// ---- used by solvers.cu --------------------------------------------------------------------------------------
int smm_dist_exchange_async(smm_dist* d, SolveState* st, cudaStream_t s, bool wait_kernel) {
    if (d->nranks == 1) return SMM_OK;
    if (!d->send.empty()) {
        long long blocks = 0;
        for (auto& sg : d->send) blocks += (sg.len + PUSH_ELEMS_PER_BLOCK - 1) / PUSH_ELEMS_PER_BLOCK;
        halo_push_kernel<<<(unsigned)blocks, PUSH_THREADS, 0, s>>>((const SegDev*)d->seg_dev, (int)d->send.size(), d->comm_dev, d->ticket,
                                                                   d->dests_dev, (int)d->dests.size(), st);
    } else {
        halo_push_empty_kernel<<<1, 1, 0, s>>>(d->comm_dev, st);
    }
    if (wait_kernel) halo_wait_kernel<<<1, 32, 0, s>>>(d->comm_dev, d->sources_dev, (int)d->sources.size(), st);
    SMM_COUNT_LAUNCH(wait_kernel ? 2 : 1);
    SMM_CUDA(cudaGetLastError());
    return SMM_OK;
}

int smm_dist_wait_async(smm_dist* d, SolveState* st, cudaStream_t s) {
    if (d->nranks == 1) return SMM_OK;
    halo_wait_kernel<<<1, 32, 0, s>>>(d->comm_dev, d->sources_dev, (int)d->sources.size(), st);
    SMM_COUNT_LAUNCH(1);
    SMM_CUDA(cudaGetLastError());
    return SMM_OK;
}

// ---- C ABI ---------------------------------------------------------------------------------------------------
int smm_solve_dist_cg_impl(smm_dist* d, const float* b_dev, const float* x0_dev, float* x_dev, int maxIterations, float eps,
                           const smm_solve_options* opts, smm_solve_info* info, cudaStream_t s);   // solvers.cu

int smm_solve_dist_impl(smm_dist* d, int solver, const float* b_dev, float* x_dev, int maxIterations, float eps,
                        const smm_solve_options* opts, smm_solve_info* info, cudaStream_t s);   // solvers.cu

extern "C" {

int smm_dist_create(int rank, int nranks, int64_t global_rows, int64_t row_begin, int64_t row_end, smm_csr_t* local, smm_dist_t** out) {
    if (!out || !local || nranks < 1 || nranks > SMM_MAX_RANKS || rank < 0 || rank >= nranks || row_begin < 0 || row_end < row_begin ||
        row_end > global_rows || local->rows != row_end - row_begin) {
        smm_set_error("smm_dist_create: bad arguments");
        return SMM_E_INVALID;
    }
    SMM_CUDA(cudaSetDevice(local->device));
    cudaStream_t s = smm_default_stream();
    smm_dist* d = new smm_dist();
    d->rank = rank; d->nranks = nranks; d->device = local->device; d->local = local;
    d->global_rows = global_rows; d->row_begin = row_begin; d->row_end = row_end;
    int* mm_dev = nullptr;
    const int rc = [&]() -> int {
        // window of the global vector the local rows read
        int mm[2] = {0x7fffffff, -1};
        SMM_CUDA(cudaMalloc(&mm_dev, 2 * sizeof(int)));
        SMM_CUDA(cudaMemcpyAsync(mm_dev, mm, sizeof mm, cudaMemcpyHostToDevice, s));
        if (local->nnz > 0) {
            minmax_col_kernel<<<1184, 256, 0, s>>>(local->positions, local->nnz, mm_dev, mm_dev + 1);
            SMM_COUNT_LAUNCH(1);
        }
        SMM_CUDA(cudaMemcpyAsync(mm, mm_dev, sizeof mm, cudaMemcpyDeviceToHost, s));
        SMM_CUDA(cudaStreamSynchronize(s));
        long long lo = row_begin, hi = row_end;
        if (local->nnz > 0) { if (mm[0] < lo) lo = mm[0]; if ((long long)mm[1] + 1 > hi) hi = (long long)mm[1] + 1; }
        if (hi > global_rows || lo < 0) { smm_set_error("smm_dist_create: column index outside the global vector"); return SMM_E_INVALID; }
        lo = row_begin - (((row_begin - lo) + 3) / 4) * 4;        // keep the owned part 16-byte aligned (may dip below 0: unused pad)
        d->lo = lo; d->hi = hi; d->own_off = row_begin - lo;
        // shared block (allocated before the matrix is touched: a failure up to here leaves `local` as it was)
        const size_t ext_bytes = (((size_t)(hi - lo) * sizeof(float)) + 255) & ~(size_t)255;
        const size_t mail_bytes = (size_t)4 * nranks * 2 * sizeof(unsigned long long);
        const size_t flag_bytes = 256;                            // flags [nranks] at +0, acknowledgements [nranks] at +128
        d->mail_off = ext_bytes; d->flag_off = ext_bytes + ((mail_bytes + 255) & ~(size_t)255);
        d->shared_bytes = d->flag_off + flag_bytes;
        SMM_CUDA(cudaMalloc(&d->shared, d->shared_bytes));
        SMM_CUDA(cudaMemsetAsync(d->shared, 0, d->shared_bytes, s));
        d->ext = reinterpret_cast<float*>(d->shared);
        SMM_CUDA(cudaMalloc(&d->comm_dev, sizeof(DistComm)));
        SMM_CUDA(cudaMalloc(&d->ticket, sizeof(unsigned int)));
        SMM_CUDA(cudaMemsetAsync(d->ticket, 0, sizeof(unsigned int), s));
        // the local CSR is re-indexed IN PLACE: columns become indices into the extended vector (documented in smm_b200.h)
        if (local->nnz > 0 && lo != 0) {
            shift_cols_kernel<<<1184, 256, 0, s>>>(local->positions, local->nnz, (int)lo);
            SMM_COUNT_LAUNCH(1);
        }
        local->cols = (int)(hi - lo);
        SMM_CUDA(cudaStreamSynchronize(s));
        return SMM_OK;
    }();
    cudaFree(mm_dev);
    if (rc != SMM_OK) {
        cudaFree(d->shared); cudaFree(d->comm_dev); cudaFree(d->ticket);
        delete d;
        return rc;
    }
    if (nranks == 1) d->connected = true;
    *out = d;
    return SMM_OK;
}

int smm_dist_info(const smm_dist_t* d, int64_t* ranges4, void* ipc_handle64) {
    if (!d) return SMM_E_INVALID;
    if (ranges4) { ranges4[0] = d->row_begin; ranges4[1] = d->row_end; ranges4[2] = d->lo; ranges4[3] = d->hi; }
    if (ipc_handle64) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        cudaIpcMemHandle_t h;
        SMM_CUDA(cudaSetDevice(d->device));
        SMM_CUDA(cudaIpcGetMemHandle(&h, d->shared));
        memcpy(ipc_handle64, &h, 64);
    }
    return SMM_OK;
}

// peer_base[] filled in (IPC-mapped or peer-accessible pointers): derive the mailbox / flag addresses, the send plan and
// the rows that read halo entries
static int dist_finish_connect(smm_dist_t* d, const int64_t* all_ranges) {
    const int P = d->nranks, me = d->rank;
    DistComm c;
    memset(&c, 0, sizeof c);
    c.rank = me; c.nranks = P;
    // every rank lays its block out as [ext | mail | flags, acks] with its own window length
    d->send.clear(); d->dests.clear(); d->sources.clear();
    std::vector<SegDev> segs;
    HaloPushDev hp;
    memset(&hp, 0, sizeof hp);
    long long first_block = 0;
    for (int r = 0; r < P; ++r) {
        const long long rb = all_ranges[4 * r], re = all_ranges[4 * r + 1], lo = all_ranges[4 * r + 2], hi = all_ranges[4 * r + 3];
        const size_t ext_bytes = (((size_t)(hi - lo) * sizeof(float)) + 255) & ~(size_t)255;
        const size_t mail_bytes = (size_t)4 * P * 2 * sizeof(unsigned long long);
        char* base = (char*)d->peer_base[r];
        c.mail[r] = reinterpret_cast<unsigned long long*>(base + ext_bytes);
        c.flags[r] = reinterpret_cast<unsigned int*>(base + ext_bytes + ((mail_bytes + 255) & ~(size_t)255));
        c.acks[r] = c.flags[r] + 32;
        c.chain[r] = reinterpret_cast<unsigned long long*>(c.flags[r] + 16);     // flags [nranks] at +0, chain word at +64, acks at +128
        if (r == me) continue;
        // what rank r needs from me: my owned rows inside its window
        const long long a = d->row_begin > lo ? d->row_begin : lo;
        const long long b = d->row_end < hi ? d->row_end : hi;
        if (b > a) {
            d->send.push_back({r, a - d->lo, a - lo, b - a});
            d->dests.push_back(r);
            SegDev sg;
            sg.dst = reinterpret_cast<float*>(base) + (a - lo);
            sg.src = d->ext + (a - d->lo);
            sg.len = b - a;
            sg.first_block = first_block;
            first_block += (sg.len + PUSH_ELEMS_PER_BLOCK - 1) / PUSH_ELEMS_PER_BLOCK;
            segs.push_back(sg);
            hp.segs[hp.nsegs].dst = sg.dst;
            hp.segs[hp.nsegs].begin = a - d->row_begin;        // relative to this rank's owned entries
            hp.segs[hp.nsegs].len = b - a;
            hp.dests[hp.nsegs] = r;
            hp.nsegs++;
        }
        // what I need from rank r: its owned rows inside my window
        const long long a2 = rb > d->lo ? rb : d->lo;
        const long long b2 = re < d->hi ? re : d->hi;
        if (b2 > a2) d->sources.push_back(r);
    }
    hp.ndests = hp.nsegs;
    if (!segs.empty()) {
        SMM_CUDA(cudaMalloc(&d->seg_dev, sizeof(SegDev) * segs.size()));
        SMM_CUDA(cudaMemcpy(d->seg_dev, segs.data(), sizeof(SegDev) * segs.size(), cudaMemcpyHostToDevice));
    }
    SMM_CUDA(cudaMalloc(&d->dests_dev, sizeof(int) * SMM_MAX_RANKS));
    SMM_CUDA(cudaMalloc(&d->sources_dev, sizeof(int) * SMM_MAX_RANKS));
    if (!d->dests.empty()) SMM_CUDA(cudaMemcpy(d->dests_dev, d->dests.data(), sizeof(int) * d->dests.size(), cudaMemcpyHostToDevice));
    if (!d->sources.empty()) SMM_CUDA(cudaMemcpy(d->sources_dev, d->sources.data(), sizeof(int) * d->sources.size(), cudaMemcpyHostToDevice));
    if (const char* e = getenv("SMM_B200_DIST_DEBUG")) c.debug = atoi(e);   // measurement only: see DistComm::debug
    SMM_CUDA(cudaMemcpy(d->comm_dev, &c, sizeof c, cudaMemcpyHostToDevice));
    // rows [0, halo_row_lo) and [halo_row_hi, rows) read halo entries: the SpMV multiplies the others while the pushes travel
    const int rows = d->local->rows;
    int bounds[2] = {0, rows};
    if (rows > 0 && d->local->nnz > 0) {
        int* bd = nullptr;
        SMM_CUDA(cudaMalloc(&bd, sizeof bounds));
        SMM_CUDA(cudaMemcpy(bd, bounds, sizeof bounds, cudaMemcpyHostToDevice));
        halo_rows_kernel<<<(rows + 255) / 256, 256>>>(d->local->start, d->local->positions, rows, (int)d->own_off, (int)(d->own_off + rows), bd, bd + 1);
        SMM_COUNT_LAUNCH(1);
        SMM_CUDA(cudaMemcpy(bounds, bd, sizeof bounds, cudaMemcpyDeviceToHost));
        cudaFree(bd);
    }
    d->halo_row_lo = bounds[0]; d->halo_row_hi = bounds[1];
    hp.comm = d->comm_dev; hp.ticket = d->ticket;
    HaloWaitDev hw;
    memset(&hw, 0, sizeof hw);
    hw.nsources = (int)d->sources.size();
    for (int k = 0; k < hw.nsources; ++k) hw.sources[k] = d->sources[k];
    hw.row_lo = bounds[0]; hw.row_hi = bounds[1]; hw.comm = d->comm_dev;
    hp.debug = c.debug;
    SMM_CUDA(cudaMalloc(&d->push_dev, sizeof hp));
    SMM_CUDA(cudaMalloc(&d->wait_dev, sizeof hw));
    SMM_CUDA(cudaMemcpy(d->push_dev, &hp, sizeof hp, cudaMemcpyHostToDevice));
    SMM_CUDA(cudaMemcpy(d->wait_dev, &hw, sizeof hw, cudaMemcpyHostToDevice));
    d->connected = true;
    return SMM_OK;
}

int smm_dist_connect(smm_dist_t* d, const int64_t* all_ranges, const void* all_handles) {
    if (!d || !all_ranges || !all_handles) return SMM_E_INVALID;
    SMM_CUDA(cudaSetDevice(d->device));
    for (int r = 0; r < d->nranks; ++r) {
        if (r == d->rank) { d->peer_base[r] = d->shared; }
        else {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const char*)all_handles + 64 * r, 64);
            SMM_CUDA(cudaIpcOpenMemHandle(&d->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess));
        }
    }
    d->ipc_mapped = true;
    return dist_finish_connect(d, all_ranges);
}

// Single-process variant: all[r] is rank r's handle, created in THIS process on its own device.  The devices are made
// peer-accessible (cudaDeviceEnablePeerAccess) and every rank addresses the others' blocks directly: no IPC handles, no
// second process, no torch.distributed.  The solvers then run one host thread per device (smm_group_*).
int smm_dist_connect_local(smm_dist_t** all, int nranks) {
    if (!all || nranks < 1 || nranks > SMM_MAX_RANKS) return SMM_E_INVALID;
    std::vector<int64_t> ranges(4 * (size_t)nranks);
    for (int r = 0; r < nranks; ++r) {
        if (!all[r] || all[r]->rank != r || all[r]->nranks != nranks) { smm_set_error("smm_dist_connect_local: handle %d is not rank %d of %d", r, r, nranks); return SMM_E_INVALID; }
        ranges[4 * r] = all[r]->row_begin; ranges[4 * r + 1] = all[r]->row_end; ranges[4 * r + 2] = all[r]->lo; ranges[4 * r + 3] = all[r]->hi;
    }
    for (int r = 0; r < nranks; ++r) {
        SMM_CUDA(cudaSetDevice(all[r]->device));
        for (int q = 0; q < nranks; ++q) {
            if (q == r || all[q]->device == all[r]->device) continue;
            int can = 0;
            SMM_CUDA(cudaDeviceCanAccessPeer(&can, all[r]->device, all[q]->device));
            if (!can) { smm_set_error("smm_dist_connect_local: device %d cannot access device %d", all[r]->device, all[q]->device); return SMM_E_CUDA; }
            const cudaError_t e = cudaDeviceEnablePeerAccess(all[q]->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return smm_cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
        }
    }
    for (int r = 0; r < nranks; ++r) {
        SMM_CUDA(cudaSetDevice(all[r]->device));
        for (int q = 0; q < nranks; ++q) all[r]->peer_base[q] = all[q]->shared;
        all[r]->ipc_mapped = false;
        SMM_TRY(dist_finish_connect(all[r], ranges.data()));
    }
    return SMM_OK;
}

// y_local = A_local * x (exchange of x's halo included); x_local_dev, y_local_dev hold this rank's owned rows.
// Calls need no barrier between them: the push waits until every destination has acknowledged the previous call's halo.
int smm_dist_spmv_dev(smm_dist_t* d, const float* x_local_dev, float* y_local_dev, void* stream) {
    if (!d || !d->connected) { smm_set_error("smm_dist_spmv_dev: not connected"); return SMM_E_STATE; }
    SMM_CUDA(cudaSetDevice(d->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : smm_default_stream();
    const long long n = d->row_end - d->row_begin;
    if (n) SMM_CUDA(cudaMemcpyAsync(d->ext + d->own_off, x_local_dev, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    const bool multi = d->nranks > 1;
    if (multi) {
        halo_credit_wait_kernel<<<1, 32, 0, s>>>(d->comm_dev, d->dests_dev, (int)d->dests.size());
        SMM_COUNT_LAUNCH(1);
    }
    const bool fused = multi && smm_spmv_rows_lanes(d->local, 0) > 0;
    SMM_TRY(smm_dist_exchange_async(d, nullptr, s, !fused));
    SpmvArgs a;
    a.m = d->local; a.op = SMM_OP_ASSIGN; a.mult = d->ext; a.out = y_local_dev;
    a.halo_wait = fused ? d->wait_dev : nullptr;
    SMM_TRY(smm_launch_spmv(a, s));
    if (multi) {
        halo_ack_kernel<<<1, 32, 0, s>>>(d->comm_dev, d->sources_dev, (int)d->sources.size());
        SMM_COUNT_LAUNCH(1);
        SMM_CUDA(cudaGetLastError());
    }
    return SMM_OK;
}

int smm_dist_solve_cg(smm_dist_t* d, const float* b_dev, const float* x0_dev, float* x_dev, int maxIterations, float eps,
                      const smm_solve_options* opts, smm_solve_info* info, void* stream) {
    if (!d || !d->connected) { smm_set_error("smm_dist_solve_cg: not connected"); return SMM_E_STATE; }
    cudaStream_t s = stream ? (cudaStream_t)stream : smm_default_stream();
    return smm_solve_dist_cg_impl(d, b_dev, x0_dev, x_dev, maxIterations, eps, opts, info, s);
}

static int dist_solve_other(smm_dist_t* d, int solver, const float* b_dev, float* x_dev, int maxIterations, float eps,
                            const smm_solve_options* opts, smm_solve_info* info, void* stream) {
    if (!d || !d->connected) { smm_set_error("smm_dist_solve: not connected"); return SMM_E_STATE; }
    cudaStream_t s = stream ? (cudaStream_t)stream : smm_default_stream();
    return smm_solve_dist_impl(d, solver, b_dev, x_dev, maxIterations, eps, opts, info, s);
}
int smm_dist_solve_bicgsym(smm_dist_t* d, const float* b_dev, float* x_dev, int maxIterations, float eps,
                           const smm_solve_options* opts, smm_solve_info* info, void* stream) {
    return dist_solve_other(d, 1, b_dev, x_dev, maxIterations, eps, opts, info, stream);
}
int smm_dist_solve_cgs(smm_dist_t* d, const float* b_dev, float* x_dev, int maxIterations, float eps,
                       const smm_solve_options* opts, smm_solve_info* info, void* stream) {
    return dist_solve_other(d, 2, b_dev, x_dev, maxIterations, eps, opts, info, stream);
}
int smm_dist_solve_bicgstab(smm_dist_t* d, const float* b_dev, float* x_dev, int maxIterations, float eps,
                            const smm_solve_options* opts, smm_solve_info* info, void* stream) {
    return dist_solve_other(d, 3, b_dev, x_dev, maxIterations, eps, opts, info, stream);
}

// measurement hook (tools/dist_kernel_times.py): the three kernels of this rank's CG iteration, timed back to back; needs
// SMM_B200_DIST_DEBUG=7 at create time (no waiting for peers, no pushes)
int smm_dist_profile_cg_iteration(smm_dist_t* d, int reps, float* ms_spmv, float* ms_r, float* ms_px, void* stream) {
    if (!d || !d->connected) return SMM_E_STATE;
    const char* e = getenv("SMM_B200_DIST_DEBUG");
    if (!e || (atoi(e) & 6) != 6) { smm_set_error("smm_dist_profile_cg_iteration: set SMM_B200_DIST_DEBUG=7 before creating the handle"); return SMM_E_STATE; }
    return smm_profile_cg_iteration_impl(d->local, d, reps, ms_spmv, ms_r, ms_px, stream);
}

int smm_dist_error(const smm_dist_t* d, int* error) {
    if (!d || !error) return SMM_E_INVALID;
    DistComm c;
    SMM_CUDA(cudaDeviceSynchronize());
    SMM_CUDA(cudaMemcpy(&c, d->comm_dev, sizeof c, cudaMemcpyDeviceToHost));
    *error = c.error;
    return SMM_OK;
}

int smm_dist_destroy(smm_dist_t* d) {
    if (!d) return SMM_OK;
    cudaSetDevice(d->device);
    cudaDeviceSynchronize();
    if (d->ipc_mapped) for (int r = 0; r < d->nranks; ++r) if (r != d->rank && d->peer_base[r]) cudaIpcCloseMemHandle(d->peer_base[r]);
    cudaFree(d->shared); cudaFree(d->comm_dev); cudaFree(d->ticket); cudaFree(d->seg_dev); cudaFree(d->dests_dev); cudaFree(d->sources_dev);
    cudaFree(d->push_dev); cudaFree(d->wait_dev);
    delete d;
    return SMM_OK;
}

}  // extern "C"
