// dist.h -- host-side state of the multi-GPU layer (shared by dist.cu and solvers.cu)
#pragma once
#include <vector>

#include "dist_device.cuh"
#include "smm_internal.cuh"

struct smm_dist {
    int rank = 0, nranks = 1, device = 0;
    smm_csr* local = nullptr;            // rows = owned rows, cols = window length, columns remapped into the window
    long long global_rows = 0, row_begin = 0, row_end = 0;
    long long lo = 0, hi = 0;            // window [lo, hi) of the global vector this rank's rows read
    long long own_off = 0;               // row_begin - lo (multiple of 4: the owned part stays 16-byte aligned)
    // one IPC-exported block: [extended vector | mailboxes | halo flags]
    void* shared = nullptr;
    size_t shared_bytes = 0, mail_off = 0, flag_off = 0;
    float* ext = nullptr;
    void* peer_base[SMM_MAX_RANKS] = {nullptr};
    DistComm* comm_dev = nullptr;
    // halo plan
    struct Seg { int peer; long long src_off, dst_off, len; };
    std::vector<Seg> send;               // my owned entries -> a peer's extended vector
    std::vector<int> dests;              // ranks I push to
    std::vector<int> sources;            // ranks that push to me
    void* seg_dev = nullptr;
    int* dests_dev = nullptr;
    int* sources_dev = nullptr;
    unsigned int* ticket = nullptr;
    HaloPushDev* push_dev = nullptr;     // the send plan for kernels that push their own output (vecops.cu)
    HaloWaitDev* wait_dev = nullptr;     // sources + the rows that read halo entries, for the SpMV that waits by itself (spmv.cu)
    int halo_row_lo = 0, halo_row_hi = 0; // rows [0, lo) and [hi, rows) may read halo entries
    bool ipc_mapped = false;             // peers mapped with CUDA IPC (one process per GPU) rather than peer access (one process)
    bool connected = false;
};

// Push this rank's boundary entries of d->ext to the peers (one launch on s).  wait_kernel: also launch the kernel that
// spins on the peers' flags; pass false when the consumer is an SpMV launched with SpmvArgs::halo_wait = d->wait_dev.
int smm_dist_exchange_async(smm_dist* d, SolveState* st, cudaStream_t s, bool wait_kernel = true);
// the wait kernel alone (the push was done by the kernel that produced the operand)
int smm_dist_wait_async(smm_dist* d, SolveState* st, cudaStream_t s);
// solvers.cu: per-kernel times of a CG iteration (dist == nullptr: the single-GPU kernels)
int smm_profile_cg_iteration_impl(const smm_csr_t* a, smm_dist* dist, int reps, float* ms_spmv, float* ms_xr, float* ms_p, void* stream);
