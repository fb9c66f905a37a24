// dist_device.cuh -- device-side communication primitives of the multi-GPU solvers (one process per GPU).
//
// All traffic is direct peer-to-peer loads/stores over NVLink into buffers every rank exports with CUDA IPC:
//   * scalar all-reduce, fused into the epilogue of the kernel that produced the partial sum: the last CTA of every
//     rank writes (value, sequence tag) as ONE 64-bit word into its slot of every peer's mailbox, then waits until
//     its own mailbox holds the current tag from every rank and adds the values in rank order -- every rank gets
//     the same bits, so all ranks take the same branches without any further agreement.  Four mailbox sets cycle,
//     and a rank cannot run two reductions ahead of a peer (it needs that peer's contribution), so a set is never
//     overwritten before it has been read.
//   * halo exchange: the boundary entries of the local vector are stored straight into the peers' extended vectors
//     and a per-source flag is raised (release, system scope) -- by the CTAs of the CG p-update themselves
//     (HaloPushDev, vecops.cu) or by a push kernel for every other operand; the consumer is the SpMV: its rows kernel
//     schedules the rows that read no halo FIRST and only the warps that reach a boundary row group spin on the flags
//     (acquire, system scope; HaloWaitDev, spmv.cu), gathering those groups' operands past L1.  Other SpMV kernels
//     are preceded by a wait kernel.
//   * chained sums: BiCGStab's ||r||^2 is ONE left-to-right sum over the whole vector in both builds of the reference
//     (H:2262-2267), not a per-rank quantity: rank r starts its part from the running sum rank r - 1 hands it (one 64-bit
//     word, value + sequence tag, into chain[r]) and hands its own on; the last rank's total reaches everybody through the
//     ordinary all-reduce with zeros from the others (x + 0 = x).
//   * back-pressure: inside the solvers a fused all-reduce sits between two exchanges, so a rank cannot push again
//     before every peer's SpMV has finished reading its halo.  The stand-alone smm_dist_spmv_dev has no such
//     reduction: there the consumer acknowledges every exchange (acks[], release) and the pusher waits for the
//     acknowledgement of the previous one before it overwrites a peer's halo.
// Every spin is bounded; on expiry the error flag is set and the solve is marked done so nothing can hang.
#pragma once
#include <stdint.h>

constexpr int SMM_MAX_RANKS = 8;
constexpr unsigned int SMM_DIST_POLL_LIMIT = 1u << 27;

struct DistComm {
    int rank, nranks;
    unsigned int red_seq;                       // reductions completed (same on every rank)
    unsigned int push_seq;                      // halo exchanges completed
    int error;                                  // 1: a bounded wait expired
    int tree_order;                             // 1: combine the ranks' values pairwise (reference-tree mode), 0: in rank order
    unsigned long long* mail[SMM_MAX_RANKS];    // mail[d]: rank d's mailbox [4 sets][nranks][2] (mail[rank] is local)
    unsigned int* flags[SMM_MAX_RANKS];         // flags[d]: rank d's flag array [nranks]; this rank writes flags[d][rank]
    unsigned int* acks[SMM_MAX_RANKS];          // acks[d]: rank d's ack array [nranks]; this rank writes acks[d][rank] (stand-alone SpMV)
    unsigned int ack_seq;                       // stand-alone SpMV calls completed
    unsigned long long* chain[SMM_MAX_RANKS];   // chain[d]: rank d's inbound word of a sum that is CHAINED through the ranks
    unsigned int chain_seq;                     // chained sums completed (same on every rank)
    int debug;                                  // measurement only (SMM_B200_DIST_DEBUG, bit mask): 1 = no all-reduce (local sums), 2 = no halo wait, 4 = no halo push
};

// halo push fused into an element-wise kernel: segment s covers elements [begin, begin + len) of the kernel's first
// output (this rank's owned entries) and goes to dst[0 .. len) in a peer's extended vector
struct HaloSeg { float* dst; long long begin; long long len; };
struct HaloPushDev {
    int nsegs, ndests;
    int debug;                                  // copy of DistComm::debug (set at connect time)
    HaloSeg segs[SMM_MAX_RANKS];
    int dests[SMM_MAX_RANKS];
    DistComm* comm;
    unsigned int* ticket;
};
// halo wait fused into the SpMV: rows [0, row_lo) and [row_hi, rows) may read halo entries
struct HaloWaitDev {
    int nsources;
    int sources[SMM_MAX_RANKS];
    int row_lo, row_hi;
    DistComm* comm;
};

#ifdef __CUDACC__
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// spin until every source rank has raised its flag for the current exchange; false when the bounded wait expired
__device__ __forceinline__ bool dist_halo_wait(const HaloWaitDev* w) {
    DistComm* c = w->comm;
    if (c->debug & 2) return true;
    const unsigned int want = c->push_seq;                  // my own push of this exchange is already counted
    const unsigned int* f = c->flags[c->rank];
    for (int k = 0; k < w->nsources; ++k) {
        unsigned int polls = 0;
        while ((int)(ld_acquire_sys_u32(f + w->sources[k]) - want) < 0) {
            if (++polls >= SMM_DIST_POLL_LIMIT) { c->error = 1; return false; }
        }
    }
    return true;
}

// Called by ONE thread per rank.  Sums t0 and t1 over all ranks: in rank order, or pairwise in reference-tree mode.
__device__ __forceinline__ void dist_allreduce2(DistComm* c, float& t0, float& t1) {
    const unsigned int seq = c->red_seq + 1u;
    if (c->debug & 1) { c->red_seq = seq; return; }
    const int P = c->nranks, me = c->rank;
    const size_t set = (size_t)(seq & 3u) * P * 2;
    const unsigned long long w0 = ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(t0);
    const unsigned long long w1 = ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(t1);
    for (int k = 0; k < P; ++k) {                           // start with the next rank: spread the NVLink writes
        const int d = (me + 1 + k) % P;
        st_sys_u64(c->mail[d] + set + 2 * me, w0);
        st_sys_u64(c->mail[d] + set + 2 * me + 1, w1);
    }
    float v0[SMM_MAX_RANKS], v1[SMM_MAX_RANKS];
    const unsigned long long* mine = c->mail[me] + set;
    for (int s = 0; s < P; ++s) {
        unsigned long long a = 0, b = 0;
        unsigned int polls = 0;
        for (;;) {
            a = ld_sys_u64(mine + 2 * s);
            b = ld_sys_u64(mine + 2 * s + 1);
            if ((unsigned int)(a >> 32) == seq && (unsigned int)(b >> 32) == seq) break;
            if (++polls >= SMM_DIST_POLL_LIMIT) { c->error = 1; break; }
        }
        v0[s] = __uint_as_float((unsigned int)a);
        v1[s] = __uint_as_float((unsigned int)b);
    }
    float s0 = 0.0f, s1 = 0.0f;
    if (c->tree_order) {
        // the ranks own the depth-log2(P) nodes of the reference's reduction tree (tbb::parallel_deterministic_reduce over
        // the whole vector, H:308-320): their values are joined pairwise, left + right, like the levels above them
        for (int w = P >> 1; w >= 1; w >>= 1)
            for (int k = 0; k < w; ++k) { v0[k] = __fadd_rn(v0[2 * k], v0[2 * k + 1]); v1[k] = __fadd_rn(v1[2 * k], v1[2 * k + 1]); }
        s0 = v0[0];
        s1 = v1[0];
    } else {
        for (int s = 0; s < P; ++s) { s0 += v0[s]; s1 += v1[s]; }
    }
    t0 = s0;
    t1 = s1;
    c->red_seq = seq;
}
#endif
