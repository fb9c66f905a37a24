// sgs_lines.cu -- line schedule of the triangular sweeps (SGSPreconditioner::apply H:1658-1713, IC0Preconditioner::apply
// H:1802-1837, and the ILU(0) apply) for matrices whose couplings are those of a natural-order grid stencil.
//
// The tile schedule (sgs_tiles.cu) gives a warp a 4 x 4 x 4 block of the grid: ten dependent steps with 1 to 12 of its 32
// lanes busy, and a hand-off through L2 per block -- dependency-bound at 23 % of the HBM roofline.  Here a LANE owns a grid
// line (the rows i = 0 .. nx-1 at fixed j, k: consecutive row indices, each depending on its predecessor), a WARP owns a
// patch of 32 neighbouring lines (8 x 4 in j, k; 32 x 1 on a 2D grid), and the lanes walk their lines skewed against each
// other: at step s the lane of line (a, b) of the patch solves row i = s - a - b.  Every coupling of a 5- or 7-point stencil
// then points to a row solved ONE step earlier -- by the same lane (i - 1), by a neighbouring lane of the warp (j - 1, k - 1:
// read back from a small ring of results in shared memory), or, on two faces of the patch, by a lane of a neighbouring
// patch, whose warp runs the same schedule a few steps ahead (read from the position-ordered result vector in global
// memory, which doubles as the ready flag like in the other schedules: NaN payload = not there yet).  So all 32 lanes work
// at every step (nx of nx + 10 steps), a patch needs no hand-off at all inside itself, and what crosses L2 is requested
// LINE_AHEAD steps before it is needed.  A warp is alone on its dependency chain, so the step is written for LATENCY: the
// per-row data (operand sources, coefficients, diagonal, row index: 32 bytes per row, lane-major, two 128-bit shared-memory
// loads) are packed in processing order and streamed into shared memory by the TMA engine in blocks of 8 steps (8 KB per
// cp.async.bulk, three blocks in flight, the block after them prefetched into L2); a row's packed data are read LINE_AHEAD
// steps early into registers, at which point its right-hand side and its out-of-patch operands are requested with
// cp.async; the eight steps of a block are unrolled so that every ring index is a constant.
//
// As with the tiles, the geometry is only a PROPOSAL (patch, lane and step of every row, computed from the grid shape the
// column offsets suggest); what makes it a schedule is verified for every stored entry on the device when the layout is
// built (line_build_kernel): an operand from the same patch must have been solved 1 .. LINE_RING-1 steps earlier, an
// operand from another patch must belong to a patch that is claimed earlier AND have a smaller time (patch offset + step),
// a row has at most W operands per sweep.  Anything else (periodic couplings, offsets that only look like a grid) makes the
// build return false and the tile or row schedule takes over.  Deadlock freedom: warps claim patches in time order from an
// atomic ticket, so a producer patch is always running or finished, and every wait points to a strictly smaller time.
// Per-row arithmetic is the reference's (operand order, two roundings per term, one division): bit-identical results.
//
// Bytes per row and sweep: 32 packed + 4 right-hand side + 4 (+ 4) results = 40 / 44 -- SURVEY 8(d) counts
// B_sgs = 8 nnz + 32 n = 88 per row and apply for the 7-point stencil.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "sgs_internal.cuh"

namespace {

constexpr int LINE_RING = 8;          // steps of results a warp keeps in shared memory for operands from its own patch
constexpr int LINE_BLOCK = 8;         // steps per bulk copy of packed rows (the unrolled body of the step loop)
constexpr int LINE_NBLK = 3;          // blocks of packed rows in shared memory
constexpr int LINE_AHEAD = 8;         // steps between reading a row's packed data (and requesting its right-hand side) and solving it
constexpr int LINE_NW = 4;            // warps per CTA: they share a patch and take turns step by step
constexpr int LINE_MARGIN = 2;        // a patch starts when the out-of-patch operands of its first LINE_AHEAD + LINE_MARGIN steps are there
constexpr int LINE_WORDS = 8;         // packed words per row: 3 operand sources, 3 coefficients, diagonal, row index
constexpr int LINE_MAX_W = 3;
constexpr int LINE_NONE = -1;         // operand slot not used
constexpr int LINE_LOCAL = -2;        // -2 - (32 * step + lane) % (32 * LINE_RING): operand from the warp's own ring
static_assert(LINE_RING == LINE_BLOCK, "the ring index of a step is its index inside the block");
static_assert(LINE_AHEAD == LINE_BLOCK, "the unrolled step loop reads the rows of the next block while it solves this one");

struct LineGeom {
    int nx, ny, nz;                   // proposed grid
    int A, B;                         // lines per patch in j and k (A * B == 32)
    int NJ, NK;                       // patches in j and k
    int S;                            // steps per patch: nx + A - 1 + B - 1, rounded up to whole blocks
    int W;                            // operands per row and sweep
    int npatch;
};

struct LineArgs {
    const uint32_t* pack;             // [patch in claim order][step in processing order][lane][LINE_WORDS]
    int npatch, S, rows;
    unsigned int sleep_first, sleep_later;
    unsigned long long* trace;        // debug (SMM_B200_SGS_TRACE): [patch in claim order][4] globaltimer at claim / start / end, and the SM
};

__host__ __device__ __forceinline__ void line_where(const LineGeom& G, const int32_t* __restrict__ rank_of, const int r, int* q, int* s, int* lane) {
    const int i = r % G.nx, jk = r / G.nx;
    const int j = jk % G.ny, k = jk / G.ny;
    const int a = j % G.A, b = k % G.B;
    *q = rank_of[(k / G.B) * G.NJ + j / G.A];
    *s = i + a + b;
    *lane = a + G.A * b;
}

// Layout + verification, one CTA per patch (rank q).  fail: bit 0 = a row has more than W operands, bit 1 = an operand of the
// same patch is not 1 .. LINE_RING-1 steps old, bit 2 = an operand of another patch is not earlier in claim order and time.
// Positions (yperm / xperm, operand codes) are 32 * (q * S + s) + lane for both sweeps; the packed rows of the backward
// sweep are stored in ITS processing order (patches and steps descending), so both kernels walk their arrays upwards.
__global__ void __launch_bounds__(256) line_build_kernel(const LineGeom G, const int32_t* __restrict__ rank_of, const int32_t* __restrict__ pJ, const int32_t* __restrict__ pK,
                                                         const int32_t* __restrict__ pT, const int32_t* __restrict__ start, const int32_t* __restrict__ positions,
                                                         const int32_t* __restrict__ diag_pos, uint32_t* __restrict__ pack0, uint32_t* __restrict__ pack1,
                                                         int32_t* __restrict__ eidx0, int32_t* __restrict__ eidx1, int* fail) {
    const int q = blockIdx.x;
    const int J = pJ[q], K = pK[q], T = pT[q];
    const int W = G.W;
    for (int idx = threadIdx.x; idx < G.S * 32; idx += blockDim.x) {
        const int s = idx >> 5, lane = idx & 31;
        const int a = lane % G.A, b = lane / G.A;
        const int j = J * G.A + a, k = K * G.B + b, i = s - a - b;
        const bool valid = b < G.B && j < G.ny && k < G.nz && i >= 0 && i < G.nx;
        const int row = valid ? (k * G.ny + j) * G.nx + i : -1;
        for (int w = 0; w < 2; ++w) {
            const size_t gk = w == 0 ? (size_t)q * G.S + s : (size_t)(G.npatch - 1 - q) * G.S + (G.S - 1 - s);
            uint32_t* pk = (w ? pack1 : pack0) + (gk * 32 + lane) * LINE_WORDS;
            int32_t* ei = (w ? eidx1 : eidx0) + (gk * 32 + lane) * 4;
            for (int t = 0; t < LINE_MAX_W; ++t) { pk[t] = (uint32_t)LINE_NONE; pk[3 + t] = 0u; ei[t] = -1; }
            pk[6] = __float_as_uint(1.0f);
            pk[7] = (uint32_t)row;
            ei[3] = -1;
            if (row < 0) continue;
            const int dg = diag_pos[row];
            ei[3] = dg;
            const int cnt = w == 0 ? dg - start[row] : start[row + 1] - 1 - dg;
            if (cnt > W) { atomicOr(fail, 1); continue; }
            for (int t = 0; t < cnt; ++t) {
                const int e = w == 0 ? start[row] + t : start[row + 1] - 1 - t;     // ascending / descending columns (H:1685, H:1704)
                int qc, sc, lc;
                line_where(G, rank_of, positions[e], &qc, &sc, &lc);
                int code;
                if (qc == q) {
                    const int age = w == 0 ? s - sc : sc - s;
                    if (age < 1 || age > LINE_RING - 1) atomicOr(fail, 2);
                    code = LINE_LOCAL - ((sc * 32 + lc) & (LINE_RING * 32 - 1));
                } else {
                    const bool ok = w == 0 ? (qc < q && pT[qc] + sc < T + s) : (qc > q && pT[qc] + sc > T + s);
                    if (!ok) atomicOr(fail, 4);
                    code = (qc * G.S + sc) * 32 + lc;
                }
                pk[t] = (uint32_t)code;
                ei[t] = e;
            }
        }
    }
}

// coefficients and diagonals of the packed rows from the matrix values (SGS) or the factor (IC(0), ILU(0))
__global__ void line_gather_kernel(const float* __restrict__ values, const int32_t* __restrict__ eidx, uint32_t* __restrict__ pack, const long long nrows4,
                                   const bool unit_diagonal) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nrows4) return;
    const int slot = (int)(t & 3);
    const int k = eidx[t];
    float v = k >= 0 ? values[k] : (slot == 3 ? 1.0f : 0.0f);
    if (slot == 3 && unit_diagonal) v = 1.0f;                  // the L factor of ILU(0): implied ones (x / 1.0f == x exactly)
    pack[(t >> 2) * LINE_WORDS + 3 + slot] = __float_as_uint(v);
}

__device__ __forceinline__ uint32_t lsmem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lsmem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void lbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lsmem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LDONE_%=;\n"
        "bra LWAIT_%=;\n"
        "LDONE_%=:\n"
        "}\n" ::"r"(lsmem(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void lbulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(lsmem(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(lsmem(bar)) : "memory");
}
__device__ __forceinline__ void lbulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lprefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void lcp_async4(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(lsmem(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void lcp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void lcp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// diagnostics (smm_debug_line_stats): steps that found an out-of-patch operand missing, their polls, polls at the gate
__device__ unsigned long long g_line_stats[4];
static_assert(LINE_MARGIN <= LINE_NW && LINE_BLOCK % LINE_NW == 0 && LINE_NW == 4, "the warps' turns");

constexpr int LINE_BLOCK_WORDS = LINE_BLOCK * 32 * LINE_WORDS;                    // 2048 words = 8 KB
constexpr size_t LINE_SMEM = (size_t)LINE_NBLK * LINE_BLOCK_WORDS * 4 + LINE_RING * 128 + 64;

// a row on its way through the register pipeline: packed data, right-hand side and out-of-patch operands as requested
__device__ __forceinline__ unsigned long long line_clock() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

template <int W>
struct LineRow { int c[W]; float v[W]; float d; int row; float init; unsigned int xb[W]; };

// LINE_NW warps per CTA share a patch and take turns step by step: warp w solves the steps k = w (mod LINE_NW).  A step's
// dependency chain (operands out of the ring -> three multiply-adds -> division -> result into the ring) is what the sweep
// waits for; everything else a step needs (reading the packed row, requesting its right-hand side and its out-of-patch
// operands LINE_AHEAD steps early, publishing) is done by its warp while the other warps solve the steps in between, so
// the chain sees one CTA barrier per step instead of a whole step's worth of one warp's instructions.
// CTAs claim patches in time order (forward: ascending rank, backward: descending -- the packed rows of the backward sweep
// are stored in that order).  IC0 = false: SGS sweeps.  IC0 = true: `sum -= f * x; x = sum / d` in both directions (IC(0), ILU(0)).
template <bool FORWARD, bool IC0, int W, int WID>
__device__ __forceinline__ void line_run(const LineArgs& A, const float* __restrict__ rhs, float* yperm, float* xperm, float* __restrict__ x,
                                         unsigned int* tickets, uint32_t* const blocks, float* const res, uint64_t* const full, unsigned int* const sh_q) {
    constexpr int MINE = LINE_BLOCK / LINE_NW;                                            // rows of a block this warp solves
    const int lane = threadIdx.x & 31;
    unsigned int* const abort_flag = tickets + 2;
    unsigned int* const ticket = tickets + (FORWARD ? 0 : 1);
    const float* const src = FORWARD ? yperm : xperm;                                     // operands are addressed by position
    float* const dst = FORWARD ? yperm : xperm;
    const int S = A.S, nblk = S / LINE_BLOCK;
    uint32_t gb = 0;                                                                      // blocks consumed so far: ring slot and phase of the next one
    for (unsigned int turn = 0;; ++turn) {
        if (threadIdx.x == 0) sh_q[turn & 1u] = atomicAdd(ticket, 1u);
        __syncthreads();
        const unsigned int q = sh_q[turn & 1u];
        if (q >= (unsigned int)A.npatch) break;
        const unsigned long long t_claim = A.trace ? line_clock() : 0ull;
        const uint32_t* const pk = A.pack + (size_t)q * S * (32 * LINE_WORDS);
        // position of this lane's row at processing step k: 32 * (patch * S + s) + lane with s = k (forward) / S - 1 - k (backward)
        const size_t pos0 = ((size_t)(FORWARD ? (int)q : A.npatch - 1 - (int)q) * S + (FORWARD ? 0 : S - 1)) * 32 + lane;
        auto pos_of = [&](int k) { return FORWARD ? pos0 + (size_t)k * 32 : pos0 - (size_t)k * 32; };
        auto issue = [&](int b) {                                                         // thread 0: block b of this patch -> ring
            const uint32_t n = gb + (uint32_t)b, slot = n % LINE_NBLK;
            lbar_expect_tx(&full[slot], LINE_BLOCK_WORDS * 4);
            lbulk_load(blocks + slot * LINE_BLOCK_WORDS, pk + (size_t)b * LINE_BLOCK_WORDS, LINE_BLOCK_WORDS * 4, &full[slot]);
        };
        auto wait_block = [&](int b) {
            const uint32_t n = gb + (uint32_t)b;
            lbar_wait(&full[n % LINE_NBLK], (n / LINE_NBLK) & 1u);
        };
        auto block_ptr = [&](int b) { return blocks + ((gb + (uint32_t)b) % LINE_NBLK) * LINE_BLOCK_WORDS + lane * LINE_WORDS; };
        // a row's packed data out of shared memory, and its requests to global memory: the right-hand side (backward: the row's
        // own forward result, at the row's position) and, unless the gate below polls for them, the out-of-patch operands
        auto load_row = [&](const uint32_t* line, const int k, const bool with_operands) {
            const uint4 lo = *reinterpret_cast<const uint4*>(line), hi = *reinterpret_cast<const uint4*>(line + 4);
            const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            LineRow<W> r;
#pragma unroll
            for (int j = 0; j < W; ++j) { r.c[j] = (int)w[j]; r.v[j] = __uint_as_float(w[3 + j]); r.xb[j] = SENTINEL; }
            r.d = __uint_as_float(w[6]);
            r.row = (int)w[7];
            r.init = 0.0f;
            if (r.row >= 0) {
                if (FORWARD) {
                    r.init = __ldg(rhs + r.row);                                          // H:1683 / H:1807
                    if (r.row + 16 < A.rows) lprefetch_l1(rhs + r.row + 16);              // the lane walks its line: the sector after the next one
                } else {
                    r.init = __ldg(yperm + pos_of(k));                                    // H:1710, H:1823 (written by the forward launch)
                }
            }
            if (with_operands) {
#pragma unroll
                for (int j = 0; j < W; ++j) if (r.c[j] >= 0) r.xb[j] = peek(src + r.c[j]);
            }
            return r;
        };
        int issued = 0;
        for (; issued < LINE_NBLK && issued < nblk; ++issued) if (threadIdx.x == 0) issue(issued);
        if (threadIdx.x == 0 && LINE_NBLK < nblk) lbulk_prefetch_l2(pk + (size_t)LINE_NBLK * LINE_BLOCK_WORDS, LINE_BLOCK_WORDS * 4);
        wait_block(0);
        LineRow<W> R[MINE];
#pragma unroll
        for (int m = 0; m < MINE; ++m) R[m] = load_row(block_ptr(0) + (WID + m * LINE_NW) * 32 * LINE_WORDS, WID + m * LINE_NW, false);
        // Gate: the out-of-patch operands of the first LINE_BLOCK + LINE_MARGIN steps are polled here, all at once -- this is where
        // a patch waits for its turn.  Once they are there the producers are LINE_MARGIN steps further than the requests of the steps
        // that follow (issued LINE_AHEAD steps early) need them to be; without the margin every one of the first requests comes back
        // empty and costs its step an L2 round trip.
        bool aborted = false;
        {
            int gc[W];                                                                    // operand sources of this warp's margin row (warps 0 .. LINE_MARGIN-1)
            unsigned int gx[W];
#pragma unroll
            for (int j = 0; j < W; ++j) { gc[j] = LINE_NONE; gx[j] = 0u; }
            if (WID < LINE_MARGIN && nblk > 1) {
                wait_block(1);
                const uint32_t* line = block_ptr(1) + WID * 32 * LINE_WORDS;
#pragma unroll
                for (int j = 0; j < W; ++j) { gc[j] = (int)line[j]; gx[j] = SENTINEL; }
            }
            unsigned int polls = 0;
            // first the cheap wait: one warp asks for the operands of the LAST of these steps (its producers publish in step
            // order) and sleeps between polls, the other warps sit in the barrier -- a patch may wait here for most of the sweep,
            // next to CTAs that are solving, and must not take their issue slots
            if (WID == LINE_MARGIN - 1) {
                for (;;) {
                    bool miss = false;
#pragma unroll
                    for (int j = 0; j < W; ++j) if (gc[j] >= 0 && gx[j] == SENTINEL) { gx[j] = peek(src + gc[j]); miss |= gx[j] == SENTINEL; }
                    if (!__any_sync(0xffffffffu, miss)) break;
                    if (!poll_pause(&polls, abort_flag, 200u, 400u)) break;             // (the full check below notices an abort)
                }
            }
            __syncthreads();
            for (;;) {
                bool miss = false;
#pragma unroll
                for (int m = 0; m < MINE; ++m)
#pragma unroll
                    for (int j = 0; j < W; ++j)
                        if (R[m].c[j] >= 0 && R[m].xb[j] == SENTINEL) { R[m].xb[j] = peek(src + R[m].c[j]); miss |= R[m].xb[j] == SENTINEL; }
#pragma unroll
                for (int j = 0; j < W; ++j) if (gc[j] >= 0 && gx[j] == SENTINEL) { gx[j] = peek(src + gc[j]); miss |= gx[j] == SENTINEL; }
                if (!__syncthreads_or(miss)) break;
                if (!poll_pause(&polls, abort_flag, A.sleep_first, A.sleep_later)) aborted = true;
                if (__syncthreads_or(aborted)) { aborted = true; break; }
            }
        }
        if (A.trace && threadIdx.x == 0) { A.trace[4ull * q] = t_claim; A.trace[4ull * q + 1] = line_clock(); }
        int blk = 0;
        unsigned int n_miss = 0, n_polls = 0;                                             // diagnostics: steps of this warp that waited, their polls
        for (; blk < nblk && !aborted; ++blk) {
            // this block's rows are in registers; the next block's are read while it is solved, and the slot this block came from is free
            // (every warp read its rows out of it before the step barriers of the previous block)
            const bool more = blk + 1 < nblk;
            if (more) wait_block(blk + 1);
            if (blk + LINE_NBLK < nblk) {
                if (threadIdx.x == 0) {
                    issue(blk + LINE_NBLK);
                    if (blk + LINE_NBLK + 1 < nblk) lbulk_prefetch_l2(pk + (size_t)(blk + LINE_NBLK + 1) * LINE_BLOCK_WORDS, LINE_BLOCK_WORDS * 4);
                }
                ++issued;
            }
            const uint32_t* const nxt = block_ptr(blk + 1);
#pragma unroll
            for (int u = 0; u < LINE_BLOCK; ++u) {
                if ((u & (LINE_NW - 1)) == WID) {
                    const int k = blk * LINE_BLOCK + u;
                    const LineRow<W> r = R[u / LINE_NW];
                    unsigned int xb[W];
                    bool miss = false;
#pragma unroll
                    for (int j = 0; j < W; ++j) {
                        xb[j] = r.c[j] >= 0 ? r.xb[j] : __float_as_uint(res[(LINE_LOCAL - r.c[j]) & (LINE_RING * 32 - 1)]);
                        miss |= r.c[j] >= 0 && xb[j] == SENTINEL;
                    }
                    if (miss) {                                                           // the producer patch is not far enough ahead: ask L2 until it is
                        unsigned int polls = 0;
                        ++n_miss;
                        for (;;) {
                            ++n_polls;
                            miss = false;
#pragma unroll
                            for (int j = 0; j < W; ++j) if (r.c[j] >= 0 && xb[j] == SENTINEL) { xb[j] = peek(src + r.c[j]); miss |= xb[j] == SENTINEL; }
                            if (!miss) break;
                            if (!poll_pause(&polls, abort_flag, A.sleep_first, A.sleep_later)) { aborted = true; break; }
                        }
                    }
                    float acc = (FORWARD || IC0) ? r.init : 0.0f;
#pragma unroll
                    for (int j = 0; j < W; ++j) {
                        if (r.c[j] != LINE_NONE) {
                            const float xv = __uint_as_float(xb[j]);
                            // forward: _smm_fma(-value, x[col], lhs), cols ascending (H:1685); backward: _smm_fma(value, x[col], lhs), cols
                            // descending (H:1704); IC0: sum -= ic0[j] * x[col] (H:1813, H:1829)
                            acc = (FORWARD || IC0) ? __fsub_rn(acc, __fmul_rn(r.v[j], xv)) : __fadd_rn(__fmul_rn(r.v[j], xv), acc);
                        }
                    }
                    const float o = (FORWARD || IC0) ? __fdiv_rn(acc, r.d)                // H:1694 / H:1818, H:1834
                                                     : __fsub_rn(r.init, __fdiv_rn(acc, r.d));   // H:1710
                    if (r.row >= 0) {
                        // H:1691-1693 `abs(diagonal) < 1e-5` with the float promoted to double: true exactly for the floats <= 1e-5f
                        if (FORWARD && !IC0 && fabsf(r.d) <= 1e-5f) atomicOr(tickets + 3, 1u);   // (reported, not fatal here)
                        res[(FORWARD ? u : LINE_BLOCK - 1 - u) * 32 + lane] = o;          // S is a multiple of the ring: step % ring = index in the block
                    }
                    __syncthreads();                                                      // the step's results are in the ring: the next warp may go
                    if (r.row >= 0) {
                        publish(dst + pos_of(k), o);
                        if (!FORWARD) x[r.row] = o;                                       // the caller's vector, natural order
                    }
                    if (more) R[u / LINE_NW] = load_row(nxt + u * 32 * LINE_WORDS, k + LINE_BLOCK, true);
                } else {
                    __syncthreads();
                }
            }
            if (__syncthreads_or(aborted)) { aborted = true; break; }
        }
        if (aborted) {                                                                    // leave only when no bulk copy into this CTA's memory is in flight
            for (int b = blk + 1; b < issued && b < nblk; ++b) wait_block(b);
            return;
        }
        gb += (uint32_t)nblk;
        if (A.trace && threadIdx.x == 0) {
            unsigned int sm;
            asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
            A.trace[4ull * q + 2] = line_clock(); A.trace[4ull * q + 3] = sm;
        }
        if (n_miss) { atomicAdd(&g_line_stats[0], (unsigned long long)n_miss); atomicAdd(&g_line_stats[1], (unsigned long long)n_polls); }
    }
}

template <bool FORWARD, bool IC0, int W>
__global__ void __launch_bounds__(LINE_NW * 32) sgs_line_kernel(const LineArgs A, const float* __restrict__ rhs, float* yperm, float* xperm, float* __restrict__ x,
                                                                unsigned int* tickets, const SolveState* st) {
    if (st != nullptr && st->done) return;
    extern __shared__ __align__(128) uint32_t line_sm[];
    uint32_t* const blocks = line_sm;                                                     // [LINE_NBLK][LINE_BLOCK][32][LINE_WORDS]
    float* const res = reinterpret_cast<float*>(blocks + LINE_NBLK * LINE_BLOCK_WORDS);   // [LINE_RING][32] results of the last steps
    uint64_t* const full = reinterpret_cast<uint64_t*>(res + LINE_RING * 32);             // [LINE_NBLK]
    unsigned int* const sh_q = reinterpret_cast<unsigned int*>(full + LINE_NBLK);         // [2] the patch claimed for this / the next turn
    if (threadIdx.x == 0) {
        for (int i = 0; i < LINE_NBLK; ++i) lbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    switch (threadIdx.x >> 5) {
        case 0: line_run<FORWARD, IC0, W, 0>(A, rhs, yperm, xperm, x, tickets, blocks, res, full, sh_q); break;
        case 1: line_run<FORWARD, IC0, W, 1>(A, rhs, yperm, xperm, x, tickets, blocks, res, full, sh_q); break;
        case 2: line_run<FORWARD, IC0, W, 2>(A, rhs, yperm, xperm, x, tickets, blocks, res, full, sh_q); break;
        default: line_run<FORWARD, IC0, W, 3>(A, rhs, yperm, xperm, x, tickets, blocks, res, full, sh_q); break;
    }
}

template <class T>
int line_upload(const std::vector<T>& h, T** d) {
    SMM_CUDA(cudaMalloc(d, sizeof(T) * (h.empty() ? 1 : h.size())));
    if (!h.empty()) SMM_CUDA(cudaMemcpy(*d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
    return SMM_OK;
}

void line_free(smm_precond* p) {
    for (int w = 0; w < 2; ++w) { cudaFree(p->line_pack[w]); cudaFree(p->line_eidx[w]); p->line_pack[w] = nullptr; p->line_eidx[w] = nullptr; }
    cudaFree(p->yperm); cudaFree(p->xperm);
    p->yperm = p->xperm = nullptr;
    p->threads_fwd = p->threads_bwd = 0;
    p->lined = false;
}

template <bool FORWARD, bool IC0, int W>
int line_launch_one(const smm_precond* p, const LineArgs& A, const float* rhs_dev, float* x_dev, SolveState* state, cudaStream_t s) {
    static int per_sm_dev[SMM_MAX_DEVICES] = {0};
    int per_sm;
    {
        std::lock_guard<std::mutex> lk(g_smm_attr_mu);
        int& r = per_sm_dev[p->m->device % SMM_MAX_DEVICES];
        if (!r) {
            SMM_CUDA(cudaFuncSetAttribute(sgs_line_kernel<FORWARD, IC0, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LINE_SMEM));
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, sgs_line_kernel<FORWARD, IC0, W>, LINE_NW * 32, LINE_SMEM) != cudaSuccess || r < 1) r = 1;
        }
        per_sm = r;
    }
    long long grid = (long long)p->m->sm_count * per_sm;
    if (grid > A.npatch) grid = A.npatch;
    sgs_line_kernel<FORWARD, IC0, W><<<(unsigned)grid, LINE_NW * 32, LINE_SMEM, s>>>(A, rhs_dev, p->yperm, p->xperm, x_dev, p->tickets, state);
    return SMM_OK;
}

template <int W>
int line_launch_w(const smm_precond* p, const LineArgs& F, const LineArgs& B, const float* rhs_dev, float* x_dev, SolveState* state, cudaStream_t s) {
    if (p->kind != 0) {
        SMM_TRY((line_launch_one<true, true, W>(p, F, rhs_dev, x_dev, state, s)));
        SMM_TRY((line_launch_one<false, true, W>(p, B, rhs_dev, x_dev, state, s)));
    } else {
        SMM_TRY((line_launch_one<true, false, W>(p, F, rhs_dev, x_dev, state, s)));
        SMM_TRY((line_launch_one<false, false, W>(p, B, rhs_dev, x_dev, state, s)));
    }
    return SMM_OK;
}

}  // namespace

// Proposal, layout and verification.  Returns false (and leaves the handle as it was) when the matrix is not of the kind,
// when the proposal does not verify, or when the device is out of memory -- the tile / row schedules need less.
bool smm_sgs_lines_build(smm_precond* p, int rows, const std::vector<int32_t>& start, const std::vector<int32_t>& pos) {
    {   // opt-in while the tile schedule is the faster one (profiles/r02_sgs_lines.txt): SMM_B200_SGS_LINES=1
        const char* e = getenv("SMM_B200_SGS_LINES");
        if (!e || atoi(e) == 0) return false;
    }
    if (rows < 64 || !p->diag_pos) return false;
    long long nx = 0, ny = 0, nz = 1;
    if (!smm_sgs_detect_grid(rows, start, pos, &nx, &ny, &nz)) return false;
    LineGeom G;
    G.nx = (int)nx; G.ny = (int)ny; G.nz = (int)nz;
    G.A = nz > 1 ? 8 : 32;
    G.B = nz > 1 ? 4 : 1;
    G.NJ = (int)((ny + G.A - 1) / G.A);
    G.NK = (int)((nz + G.B - 1) / G.B);
    G.S = (G.nx + G.A - 1 + G.B - 1 + LINE_BLOCK - 1) / LINE_BLOCK * LINE_BLOCK;
    G.W = nz > 1 ? 3 : 2;                                      // three (two) distinct offsets: at most that many couplings per triangle
    const long long npatch = (long long)G.NJ * G.NK;
    const long long nsteps = npatch * G.S;
    if (nsteps * 32 >= (1ll << 31) || npatch >= (1ll << 24)) return false;       // positions are int32
    G.npatch = (int)npatch;
    // patches in time order: offset T = A J + B K (the hyperplane index of the patch's first row); rank = claim order
    std::vector<int32_t> by_rank((size_t)npatch), rank_of((size_t)npatch), pJ((size_t)npatch), pK((size_t)npatch), pT((size_t)npatch);
    std::iota(by_rank.begin(), by_rank.end(), 0);
    auto time_of = [&](int id) { return G.A * (id % G.NJ) + G.B * (id / G.NJ); };
    std::stable_sort(by_rank.begin(), by_rank.end(), [&](int a, int b) { return time_of(a) < time_of(b); });
    for (int q = 0; q < (int)npatch; ++q) {
        const int id = by_rank[(size_t)q];
        rank_of[(size_t)id] = q; pJ[(size_t)q] = id % G.NJ; pK[(size_t)q] = id / G.NJ; pT[(size_t)q] = time_of(id);
    }
    int32_t *d_rank = nullptr, *d_J = nullptr, *d_K = nullptr, *d_T = nullptr;
    int* d_fail = nullptr;
    bool ok = line_upload(rank_of, &d_rank) == SMM_OK && line_upload(pJ, &d_J) == SMM_OK && line_upload(pK, &d_K) == SMM_OK && line_upload(pT, &d_T) == SMM_OK &&
              cudaMalloc(&d_fail, sizeof(int)) == cudaSuccess && cudaMemset(d_fail, 0, sizeof(int)) == cudaSuccess;
    for (int w = 0; w < 2 && ok; ++w)
        ok = cudaMalloc(&p->line_pack[w], sizeof(uint32_t) * (size_t)nsteps * 32 * LINE_WORDS) == cudaSuccess &&
             cudaMalloc(&p->line_eidx[w], sizeof(int32_t) * (size_t)nsteps * 32 * 4) == cudaSuccess;
    ok = ok && cudaMalloc(&p->yperm, sizeof(float) * (size_t)nsteps * 32) == cudaSuccess && cudaMalloc(&p->xperm, sizeof(float) * (size_t)nsteps * 32) == cudaSuccess;
    int fail = -1;
    if (ok) {
        line_build_kernel<<<(unsigned)npatch, 256>>>(G, d_rank, d_J, d_K, d_T, p->m->start, p->m->positions, p->diag_pos, p->line_pack[0], p->line_pack[1],
                                                      p->line_eidx[0], p->line_eidx[1], d_fail);
        SMM_COUNT_LAUNCH(1);
        ok = cudaMemcpy(&fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess;
    }
    cudaFree(d_rank); cudaFree(d_J); cudaFree(d_K); cudaFree(d_T); cudaFree(d_fail);
    if (!ok || fail != 0) {
        cudaGetLastError();
        line_free(p);
        return false;
    }
    p->lined = true;
    p->line_w = G.W;
    p->line_steps = G.S;
    p->line_patches = (int)npatch;
    p->line_levels = G.NJ + G.NK - 1;
    p->threads_fwd = p->threads_bwd = nsteps * 32;
    return true;
}

// refresh the packed coefficients after the matrix values changed (or once after the build)
int smm_sgs_lines_gather(const smm_precond* p, cudaStream_t s) {
    const long long n = (long long)p->line_patches * p->line_steps * 32 * 4;
    for (int w = 0; w < 2; ++w)
        line_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p->kind != 0 ? p->factor : p->m->values, p->line_eidx[w], p->line_pack[w], n, p->kind == 2 && w == 0);
    SMM_COUNT_LAUNCH(2);
    return SMM_OK;
}

extern "C" int smm_debug_line_stats(unsigned long long* out4) {
    SMM_CUDA(cudaDeviceSynchronize());
    SMM_CUDA(cudaMemcpyFromSymbol(out4, g_line_stats, sizeof(unsigned long long) * 4));
    const unsigned long long zero[4] = {0, 0, 0, 0};
    SMM_CUDA(cudaMemcpyToSymbol(g_line_stats, zero, sizeof(zero)));
    return SMM_OK;
}

int smm_sgs_lines_launch(const smm_precond* p, const float* rhs_dev, float* x_dev, SolveState* state, unsigned int sleep_first, unsigned int sleep_later, cudaStream_t s) {
    // debug: SMM_B200_SGS_TRACE=<file> records per-patch timestamps of the forward sweep of every apply (last one kept)
    static const char* trace_path = getenv("SMM_B200_SGS_TRACE");
    static unsigned long long* trace = nullptr;
    static long long trace_cap = 0;
    if (trace_path && trace_cap < p->line_patches) {
        cudaFree(trace);
        trace = nullptr;
        SMM_CUDA(cudaMalloc(&trace, sizeof(unsigned long long) * 4 * (size_t)p->line_patches));
        trace_cap = p->line_patches;
    }
    LineArgs F{p->line_pack[0], p->line_patches, p->line_steps, p->rows, sleep_first, sleep_later, trace};
    LineArgs B{p->line_pack[1], p->line_patches, p->line_steps, p->rows, sleep_first, sleep_later, nullptr};
    SMM_TRY(p->line_w == 2 ? line_launch_w<2>(p, F, B, rhs_dev, x_dev, state, s) : line_launch_w<3>(p, F, B, rhs_dev, x_dev, state, s));
    if (trace) {
        std::vector<unsigned long long> h(4 * (size_t)p->line_patches);
        SMM_CUDA(cudaStreamSynchronize(s));
        SMM_CUDA(cudaMemcpy(h.data(), trace, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost));
        if (FILE* f = fopen(trace_path, "wb")) { fwrite(h.data(), sizeof(unsigned long long), h.size(), f); fclose(f); }
    }
    return SMM_OK;
}
