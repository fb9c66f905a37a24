// sgs_tiles.cu -- tile-level schedule for the triangular sweeps of sgs.cu (SGS, IC(0), ILU(0) apply).
//
// The row-level schedule of sgs.cu pays one producer->L2->consumer hand-off (340-530 ns, tools/hop_latency.cu) per
// dependency level, and a 7-point stencil on an N^3 grid has 3N-2 of them.  Here rows are grouped into TILES of up to
// 64 rows that one warp solves in shared memory, so that only the hand-offs BETWEEN tiles go through L2: a 4x4x4 tile
// has 10 internal levels (about 120 ns each when the warps of an SM solve together) and the tile graph of the same
// grid has 3N/4-2 levels.
//
//   * Any grouping is legal as long as the tile graph stays acyclic; the per-row arithmetic (operand order, two
//     roundings per term, one division) is that of the row-level kernel, so the result has the same bits for any
//     grouping.  build() takes a PROPOSAL (geometric tiles when the column offsets of the matrix are those of a
//     natural-order 2D / 3D grid stencil) and VERIFIES it generically: tiles of <= 64 rows, <= 4 stored operands
//     per row and sweep, <= 3 rows of the same tile consuming a row, an acyclic tile graph (Kahn).  Anything else falls
//     back to the row-level schedule.
//   * Layout: tiles sorted by tile level; position = 64 * tile + index, rows inside a tile sorted by internal level.
//     Intermediate vectors are stored by position (as in sgs.cu), entries as [tile][operand slot][64].
//   * Kernel: one warp per tile, lane l holds rows l and l + 32 of the tile in registers (loaded, with the next
//     tile's, ahead of time).  Operands from other tiles are polled by position -- all of the tile's at once, and again
//     (every lane, everything it still misses) until all have been published; then step s solves the rows of internal
//     level s, operands inside the tile coming from shared memory (pushed there by their producer).  A tile publishes
//     its 64 results together after its last step.  The step loop is kept to about 30 instructions: the warps of an SM
//     solve at the same time and share its issue slots.
//   * Schedule: tiles are grouped into CHAINS (for a grid: the tiles of one (J, K) column, marched along i) that ONE warp
//     solves from end to end; chains are handed out in dependency order by an atomic ticket, so every awaited producer
//     belongs to a chain that a running warp already owns (no deadlock).  A warp that marches along a chain finds the
//     operands of its own chain already published, neighbouring chains settle one tile plus a hand-off apart, and no
//     tile ever waits for a free warp (with tiles handed out level by level the middle levels of a 256^3 grid hold more
//     tiles than there are resident warps).  The proposal's chains are verified generically (dependencies inside a chain
//     point backwards, the chain graph is acyclic); anything else falls back to single-tile chains in tile-level order.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <atomic>
#include <future>
#include <thread>
#include <vector>

#include <cooperative_groups.h>

#include "sgs_internal.cuh"

namespace cg = cooperative_groups;

namespace {

#ifndef SMM_TILE_MIN_CTAS
#define SMM_TILE_MIN_CTAS 4      // resident 4-warp CTAs per SM the register allocation must allow
#endif
// Cluster schedule: a BLOCK of CLUSTER_CHAINS neighbouring chains is solved by one thread-block cluster (8 CTAs x 4 warps,
// one warp per chain); operands that cross from one chain of the block to another travel through distributed shared memory.
constexpr int CLUSTER_CTAS = 8;
constexpr int CLUSTER_WARPS = 4;
constexpr int CLUSTER_CHAINS = CLUSTER_CTAS * CLUSTER_WARPS;
constexpr int INBOX_SLOTS = 32;                      // operands per tile that may arrive from other chains of the block
constexpr int CLUSTER_MAX_CHAIN = 160;               // longest chain the inbox is sized for (4 warps x 160 tiles x 128 B = 80 KB per CTA)
// operand codes in ecol for the cluster schedule (>= 0: position in the intermediate vector, polled through L2)
constexpr int E_NONE = -1;                           // no operand (0 * 0)
constexpr int E_LOCAL = -2;                          // pushed into the warp's own staging (same tile, or the previous tile of the chain)
constexpr int E_INBOX = -16;                         // -16 - slot: arrives in the warp's inbox [tile of the chain][slot]
// push2 word of a producer row: [7:0] staging slot of the consumer in the NEXT tile of the chain (0xFF: none),
// [18:8] and [29:19]: (target warp of the cluster block << 5 | inbox slot) (0x7FF: none)
constexpr uint32_t PUSH2_NONE = 0xFFu | (0x7FFu << 8) | (0x7FFu << 19);

struct TileArgs {
    const uint8_t* nsteps;      // [tiles]
    const uint8_t* row_step;    // [tiles * 64] the step in which the row is solved (255: padding)
    const uint32_t* push;       // [tiles * 64] up to three operand slots (row x 4 + operand) of rows of the SAME tile that consume this row
    const int32_t* order;       // [tiles * 64] row or -1
    const int32_t* ypos;        // backward: position of the row in yperm
    const int32_t* ecol;        // [tiles][width][64] operand position or -1
    const float* eval;
    const float* dval;          // [tiles * 64]
    long long ntiles;
    long long nchains;          // ntiles / chain_len
    int chain_len;              // tiles per chain (consecutive tile indices)
    int nblocks;                // cluster schedule: blocks of CLUSTER_CHAINS chains (0: not laid out for clusters)
    const uint32_t* push2;      // cluster schedule: [tiles * 64] pushes that leave the tile (PUSH2_*)
    int width;
    unsigned int sleep_first, sleep_later;
    unsigned long long* trace;  // debug (SMM_B200_SGS_TRACE): [tiles][4] globaltimer at claim / first step / solved, and the SM
};

struct TileHead { int row[2]; int yp[2]; };
struct TileBody {
    int nsteps, step[2];
    unsigned int push[2];
    float d[2], init[2];
    int c[2][TILE_MAX_W];
    float v[2][TILE_MAX_W];
};

__device__ __forceinline__ unsigned long long tile_clock() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

template <bool FORWARD, bool IC0, int TILE_WARPS>
__global__ void __launch_bounds__(TILE_WARPS * 32, SMM_TILE_MIN_CTAS * 4 / TILE_WARPS) sgs_tile_kernel(const TileArgs A, const float* __restrict__ rhs, float* yperm, float* xperm,
                                                                  float* __restrict__ x, unsigned int* tickets, const SolveState* st) {
    if (st != nullptr && st->done) return;
    // operand staging: slot 4 * r + e holds operand e of row r of the warp's current tile.  Operands from other tiles
    // are put there by the row's own lane once they have been published; operands from the same tile are PUSHED
    // there by the lane that solves them, so a row needs a single 128-bit shared-memory load when its step comes
    __shared__ __align__(16) float stage[TILE_WARPS][TILE * TILE_MAX_W];
    unsigned int* abort_flag = tickets + 2;
    unsigned int* ticket = tickets + (FORWARD ? 0 : 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* src = FORWARD ? yperm : xperm;
    float* dst = FORWARD ? yperm : xperm;
    float* mine = stage[warp];

    auto load_head = [&](const long long tile) {
        TileHead h;
        h.row[0] = h.row[1] = -1;
        h.yp[0] = h.yp[1] = 0;
        if (tile < A.ntiles) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                h.row[k] = A.order[tile * TILE + k * 32 + lane];
                if (!FORWARD) h.yp[k] = A.ypos[tile * TILE + k * 32 + lane];
            }
        }
        return h;
    };
    auto load_body = [&](const long long tile, const TileHead& h) {
        TileBody b;
        const bool live = tile < A.ntiles;
        b.nsteps = live ? A.nsteps[tile] : 0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const long long at = tile * TILE + k * 32 + lane;
            b.step[k] = live ? A.row_step[at] : 255;
            b.push[k] = live ? A.push[at] : 0u;
            b.d[k] = live ? A.dval[at] : 1.0f;
#pragma unroll
            for (int e = 0; e < TILE_MAX_W; ++e) {
                const bool in = live && e < A.width;
                b.c[k][e] = in ? A.ecol[(tile * A.width + e) * TILE + k * 32 + lane] : -1;
                b.v[k][e] = in ? A.eval[(tile * A.width + e) * TILE + k * 32 + lane] : 0.0f;
            }
            // forward: the right-hand side (H:1683 / H:1807); backward: the row's own forward result (previous launch)
            b.init[k] = h.row[k] >= 0 ? (FORWARD ? rhs[h.row[k]] : yperm[h.yp[k]]) : 0.0f;
        }
        return b;
    };
    auto solve_tile = [&](const long long tile_ll, const TileHead& h, const TileBody& b) {
        const int tile = (int)tile_ll;
        if (A.trace && lane == 0) A.trace[4ll * tile] = tile_clock();
        // Operands from other tiles (both rows of this lane) are requested up front, all at once, and whatever has not been
        // published yet is requested again until everything is there
        unsigned int pend = 0u;                                // bit 4 k + e: operand e of row k is still awaited
        unsigned int first[2 * TILE_MAX_W];
#pragma unroll
        for (int q = 0; q < 2 * TILE_MAX_W; ++q) {                                // all requests first
            const int c = b.c[q / TILE_MAX_W][q % TILE_MAX_W];
            first[q] = (c >= 0 && (c >> 6) != tile) ? peek(src + c) : 0u;         // no operand: 0 * 0 leaves the sum unchanged
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
#pragma unroll
            for (int e = 0; e < TILE_MAX_W; ++e) if (first[4 * k + e] == SENTINEL) pend |= 1u << (4 * k + e);
            // own-tile slots are overwritten by their producer before the row's step
            *reinterpret_cast<float4*>(mine + 4 * (k * 32 + lane)) = make_float4(__uint_as_float(first[4 * k]), __uint_as_float(first[4 * k + 1]),
                                                                                   __uint_as_float(first[4 * k + 2]), __uint_as_float(first[4 * k + 3]));
            if (FORWARD && !IC0 && h.row[k] >= 0 && fabsf(b.d[k]) < 1e-5) atomicOr(tickets + 3, 1u);   // H:1691-1693 (reported, not fatal here)
        }
        // Everything the tile needs from other tiles must be there before its first step: the step loop below stays free of
        // any waiting logic (every instruction in it is paid ten times per tile by warps that share an SM's issue slots, and a
        // wait inside it costs an L2 round trip per step once chains run close behind each other: 1.29 ms instead of 1.13 ms
        // per apply on 256^3).  Predecessors publish their rows together after their last step.
        unsigned int polls = 0;
        while (__any_sync(0xFFFFFFFFu, pend != 0u)) {
            if (!poll_pause(&polls, abort_flag, A.sleep_first, A.sleep_later)) pend = 0u;
            unsigned int bits[2 * TILE_MAX_W];
#pragma unroll
            for (int q = 0; q < 2 * TILE_MAX_W; ++q)                              // all requests first: one L2 round trip per round
                bits[q] = (pend >> q) & 1u ? peek(src + b.c[q / TILE_MAX_W][q % TILE_MAX_W]) : SENTINEL;
#pragma unroll
            for (int q = 0; q < 2 * TILE_MAX_W; ++q) {
                if (bits[q] != SENTINEL) { pend &= ~(1u << q); mine[4 * ((q / TILE_MAX_W) * 32 + lane) + (q % TILE_MAX_W)] = __uint_as_float(bits[q]); }
            }
        }
        __syncwarp();
        if (A.trace && lane == 0) A.trace[4ll * tile + 1] = tile_clock();         // operands complete
        float* const out = dst + ((long long)tile * TILE + lane);
        float solved[2] = {0.0f, 0.0f};
        // one row of the lane in one step: operands out of the staging slots, the sum in operand order, the division, and
        // the result handed to the (up to three) rows of this tile that use it
        auto solve_row = [&](const int k) {
            const float4 xo = *reinterpret_cast<const float4*>(mine + 4 * (k * 32 + lane));
            // same operand order and roundings as the row-level kernel (H:1685, H:1704, H:1813, H:1829); only the
            // additions and the division are on the step's dependent chain
            const float p0 = __fmul_rn(b.v[k][0], xo.x), p1 = __fmul_rn(b.v[k][1], xo.y), p2 = __fmul_rn(b.v[k][2], xo.z), p3 = __fmul_rn(b.v[k][3], xo.w);
            float acc = (FORWARD || IC0) ? b.init[k] : 0.0f;                      // H:1683 / T sum = x[row], H:1823 / H:1702
            if (FORWARD || IC0) { acc = __fsub_rn(acc, p0); acc = __fsub_rn(acc, p1); acc = __fsub_rn(acc, p2); acc = __fsub_rn(acc, p3); }
            else { acc = __fadd_rn(p0, acc); acc = __fadd_rn(p1, acc); acc = __fadd_rn(p2, acc); acc = __fadd_rn(p3, acc); }
            const float res = (FORWARD || IC0) ? __fdiv_rn(acc, b.d[k])           // H:1694 / H:1818, H:1834
                                               : __fsub_rn(b.init[k], __fdiv_rn(acc, b.d[k]));   // H:1710
            const unsigned int pu = b.push[k];
            mine[pu & 255u] = res; mine[(pu >> 8) & 255u] = res; mine[(pu >> 16) & 255u] = res;
            solved[k] = res;
        };
        // rows l (k = 0) belong to the early steps and rows l + 32 (k = 1) to the late ones (a tile's rows are sorted by
        // step): three loops, so that a step only tests the rows that can be due
        const int first1 = __reduce_min_sync(0xFFFFFFFFu, b.step[1]);              // 255 when the tile has no such row
        const int last0 = __reduce_max_sync(0xFFFFFFFFu, b.step[0] == 255 ? -1 : b.step[0]);
        int s = 0;
        for (; s < b.nsteps && s < first1; ++s) {
            if (s == b.step[0]) solve_row(0);
            __syncwarp();
        }
        for (; s < b.nsteps && s <= last0; ++s) {
            if (s == b.step[0]) solve_row(0);
            if (s == b.step[1]) solve_row(1);
            __syncwarp();
        }
        for (; s < b.nsteps; ++s) {
            if (s == b.step[1]) solve_row(1);
            __syncwarp();
        }
        // the tile's rows are published together, after its last step: the stores stay off the step chain, and a
        // successor that waits for this tile finds all of it in one poll
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (h.row[k] >= 0) {
                publish(out + k * 32, solved[k]);
                if (!FORWARD) x[h.row[k]] = solved[k];
            }
        }
        if (A.trace && lane == 0) {
            unsigned int sm;
            asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
            A.trace[4ll * tile + 2] = tile_clock(); A.trace[4ll * tile + 3] = sm;
        }
    };

    // this warp's stream of tiles: chain after chain (claimed in dependency order), every chain from its first tile to its last
    long long chain = 0;
    int at = 0;
    auto claim = [&]() {
        unsigned int v = 0u;
        if (lane == 0) v = atomicAdd(ticket, 1u);
        return (long long)__shfl_sync(0xFFFFFFFFu, v, 0);
    };
    auto next_tile = [&]() {
        if (chain >= A.nchains) return A.ntiles;
        const long long t = chain * A.chain_len + at;
        if (++at == A.chain_len) { at = 0; chain = claim(); }
        return t;
    };
    chain = claim();
    long long t0 = next_tile(), t1 = next_tile();
    TileHead h0 = load_head(t0), h1 = load_head(t1);
    TileBody r0 = load_body(t0, h0);
    while (t0 < A.ntiles) {
        const long long t2 = next_tile();
        const TileHead h2 = load_head(t2);
        const TileBody r1 = load_body(t1, h1);
        solve_tile(t0, h0, r0);
        t0 = t1; t1 = t2; h0 = h1; h1 = h2; r0 = r1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Cluster schedule.  The trace of the chain schedule above (tools/sgs_trace.py, 256^3) shows where a sweep's time goes:
// a tile is solved in 1.2 us, but the hand-off of its results through L2 (publish -> visible -> polled) costs another
// 1.1 us, per tile along a chain and per hop from chain to chain: 126 hops x 2.5 us + 64 tiles x 2.4 us = 0.47 ms.
// Here a thread-block CLUSTER (8 CTAs x 4 warps) takes a block of 32 neighbouring chains, one warp each, and results move
// by PUSH through shared memory: to the rows of the same tile and of the chain's next tile in the warp's own staging
// (double-buffered), to the other chains of the block straight into the consumer warp's INBOX over distributed shared
// memory (mapa + st.shared::cluster; one slot per operand and tile of the chain, written once per sweep, so no flow
// control is needed).  A consumer's operand fetch is a shared-memory load, and waiting is re-loading it: cheap enough to
// wait row by row, so the chains of a block run three steps plus a DSMEM hop apart instead of a tile plus an L2 round
// trip.  Only operands from OTHER blocks are still polled through L2 (whole-tile wait, as above).  Blocks are claimed in
// dependency order by the resident clusters (no deadlock); the per-row arithmetic is unchanged (same bits).
// ---------------------------------------------------------------------------------------------------------------
struct ClusterBody {
    int nsteps, step[2];
    unsigned int push[2], push2[2];
    float d[2], init[2];
    int c[2][TILE_MAX_W];       // operand codes (E_*) or positions
    float v[2][TILE_MAX_W];
};

__device__ __forceinline__ float lds_volatile(uint32_t addr) {
    float v;
    asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void st_cluster_f32(uint32_t local_addr, unsigned int cta_rank, float v) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(cta_rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}

template <bool FORWARD, bool IC0>
__global__ void __cluster_dims__(CLUSTER_CTAS, 1, 1) __launch_bounds__(CLUSTER_WARPS * 32, SMM_TILE_MIN_CTAS * 4 / CLUSTER_WARPS)
sgs_cluster_kernel(const TileArgs A, const float* __restrict__ rhs, float* yperm, float* xperm, float* __restrict__ x, unsigned int* tickets, const SolveState* st) {
    if (st != nullptr && st->done) return;
    extern __shared__ __align__(16) float cluster_smem[];      // stage [warp][2][256] | inbox [warp][chain_len][INBOX_SLOTS]
    __shared__ unsigned int sh_block;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned int crank = cluster.block_rank();
    unsigned int* abort_flag = tickets + 2;
    unsigned int* ticket = tickets + (FORWARD ? 0 : 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* src = FORWARD ? yperm : xperm;
    float* dst = FORWARD ? yperm : xperm;
    const uint32_t stage_addr = (uint32_t)__cvta_generic_to_shared(cluster_smem) + (uint32_t)warp * 2u * 1024u;
    const uint32_t inbox_all = (uint32_t)__cvta_generic_to_shared(cluster_smem) + CLUSTER_WARPS * 2u * 1024u;   // same offset in every CTA
    const uint32_t inbox_bytes = (uint32_t)A.chain_len * INBOX_SLOTS * 4u;                                       // per warp
    float* const my_inbox = cluster_smem + CLUSTER_WARPS * 512 + (size_t)warp * A.chain_len * INBOX_SLOTS;

    auto load_head = [&](const long long tile, const bool live) {
        TileHead h;
        h.row[0] = h.row[1] = -1;
        h.yp[0] = h.yp[1] = 0;
        if (live) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                h.row[k] = A.order[tile * TILE + k * 32 + lane];
                if (!FORWARD) h.yp[k] = A.ypos[tile * TILE + k * 32 + lane];
            }
        }
        return h;
    };
    auto load_body = [&](const long long tile, const bool live, const TileHead& h) {
        ClusterBody b;
        b.nsteps = live ? A.nsteps[tile] : 0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const long long at = tile * TILE + k * 32 + lane;
            b.step[k] = live ? A.row_step[at] : 255;
            b.push[k] = live ? A.push[at] : 0u;
            b.push2[k] = live ? A.push2[at] : PUSH2_NONE;
            b.d[k] = live ? A.dval[at] : 1.0f;
#pragma unroll
            for (int e = 0; e < TILE_MAX_W; ++e) {
                const bool in = live && e < A.width;
                b.c[k][e] = in ? A.ecol[(tile * A.width + e) * TILE + k * 32 + lane] : E_NONE;
                b.v[k][e] = in ? A.eval[(tile * A.width + e) * TILE + k * 32 + lane] : 0.0f;
            }
            b.init[k] = h.row[k] >= 0 ? (FORWARD ? rhs[h.row[k]] : yperm[h.yp[k]]) : 0.0f;
        }
        return b;
    };
    // tile `at` of the warp's chain
    auto solve_tile = [&](const long long tile, const int at, const TileHead& h, const ClusterBody& b) {
        const uint32_t mine = stage_addr + (uint32_t)(at & 1) * 1024u, next = stage_addr + (uint32_t)((at & 1) ^ 1) * 1024u;
        const uint32_t inbox_row = inbox_all + (uint32_t)warp * inbox_bytes + (uint32_t)at * (INBOX_SLOTS * 4u);
        if (A.trace && lane == 0) A.trace[4ll * tile] = tile_clock();
        // operands that come through L2 (other blocks): requested all at once, re-requested until everything is there
        unsigned int pend = 0u;
        unsigned int first[2 * TILE_MAX_W];
#pragma unroll
        for (int q = 0; q < 2 * TILE_MAX_W; ++q) {
            const int c = b.c[q / TILE_MAX_W][q % TILE_MAX_W];
            first[q] = c >= 0 ? peek(src + c) : 0u;
        }
        // up to two operands per row arrive in the inbox (other chains of the block): their shared-memory addresses and
        // which operand they are; everything else is in the row's four staging slots (one conflict-free 128-bit load)
        uint32_t ibaddr[2][2] = {{0u, 0u}, {0u, 0u}};
        int ibe[2][2] = {{-1, -1}, {-1, -1}};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
#pragma unroll
            for (int e = 0; e < TILE_MAX_W; ++e) {
                const int c = b.c[k][e];
                const uint32_t slot = mine + 4u * (uint32_t)(4 * (k * 32 + lane) + e);
                if (c <= E_INBOX) {
                    const uint32_t a_ = inbox_row + 4u * (uint32_t)(E_INBOX - c);
                    if (ibe[k][0] < 0) { ibaddr[k][0] = a_; ibe[k][0] = e; } else { ibaddr[k][1] = a_; ibe[k][1] = e; }
                }
                if (c >= 0 && first[4 * k + e] == SENTINEL) pend |= 1u << (4 * k + e);
                if (c >= 0 || c == E_NONE) sts_f32(slot, __uint_as_float(first[4 * k + e]));     // pushed slots (E_LOCAL) are left alone
            }
            if (FORWARD && !IC0 && h.row[k] >= 0 && fabsf(b.d[k]) < 1e-5) atomicOr(tickets + 3, 1u);   // H:1691-1693 (reported, not fatal here)
        }
        unsigned int polls = 0;
        while (__any_sync(0xFFFFFFFFu, pend != 0u)) {
            if (!poll_pause(&polls, abort_flag, A.sleep_first, A.sleep_later)) pend = 0u;
            unsigned int bits[2 * TILE_MAX_W];
#pragma unroll
            for (int q = 0; q < 2 * TILE_MAX_W; ++q)
                bits[q] = (pend >> q) & 1u ? peek(src + b.c[q / TILE_MAX_W][q % TILE_MAX_W]) : SENTINEL;
#pragma unroll
            for (int q = 0; q < 2 * TILE_MAX_W; ++q) {
                if (bits[q] != SENTINEL) { pend &= ~(1u << q); sts_f32(mine + 4u * (uint32_t)(4 * ((q / TILE_MAX_W) * 32 + lane) + (q % TILE_MAX_W)), __uint_as_float(bits[q])); }
            }
        }
        __syncwarp();
        if (A.trace && lane == 0) A.trace[4ll * tile + 1] = tile_clock();         // L2 operands complete
        float solved[2] = {0.0f, 0.0f};
        unsigned int spins = 0;
        // One step: a row that is due takes the operands that arrive from neighbour chains out of its inbox -- re-reading them,
        // with a short pause, for as long as one has not arrived -- and the rest out of its staging slots; then the sum in
        // operand order, the division, and the result pushed to every consumer that does not go through L2.
        auto step_rows = [&](const bool due0, const bool due1) {
            float in0[2] = {0.0f, 0.0f}, in1[2] = {0.0f, 0.0f};
            for (;;) {
                bool missing = false;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (k == 0 ? due0 : due1) {
                        if (ibe[k][0] >= 0) { in0[k] = lds_volatile(ibaddr[k][0]); missing |= __float_as_uint(in0[k]) == SENTINEL; }
                        if (ibe[k][1] >= 0) { in1[k] = lds_volatile(ibaddr[k][1]); missing |= __float_as_uint(in1[k]) == SENTINEL; }
                    }
                }
                if (!__any_sync(0xFFFFFFFFu, missing)) break;
                __nanosleep(32);
                if ((++spins & 1023u) == 0u) {
                    if (peek(reinterpret_cast<const float*>(abort_flag)) != 0u || spins >= (POLL_LIMIT << 2)) { atomicExch(abort_flag, 1u); break; }
                }
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (k == 0 ? due0 : due1) {
                    float4 xo;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(xo.x), "=f"(xo.y), "=f"(xo.z), "=f"(xo.w) : "r"(mine + 16u * (uint32_t)(k * 32 + lane)) : "memory");
                    const int e0 = ibe[k][0], e1 = ibe[k][1];
                    xo.x = e0 == 0 ? in0[k] : (e1 == 0 ? in1[k] : xo.x);
                    xo.y = e0 == 1 ? in0[k] : (e1 == 1 ? in1[k] : xo.y);
                    xo.z = e0 == 2 ? in0[k] : (e1 == 2 ? in1[k] : xo.z);
                    xo.w = e0 == 3 ? in0[k] : (e1 == 3 ? in1[k] : xo.w);
                    // same operand order and roundings as the row-level kernel (H:1685, H:1704, H:1813, H:1829)
                    const float p0 = __fmul_rn(b.v[k][0], xo.x), p1 = __fmul_rn(b.v[k][1], xo.y), p2 = __fmul_rn(b.v[k][2], xo.z), p3 = __fmul_rn(b.v[k][3], xo.w);
                    float acc = (FORWARD || IC0) ? b.init[k] : 0.0f;
                    if (FORWARD || IC0) { acc = __fsub_rn(acc, p0); acc = __fsub_rn(acc, p1); acc = __fsub_rn(acc, p2); acc = __fsub_rn(acc, p3); }
                    else { acc = __fadd_rn(p0, acc); acc = __fadd_rn(p1, acc); acc = __fadd_rn(p2, acc); acc = __fadd_rn(p3, acc); }
                    const float res = (FORWARD || IC0) ? __fdiv_rn(acc, b.d[k]) : __fsub_rn(b.init[k], __fdiv_rn(acc, b.d[k]));
                    const unsigned int pu = b.push[k], p2w = b.push2[k];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {                                                     // other chains of the block first (the longest way): their inbox, over DSMEM
                        const unsigned int r = (p2w >> (8 + 11 * j)) & 0x7FFu;
                        if (r != 0x7FFu) {
                            const unsigned int tw = r >> 5;
                            st_cluster_f32(inbox_all + (tw & (CLUSTER_WARPS - 1)) * inbox_bytes + (uint32_t)at * (INBOX_SLOTS * 4u) + 4u * (r & 31u), tw / CLUSTER_WARPS, res);
                        }
                    }
                    sts_f32(mine + 4u * (pu & 255u), res); sts_f32(mine + 4u * ((pu >> 8) & 255u), res); sts_f32(mine + 4u * ((pu >> 16) & 255u), res);
                    if ((p2w & 0xFFu) != 0xFFu) sts_f32(next + 4u * (p2w & 0xFFu), res);              // the chain's next tile
                    solved[k] = res;
                }
            }
        };
        const int first1 = __reduce_min_sync(0xFFFFFFFFu, b.step[1]);
        const int last0 = __reduce_max_sync(0xFFFFFFFFu, b.step[0] == 255 ? -1 : b.step[0]);
        int s = 0;
        for (; s < b.nsteps && s < first1; ++s) { step_rows(s == b.step[0], false); __syncwarp(); }
        for (; s < b.nsteps && s <= last0; ++s) { step_rows(s == b.step[0], s == b.step[1]); __syncwarp(); }
        for (; s < b.nsteps; ++s) { step_rows(false, s == b.step[1]); __syncwarp(); }
        if (A.trace && lane == 0) {
            unsigned int sm;
            asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
            A.trace[4ll * tile + 2] = tile_clock(); A.trace[4ll * tile + 3] = sm;
        }
        // every row is also published to the intermediate vector: other blocks poll it, the backward sweep starts from it
        float* const out = dst + (tile * TILE + lane);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (h.row[k] >= 0) {
                publish(out + k * 32, solved[k]);
                if (!FORWARD) x[h.row[k]] = solved[k];
            }
        }
    };

    for (;;) {
        if (crank == 0 && threadIdx.x == 0) sh_block = atomicAdd(ticket, 1u);
        cluster.sync();                                        // the previous block is finished everywhere; the claim is visible
        const unsigned int blk = *cluster.map_shared_rank(&sh_block, 0);
        if (blk >= (unsigned int)A.nblocks) {
            cluster.sync();                                    // nobody may leave while a peer still reads its shared memory (the claim above)
            break;
        }
        for (int i = lane; i < A.chain_len * INBOX_SLOTS; i += 32) my_inbox[i] = __uint_as_float(SENTINEL);
        cluster.sync();                                        // every inbox of the cluster is empty before anybody pushes
        const long long t0 = (((long long)blk * CLUSTER_CTAS + crank) * CLUSTER_WARPS + warp) * A.chain_len;
        TileHead h0 = load_head(t0, true), h1 = load_head(t0 + 1, A.chain_len > 1);
        ClusterBody r0 = load_body(t0, true, h0);
        for (int at = 0; at < A.chain_len; ++at) {
            const bool l1 = at + 1 < A.chain_len, l2 = at + 2 < A.chain_len;
            const TileHead h2 = load_head(t0 + at + 2, l2);
            const ClusterBody r1 = load_body(t0 + at + 1, l1, h1);
            solve_tile(t0 + at, at, h0, r0);
            h0 = h1; h1 = h2; r0 = r1;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host: proposal + generic verification + layout
// ---------------------------------------------------------------------------------------------------------------
// vectors whose resize() leaves the new elements uninitialised: the big layout arrays are filled by several threads right
// after (first touch in parallel instead of a single-threaded fill of several hundred MB)
template <class T>
struct raw_allocator : std::allocator<T> {
    template <class U> struct rebind { using other = raw_allocator<U>; };
    template <class U, class... Args>
    void construct(U* p, Args&&... args) {
        if constexpr (sizeof...(Args) == 0) ::new (static_cast<void*>(p)) U;
        else ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...);
    }
};
template <class T> using raw_vector = std::vector<T, raw_allocator<T>>;

// f(begin, end) over [0, n) on the host's threads (set-up code)
template <class F>
void for_range(long long n, F f) {
    const unsigned int hw = std::thread::hardware_concurrency();
    const int nt = n < (1 << 20) ? 1 : (int)std::min<unsigned int>(hw ? hw : 1, 16);
    if (nt <= 1) { f(0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back([&, t] { f(n * t / nt, n * (t + 1) / nt); });
    for (std::thread& x : th) x.join();
}
template <class V, class T>
void fill_parallel(V& v, size_t n, T value) {
    v.resize(n);
    auto* p = v.data();
    for_range((long long)n, [&](long long a, long long b) { std::fill(p + a, p + b, value); });
}

struct SweepLayout {
    std::vector<int32_t> level_ptr;                  // tiles in tile-level order: level l = tiles [level_ptr[l], level_ptr[l + 1])
    raw_vector<int32_t> order, ecol, eidx, where;    // where[row] = position
    raw_vector<uint8_t> steps;
    raw_vector<uint32_t> push;
    int levels = 0;
    int chain_len = 1;                               // tiles per chain (consecutive tile indices); 1: tiles in tile-level order
    // cluster schedule (sgs_cluster_kernel): CLUSTER_CHAINS chains per block, a block per thread-block cluster at a time
    int nblocks = 0;                                 // 0: not laid out for clusters
    long long ntiles = 0;                            // tiles the arrays hold (padding chains of incomplete blocks included)
    raw_vector<uint32_t> push2;                      // [tiles * 64] where a row's result goes outside its tile (see PUSH2_*)
};

struct ClusterPlan {                                  // proposal: which block a chain belongs to and which warp of the cluster takes it
    int nblocks = 0;
    std::vector<int32_t> block_of_chain, warp_of_chain;
};

}  // namespace

// Natural-order grid stencil?  (column offsets {1, nx} or {1, nx, nx * ny} with the row count a multiple of the largest.)
bool smm_sgs_detect_grid(int rows, const std::vector<int32_t>& start, const std::vector<int32_t>& pos, long long* pnx, long long* pny, long long* pnz) {
    // distinct |col - row| > 0 over a sample of the rows (head, middle, tail): this is only a proposal, what it leads to
    // is verified on every row by layout_sweep
    std::vector<long long> offs;
    const int sample = 1 << 16;
    for (int r = 0; r < rows; ++r) {
        if (r >= sample && r < rows - sample && !(r >= rows / 2 && r < rows / 2 + sample)) { r = (r < rows / 2 ? rows / 2 : rows - sample) - 1; continue; }
        for (int k = start[r]; k < start[r + 1]; ++k) {
            long long d = (long long)pos[k] - r;
            if (d < 0) d = -d;
            if (d == 0) continue;
            if (std::find(offs.begin(), offs.end(), d) == offs.end()) {
                if (offs.size() == 3) return false;
                offs.push_back(d);
            }
        }
    }
    return smm_sgs_grid_from_offsets(rows, offs, pnx, pny, pnz);
}

// the grid a set of distinct |col - row| > 0 offsets suggests: {1, nx} or {1, nx, nx * ny}
bool smm_sgs_grid_from_offsets(int rows, std::vector<long long> offs, long long* pnx, long long* pny, long long* pnz) {
    std::sort(offs.begin(), offs.end());
    if (offs.size() < 2 || offs.size() > 3 || offs[0] != 1) return false;
    long long nx = offs[1], ny = 0, nz = 1;
    if (offs.size() == 2) {
        if (rows % nx) return false;
        ny = rows / nx;
    } else {
        if (offs[2] % nx || rows % offs[2]) return false;
        ny = offs[2] / nx;
        nz = rows / offs[2];
    }
    *pnx = nx; *pny = ny; *pnz = nz;
    return true;
}

namespace {

// cluster ids for a natural-order grid stencil, or empty when the column offsets are not of that kind
// *chain_len: clusters [c * chain_len, (c + 1) * chain_len) are proposed as chain c (the tiles of one grid column along i)
std::vector<int32_t> propose_grid_tiles(int rows, const std::vector<int32_t>& start, const std::vector<int32_t>& pos, int* nclusters, int* chain_len = nullptr,
                                        ClusterPlan* plan = nullptr) {
    long long nx = 0, ny = 0, nz = 1;
    if (!smm_sgs_detect_grid(rows, start, pos, &nx, &ny, &nz)) return {};
    int ti, tj, tk;
    smm_sgs_tile_shape(nz, &ti, &tj, &tk);
    const long long TI = (nx + ti - 1) / ti, TJ = (ny + tj - 1) / tj, TK = (nz + tk - 1) / tk;
    if (TI * TJ * TK >= (1ll << 25)) return {};                // positions are int32: 64 * tiles < 2^31
    std::vector<int32_t> cl((size_t)rows);
    for_range(rows, [&](long long r0, long long r1) {
        for (long long r = r0; r < r1; ++r) {
            const long long i = r % nx, j = (r / nx) % ny, k = r / (nx * ny);
            cl[(size_t)r] = (int32_t)(((k / tk) * TJ + j / tj) * TI + i / ti);
        }
    });
    *nclusters = (int)(TI * TJ * TK);
    if (chain_len) *chain_len = (int)TI;
    if (plan) {
        // blocks of 4 x 8 neighbouring columns (32 x 1 for a 2D grid): most hand-offs between chains stay inside a block
        const long long BJ = TK > 1 ? 4 : CLUSTER_CHAINS, BK = TK > 1 ? 8 : 1;
        const long long NBJ = (TJ + BJ - 1) / BJ, NBK = (TK + BK - 1) / BK;
        plan->nblocks = (int)(NBJ * NBK);
        plan->block_of_chain.resize((size_t)(TJ * TK));
        plan->warp_of_chain.resize((size_t)(TJ * TK));
        for (long long K = 0; K < TK; ++K) for (long long J = 0; J < TJ; ++J) {
            plan->block_of_chain[(size_t)(K * TJ + J)] = (int32_t)((K / BK) * NBJ + J / BJ);
            plan->warp_of_chain[(size_t)(K * TJ + J)] = (int32_t)((K % BK) * BJ + J % BJ);
        }
    }
    return cl;
}

// f(cluster) for every cluster, on a few threads (set-up code; the clusters are independent of each other)
template <class F>
void for_clusters(int ncl, F f) {
    const unsigned int hw = std::thread::hardware_concurrency();
    int nt = (int)std::min<unsigned int>(hw > 3 ? hw / 2 : 1, 8);       // the two sweeps are laid out side by side
    if (ncl < 4096) nt = 1;
    if (nt <= 1) { for (int a = 0; a < ncl; ++a) f(a); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&, t] { for (int a = (int)((long long)ncl * t / nt); a < (int)((long long)ncl * (t + 1) / nt); ++a) f(a); });
    for (std::thread& x : th) x.join();
}

// Lay one sweep out by tiles.  Returns false when the proposal does not verify.
bool layout_sweep(bool forward, int rows, const std::vector<int32_t>& start, const std::vector<int32_t>& pos, const std::vector<int32_t>& diag,
                  const std::vector<int32_t>& cl, int ncl, int width, SweepLayout* out, int chain_len = 1, const ClusterPlan* plan = nullptr) {
    auto dep_begin = [&](int r) { return forward ? start[r] : diag[r] + 1; };
    auto dep_end = [&](int r) { return forward ? diag[r] : start[r + 1]; };
    static const bool trace_on = [] { const char* e = getenv("SMM_B200_SETUP_TRACE"); return e && atoi(e) != 0; }();
    auto tl = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!trace_on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[smm set-up]   %s sweep: %-28s %7.3f s\n", forward ? "forward " : "backward", what, std::chrono::duration<double>(now - tl).count());
        tl = now;
    };
    // rows of every cluster, ascending
    std::vector<int32_t> cptr((size_t)ncl + 1, 0);
    std::vector<int32_t> crow((size_t)rows);
    {
        // counting sort of the rows by cluster, stable, on a few threads: thread t counts its slice of the rows, the counts
        // are prefix-summed cluster by cluster and slice by slice, and every thread scatters its slice
        const unsigned int hw = std::thread::hardware_concurrency();
        const int nt = rows < (1 << 20) ? 1 : (int)std::min<unsigned int>(hw > 3 ? hw / 2 : 1, 8);
        std::vector<std::vector<uint8_t>> cnt((size_t)nt);       // a cluster holds at most TILE rows: 8 bits, saturating
        std::atomic<bool> too_many(false);
        auto slice = [&](int t) { return std::make_pair((long long)rows * t / nt, (long long)rows * (t + 1) / nt); };
        {
            std::vector<std::thread> th;
            for (int t = 0; t < nt; ++t) th.emplace_back([&, t] {
                cnt[t].assign((size_t)ncl, 0);
                const auto [r0, r1] = slice(t);
                for (long long r = r0; r < r1; ++r) { uint8_t& c = cnt[t][(size_t)cl[r]]; if (c == 255) too_many = true; else ++c; }
            });
            for (std::thread& x : th) x.join();
        }
        if (too_many) return false;
        std::vector<std::vector<int32_t>> base((size_t)nt, std::vector<int32_t>());
        for (int t = 0; t < nt; ++t) base[t].resize((size_t)ncl);
        int32_t run = 0;
        for (int a = 0; a < ncl; ++a) {
            cptr[a] = run;
            int total = 0;
            for (int t = 0; t < nt; ++t) { base[t][a] = run + total; total += cnt[t][a]; }
            if (total > TILE) return false;
            run += total;
        }
        cptr[ncl] = run;
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back([&, t] {
            const auto [r0, r1] = slice(t);
            std::vector<int32_t>& cur = base[t];
            for (long long r = r0; r < r1; ++r) crow[(size_t)cur[(size_t)cl[r]]++] = (int32_t)r;
        });
        for (std::thread& x : th) x.join();
    }
    mark("rows by tile (counting sort)");
    // distinct predecessor tiles (in the order the rows of the tile meet them)
    std::vector<int32_t> pred((size_t)ncl * MAX_PREDS, -1);
    std::vector<uint8_t> npred((size_t)ncl, 0);
    std::atomic<bool> bad(false);
    for_clusters(ncl, [&](int a) {
        int32_t* pa = &pred[(size_t)a * MAX_PREDS];
        int n = 0;
        for (int q = cptr[a]; q < cptr[a + 1]; ++q) {
            const int r = crow[q];
            for (int k = dep_begin(r); k < dep_end(r); ++k) {
                const int b = cl[pos[k]];
                if (b == a) continue;
                int i = 0;
                while (i < n && pa[i] != b) ++i;
                if (i == n) {
                    if (n == MAX_PREDS) { bad = true; return; }
                    pa[n++] = b;
                }
            }
        }
        npred[a] = (uint8_t)n;
    });
    if (bad) return false;
    mark("predecessor tiles");
    // Kahn: tile levels, cycle check
    std::vector<int32_t> sptr((size_t)ncl + 1, 0);
    for (int a = 0; a < ncl; ++a) for (int i = 0; i < npred[a]; ++i) sptr[(size_t)pred[(size_t)a * MAX_PREDS + i] + 1]++;
    for (int a = 0; a < ncl; ++a) sptr[(size_t)a + 1] += sptr[a];
    std::vector<int32_t> succ((size_t)sptr[ncl]);
    {
        std::vector<int32_t> cur(sptr.begin(), sptr.end() - 1);
        for (int a = 0; a < ncl; ++a) for (int i = 0; i < npred[a]; ++i) succ[(size_t)cur[pred[(size_t)a * MAX_PREDS + i]]++] = a;
    }
    std::vector<int32_t> level((size_t)ncl, 0), indeg((size_t)ncl), queue;
    queue.reserve((size_t)ncl);
    for (int a = 0; a < ncl; ++a) { indeg[a] = npred[a]; if (!indeg[a]) queue.push_back(a); }
    int maxl = -1;
    for (size_t q = 0; q < queue.size(); ++q) {
        const int a = queue[q];
        maxl = std::max(maxl, level[a]);
        for (int i = sptr[a]; i < sptr[a + 1]; ++i) {
            const int b = succ[i];
            level[b] = std::max(level[b], level[a] + 1);
            if (--indeg[b] == 0) queue.push_back(b);
        }
    }
    if ((int)queue.size() != ncl) return false;                // the tile graph has a cycle
    out->levels = maxl + 1;
    // Tile order.  Proposed chains (clusters [c * chain_len, (c + 1) * chain_len), walked upwards by the forward sweep and
    // downwards by the backward one) are verified: a dependency inside a chain must point to an earlier tile of the chain,
    // and the graph of the chains must be acyclic; chains are then ordered by their level in that graph, a chain's tiles
    // are consecutive.  Otherwise: single-tile chains in tile-level order.
    std::vector<int32_t> tile_of((size_t)ncl);
    out->chain_len = 1;
    if (chain_len > 1 && ncl % chain_len == 0) {
        const int nch = ncl / chain_len;
        auto chain_of = [&](int a) { return a / chain_len; };
        auto at_of = [&](int a) { return forward ? a % chain_len : chain_len - 1 - a % chain_len; };
        std::vector<int32_t> cpred((size_t)nch * MAX_PREDS, -1);
        std::vector<uint8_t> ncp((size_t)nch, 0);
        bool ok = true;
        for (int a = 0; a < ncl && ok; ++a) {
            const int ca = chain_of(a);
            for (int i = 0; i < npred[a] && ok; ++i) {
                const int b = pred[(size_t)a * MAX_PREDS + i], cb = chain_of(b);
                if (cb == ca) { ok = at_of(b) < at_of(a); continue; }
                int32_t* pc = &cpred[(size_t)ca * MAX_PREDS];
                int j = 0;
                while (j < ncp[ca] && pc[j] != cb) ++j;
                if (j == ncp[ca]) {
                    if (ncp[ca] == MAX_PREDS) ok = false; else pc[ncp[ca]++] = cb;
                }
            }
        }
        std::vector<int32_t> clevel((size_t)nch, 0);
        if (ok) {                                              // Kahn on the chain graph
            std::vector<int32_t> cs((size_t)nch + 1, 0), indeg2((size_t)nch), q2;
            for (int c = 0; c < nch; ++c) for (int i = 0; i < ncp[c]; ++i) cs[(size_t)cpred[(size_t)c * MAX_PREDS + i] + 1]++;
            for (int c = 0; c < nch; ++c) cs[(size_t)c + 1] += cs[c];
            std::vector<int32_t> csucc((size_t)cs[nch]);
            {
                std::vector<int32_t> cur(cs.begin(), cs.end() - 1);
                for (int c = 0; c < nch; ++c) for (int i = 0; i < ncp[c]; ++i) csucc[(size_t)cur[cpred[(size_t)c * MAX_PREDS + i]]++] = c;
            }
            q2.reserve((size_t)nch);
            for (int c = 0; c < nch; ++c) { indeg2[c] = ncp[c]; if (!indeg2[c]) q2.push_back(c); }
            for (size_t q = 0; q < q2.size(); ++q) {
                const int c = q2[q];
                for (int i = cs[c]; i < cs[c + 1]; ++i) {
                    const int d = csucc[i];
                    clevel[d] = std::max(clevel[d], clevel[c] + 1);
                    if (--indeg2[d] == 0) q2.push_back(d);
                }
            }
            ok = (int)q2.size() == nch;
        }
        if (ok) {
            int maxc = 0;
            for (int c = 0; c < nch; ++c) maxc = std::max(maxc, clevel[c]);
            std::vector<int32_t> lp((size_t)maxc + 2, 0), rank((size_t)nch);
            for (int c = 0; c < nch; ++c) lp[(size_t)clevel[c] + 1]++;
            for (int l = 0; l <= maxc; ++l) lp[(size_t)l + 1] += lp[l];
            if (forward) { for (int c = 0; c < nch; ++c) rank[c] = lp[clevel[c]]++; }
            else { for (int c = nch - 1; c >= 0; --c) rank[c] = lp[clevel[c]]++; }
            for (int a = 0; a < ncl; ++a) tile_of[a] = rank[chain_of(a)] * chain_len + at_of(a);
            out->chain_len = chain_len;
            // Cluster schedule: the proposed blocks of chains are verified -- the graph of the blocks must be acyclic -- and
            // ordered by their level in it; chain rank = 32 * block rank + warp.  (Chains of one block run concurrently, one
            // per warp of the cluster, so dependencies inside a block need no order.)
            if (plan && plan->nblocks > 0 && (int)plan->block_of_chain.size() == nch) {
                const int nb = plan->nblocks;
                std::vector<std::vector<int32_t>> bsucc((size_t)nb);
                std::vector<int32_t> bdeg((size_t)nb, 0), blevel((size_t)nb, 0), bq;
                for (int c = 0; c < nch; ++c) for (int i = 0; i < ncp[c]; ++i) {
                    const int bb = plan->block_of_chain[cpred[(size_t)c * MAX_PREDS + i]], b = plan->block_of_chain[c];
                    if (bb != b && std::find(bsucc[bb].begin(), bsucc[bb].end(), b) == bsucc[bb].end()) { bsucc[bb].push_back(b); bdeg[b]++; }
                }
                for (int b = 0; b < nb; ++b) if (!bdeg[b]) bq.push_back(b);
                for (size_t q = 0; q < bq.size(); ++q) for (int b : bsucc[bq[q]]) { blevel[b] = std::max(blevel[b], blevel[bq[q]] + 1); if (--bdeg[b] == 0) bq.push_back(b); }
                bool okb = (int)bq.size() == nb;
                std::vector<uint8_t> seen((size_t)nb * CLUSTER_CHAINS, 0);
                for (int c = 0; c < nch && okb; ++c) {          // every chain its own warp of its block
                    const int w = plan->warp_of_chain[c];
                    okb = w >= 0 && w < CLUSTER_CHAINS && !seen[(size_t)plan->block_of_chain[c] * CLUSTER_CHAINS + w];
                    if (okb) seen[(size_t)plan->block_of_chain[c] * CLUSTER_CHAINS + w] = 1;
                }
                if (okb) {
                    std::vector<int32_t> order_b((size_t)nb), brank((size_t)nb);
                    for (int b = 0; b < nb; ++b) order_b[b] = b;
                    std::stable_sort(order_b.begin(), order_b.end(), [&](int x, int y) {
                        if (blevel[x] != blevel[y]) return blevel[x] < blevel[y];
                        return forward ? x < y : x > y;
                    });
                    for (int i = 0; i < nb; ++i) brank[order_b[i]] = i;
                    for (int a = 0; a < ncl; ++a) {
                        const int c = chain_of(a);
                        tile_of[a] = (brank[plan->block_of_chain[c]] * CLUSTER_CHAINS + plan->warp_of_chain[c]) * chain_len + at_of(a);
                    }
                    out->nblocks = nb;
                }
            }
        }
    }
    mark("tile levels and order");
    const int ntl = out->nblocks > 0 ? out->nblocks * CLUSTER_CHAINS * out->chain_len : ncl;   // tiles the arrays hold
    out->ntiles = ntl;
    if (out->chain_len == 1) {                                 // tiles in level order (stable in the cluster id; the backward sweep runs the ids downwards)
        std::vector<int32_t> lptr((size_t)out->levels + 1, 0);
        for (int a = 0; a < ncl; ++a) lptr[(size_t)level[a] + 1]++;
        for (int l = 0; l < out->levels; ++l) lptr[(size_t)l + 1] += lptr[l];
        std::vector<int32_t> cur(lptr.begin(), lptr.end() - 1);
        out->level_ptr = lptr;
        if (forward) { for (int a = 0; a < ncl; ++a) tile_of[a] = cur[level[a]]++; }
        else { for (int a = ncl - 1; a >= 0; --a) tile_of[a] = cur[level[a]]++; }
    }
    fill_parallel(out->order, (size_t)ntl * TILE, (int32_t)-1);
    fill_parallel(out->where, (size_t)rows, (int32_t)0);
    fill_parallel(out->steps, (size_t)ntl * TILE + (size_t)ntl, (uint8_t)255);   // [tiles * 64] step of every row, then [tiles] number of steps
    if (out->nblocks > 0) std::fill(out->steps.begin() + (size_t)ntl * TILE, out->steps.end(), (uint8_t)0);   // padding tiles: no steps
    raw_vector<int8_t> ilev;
    fill_parallel(ilev, (size_t)rows, (int8_t)0);
    for_clusters(ncl, [&](int a) {
        const int n = cptr[a + 1] - cptr[a];
        const int32_t* R = &crow[(size_t)cptr[a]];
        int nl = 0;
        for (int q = 0; q < n; ++q) {                          // internal levels, in dependency order
            const int r = forward ? R[q] : R[n - 1 - q];
            int l = 0;
            for (int k = dep_begin(r); k < dep_end(r); ++k) if (cl[pos[k]] == a) l = std::max(l, ilev[pos[k]] + 1);
            if (l >= MAX_STEPS) { bad = true; return; }
            ilev[r] = (int8_t)l;
            nl = std::max(nl, l + 1);
        }
        int32_t tmp[TILE];                                     // rows by internal level, stable in the sweep's row order
        int at[MAX_STEPS + 1] = {0};
        for (int q = 0; q < n; ++q) at[ilev[R[q]] + 1]++;
        for (int l = 0; l < nl; ++l) at[l + 1] += at[l];
        for (int q = 0; q < n; ++q) { const int r = forward ? R[q] : R[n - 1 - q]; tmp[at[ilev[r]]++] = r; }
        const int t = tile_of[a];
        out->steps[(size_t)ntl * TILE + t] = (uint8_t)nl;       // a step = an internal level (lane l solves rows l and l + 32)
        for (int i = 0; i < n; ++i) {
            out->steps[(size_t)t * TILE + i] = (uint8_t)ilev[tmp[i]];
            out->order[(size_t)t * TILE + i] = tmp[i];
            out->where[(size_t)tmp[i]] = t * TILE + i;
        }
    });
    if (bad) return false;
    mark("rows inside the tiles (steps)");
    // entries: [tile][slot][64], operand order = the reference's (ascending columns forward, descending backward);
    // push lists: where inside the tile's operand staging a row's result has to go (bytes 0..2; a byte that is not needed
    // points at the row's own first slot, which nobody reads once the row is solved)
    fill_parallel(out->ecol, (size_t)ntl * width * TILE, (int32_t)-1);
    fill_parallel(out->eidx, (size_t)ntl * width * TILE, (int32_t)-1);
    fill_parallel(out->push, (size_t)ntl * TILE, 0u);
    const bool clustered = out->nblocks > 0;
    const int clen = out->chain_len;
    if (clustered) fill_parallel(out->push2, (size_t)ntl * TILE, PUSH2_NONE);
    // claim a field of a producer row's push2 word (other consumer tiles are laid out by other threads): CAS
    auto claim_push2 = [&](uint32_t* word, int shift, uint32_t mask, uint32_t value) {
        uint32_t cur = __atomic_load_n(word, __ATOMIC_RELAXED);
        for (;;) {
            if (((cur >> shift) & mask) != mask) return false;                    // taken
            const uint32_t want = (cur & ~(mask << shift)) | (value << shift);
            if (__atomic_compare_exchange_n(word, &cur, want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return true;
        }
    };
    for_clusters(ncl, [&](int a) {                             // in-tile pushes touch the tile's own rows only; push2 words are claimed atomically
        const int t = tile_of[a];
        uint8_t npush[TILE] = {0};
        int ninbox = 0;
        for (int i = 0; i < TILE; ++i) out->push[(size_t)t * TILE + i] = (uint32_t)(i * TILE_MAX_W) * 0x01010101u;
        for (int q = cptr[a]; q < cptr[a + 1]; ++q) {
            const int r = crow[q];
            const int i = out->where[r] & (TILE - 1);
            const int cnt = dep_end(r) - dep_begin(r);
            for (int e = 0; e < cnt; ++e) {
                const int srci = forward ? start[r] + e : start[r + 1] - 1 - e;
                const size_t at = ((size_t)t * width + e) * TILE + i;
                const int w = out->where[pos[srci]];
                out->ecol[at] = w;
                out->eidx[at] = srci;
                if ((w >> 6) == t) {                           // produced inside the tile: the producer pushes it
                    const int j = w & (TILE - 1);
                    if (npush[j] == 3) { bad = true; return; } // a row with more than three consumers inside its tile
                    const int sh = 8 * npush[j]++;
                    uint32_t& pu = out->push[(size_t)t * TILE + j];
                    pu = (pu & ~(0xFFu << sh)) | ((uint32_t)(i * TILE_MAX_W + e) << sh);
                    if (clustered) out->ecol[at] = E_LOCAL;
                } else if (clustered) {
                    // the previous tile of the same chain (same warp: its staging), or the same tile index of another chain
                    // of the block (another warp of the cluster: its inbox); everything else stays a polled position
                    const int tp = w >> 6, chain_t = t / clen, chain_p = tp / clen;
                    uint32_t* word = &out->push2[(size_t)w];
                    if (chain_p == chain_t && tp == t - 1) {
                        if (claim_push2(word, 0, 0xFFu, (uint32_t)(i * TILE_MAX_W + e))) out->ecol[at] = E_LOCAL;
                    } else if (chain_p / CLUSTER_CHAINS == chain_t / CLUSTER_CHAINS && tp % clen == t % clen && ninbox < INBOX_SLOTS) {
                        const uint32_t v = ((uint32_t)(chain_t % CLUSTER_CHAINS) << 5) | (uint32_t)ninbox;
                        if (claim_push2(word, 8, 0x7FFu, v) || claim_push2(word, 19, 0x7FFu, v)) out->ecol[at] = E_INBOX - ninbox++;
                    }
                }
            }
        }
    });
    if (bad) return false;
    mark("entries and push lists");
    if (clustered) {                                           // the two remote fields in a canonical order (they were claimed by racing threads)
        for (uint32_t& v : out->push2) {
            const uint32_t r0 = (v >> 8) & 0x7FFu, r1 = (v >> 19) & 0x7FFu;
            if (r0 > r1) v = (v & 0xFFu) | (r1 << 8) | (r0 << 19);
        }
    }
    return true;
}

template <class V, class A>
int upload(const std::vector<V, A>& h, V** d) {
    SMM_CUDA(cudaMalloc(d, sizeof(V) * (h.empty() ? 1 : h.size())));
    if (!h.empty()) SMM_CUDA(cudaMemcpy(*d, h.data(), sizeof(V) * h.size(), cudaMemcpyHostToDevice));
    return SMM_OK;
}

}  // namespace

bool smm_sgs_tiles_build(smm_precond* p, int rows, const std::vector<int32_t>& start, const std::vector<int32_t>& pos, const std::vector<int32_t>& diag) {
    if (const char* e = getenv("SMM_B200_SGS_TILES")) { if (atoi(e) == 0) return false; }
    if (rows < 2 * TILE) return false;
    int width = 0;
    for (int r = 0; r < rows; ++r) width = std::max(width, std::max(diag[r] - start[r], start[r + 1] - 1 - diag[r]));
    if (width > TILE_MAX_W) return false;
    if (width == 0) width = 1;
    int ncl = 0, chain_len = 1;
    ClusterPlan plan;
    const std::vector<int32_t> cl = propose_grid_tiles(rows, start, pos, &ncl, &chain_len, &plan);
    if (cl.empty()) return false;
    // Schedule.  Default: tiles in tile-level order, one warp per tile -- the fastest of the three on every grid measured
    // (256^3 apply: 0.96 ms; chains 0.99 ms; clusters 1.43-1.94 ms, profiles/r02_sgs_schedules.txt).  The other two are kept
    // selectable for measurements: SMM_B200_SGS_CHAINS=1 (a warp marches along a column of tiles), SMM_B200_SGS_CLUSTERS=1
    // (chains in blocks of 32, hand-offs inside a block pushed through shared memory / DSMEM by a thread-block cluster).
    const char* e_chains = getenv("SMM_B200_SGS_CHAINS");
    const char* e_clusters = getenv("SMM_B200_SGS_CLUSTERS");
    bool want_clusters = e_clusters && atoi(e_clusters) != 0 && chain_len > 1 && chain_len <= CLUSTER_MAX_CHAIN;
    if (!want_clusters && !(e_chains && atoi(e_chains) != 0)) chain_len = 1;
    SweepLayout L[2];                                          // the two sweeps are laid out side by side (set-up time)
    for (int attempt = 0; attempt < 2; ++attempt) {
        const ClusterPlan* pl = want_clusters ? &plan : nullptr;
        L[0] = SweepLayout(); L[1] = SweepLayout();
        std::future<bool> bwd = std::async(std::launch::async, [&] { return layout_sweep(false, rows, start, pos, diag, cl, ncl, width, &L[1], chain_len, pl); });
        const bool fwd_ok = layout_sweep(true, rows, start, pos, diag, cl, ncl, width, &L[0], chain_len, pl);
        if (!bwd.get() || !fwd_ok) return false;
        // both sweeps must agree on the schedule kind and on the padded size (they share the position space)
        if ((L[0].nblocks > 0) == (L[1].nblocks > 0) && L[0].ntiles == L[1].ntiles) break;
        if (!want_clusters) return false;
        want_clusters = false;
    }
    raw_vector<int32_t> yp;
    yp.resize(L[1].order.size());
    for_range((long long)yp.size(), [&](long long a, long long b) {
        for (long long t = a; t < b; ++t) yp[(size_t)t] = L[1].order[(size_t)t] >= 0 ? L[0].where[(size_t)L[1].order[(size_t)t]] : 0;
    });
    const size_t npos = (size_t)L[0].ntiles * TILE;
    p->tile_blocks = L[0].nblocks;
    p->threads_fwd = p->threads_bwd = (long long)npos;
    p->tile_width = width;
    p->tile_levels[0] = L[0].levels;
    p->tile_levels[1] = L[1].levels;
    p->tile_chain[0] = L[0].chain_len;
    p->tile_chain[1] = L[1].chain_len;
    p->tile_level_ptr = L[0].level_ptr;
    bool ok = upload(L[0].order, &p->order_fwd) == SMM_OK && upload(L[1].order, &p->order_bwd) == SMM_OK && upload(yp, &p->ypos) == SMM_OK &&
              cudaMalloc(&p->yperm, sizeof(float) * npos) == cudaSuccess && cudaMalloc(&p->xperm, sizeof(float) * npos) == cudaSuccess;
    for (int w = 0; w < 2 && ok; ++w) {
        p->esize[w] = (long long)L[w].ecol.size();
        ok = upload(L[w].ecol, &p->ecol[w]) == SMM_OK && upload(L[w].eidx, &p->eidx[w]) == SMM_OK && upload(L[w].steps, &p->tile_steps[w]) == SMM_OK &&
             upload(L[w].push, &p->tile_push[w]) == SMM_OK && (L[w].push2.empty() || upload(L[w].push2, &p->tile_push2[w]) == SMM_OK) &&
             cudaMalloc(&p->eval[w], sizeof(float) * L[w].ecol.size()) == cudaSuccess &&
             cudaMalloc(&p->dval[w], sizeof(float) * npos) == cudaSuccess;
    }
    if (!ok) {                                                  // out of memory: leave the handle as it was, the row-level schedule needs less
        cudaGetLastError();
        cudaFree(p->order_fwd); cudaFree(p->order_bwd); cudaFree(p->ypos); cudaFree(p->yperm); cudaFree(p->xperm);
        p->order_fwd = p->order_bwd = p->ypos = nullptr;
        p->yperm = p->xperm = nullptr;
        for (int w = 0; w < 2; ++w) {
            cudaFree(p->ecol[w]); cudaFree(p->eidx[w]); cudaFree(p->tile_steps[w]); cudaFree(p->tile_push[w]); cudaFree(p->tile_push2[w]); cudaFree(p->eval[w]); cudaFree(p->dval[w]);
            p->ecol[w] = p->eidx[w] = nullptr; p->tile_steps[w] = nullptr; p->tile_push[w] = nullptr; p->tile_push2[w] = nullptr; p->eval[w] = p->dval[w] = nullptr;
            p->esize[w] = 0;
        }
        p->threads_fwd = p->threads_bwd = 0;
        p->tile_blocks = 0;
        return false;
    }
    p->tiled = true;
    return true;
}

namespace {
template <int TILE_WARPS>
void launch_tiles(const smm_precond* p, const TileArgs& F, const TileArgs& B, const float* rhs_dev, float* x_dev, SolveState* state, long long cap,
                  cudaStream_t s) {
    const long long nchains = F.nchains > B.nchains ? F.nchains : B.nchains;
    const long long nblocks = (nchains + TILE_WARPS - 1) / TILE_WARPS;   // one warp per chain at a time
    // persistent grid: no more CTAs than can be resident (the rest would only find the tickets used up)
    static int resident_dev[SMM_MAX_DEVICES] = {0};            // occupancy is a per-device property
    int resident;
    {
        std::lock_guard<std::mutex> lk(g_smm_attr_mu);
        int& r = resident_dev[p->m->device % SMM_MAX_DEVICES];
        if (!r) {
            int per_sm = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sgs_tile_kernel<false, false, TILE_WARPS>, TILE_WARPS * 32, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
            r = per_sm * p->m->sm_count;
        }
        resident = r;
    }
    if (cap > resident) cap = resident;
    const unsigned grid = (unsigned)(nblocks < cap ? nblocks : cap);
    if (p->kind != 0) {
        sgs_tile_kernel<true, true, TILE_WARPS><<<grid, TILE_WARPS * 32, 0, s>>>(F, rhs_dev, p->yperm, p->xperm, x_dev, p->tickets, state);
        sgs_tile_kernel<false, true, TILE_WARPS><<<grid, TILE_WARPS * 32, 0, s>>>(B, rhs_dev, p->yperm, p->xperm, x_dev, p->tickets, state);
    } else {
        sgs_tile_kernel<true, false, TILE_WARPS><<<grid, TILE_WARPS * 32, 0, s>>>(F, rhs_dev, p->yperm, p->xperm, x_dev, p->tickets, state);
        sgs_tile_kernel<false, false, TILE_WARPS><<<grid, TILE_WARPS * 32, 0, s>>>(B, rhs_dev, p->yperm, p->xperm, x_dev, p->tickets, state);
    }
}
}  // namespace

namespace {
size_t cluster_smem_bytes(int chain_len) { return (size_t)CLUSTER_WARPS * 2 * 1024 + (size_t)CLUSTER_WARPS * chain_len * INBOX_SLOTS * sizeof(float); }

template <bool FORWARD, bool IC0>
int launch_cluster_sweep(const smm_precond* p, const TileArgs& A, const float* rhs_dev, float* x_dev, SolveState* state, cudaStream_t s) {
    const size_t smem = cluster_smem_bytes(A.chain_len);
    // persistent grid: as many clusters as can be resident (the rest would only find the ticket used up)
    static int resident_dev[SMM_MAX_DEVICES][2][2] = {};
    int resident;
    {
        std::lock_guard<std::mutex> lk(g_smm_attr_mu);
        int& r = resident_dev[p->m->device % SMM_MAX_DEVICES][FORWARD ? 1 : 0][IC0 ? 1 : 0];
        if (!r) {
            SMM_CUDA(cudaFuncSetAttribute(sgs_cluster_kernel<FORWARD, IC0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cluster_smem_bytes(CLUSTER_MAX_CHAIN)));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(CLUSTER_CTAS * 1024); cfg.blockDim = dim3(CLUSTER_WARPS * 32); cfg.dynamicSmemBytes = cluster_smem_bytes(CLUSTER_MAX_CHAIN);
            cudaLaunchAttribute at;
            at.id = cudaLaunchAttributeClusterDimension;
            at.val.clusterDim.x = CLUSTER_CTAS; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
            cfg.attrs = &at; cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, sgs_cluster_kernel<FORWARD, IC0>, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = p->m->sm_count / CLUSTER_CTAS; }
            r = n;
        }
        resident = r;
    }
    // the occupancy was queried for the largest inbox; shorter chains leave room for more clusters, bounded by the warps per SM
    long long fit = resident;
    if (A.chain_len < CLUSTER_MAX_CHAIN) {
        const long long by_smem = (long long)((227 * 1024) / (smem + 1024)), by_warps = 2048 / (CLUSTER_WARPS * 32), by_regs = SMM_TILE_MIN_CTAS * 4 / CLUSTER_WARPS + 1;
        const long long per_sm = std::min(by_smem, std::min(by_warps, by_regs));
        fit = std::max<long long>(resident, per_sm * p->m->sm_count / CLUSTER_CTAS * 7 / 8);
    }
    const unsigned clusters = (unsigned)std::min<long long>(fit, A.nblocks);
    sgs_cluster_kernel<FORWARD, IC0><<<clusters * CLUSTER_CTAS, CLUSTER_WARPS * 32, smem, s>>>(A, rhs_dev, p->yperm, p->xperm, x_dev, p->tickets, state);
    return SMM_OK;
}

int launch_clusters(const smm_precond* p, const TileArgs& F, const TileArgs& B, const float* rhs_dev, float* x_dev, SolveState* state, cudaStream_t s) {
    if (p->kind != 0) {
        SMM_TRY((launch_cluster_sweep<true, true>(p, F, rhs_dev, x_dev, state, s)));
        SMM_TRY((launch_cluster_sweep<false, true>(p, B, rhs_dev, x_dev, state, s)));
    } else {
        SMM_TRY((launch_cluster_sweep<true, false>(p, F, rhs_dev, x_dev, state, s)));
        SMM_TRY((launch_cluster_sweep<false, false>(p, B, rhs_dev, x_dev, state, s)));
    }
    return SMM_OK;
}
}  // namespace

int smm_sgs_tiles_launch(const smm_precond* p, const float* rhs_dev, float* x_dev, SolveState* state, int ctas_per_sm, unsigned int sleep_first,
                         unsigned int sleep_later, cudaStream_t s) {
    const long long ntiles = p->threads_fwd / TILE;
    const long long cap = (long long)p->m->sm_count * ctas_per_sm;
    // debug: SMM_B200_SGS_TRACE=<file> records per-tile timestamps of the forward sweep of every apply (last one kept)
    static const char* trace_path = getenv("SMM_B200_SGS_TRACE");
    static unsigned long long* trace = nullptr;
    static long long trace_cap = 0;
    if (trace_path && trace_cap < ntiles) {
        cudaFree(trace);
        trace = nullptr;
        SMM_CUDA(cudaMalloc(&trace, sizeof(unsigned long long) * 4 * (size_t)ntiles));
        trace_cap = ntiles;
    }
    const uint8_t* nsf = p->tile_steps[0] + ntiles * TILE;
    const uint8_t* nsb = p->tile_steps[1] + ntiles * TILE;
    const int cf = p->tile_chain[0] > 0 ? p->tile_chain[0] : 1, cb = p->tile_chain[1] > 0 ? p->tile_chain[1] : 1;
    TileArgs F{nsf, p->tile_steps[0], p->tile_push[0], p->order_fwd, nullptr, p->ecol[0], p->eval[0], p->dval[0], ntiles, ntiles / cf, cf, p->tile_blocks,
               p->tile_push2[0], p->tile_width, sleep_first, sleep_later, trace};
    TileArgs B{nsb, p->tile_steps[1], p->tile_push[1], p->order_bwd, p->ypos, p->ecol[1], p->eval[1], p->dval[1], ntiles, ntiles / cb, cb, p->tile_blocks,
               p->tile_push2[1], p->tile_width, sleep_first, sleep_later, nullptr};
    if (p->tile_blocks > 0) SMM_TRY(launch_clusters(p, F, B, rhs_dev, x_dev, state, s));
    else launch_tiles<4>(p, F, B, rhs_dev, x_dev, state, cap, s);   // 1, 2 and 8 tiles per CTA claim measured the same
    SMM_CUDA(cudaGetLastError());
    if (trace) {
        std::vector<unsigned long long> h(4 * (size_t)ntiles);
        SMM_CUDA(cudaStreamSynchronize(s));
        SMM_CUDA(cudaMemcpy(h.data(), trace, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost));
        if (FILE* f = fopen(trace_path, "wb")) { fwrite(h.data(), sizeof(unsigned long long), h.size(), f); fclose(f); }
    }
    return SMM_OK;
}
