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
#include <atomic>
#include <future>
#include <thread>
#include <vector>

#include "sgs_internal.cuh"

namespace {

#ifndef SMM_TILE_MIN_CTAS
#define SMM_TILE_MIN_CTAS 4      // resident 4-warp CTAs per SM the register allocation must allow
#endif
constexpr int TILE = 64;
constexpr int TILE_MAX_W = 4;
constexpr int MAX_STEPS = 64;
constexpr int MAX_PREDS = 8;

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
// host: proposal + generic verification + layout
// ---------------------------------------------------------------------------------------------------------------
struct SweepLayout {
    std::vector<int32_t> order, ecol, eidx, where;   // where[row] = position
    std::vector<uint8_t> steps;
    std::vector<uint32_t> push;
    int levels = 0;
    int chain_len = 1;                               // tiles per chain (consecutive tile indices); 1: tiles in tile-level order
};

// cluster ids for a natural-order grid stencil, or empty when the column offsets are not of that kind
// *chain_len: clusters [c * chain_len, (c + 1) * chain_len) are proposed as chain c (the tiles of one grid column along i)
std::vector<int32_t> propose_grid_tiles(int rows, const std::vector<int32_t>& start, const std::vector<int32_t>& pos, int* nclusters, int* chain_len = nullptr) {
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
                if (offs.size() == 3) return {};
                offs.push_back(d);
            }
        }
    }
    std::sort(offs.begin(), offs.end());
    if (offs.size() < 2 || offs[0] != 1) return {};
    long long nx = offs[1], ny = 0, nz = 1;
    if (offs.size() == 2) {
        if (rows % nx) return {};
        ny = rows / nx;
    } else {
        if (offs[2] % nx || rows % offs[2]) return {};
        ny = offs[2] / nx;
        nz = rows / offs[2];
    }
    const int ti = nz > 1 ? 4 : 8, tj = nz > 1 ? 4 : 8, tk = nz > 1 ? 4 : 1;
    const long long TI = (nx + ti - 1) / ti, TJ = (ny + tj - 1) / tj, TK = (nz + tk - 1) / tk;
    if (TI * TJ * TK >= (1ll << 25)) return {};                // positions are int32: 64 * tiles < 2^31
    std::vector<int32_t> cl((size_t)rows);
    for (int r = 0; r < rows; ++r) {
        const long long i = r % nx, j = (r / nx) % ny, k = r / (nx * ny);
        cl[(size_t)r] = (int32_t)(((k / tk) * TJ + j / tj) * TI + i / ti);
    }
    *nclusters = (int)(TI * TJ * TK);
    if (chain_len) *chain_len = (int)TI;
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
                  const std::vector<int32_t>& cl, int ncl, int width, SweepLayout* out, int chain_len = 1) {
    auto dep_begin = [&](int r) { return forward ? start[r] : diag[r] + 1; };
    auto dep_end = [&](int r) { return forward ? diag[r] : start[r + 1]; };
    // rows of every cluster, ascending
    std::vector<int32_t> cptr((size_t)ncl + 1, 0);
    for (int r = 0; r < rows; ++r) cptr[(size_t)cl[r] + 1]++;
    for (int a = 0; a < ncl; ++a) {
        if (cptr[(size_t)a + 1] > TILE) return false;
        cptr[(size_t)a + 1] += cptr[a];
    }
    std::vector<int32_t> crow((size_t)rows);
    {
        std::vector<int32_t> cur(cptr.begin(), cptr.end() - 1);
        for (int r = 0; r < rows; ++r) crow[(size_t)cur[cl[r]]++] = r;
    }
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
        }
    }
    if (out->chain_len == 1) {                                 // tiles in level order (stable in the cluster id; the backward sweep runs the ids downwards)
        std::vector<int32_t> lptr((size_t)out->levels + 1, 0);
        for (int a = 0; a < ncl; ++a) lptr[(size_t)level[a] + 1]++;
        for (int l = 0; l < out->levels; ++l) lptr[(size_t)l + 1] += lptr[l];
        std::vector<int32_t> cur(lptr.begin(), lptr.end() - 1);
        if (forward) { for (int a = 0; a < ncl; ++a) tile_of[a] = cur[level[a]]++; }
        else { for (int a = ncl - 1; a >= 0; --a) tile_of[a] = cur[level[a]]++; }
    }
    out->order.assign((size_t)ncl * TILE, -1);
    out->where.assign((size_t)rows, 0);
    out->steps.assign((size_t)ncl * TILE + (size_t)ncl, 255);    // [tiles * 64] step of every row, then [tiles] number of steps
    std::vector<int8_t> ilev((size_t)rows, 0);
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
        out->steps[(size_t)ncl * TILE + t] = (uint8_t)nl;       // a step = an internal level (lane l solves rows l and l + 32)
        for (int i = 0; i < n; ++i) {
            out->steps[(size_t)t * TILE + i] = (uint8_t)ilev[tmp[i]];
            out->order[(size_t)t * TILE + i] = tmp[i];
            out->where[(size_t)tmp[i]] = t * TILE + i;
        }
    });
    if (bad) return false;
    // entries: [tile][slot][64], operand order = the reference's (ascending columns forward, descending backward);
    // push lists: where inside the tile's operand staging a row's result has to go (bytes 0..2; a byte that is not needed
    // points at the row's own first slot, which nobody reads once the row is solved)
    out->ecol.assign((size_t)ncl * width * TILE, -1);
    out->eidx.assign((size_t)ncl * width * TILE, -1);
    out->push.assign((size_t)ncl * TILE, 0u);
    for_clusters(ncl, [&](int a) {                             // a row is only pushed to rows of its own tile: no sharing between tiles
        const int t = tile_of[a];
        uint8_t npush[TILE] = {0};
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
                }
            }
        }
    });
    if (bad) return false;
    return true;
}

template <class V>
int upload(const std::vector<V>& h, V** d) {
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
    const std::vector<int32_t> cl = propose_grid_tiles(rows, start, pos, &ncl, &chain_len);
    if (cl.empty()) return false;
    if (const char* e = getenv("SMM_B200_SGS_CHAINS")) { if (atoi(e) == 0) chain_len = 1; }   // A/B: tiles in tile-level order
    SweepLayout L[2];                                          // the two sweeps are laid out side by side (set-up time)
    std::future<bool> bwd = std::async(std::launch::async, [&] { return layout_sweep(false, rows, start, pos, diag, cl, ncl, width, &L[1], chain_len); });
    const bool fwd_ok = layout_sweep(true, rows, start, pos, diag, cl, ncl, width, &L[0], chain_len);
    if (!bwd.get() || !fwd_ok) return false;
    std::vector<int32_t> yp(L[1].order.size(), 0);
    for (size_t t = 0; t < yp.size(); ++t) if (L[1].order[t] >= 0) yp[t] = L[0].where[(size_t)L[1].order[t]];
    const size_t npos = (size_t)ncl * TILE;
    p->threads_fwd = p->threads_bwd = (long long)npos;
    p->tile_width = width;
    p->tile_levels[0] = L[0].levels;
    p->tile_levels[1] = L[1].levels;
    p->tile_chain[0] = L[0].chain_len;
    p->tile_chain[1] = L[1].chain_len;
    bool ok = upload(L[0].order, &p->order_fwd) == SMM_OK && upload(L[1].order, &p->order_bwd) == SMM_OK && upload(yp, &p->ypos) == SMM_OK &&
              cudaMalloc(&p->yperm, sizeof(float) * npos) == cudaSuccess && cudaMalloc(&p->xperm, sizeof(float) * npos) == cudaSuccess;
    for (int w = 0; w < 2 && ok; ++w) {
        p->esize[w] = (long long)L[w].ecol.size();
        ok = upload(L[w].ecol, &p->ecol[w]) == SMM_OK && upload(L[w].eidx, &p->eidx[w]) == SMM_OK && upload(L[w].steps, &p->tile_steps[w]) == SMM_OK &&
             upload(L[w].push, &p->tile_push[w]) == SMM_OK && cudaMalloc(&p->eval[w], sizeof(float) * L[w].ecol.size()) == cudaSuccess &&
             cudaMalloc(&p->dval[w], sizeof(float) * npos) == cudaSuccess;
    }
    if (!ok) {                                                  // out of memory: leave the handle as it was, the row-level schedule needs less
        cudaGetLastError();
        cudaFree(p->order_fwd); cudaFree(p->order_bwd); cudaFree(p->ypos); cudaFree(p->yperm); cudaFree(p->xperm);
        p->order_fwd = p->order_bwd = p->ypos = nullptr;
        p->yperm = p->xperm = nullptr;
        for (int w = 0; w < 2; ++w) {
            cudaFree(p->ecol[w]); cudaFree(p->eidx[w]); cudaFree(p->tile_steps[w]); cudaFree(p->tile_push[w]); cudaFree(p->eval[w]); cudaFree(p->dval[w]);
            p->ecol[w] = p->eidx[w] = nullptr; p->tile_steps[w] = nullptr; p->tile_push[w] = nullptr; p->eval[w] = p->dval[w] = nullptr;
            p->esize[w] = 0;
        }
        p->threads_fwd = p->threads_bwd = 0;
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
    TileArgs F{nsf, p->tile_steps[0], p->tile_push[0], p->order_fwd, nullptr, p->ecol[0], p->eval[0], p->dval[0], ntiles, ntiles / cf, cf, p->tile_width, sleep_first, sleep_later, trace};
    TileArgs B{nsb, p->tile_steps[1], p->tile_push[1], p->order_bwd, p->ypos, p->ecol[1], p->eval[1], p->dval[1], ntiles, ntiles / cb, cb, p->tile_width, sleep_first, sleep_later, nullptr};
    launch_tiles<4>(p, F, B, rhs_dev, x_dev, state, cap, s);   // 1, 2 and 8 tiles per CTA claim measured the same
    SMM_CUDA(cudaGetLastError());
    if (trace) {
        std::vector<unsigned long long> h(4 * (size_t)ntiles);
        SMM_CUDA(cudaStreamSynchronize(s));
        SMM_CUDA(cudaMemcpy(h.data(), trace, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost));
        if (FILE* f = fopen(trace_path, "wb")) { fwrite(h.data(), sizeof(unsigned long long), h.size(), f); fclose(f); }
    }
    return SMM_OK;
}
