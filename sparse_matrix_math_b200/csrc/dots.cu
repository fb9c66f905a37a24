// dots.cu -- stand-alone dot products in the reference's summation orders (parity modes), plus the C-ABI smm_dot.
//
//   SMM_REDUCE_REFERENCE_TREE   Vector::operator* of the SMM_MULTITHREADING build (H:308-320):
//       tbb::parallel_deterministic_reduce over blocked_range<int>(0,n,8192), identity 0.0f, std::plus.
//       The range is halved at begin+(end-begin)/2 while its size exceeds the grain; each leaf is summed left to
//       right from 0; joins are left+right.  All nodes at depth D' = min{d : floor(n/2^d) <= 8192} exist, and each
//       is either a leaf or (size 8193) splits exactly once more, so: one WARP per depth-D' node walks down from
//       the root to find its range, sums its one or two leaves (coalesced loads by all lanes, the sequential adds by
//       lane 0 out of shared memory), and the 2^D' node values are then combined by a perfect pairwise tree -- the
//       same additions in the same order as the reference.
//   SMM_REDUCE_REFERENCE_SERIAL the serial build (H:322-326): left to right.  One thread adds, the rest of its CTA
//       streams products into shared memory ahead of it.
//   SMM_REDUCE_FAST             vecops.cu's fused two-stage reduction (VEC_DOT2).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "epilogue.cuh"
#include "smm_internal.cuh"

namespace {

constexpr int TBB_GRAIN = 8192;

constexpr int LEAF_CHUNK = 512;        // products staged per round: 16 per lane
constexpr int TREE_WARPS = 4;

// One warp sums one leaf: all lanes stream a[], b[] with coalesced loads (the next chunk is already in flight in
// registers), multiply, and park the products in shared memory; lane 0 then adds them left to right from 0, which
// is the order of the reference's leaf loop (H:312-316).  The 4-cycle FADD chain of lane 0 (8192 adds per leaf) is
// the critical path; the loads hide under it.
__device__ __forceinline__ float leaf_sum(const float* __restrict__ a, const float* __restrict__ b, long long lo, long long hi, float* buf) {
    const int lane = threadIdx.x & 31;
    float cur = 0.0f;                                        // identity, H:312
    float ra[LEAF_CHUNK / 32], rb[LEAF_CHUNK / 32];
    auto fetch = [&](long long base) {
#pragma unroll
        for (int k = 0; k < LEAF_CHUNK / 32; ++k) {
            const long long j = base + k * 32 + lane;
            const bool in = j < hi;
            ra[k] = in ? __ldg(a + j) : 0.0f;
            rb[k] = in ? __ldg(b + j) : 0.0f;
        }
    };
    fetch(lo);
    for (long long base = lo; base < hi; base += LEAF_CHUNK) {
#pragma unroll
        for (int k = 0; k < LEAF_CHUNK / 32; ++k) buf[k * 32 + lane] = __fmul_rn(ra[k], rb[k]);
        __syncwarp();
        if (base + LEAF_CHUNK < hi) fetch(base + LEAF_CHUNK);
        if (lane == 0) {
            const long long left = hi - base;
            const int m = left < LEAF_CHUNK ? (int)left : LEAF_CHUNK;
            int t = 0;
            for (; t + 4 <= m; t += 4) {
                const float4 q = *reinterpret_cast<const float4*>(buf + t);
                cur = __fadd_rn(cur, q.x); cur = __fadd_rn(cur, q.y); cur = __fadd_rn(cur, q.z); cur = __fadd_rn(cur, q.w);   // H:314-316
            }
            for (; t < m; ++t) cur = __fadd_rn(cur, buf[t]);
        }
        __syncwarp();
    }
    return cur;                                              // valid in lane 0
}

struct TreeParams {
    long long n;
    int depth;             // D'
    int ndots;             // 1 or 2
    const float* a[2];
    const float* b[2];
    float* nodes;          // [2 dots][2 ping-pong][2^D']
    unsigned int* ticket;
    SolveState* state;
    int finish;
    float* out_dev;        // optional: totals written here too
    float* upd;            // dot_tree_rows_kernel<R, true>: r (== a[0]) is first replaced by r - alpha * b[0], the dot is r.r of the new r
};

// every CTA has stored its depth-D' node values: the last one to arrive combines them by a perfect pairwise tree (ping-pong
// between the two halves of each dot's node buffer) -- the joins of the reference's reduction, left + right -- and hands
// the totals to the scalar step.  Called by all threads of the CTA.
__device__ __forceinline__ void tree_finish(const TreeParams& P) {
    __shared__ int sh_last;
    const long long nn = 1ll << P.depth;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) sh_last = (atomicAdd(P.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!sh_last) return;
    __threadfence();
    float totals[2] = {0.f, 0.f};
    for (int d = 0; d < P.ndots; ++d) {
        float* cur = P.nodes + (size_t)d * 2 * nn;
        float* nxt = cur + nn;
        for (long long w = nn >> 1; w >= 1; w >>= 1) {
            for (long long k = threadIdx.x; k < w; k += blockDim.x) {
                const float2 pr = __ldcg(reinterpret_cast<const float2*>(cur) + k);
                nxt[k] = __fadd_rn(pr.x, pr.y);
            }
            __threadfence_block();
            __syncthreads();
            float* t = cur; cur = nxt; nxt = t;
        }
        totals[d] = __ldcg(cur);
    }
    if (threadIdx.x == 0) {
        *P.ticket = 0u;
        if (P.out_dev) { P.out_dev[0] = totals[0]; P.out_dev[1] = totals[1]; }
        if (P.state) smm_finish(P.finish, P.state, totals[0], totals[1]);
    }
}

__global__ void __launch_bounds__(TREE_WARPS * 32) dot_tree_kernel(const TreeParams P) {
    if (P.state != nullptr && P.state->done) return;
    __shared__ __align__(16) float sh_buf[TREE_WARPS][LEAF_CHUNK];
    const long long nn = 1ll << P.depth;
    const int warp = threadIdx.x >> 5;
    const long long job = (long long)blockIdx.x * TREE_WARPS + warp;     // (dot, depth-D' node)
    if (job < nn * P.ndots) {
        const int d = (int)(job / nn);
        const long long i = job - (long long)d * nn;
        long long lo = 0, hi = P.n;
        for (int level = P.depth - 1; level >= 0; --level) {
            const long long mid = lo + (hi - lo) / 2;
            if ((i >> level) & 1) lo = mid; else hi = mid;
        }
        float v;
        if (hi - lo > TBB_GRAIN) {
            const long long mid = lo + (hi - lo) / 2;
            const float l = leaf_sum(P.a[d], P.b[d], lo, mid, sh_buf[warp]);
            v = __fadd_rn(l, leaf_sum(P.a[d], P.b[d], mid, hi, sh_buf[warp]));
        } else {
            v = leaf_sum(P.a[d], P.b[d], lo, hi, sh_buf[warp]);
        }
        if ((threadIdx.x & 31) == 0) P.nodes[(size_t)d * 2 * nn + i] = v;
    }
    tree_finish(P);
}

// ---------------------------------------------------------------------------------------------------
// Long vectors: one LANE per depth-D' node, R nodes per warp.
//
// dot_tree_kernel spends one warp on a node and has lane 0 add its 8192 products: about 1.4 warp instructions per element,
// which is what bounds it (377 us for 134 M elements; the fused fast-mode dot streams the same bytes in 193 us).  Here a warp
// takes R consecutive nodes; their windows of a[] and b[] are brought in by cp.async (16-byte chunks, three stages in
// flight per warp, rows of 1024 / R elements padded by 4 so that the per-lane 128-bit reads are conflict-free), and lane l
// adds the products of node l left to right from 0 -- the reference's leaf loop (H:312-316), R chains per warp instead of
// one, about 0.2 warp instructions per element for R = 16.  A node of 8193 elements is two leaves (H:308-320: the range is
// halved while it exceeds the grain): the lane restarts from 0 at the split point and joins left + right at the end.
// ---------------------------------------------------------------------------------------------------
constexpr int ROWS_TILE = 1024;        // elements per stage and array, all R rows together (the wide form; the narrow forms stage 64 per row)
constexpr int ROWS_STAGES = 3;
#ifndef SMM_DOT_NARROW_DEFAULT
#define SMM_DOT_NARROW_DEFAULT 0     // 0: wide form; 8 / 4: narrow form on long vectors (see tree_rows_narrow)
#endif

__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src, int src_bytes) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}

// UPD (ConjugateGradient in the reference-order mode): the r update of H:2366-2368 rides on the dot that follows it -- the
// lane that adds node l's squares first forms r_i = fma(-alpha, Ap_i, r_i) (two roundings, like vecops.cu's FCgR) from the
// staged windows of r and Ap, leaves the new r in the staging buffer, and the warp stores the finished stage back with
// coalesced 16-byte stores (only the elements of its own nodes: a window starts at the 16-byte boundary below its node).
// One pass over r and Ap instead of an update kernel (12 n bytes) plus a dot (4 n).
template <int R, int TILE, bool UPD>
__global__ void __launch_bounds__(TREE_WARPS * 32) dot_tree_rows_kernel(const TreeParams P) {
    if (P.state != nullptr && P.state->done) return;
    extern __shared__ __align__(16) float rows_smem[];
    constexpr int C = TILE / R, CP = C + 4, ARR = R * CP;                    // elements per row, padded row, one array of a stage
    constexpr int CHUNKS_PER_ROW = C / 4;
    const long long nn = 1ll << P.depth;
    const long long groups = nn / R;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* const mine = rows_smem + (size_t)warp * ROWS_STAGES * 2 * ARR;
    const long long job = (long long)blockIdx.x * TREE_WARPS + warp;         // (dot, group of R nodes)
    if (job < groups * P.ndots) {
        const int d = (int)(job / groups);
        const long long node = (job - (long long)d * groups) * R + (lane % R);
        const float* a = P.a[d];                                               // UPD: also written (through P.upd)
        const float* __restrict__ b = P.b[d];
        const float nalpha = UPD ? -P.state->alpha : 0.0f;
        long long lo = 0, hi = P.n;
        for (int level = P.depth - 1; level >= 0; --level) {
            const long long mid = lo + (hi - lo) / 2;
            if ((node >> level) & 1) lo = mid; else hi = mid;
        }
        const long long split = hi - lo > TBB_GRAIN ? lo + (hi - lo) / 2 : -1;   // two leaves
        const long long a4 = lo & ~3ll;                                        // 16-byte aligned start of the node's window
        int nstages = (int)((hi - a4 + C - 1) / C);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nstages = max(nstages, __shfl_xor_sync(0xFFFFFFFFu, nstages, o));
        auto issue = [&](const int t) {
            if (t < nstages) {
                float* sa = mine + (size_t)(t % ROWS_STAGES) * 2 * ARR;
                float* sb = sa + ARR;
#pragma unroll
                for (int k = 0; k < (R * CHUNKS_PER_ROW) / 32; ++k) {
                    const int q = lane + 32 * k, row = q / CHUNKS_PER_ROW, col = (q % CHUNKS_PER_ROW) * 4;
                    const long long g0 = __shfl_sync(0xFFFFFFFFu, a4, row) + (long long)t * C + col;
                    const long long left = P.n - g0;
                    if (left > 0) {
                        const int bytes = left >= 4 ? 16 : (int)left * 4;
                        cp_async16(sa + row * CP + col, a + g0, bytes);
                        cp_async16(sb + row * CP + col, b + g0, bytes);
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        issue(0);
        issue(1);
        float acc = 0.0f, first = 0.0f;                                       // identity, H:312
        for (int t = 0; t < nstages; ++t) {
            issue(t + 2);
            asm volatile("cp.async.wait_group 2;" ::: "memory");
            __syncwarp();
            if (lane < R) {
                float* sa = mine + (size_t)(t % ROWS_STAGES) * 2 * ARR + lane * CP;
                const float* sb = sa + ARR;
                const long long g0 = a4 + (long long)t * C;
                if (g0 >= lo && g0 + C <= hi && !(split >= g0 && split < g0 + C)) {
#pragma unroll 4
                    for (int c = 0; c < C; c += 4) {
                        float4 va = *reinterpret_cast<const float4*>(sa + c);
                        float4 vb = *reinterpret_cast<const float4*>(sb + c);
                        if (UPD) {                                                  // r = fma(-alpha, Ap, r), then r.r
                            va.x = smm_fma2(nalpha, vb.x, va.x); va.y = smm_fma2(nalpha, vb.y, va.y);
                            va.z = smm_fma2(nalpha, vb.z, va.z); va.w = smm_fma2(nalpha, vb.w, va.w);
                            *reinterpret_cast<float4*>(sa + c) = va;
                            vb = va;
                        }
                        acc = __fadd_rn(acc, __fmul_rn(va.x, vb.x)); acc = __fadd_rn(acc, __fmul_rn(va.y, vb.y));   // H:314-316
                        acc = __fadd_rn(acc, __fmul_rn(va.z, vb.z)); acc = __fadd_rn(acc, __fmul_rn(va.w, vb.w));
                    }
                } else {
                    for (int c = 0; c < C; ++c) {
                        const long long g = g0 + c;
                        if (g < lo || g >= hi) continue;
                        if (g == split) { first = acc; acc = 0.0f; }
                        float va = sa[c], vb = sb[c];
                        if (UPD) { va = smm_fma2(nalpha, vb, va); sa[c] = va; vb = va; }
                        acc = __fadd_rn(acc, __fmul_rn(va, vb));
                    }
                }
            }
            __syncwarp();
            if (UPD) {                                                              // the finished stage of r goes back, every node its own elements
                const float* sa = mine + (size_t)(t % ROWS_STAGES) * 2 * ARR;
#pragma unroll
                for (int k = 0; k < (R * CHUNKS_PER_ROW) / 32; ++k) {
                    const int q = lane + 32 * k, row = q / CHUNKS_PER_ROW, col = (q % CHUNKS_PER_ROW) * 4;
                    const long long g0 = __shfl_sync(0xFFFFFFFFu, a4, row) + (long long)t * C + col;
                    const long long rlo = __shfl_sync(0xFFFFFFFFu, lo, row), rhi = __shfl_sync(0xFFFFFFFFu, hi, row);
                    const float4 v = *reinterpret_cast<const float4*>(sa + row * CP + col);
                    if (g0 >= rlo && g0 + 4 <= rhi) *reinterpret_cast<float4*>(P.upd + g0) = v;
                    else {
                        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int c = 0; c < 4; ++c) if (g0 + c >= rlo && g0 + c < rhi) P.upd[g0 + c] = e[c];
                    }
                }
                __syncwarp();
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (lane < R) P.nodes[(size_t)d * 2 * nn + node] = split >= 0 ? __fadd_rn(first, acc) : acc;
    }
    tree_finish(P);
}

struct SerialParams {
    long long n;
    int ndots;
    const float* a[2];
    const float* b[2];
    const float* pre[2];   // total already computed by sum_squares_serial_kernel (a == b), or null
    SolveState* state;
    int finish;
    float* out_dev;
};

constexpr int SER_THREADS = 256;
constexpr int SER_CHUNK = 4096;

__global__ void __launch_bounds__(SER_THREADS) dot_serial_kernel(const SerialParams P) {
    if (P.state != nullptr && P.state->done) return;
    __shared__ __align__(16) float buf[2][SER_CHUNK];
    float totals[2] = {0.f, 0.f};
    for (int d = 0; d < P.ndots; ++d) {
        if (P.pre[d] != nullptr) { totals[d] = *P.pre[d]; continue; }   // uniform across the CTA
        const float* a = P.a[d];
        const float* b = P.b[d];
        float cur = 0.0f;
        const long long nchunks = (P.n + SER_CHUNK - 1) / SER_CHUNK;
        // prologue: chunk 0
        for (int t = threadIdx.x; t < SER_CHUNK; t += SER_THREADS) {
            const long long j = t;
            buf[0][t] = j < P.n ? __fmul_rn(a[j], b[j]) : 0.0f;
        }
        __syncthreads();
        for (long long c = 0; c < nchunks; ++c) {
            const int pb = (int)(c & 1);
            if (threadIdx.x >= 32) {                          // warps 1..7 prefetch the next chunk
                if (c + 1 < nchunks) {
                    for (int t = threadIdx.x - 32; t < SER_CHUNK; t += SER_THREADS - 32) {
                        const long long j = (c + 1) * SER_CHUNK + t;
                        buf[pb ^ 1][t] = j < P.n ? __fmul_rn(a[j], b[j]) : 0.0f;
                    }
                }
            } else if (threadIdx.x == 0) {                    // one thread adds left to right, H:322-326
                const long long left = P.n - c * SER_CHUNK;
                const int m = left < SER_CHUNK ? (int)left : SER_CHUNK;
                int t = 0;
                for (; t + 4 <= m; t += 4) {                  // one 128-bit load per four additions: the FADD chain is the only cost
                    const float4 q = *reinterpret_cast<const float4*>(&buf[pb][t]);
                    cur = __fadd_rn(cur, q.x); cur = __fadd_rn(cur, q.y); cur = __fadd_rn(cur, q.z); cur = __fadd_rn(cur, q.w);
                }
                for (; t < m; ++t) cur = __fadd_rn(cur, buf[pb][t]);
            }
            __syncthreads();
        }
        totals[d] = cur;
    }
    if (threadIdx.x == 0) {
        if (P.out_dev) { P.out_dev[0] = totals[0]; P.out_dev[1] = totals[1]; }
        if (P.state) smm_finish(P.finish, P.state, totals[0], totals[1]);
    }
}

// ---------------------------------------------------------------------------------------------------
// Left-to-right float sum of SQUARES, exactly, in parallel.
//
// The reference adds ||r||^2 serially in both of its builds (BiCGStab, H:2262-2267), and its serial build adds every
// dot product that way; one thread doing the same costs 4 cycles per element (the FADD chain).  When every term is a
// square the running sum s only grows, and while it stays inside one binade [2^e, 2^(e+1)) adding a term is integer
// arithmetic on its significand m (s = m ulp): with p = (k + f) ulp, RN(s + p) = (m + k) ulp rounded up when f > 1/2,
// or when f = 1/2 and m + k is odd (ties to even).  So a term is a map m -> m + inc[parity of m], and such maps
// compose associatively: a thread folds its 8 terms into one map, a cluster of 8 CTAs scans the maps (warp shuffles,
// shared memory, then DSMEM between the CTAs), and the first thread at which m would reach 2^24 (the sum leaves the
// binade; also any infinity / NaN term) replays its own terms with real float additions, which hands the next
// round its binade; the 64 K-term window is then re-scanned in place.  Everything before that thread is exact by
// construction, so the result has the reference's bits for any input.
// ---------------------------------------------------------------------------------------------------
constexpr int SQ_THREADS = 1024;
constexpr int SQ_CTAS = 8;                                   // one thread-block cluster: the CTAs trade their maps over DSMEM
constexpr int SQ_EPT = 8;                                    // terms per thread and window
constexpr int SQ_WINDOW = SQ_CTAS * SQ_THREADS * SQ_EPT;
constexpr int SQ_SMEM_BYTES = SQ_THREADS * (SQ_EPT + 1) * (int)sizeof(unsigned int);
constexpr unsigned int SQ_CAP = 1u << 26;                    // "leaves the binade" (saturating; parities beyond it are meaningless)
constexpr unsigned int SQ_NONE = 0xFFFFFFFFu;

struct Inc2 { unsigned int e, o; };                          // increment of the significand when it is even / odd

__device__ __forceinline__ Inc2 inc_then(const Inc2 a, const Inc2 b) {      // first a, then b
    Inc2 c;
    c.e = min(a.e + ((a.e & 1u) ? b.o : b.e), SQ_CAP);
    c.o = min(a.o + ((a.o & 1u) ? b.e : b.o), SQ_CAP);
    return c;
}

// p >= 0 given by its bits; e_eff = max(biased exponent of the running sum, 1).  Branch-free: with p = P 2^-sh ulp,
// k = P >> sh, and the discarded bits decide the rounding; sh >= 25 gives k = 0 and less than half an ulp by itself.
__device__ __forceinline__ Inc2 term_inc(const unsigned int pbits, const int e_eff) {
    const unsigned int ep = (pbits >> 23) & 0xFFu, mant = pbits & 0x7FFFFFu;
    const unsigned int P = ep ? (mant | 0x800000u) : mant;
    const int sh_raw = e_eff - (int)(ep ? ep : 1u);
    const unsigned int sh = (unsigned int)min(max(sh_raw, 0), 31);
    const unsigned int k = P >> sh, rem = P & ((1u << sh) - 1u), half = (1u << sh) >> 1;
    const unsigned int up = rem > half ? 1u : 0u;
    const unsigned int tie = (rem == half && sh != 0u) ? 1u : 0u;          // ties go to the even significand
    Inc2 r;
    r.e = k + up + (tie & k);
    r.o = k + up + (tie & ~k);
    // infinity / NaN terms and terms that alone exceed the binade are left to real arithmetic
    if (ep == 255u || (sh_raw < 0 && P != 0u)) r.e = r.o = SQ_CAP;
    return r;
}

struct SquaresParams {
    long long n;
    const float* r;
    float* out_dev;        // optional: the total
    SolveState* state;     // optional: smm_finish(finish, state, total, 0)
    int finish;
};

__global__ void __cluster_dims__(SQ_CTAS, 1, 1) __launch_bounds__(SQ_THREADS, 1) sum_squares_serial_kernel(const SquaresParams P) {
    // (every CTA of the cluster takes the same branch: the cluster barriers below stay matched)
    if (P.state != nullptr && P.state->done) return;
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ unsigned int sq_terms[];               // [SQ_THREADS][SQ_EPT + 1] products (bits)
    __shared__ Inc2 sh_warp[32];
    __shared__ Inc2 sh_cta[SQ_CTAS];                         // every CTA's map, written into every CTA (DSMEM)
    __shared__ unsigned int sh_first[SQ_CTAS];               // every CTA's first thread that leaves the binade
    __shared__ unsigned int sh_local_first, sh_bits, sh_flag;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)cluster.block_rank();
    const unsigned int g = (unsigned int)rank * SQ_THREADS + tid;              // thread index inside the window
    unsigned int sbits = 0u;                                 // the running sum (H:2262: res = 0)
    // Multi-GPU: the sum runs through the ranks in order -- this rank's rows continue the sum of the ranks before it
    DistComm* const chain = (P.state != nullptr) ? P.state->comm : nullptr;
    if (chain != nullptr && chain->rank > 0) {
        if (rank == 0 && tid == 0) {
            const unsigned int want = chain->chain_seq + 1u;
            unsigned long long w = 0ull;
            unsigned int polls = 0;
            for (;;) {
                w = ld_sys_u64(chain->chain[chain->rank]);
                if ((unsigned int)(w >> 32) == want) break;
                if (++polls >= SMM_DIST_POLL_LIMIT) { chain->error = 1; break; }
            }
            for (int c = 0; c < SQ_CTAS; ++c) *cluster.map_shared_rank(&sh_bits, c) = (unsigned int)w;
        }
        cluster.sync();
        sbits = sh_bits;
        cluster.sync();                                      // sh_bits is written again inside the loop
    }
    long long base = 0;
    bool open_ended = false;                                 // the sum has become infinite or NaN
    // values of this CTA's part of the window, coalesced, all loads of a thread in flight; the NEXT window is requested
    // as soon as this one has been parked in shared memory
    float v[SQ_EPT];
    auto fetch = [&](long long from) {
#pragma unroll
        for (int it = 0; it < SQ_EPT; ++it) {
            const long long j = from + (long long)rank * (SQ_THREADS * SQ_EPT) + it * SQ_THREADS + tid;
            v[it] = j < P.n ? __ldg(P.r + j) : 0.0f;
        }
    };
    fetch(0);
    while (base < P.n && !open_ended) {
        // products parked so that thread t finds its terms in consecutive banks
#pragma unroll
        for (int it = 0; it < SQ_EPT; ++it) {
            const int w = it * SQ_THREADS + tid;
            sq_terms[(w / SQ_EPT) * (SQ_EPT + 1) + (w % SQ_EPT)] = __float_as_uint(__fmul_rn(v[it], v[it]));
        }
        __syncthreads();
        if (base + SQ_WINDOW < P.n) fetch(base + SQ_WINDOW);
        unsigned int done = 0;                               // threads [0, done) of this window are already in the sum
        while (done < (unsigned int)(SQ_CTAS * SQ_THREADS)) {
            const unsigned int se = (sbits >> 23) & 0xFFu;
            if (se == 255u) { open_ended = true; break; }    // infinity or NaN: see below
            const int e_eff = se ? (int)se : 1;
            const unsigned int m = se ? ((sbits & 0x7FFFFFu) | 0x800000u) : sbits;
            if (tid == 0) sh_local_first = SQ_NONE;
            Inc2 mine = {0u, 0u};
            if (g >= done) {
#pragma unroll
                for (int j = 0; j < SQ_EPT; ++j) mine = inc_then(mine, term_inc(sq_terms[tid * (SQ_EPT + 1) + j], e_eff));
            }
            // scan of the maps in thread order: warp, CTA, cluster
            Inc2 incl = mine;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                Inc2 prev;
                prev.e = __shfl_up_sync(0xFFFFFFFFu, incl.e, off);
                prev.o = __shfl_up_sync(0xFFFFFFFFu, incl.o, off);
                if (lane >= off) incl = inc_then(prev, incl);
            }
            if (lane == 31) sh_warp[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                Inc2 w = sh_warp[lane];
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    Inc2 prev;
                    prev.e = __shfl_up_sync(0xFFFFFFFFu, w.e, off);
                    prev.o = __shfl_up_sync(0xFFFFFFFFu, w.o, off);
                    if (lane >= off) w = inc_then(prev, w);
                }
                sh_warp[lane] = w;                           // inclusive over warps
                Inc2 total;                                  // this CTA's map, handed to every CTA of the cluster
                total.e = __shfl_sync(0xFFFFFFFFu, w.e, 31);
                total.o = __shfl_sync(0xFFFFFFFFu, w.o, 31);
                if (lane < SQ_CTAS) *cluster.map_shared_rank(&sh_cta[rank], lane) = total;
            }
            cluster.sync();
            Inc2 before = {0u, 0u};                          // everything ahead of this thread: CTAs, warps, lanes
            for (int c = 0; c < rank; ++c) before = inc_then(before, sh_cta[c]);
            if (warp > 0) before = inc_then(before, sh_warp[warp - 1]);
            {
                Inc2 prev;
                prev.e = __shfl_up_sync(0xFFFFFFFFu, incl.e, 1);
                prev.o = __shfl_up_sync(0xFFFFFFFFu, incl.o, 1);
                if (lane > 0) before = inc_then(before, prev);
            }
            const unsigned int excl = (m & 1u) ? before.o : before.e;
            const unsigned int own = ((m + excl) & 1u) ? mine.o : mine.e;
            const bool leaves = excl >= SQ_CAP || own >= SQ_CAP || m + excl + own >= (1u << 24);
            if (leaves) atomicMin(&sh_local_first, g);
            __syncthreads();
            if (tid < SQ_CTAS) *cluster.map_shared_rank(&sh_first[rank], tid) = sh_local_first;
            cluster.sync();
            unsigned int first = SQ_NONE;
#pragma unroll
            for (int c = 0; c < SQ_CTAS; ++c) first = min(first, sh_first[c]);
            unsigned int result = 0u;
            bool writer = false;
            if (first == SQ_NONE) {                          // the rest of the window stays inside the binade
                if (g == (unsigned int)(SQ_CTAS * SQ_THREADS - 1)) {
                    const unsigned int m2 = m + excl + own;  // < 2^24
                    result = m2 >= 0x800000u ? (((unsigned int)e_eff << 23) | (m2 & 0x7FFFFFu)) : m2;
                    writer = true;
                }
                done = SQ_CTAS * SQ_THREADS;
            } else {
                if (g == first) {                            // real additions from the exact sum ahead of this thread
                    const unsigned int m2 = m + excl;        // < 2^24
                    float cur = __uint_as_float(m2 >= 0x800000u ? (((unsigned int)e_eff << 23) | (m2 & 0x7FFFFFu)) : m2);
                    for (int j = 0; j < SQ_EPT; ++j) cur = __fadd_rn(cur, __uint_as_float(sq_terms[tid * (SQ_EPT + 1) + j]));   // H:2266 (absent terms are +0)
                    result = __float_as_uint(cur);
                    writer = true;
                }
                done = first + 1u;
            }
            if (writer) {
#pragma unroll
                for (int c = 0; c < SQ_CTAS; ++c) *cluster.map_shared_rank(&sh_bits, c) = result;
            }
            cluster.sync();
            sbits = sh_bits;
        }
        if (!open_ended) base += SQ_WINDOW;
        else base += (long long)done * SQ_EPT;
        __syncthreads();                                     // sq_terms is refilled
    }
    cluster.sync();                                          // nobody leaves while a peer may still write into it
    if (rank != 0) return;
    // +infinity stays +infinity unless a NaN term follows (NaN stays NaN): r*r is NaN only for a NaN r
    if (((sbits >> 23) & 0xFFu) == 255u && (sbits & 0x7FFFFFu) == 0u && base < P.n) {
        if (tid == 0) sh_flag = 0u;
        __syncthreads();
        bool nan = false;
        for (long long j = base + tid; j < P.n; j += SQ_THREADS) { const float x = __ldg(P.r + j); nan |= x != x; }
        if (nan) sh_flag = 1u;
        __syncthreads();
        if (sh_flag) sbits = 0x7FFFFFFFu;
    }
    if (tid == 0) {
        float total = __uint_as_float(sbits);
        if (chain != nullptr) {
            const unsigned int seq = chain->chain_seq + 1u;
            if (chain->rank + 1 < chain->nranks) {             // hand the running sum on; the total comes back through the all-reduce
                st_sys_u64(chain->chain[chain->rank + 1], ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(total));
                total = 0.0f;
            }
            chain->chain_seq = seq;
        }
        if (P.out_dev) P.out_dev[0] = total;
        if (P.state) smm_finish(P.finish, P.state, total, 0.0f);
    }
}

typedef smm_dot_scratch DotScratch;
DotScratch g_scratch[SMM_MAX_DEVICES];     // stand-alone smm_dot (no handle): one per device, used under g_scratch_mu
std::mutex g_scratch_mu;
}  // namespace
int smm_tree_depth(long long n);
namespace {

int scratch_for(long long nn, smm_workspace* ws, DotScratch** out) {
    int dev = 0;
    SMM_CUDA(cudaGetDevice(&dev));
    DotScratch& sc = ws ? ws->dot : g_scratch[dev % SMM_MAX_DEVICES];
    if (!sc.ticket) {
        SMM_CUDA(cudaMalloc(&sc.ticket, sizeof(unsigned int)));
        SMM_CUDA(cudaMemset(sc.ticket, 0, sizeof(unsigned int)));
        SMM_CUDA(cudaMalloc(&sc.out, 2 * sizeof(float)));
        SMM_CUDA(cudaMalloc(&sc.sq, 2 * sizeof(float)));
        SMM_CUDA(cudaFuncSetAttribute(sum_squares_serial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SQ_SMEM_BYTES));
    }
    const size_t need = (size_t)nn * 4;
    if (sc.nodes_cap < need) {
        cudaFree(sc.nodes);
        sc.nodes = nullptr; sc.nodes_cap = 0;
        SMM_CUDA(cudaMalloc(&sc.nodes, need * sizeof(float)));
        sc.nodes_cap = need;
    }
    *out = &sc;
    return SMM_OK;
}

}  // namespace

void smm_dot_scratch_free(smm_dot_scratch* sc) {
    cudaFree(sc->nodes); cudaFree(sc->ticket); cudaFree(sc->out); cudaFree(sc->sq);
    *sc = smm_dot_scratch();
}

// allocate the node scratch for vectors of length n now (cudaMalloc is not allowed while a stream is capturing)
int smm_dot_ref_prepare(long long n, smm_workspace* ws) {
    DotScratch* sc = nullptr;
    return scratch_for(1ll << smm_tree_depth(n), ws, &sc);
}

int smm_tree_depth(long long n) {
    int d = 0;
    while ((n >> d) > TBB_GRAIN) ++d;                        // floor(n / 2^d) <= grain
    return d;
}

// t0 = a0.b0 [, t1 = a1.b1] in the requested reference order; the totals go to smm_finish(finish, state, t0, t1)
// and/or out_dev[0..1].  mode: SMM_REDUCE_REFERENCE_TREE or SMM_REDUCE_REFERENCE_SERIAL.
namespace {
// lanes-per-node variant of the tree dot for long, 16-byte aligned vectors: R nodes per warp, or 0 when it does not apply
int tree_rows_R(long long n, int ndots, const float* a0, const float* b0, const float* a1, const float* b1) {
    const int depth = smm_tree_depth(n);
    const long long jobs = (1ll << depth) * ndots;
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(a0) | reinterpret_cast<uintptr_t>(b0) | reinterpret_cast<uintptr_t>(a1) | reinterpret_cast<uintptr_t>(b1)) & 15) == 0;
    static const bool rows_off = [] { const char* e = getenv("SMM_B200_DOT_ROWS"); return e && atoi(e) == 0; }();
    int R = 0;
    if (aligned16 && !rows_off) for (int r = 16; r >= 2 && !R; r >>= 1) if ((1ll << depth) >= r && jobs / r >= 592) R = r;
    return R;
}

// 0: the wide form with R rows; 8 / 4: the narrow form with that many rows per warp (long vectors only: every resident warp slot
// must still find a job).  SMM_B200_DOT_NARROW=0 / 8 / 4 forces the choice (measurements).
int tree_rows_narrow(int R, long long jobs) {
    static const int forced = [] { const char* e = getenv("SMM_B200_DOT_NARROW"); return e ? atoi(e) : -1; }();
    if (R != 16) return 0;
    if (forced == 0) return 0;
    if (forced == 8 || forced == 4) return jobs / forced >= 592 ? forced : 0;
    return SMM_DOT_NARROW_DEFAULT;
}

template <int R, int TILE, bool UPD>
int launch_rows(const TreeParams& P, long long jobs, cudaStream_t s) {
    const size_t smem = (size_t)TREE_WARPS * ROWS_STAGES * 2 * (TILE + 4 * R) * sizeof(float);
    {
        static bool attr_done[SMM_MAX_DEVICES] = {false};              // per instantiation and device
        int dev = 0;
        SMM_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lk(g_smm_attr_mu);
        if (!attr_done[dev % SMM_MAX_DEVICES]) {
            SMM_CUDA(cudaFuncSetAttribute(dot_tree_rows_kernel<R, TILE, UPD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_done[dev % SMM_MAX_DEVICES] = true;
        }
    }
    const unsigned grid = (unsigned)((jobs / R + TREE_WARPS - 1) / TREE_WARPS);
    dot_tree_rows_kernel<R, TILE, UPD><<<grid, TREE_WARPS * 32, smem, s>>>(P);
    return SMM_OK;
}
}  // namespace

// r = r - alpha * ap (alpha from the state) and t0 = r.r of the new r in the reference's tree order, in one pass: only in
// the lane-per-node form of the tree dot
bool smm_dot_ref_update_applies(int mode, long long n, const float* r, const float* ap) {
    const char* e = getenv("SMM_B200_DOT_UPDATE");            // read per call: the tests run both forms
    const bool off = e && atoi(e) == 0;
    return !off && mode == SMM_REDUCE_REFERENCE_TREE && tree_rows_R(n, 1, r, ap, r, ap) != 0;
}

int smm_launch_dot_ref(int mode, long long n, int ndots, const float* a0, const float* b0, const float* a1, const float* b1,
                       SolveState* state, int finish, float* out_dev, cudaStream_t s, smm_workspace* ws, float* update_r) {
    if (update_r != nullptr && !(mode == SMM_REDUCE_REFERENCE_TREE && ndots == 1 && update_r == a0 && state != nullptr && tree_rows_R(n, 1, a0, b0, a0, b0) != 0)) {
        smm_set_error("dot with a fused r update: not applicable here");
        return SMM_E_INVALID;
    }
    if (mode == SMM_REDUCE_REFERENCE_TREE) {
        TreeParams P;
        P.n = n; P.depth = smm_tree_depth(n); P.ndots = ndots;
        P.upd = update_r;
        P.a[0] = a0; P.b[0] = b0; P.a[1] = a1; P.b[1] = b1;
        DotScratch* sc = nullptr;
        SMM_TRY(scratch_for(1ll << P.depth, ws, &sc));
        P.nodes = sc->nodes; P.ticket = sc->ticket; P.state = state; P.finish = finish; P.out_dev = out_dev;
        const long long jobs = (1ll << P.depth) * ndots;       // (dot, depth-D' node)
        // long vectors: a lane per node, R nodes per warp (as many as still leave four warps for every SM)
        const int R = tree_rows_R(n, ndots, a0, b0, a1, b1);
        if (R) {
            // narrow form (64 elements per row and stage instead of 1024 / R): a quarter / half of the shared memory per warp, so
            // two / four times the resident warps -- the lane that adds a node's products issues ~5 instructions per element
            // and a warp is bound by that chain, not by the loads (ncu: 7 warps per SM, 0.21 instructions per cycle and warp)
            const int narrow = tree_rows_narrow(R, jobs);
            int rc = SMM_OK;
            if (update_r) {
                if (narrow == 8) rc = launch_rows<8, 512, true>(P, jobs, s);
                else if (narrow == 4) rc = launch_rows<4, 256, true>(P, jobs, s);
                else if (R == 16) rc = launch_rows<16, ROWS_TILE, true>(P, jobs, s);
                else if (R == 8) rc = launch_rows<8, ROWS_TILE, true>(P, jobs, s);
                else if (R == 4) rc = launch_rows<4, ROWS_TILE, true>(P, jobs, s);
                else rc = launch_rows<2, ROWS_TILE, true>(P, jobs, s);
            } else {
                if (narrow == 8) rc = launch_rows<8, 512, false>(P, jobs, s);
                else if (narrow == 4) rc = launch_rows<4, 256, false>(P, jobs, s);
                else if (R == 16) rc = launch_rows<16, ROWS_TILE, false>(P, jobs, s);
                else if (R == 8) rc = launch_rows<8, ROWS_TILE, false>(P, jobs, s);
                else if (R == 4) rc = launch_rows<4, ROWS_TILE, false>(P, jobs, s);
                else rc = launch_rows<2, ROWS_TILE, false>(P, jobs, s);
            }
            SMM_TRY(rc);
        } else
        dot_tree_kernel<<<(unsigned)((jobs + TREE_WARPS - 1) / TREE_WARPS), TREE_WARPS * 32, 0, s>>>(P);
    } else {
        // left to right.  Dots of a vector with itself are sums of squares: exact in parallel (sum_squares_serial_kernel)
        DotScratch* sc = nullptr;
        SMM_TRY(scratch_for(1, ws, &sc));
        const float* aa[2] = {a0, a1};
        const float* bb[2] = {b0, b1};
        if (ndots == 1 && a0 == b0) {
            SquaresParams Q{n, a0, out_dev, state, finish};
            sum_squares_serial_kernel<<<SQ_CTAS, SQ_THREADS, SQ_SMEM_BYTES, s>>>(Q);
        } else {
            SerialParams P;
            P.n = n; P.ndots = ndots;
            for (int d = 0; d < 2; ++d) {
                P.a[d] = aa[d]; P.b[d] = bb[d]; P.pre[d] = nullptr;
                if (d < ndots && aa[d] == bb[d]) {
                    SquaresParams Q{n, aa[d], sc->sq + d, nullptr, FIN_NONE};
                    sum_squares_serial_kernel<<<SQ_CTAS, SQ_THREADS, SQ_SMEM_BYTES, s>>>(Q);
                    SMM_COUNT_LAUNCH(1);
                    P.pre[d] = sc->sq + d;
                }
            }
            P.state = state; P.finish = finish; P.out_dev = out_dev;
            dot_serial_kernel<<<1, SER_THREADS, 0, s>>>(P);
        }
    }
    SMM_COUNT_LAUNCH(1);
    SMM_CUDA(cudaGetLastError());
    return SMM_OK;
}

// ---------------------------------------------------------------------------------------------------
// C ABI: Vector::operator* (H:305-328)
// ---------------------------------------------------------------------------------------------------
namespace {
smm_workspace* g_dot_ws[64] = {nullptr};

int dot_workspace(smm_workspace** out) {
    int dev = 0;
    SMM_CUDA(cudaGetDevice(&dev));
    if (!g_dot_ws[dev]) {
        smm_csr fake;                                          // matrix-free workspace
        fake.device = dev;
        fake.num_blocks = 0;
        smm_workspace* ws = nullptr;
        SMM_TRY(smm_workspace_get(&fake, &ws));
        g_dot_ws[dev] = ws;
        fake.ws = nullptr;
    }
    *out = g_dot_ws[dev];
    return SMM_OK;
}
}  // namespace

extern "C" {

int smm_dot_dev(int64_t n, const float* a_dev, const float* b_dev, int reduction_mode, float* out_host, void* stream) {
    if (n < 0 || !out_host || (n && (!a_dev || !b_dev))) { smm_set_error("smm_dot: bad arguments"); return SMM_E_INVALID; }
    cudaStream_t s = stream ? (cudaStream_t)stream : smm_default_stream();
    float res[2] = {0.f, 0.f};
    if (reduction_mode == SMM_REDUCE_FAST) {
        smm_workspace* ws = nullptr;
        SMM_TRY(dot_workspace(&ws));
        SMM_CUDA(cudaMemsetAsync(ws->state, 0, sizeof(SolveState), s));
        VecArgs v;
        v.n = n; v.in[0] = a_dev; v.in[1] = b_dev; v.state = ws->state; v.finish = FIN_STORE; v.slot = 0; v.ws = ws;
        SMM_TRY(smm_launch_vec(VEC_DOT2, v, s));
        SMM_CUDA(cudaMemcpyAsync(res, (const char*)ws->state + offsetof(SolveState, scratch), 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
        SMM_CUDA(cudaStreamSynchronize(s));
    } else if (reduction_mode == SMM_REDUCE_REFERENCE_TREE || reduction_mode == SMM_REDUCE_REFERENCE_SERIAL) {
        std::lock_guard<std::mutex> lk(g_scratch_mu);          // the per-device scratch: one stand-alone dot at a time
        DotScratch* sc = nullptr;
        SMM_TRY(scratch_for(1ll << smm_tree_depth(n), nullptr, &sc));
        SMM_TRY(smm_launch_dot_ref(reduction_mode, n, 1, a_dev, b_dev, a_dev, b_dev, nullptr, FIN_NONE, sc->out, s, nullptr));
        SMM_CUDA(cudaMemcpyAsync(res, sc->out, 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
        SMM_CUDA(cudaStreamSynchronize(s));
    } else {
        smm_set_error("smm_dot: unknown reduction mode %d", reduction_mode);
        return SMM_E_INVALID;
    }
    *out_host = res[0];
    return SMM_OK;
}

int smm_dot(int64_t n, const float* a, const float* b, int reduction_mode, float* out) {
    if (n < 0 || !out || (n && (!a || !b))) { smm_set_error("smm_dot: bad arguments"); return SMM_E_INVALID; }
    float *da = nullptr, *db = nullptr;
    const size_t bytes = sizeof(float) * (size_t)(n ? n : 1);
    const bool same = a == b;                                  // a vector with itself: one device copy (and the sum-of-squares path)
    SMM_CUDA(cudaMalloc(&da, bytes));
    if (!same && cudaMalloc(&db, bytes) != cudaSuccess) { cudaFree(da); return smm_cuda_fail(cudaGetLastError(), "cudaMalloc", __FILE__, __LINE__); }
    int rc = SMM_OK;
    if (n && (cudaMemcpy(da, a, bytes, cudaMemcpyHostToDevice) != cudaSuccess || (!same && cudaMemcpy(db, b, bytes, cudaMemcpyHostToDevice) != cudaSuccess)))
        rc = smm_cuda_fail(cudaGetLastError(), "memcpy", __FILE__, __LINE__);
    if (rc == SMM_OK) rc = smm_dot_dev(n, da, same ? da : db, reduction_mode, out, nullptr);
    cudaFree(da);
    cudaFree(db);
    return rc;
}

}  // extern "C"
